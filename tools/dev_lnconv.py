"""Dev check of the dense-conv forms (rf_lnconv.cu): bf16 forward vs the fp32 parity engine on odd-sized frames."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf

def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    peak = max(b.abs().max().item(), 1e-6)
    return 10 * torch.log10(torch.tensor(peak * peak / max(mse, 1e-30))).item()

dev = torch.device("cuda", 0)
for (H, W, scale, B) in [(64, 64, 1.0, 1), (272, 400, 1.5, 2), (544, 816, 1.0, 1)]:
    m32 = rf.RawFormer(model_size="S", precision="fp32")
    sd = T.make_state_dict(m32, seed=1234, scale=scale)
    m32.load_state_dict(sd); m32 = m32.to(dev).eval()
    m16 = rf.RawFormer(model_size="S", precision="bf16")
    m16.load_state_dict(sd); m16 = m16.to(dev).eval()
    x = torch.from_numpy(T.gen_input("rand", (B, 1, H, W), 3)).to(dev)
    with torch.no_grad():
        y32 = m32(x); y16 = m16(x); y16b = m16(x)
    torch.cuda.synchronize()
    print(f"{H}x{W} scale {scale} B {B}: psnr(bf16, fp32) = {psnr(y16, y32):.2f} dB, max abs {float((y16-y32).abs().max()):.4f}, "
          f"finite {bool(torch.isfinite(y16).all())}, repeat-identical {bool(torch.equal(y16, y16b))}", flush=True)
