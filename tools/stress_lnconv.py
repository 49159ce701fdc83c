"""Stress of the persistent dense-conv kernels (rf_lnconv.cu) over frame sizes that give 1, 2, 3, ... tiles per CTA (the
pipelines' prologues, ring wrap-arounds and tails), eager and as CUDA graphs, repeated: every run must finish (run it under
`timeout`), be finite, repeat bit-identically and stay within the bf16 bar of the fp32 parity engine."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    peak = max(b.abs().max().item(), 1e-6)
    return 99.0 if mse == 0 else 10 * torch.log10(torch.tensor(peak * peak / mse)).item()


dev = torch.device("cuda", 0)
sizes = [(32, 32, 1), (64, 64, 1), (64, 96, 3), (160, 176, 1), (320, 304, 1), (384, 400, 2), (512, 608, 1), (544, 816, 1),
         (1024, 1024, 1), (1424, 2128, 1), (2848, 4256, 1)]
worst = 99.0
for variant in ("flca", "ml"):
    cls = rf.RawFormer if variant == "flca" else rf.multilevel.RawFormer
    m32 = cls(model_size="S", precision="fp32")
    sd = T.make_state_dict(m32, seed=77, scale=1.0)
    m32.load_state_dict(sd); m32 = m32.to(dev).eval()
    m16 = cls(model_size="S", precision="bf16")
    m16.load_state_dict(sd); m16 = m16.to(dev).eval()
    for (H, W, B) in sizes:
        if variant == "ml" and H * W > 1100 * 1100:
            continue
        x = torch.from_numpy(T.gen_input("rand", (B, 1, H, W), H + W)).to(dev)
        t0 = time.time()
        with torch.no_grad():
            ref = m32(x)
            outs = [m16(x).clone() for _ in range(4)]
            m16.enable_cuda_graphs()
            outs += [m16(x).clone() for _ in range(4)]
            m16.enable_cuda_graphs(False)
        torch.cuda.synchronize()
        same = all(torch.equal(outs[0], o) for o in outs[1:])
        p = psnr(outs[0], ref)
        worst = min(worst, p)
        ok = bool(torch.isfinite(outs[0]).all()) and same and p >= 50.0
        print(f"{variant} {H}x{W} B{B}: psnr {p:.1f} dB, identical {same}, {time.time() - t0:.1f} s {'ok' if ok else 'FAIL'}", flush=True)
        if not ok:
            sys.exit(1)
print("stress ok, worst psnr", round(worst, 1))
