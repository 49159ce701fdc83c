"""Diagnostic: eager vs graph-replayed LocalBands against the whole-frame forward (run on a B200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf


def psnr(a, b):
    rng = float(b.max() - b.min())
    mse = float(((a - b).double() ** 2).mean())
    return 99.0 if mse == 0 else 10 * __import__("math").log10(rng * rng / mse)


dev = torch.device("cuda", 0)
m = rf.RawFormer(dim=32, precision="bf16")
m.load_state_dict(T.make_state_dict(m, seed=5, scale=2.0), strict=True)
m = m.to(dev).eval()
H, W, n = 192, 128, 3
frames = [torch.from_numpy(T.gen_input(k, (1, 1, H, W), s)).to(dev) for k, s in (("rand", 21), ("dark", 22), ("rand", 21))]
eager, graphs = rf.LocalBands(m, H, W, n), rf.LocalBands(m, H, W, n, graphs=True)
buf = frames[0].clone()
for i, f in enumerate(frames):
    buf.copy_(f)
    with torch.no_grad():
        whole = m(buf).clone()
    a1 = eager(buf).clone()
    a2 = eager(buf).clone()
    b1 = graphs(buf).clone()
    b2 = graphs(buf).clone()
    print(f"frame {i}: range {float(whole.max() - whole.min()):.3f}  eager/eager {psnr(a1, a2):.1f}  eager/whole {psnr(a1, whole):.1f}  "
          f"graph/whole {psnr(b1, whole):.1f} {psnr(b2, whole):.1f}  graph/eager {psnr(b1, a1):.1f}  maxdiff g/e {float((b1 - a1).abs().max()):.4f} "
          f"e/e {float((a1 - a2).abs().max()):.4f} e/w {float((a1 - whole).abs().max()):.4f}", flush=True)
