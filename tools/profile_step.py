"""One warm-up + N forwards of RawFormer on a full SID Sony frame (for ncu / quick timing)."""
import argparse, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
ap = argparse.ArgumentParser()
ap.add_argument("--size", default="S"); ap.add_argument("--precision", default="bf16"); ap.add_argument("--variant", default="flca")
ap.add_argument("--iters", type=int, default=1); ap.add_argument("--h", type=int, default=2848); ap.add_argument("--w", type=int, default=4256)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cls = rf.RawFormer if a.variant == "flca" else rf.multilevel.RawFormer
m = cls(model_size=a.size, precision=a.precision)
m.load_state_dict(T.make_state_dict(m, seed=1234)); m = m.to(dev).eval()
x = torch.rand(1, 1, a.h, a.w, device=dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): y = m(x)
    e1.record(); torch.cuda.synchronize()
print("ms/frame", e0.elapsed_time(e1) / a.iters, float(y.abs().max()))
import time
with torch.no_grad():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): y = m(x)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms/frame", (t1 - t0) / 5 * 1e3, "total ms/frame", (t2 - t0) / 5 * 1e3)
