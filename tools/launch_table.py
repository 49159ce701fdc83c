"""Per-launch table of one full-frame forward (CUDA events around every launch): name, ms, algorithmic GB/s, TFLOP/s."""
import argparse, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
ap = argparse.ArgumentParser()
ap.add_argument("--size", default="S"); ap.add_argument("--variant", default="flca"); ap.add_argument("--min-ms", type=float, default=0.02)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cls = rf.RawFormer if a.variant == "flca" else rf.multilevel.RawFormer
m = cls(model_size=a.size, precision="bf16")
m.load_state_dict(T.make_state_dict(m, seed=1234)); m = m.to(dev).eval()
x = torch.rand(1, 1, 2848, 4256, device=dev)
with torch.no_grad():
    for _ in range(3): m(x)
    _, L = m.forward_profiled(x)
    _, L = m.forward_profiled(x)
tot = sum(l["ms"] for l in L)
print(f"# {len(L)} launches, sum {tot:.3f} ms")
for i, l in enumerate(L):
    if l["ms"] >= a.min_ms:
        gbs = l["bytes"] / l["ms"] / 1e6 if l["ms"] > 0 else 0
        tf = l["flops"] / l["ms"] / 1e9 if l["ms"] > 0 else 0
        print(f"{i:4d} {l['name']:18s} {l['ms']*1e3:8.1f} us  {l['bytes']/1e6:8.1f} MB  {gbs:7.0f} GB/s  {tf:7.1f} TFLOP/s")
