"""Run-to-run reproducibility of ONE TransformerBlock (bf16, dim 64) on a large feature map, with the shared workspace scrambled
between the runs; prints where two runs differ."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bayer_low_light_image_enhancement_b200 as rf
from bayer_low_light_image_enhancement_b200 import modules as M
dev = torch.device("cuda", 0)
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 64
H, W = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(1)
blk = M.TransformerBlock(dim, 8, 2, True)
blk.precision = "bf16"
blk = blk.to(dev).eval()
for p in blk.parameters():
    if p.dim() > 1: torch.nn.init.normal_(p, std=0.2)
if hasattr(blk, "set_precision"): blk.set_precision("bf16")
x = torch.randn(1, dim, H, W, device=dev) * 0.5
small = torch.randn(1, dim, 96, 160, device=dev)
outs = []
with torch.no_grad():
    for k in range(4):
        outs.append(blk(x).clone())
        blk(small * (k + 1))                       # scrambles the head of the shared workspace
torch.cuda.synchronize()
print("dtype code", blk._dtype())
for i, o in enumerate(outs[1:], 1):
    d = (o - outs[0]).abs()
    nz = int((d > 0).sum())
    if nz:
        idx = (d > 0).nonzero()
        ys, xs, cs = idx[:, 2], idx[:, 3], idx[:, 1]
        print(f"run {i}: {nz} of {d.numel()} differ, max {float(d.max()):.3e}; y {int(ys.min())}..{int(ys.max())} x {int(xs.min())}..{int(xs.max())} c {int(cs.min())}..{int(cs.max())}; first {idx[0].tolist()}")
    else:
        print(f"run {i}: identical")
