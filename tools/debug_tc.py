"""A/B check of the tcgen05 contraction kernels against the CUDA-core kernels (both bf16 storage) per sub-module."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
from bayer_low_light_image_enhancement_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)

def stat(name, a, b):
    a = a.float().cpu().numpy(); b = b.float().cpu().numpy()
    d = np.abs(a - b); rng = float(b.max() - b.min()) + 1e-12
    print(f"{name:50s} max-abs {d.max():.4e}  rel-to-range {d.max()/rng:.4e}  mean {d.mean():.3e} finite={np.isfinite(a).all()}", flush=True)

def ab(name, fn):
    lib.rf_set_tcgen05(0); ref = fn()
    lib.rf_set_tcgen05(1); out = fn()
    torch.cuda.synchronize()
    stat(name, out, ref)

for C, hw in ((32, (16, 24)), (64, (12, 20)), (96, (9, 14)), (48, (9, 14)), (128, (16, 16)), (256, (6, 10)), (512, (4, 6)), (32, (64, 96))):
    blk = rf.Conv_Transformer(C)
    blk.load_state_dict(T.make_state_dict(blk, seed=C, scale=1.5))
    blk = blk.to(dev).eval(); blk.precision = "bf16"
    for m in blk.modules():
        if hasattr(m, "precision"): m.precision = "bf16"
    x = torch.randn(2, C, *hw, device=dev)
    xds = torch.rand(2, 4, hw[0] * 2, hw[1] * 2, device=dev)
    y, cr, cb = rf.BayerLumaChroma().to(dev)(xds)
    ab(f"C={C} {hw} ffn (plain gemm x2)", lambda: blk.Transformer.ffn(x))
    ab(f"C={C} {hw} attn (gemm + per-image gemm)", lambda: blk.Transformer.attn(x))
    ab(f"C={C} {hw} transformer", lambda: blk.Transformer(x))
    ab(f"C={C} {hw} conv_transformer (cat + conv3x3)", lambda: blk(x, y, cr, cb))
    if C % 16 == 0 and hw[0] % 2 == 0 and hw[1] % 2 == 0:
        ds = rf.Downsample(C); ds.load_state_dict(T.make_state_dict(ds, seed=3, scale=1.5)); ds = ds.to(dev); ds.precision = "bf16"
        ab(f"C={C} {hw} downsample (conv3x3 + unshuffle)", lambda: ds(x))

for size in ("S", "B", "L"):
    m = rf.RawFormer(model_size=size, precision="bf16")
    m.load_state_dict(T.make_state_dict(m, seed=5, scale=1.5)); m = m.to(dev).eval()
    x = torch.rand(1, 1, 128, 192, device=dev)
    ab(f"RawFormer-{size} 128x192 whole model", lambda: m(x))
print("done")
