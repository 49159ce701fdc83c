"""Pipeline-wraparound check of the bf16 GEMM / depthwise kernels: sub-modules on an image large enough that every
persistent CTA walks many tiles; bf16 mode against fp32 mode (PSNR)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
blk = T.build_block("flca", C)
blk.load_state_dict(T.make_state_dict(blk, seed=50 + C, scale=1.5), strict=True)
blk = blk.to(dev).eval()
feat = torch.randn(1, C, hw, hw + 16, device=dev)
x_ds = torch.rand(1, 4, hw, hw + 16, device=dev)
y, cr, cb = rf.BayerLumaChroma().to(dev)(x_ds)
def run(prec):
    for m in blk.modules():
        if hasattr(m, "precision"): m.precision = prec
    with torch.no_grad():
        o = {"ffn": blk.Transformer.ffn(feat), "attn": blk.Transformer.attn(feat), "trans": blk.Transformer(feat),
             "out": blk(feat, y, cr, cb)}
    torch.cuda.synchronize()
    return {k: v.float().cpu().numpy() for k, v in o.items()}
ref = run("fp32"); got = run("bf16")
for k in ref:
    rng = float(ref[k].max() - ref[k].min()); mse = float(np.mean((ref[k] - got[k]) ** 2))
    print(k, "PSNR %.1f dB" % (10 * np.log10(rng * rng / max(mse, 1e-30))), "finite", bool(np.isfinite(got[k]).all()))
