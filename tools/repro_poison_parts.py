"""Poison test of the block-level entry points (bf16): zeroed vs 0xFF-filled shared workspace."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bayer_low_light_image_enhancement_b200 import _lib, modules as M
dev = torch.device("cuda", 0)
C, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
x = torch.randn(1, C, H, W, device=dev) * 0.5
y = torch.rand(1, 1, 2 * H // 2, 2 * W // 2, device=dev); cr = torch.rand_like(y) - 0.5; cb = torch.rand_like(y) - 0.5
parts = {
    "FLCA": (M.FLCA(C), lambda m: m(x, y, cr, cb)),
    "Attention": (M.Attention(C, 8, True), lambda m: m(x)),
    "conv_ffn": (M.conv_ffn(C, 2 * C, C), lambda m: m(x)),
    "TransformerBlock": (M.TransformerBlock(C, 8, 2, True), lambda m: m(x)),
    "Conv_Transformer": (M.Conv_Transformer(C), lambda m: m(x, y, cr, cb)),
    "Downsample": (M.Downsample(C), lambda m: m(x)),
}
for name, (mod, fn) in parts.items():
    mod.precision = "bf16"
    mod = mod.to(dev).eval()
    try:
        with torch.no_grad():
            fn(mod); torch.cuda.synchronize()
            ws = _lib._shared_ws.get(1, dev) if hasattr(_lib, "_shared_ws") else None
            res = []
            for pat in (0x00, 0xFF):
                ws.fill_(pat); torch.cuda.synchronize()
                res.append(fn(mod).clone())
        torch.cuda.synchronize()
        bad = int((~torch.isfinite(res[1])).sum()); d = torch.nan_to_num((res[1] - res[0]).abs())
        print(f"{name}: workspace {ws.numel() / 1e6:.0f} MB, poisoned run non-finite {bad} of {res[1].numel()}, differing {int((d > 0).sum())}")
    except Exception as e:
        print(f"{name}: {type(e).__name__}: {str(e)[:150]}")
