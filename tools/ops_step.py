"""One call of every stand-alone wavelet / shuffle operator at the README shape (for ncu)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayer_low_light_image_enhancement_b200 as rf
dev = torch.device("cuda", 0)
x = torch.randn(1, 32, 1424, 2128, device=dev)
sub = torch.randn(1, 128, 712, 1064, device=dev)
with torch.no_grad():
    for _ in range(2):
        a = rf.CustomDWT().to(dev)(x); b = rf.CustomIDWT().to(dev)(sub)
        c = rf.dwt_init(x); d = rf.iwt_init(sub.view(4, 32, 712, 1064))
        e = rf.HaarDWT().to(dev)(x); f = rf.downshuffle(x, 2); g = rf.PixelShuffle(2)(sub)
torch.cuda.synchronize()
print("ok", float(a.abs().max()), float(g.abs().max()))
