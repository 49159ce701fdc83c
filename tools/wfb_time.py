"""Timing of the WFB block pieces at the LL band of a full SID Sony frame (C = 32, 712 x 1064): FEB, FFAB, Illumination_Estimator,
rfft2 / irfft2 alone.  (the dense-DFT transform is a first, untuned implementation.)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
from bayer_low_light_image_enhancement_b200 import wfb
dev = torch.device("cuda", 0)
C, H, W = 32, 712, 1064
x = torch.randn(1, C, H, W, device=dev)

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

with torch.no_grad():
    spec = wfb.rfft2_ortho(x)
    print(f"rfft2 [{C},{H},{W}]: {timeit(lambda: wfb.rfft2_ortho(x)):.2f} ms; irfft2: {timeit(lambda: wfb.irfft2_ortho(spec, W)):.2f} ms")
    for name, m in (("FEB", rf.FEB(C)), ("FFAB", rf.FFAB(C)), ("Illumination_Estimator", rf.Illumination_Estimator(C, C + 1, C))):
        m.load_state_dict(T.make_state_dict(m, seed=3)); m = m.to(dev).eval()
        print(f"{name}({C}) on [1,{C},{H},{W}]: {timeit(lambda: m(x)):.2f} ms")
