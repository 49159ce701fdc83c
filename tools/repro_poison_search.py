"""Binary search for the workspace bytes whose (stale) content reaches the output: poison [lo, hi) with 0xFF, zero elsewhere."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
from bayer_low_light_image_enhancement_b200 import _lib
dev = torch.device("cuda", 0)
H, W, size = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
m = rf.RawFormer(model_size=size, precision="bf16")
m.load_state_dict(T.make_state_dict(m, seed=77, scale=1.0)); m = m.to(dev).eval()
x = torch.from_numpy(T.gen_input("rand", (1, 1, H, W), H + W)).to(dev)
nbytes = _lib.load().rf_rawformer_workspace_bytes(m.dim, m._dtype(), 0, 1, H, W)
def bad(lo, hi):
    ws.zero_(); ws[lo:hi].fill_(0xFF); torch.cuda.synchronize()
    o = m(x); torch.cuda.synchronize()
    return bool((~torch.isfinite(o)).any())
with torch.no_grad():
    m(x); ws = _lib.shared_workspace(nbytes, dev)
    print("workspace bytes", ws.numel(), "whole poisoned ->", bad(0, ws.numel()))
    lo, hi = 0, ws.numel()
    while hi - lo > 256:
        mid = (lo + hi) // 2 // 256 * 256
        if bad(lo, mid): hi = mid
        elif bad(mid, hi): lo = mid
        else:
            print("neither half alone reproduces it at", lo, mid, hi); break
    print("stale bytes reach the output from workspace range", lo, hi)
    # extent of the region: grow to the right / left while still bad
    for step in (1 << 20, 1 << 16, 1 << 12, 256):
        while lo - step >= 0 and bad(lo - step, lo): lo -= step
    for step in (1 << 20, 1 << 16, 1 << 12, 256):
        while hi + step <= ws.numel() and bad(hi, hi + step): hi += step
    print("contiguous poisoned-sensitive region", lo, hi, "bytes", hi - lo)
