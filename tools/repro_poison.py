"""Poison test: fill the shared workspace with a byte pattern before a forward and compare with a forward over a zeroed
workspace: a difference means some kernel reads workspace memory that the same forward has not written."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
from bayer_low_light_image_enhancement_b200 import _lib
dev = torch.device("cuda", 0)
H, W = int(sys.argv[1]), int(sys.argv[2])
size = sys.argv[3] if len(sys.argv) > 3 else "S"
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
variant = sys.argv[5] if len(sys.argv) > 5 else "flca"
m = (rf.RawFormer if variant == "flca" else rf.multilevel.RawFormer)(model_size=size, precision=prec)
m.load_state_dict(T.make_state_dict(m, seed=77, scale=1.0)); m = m.to(dev).eval()
x = torch.from_numpy(T.gen_input("rand", (1, 1, H, W), H + W)).to(dev)
nbytes = _lib.load().rf_rawformer_workspace_bytes(m.dim, m._dtype(), m.variant, 1, H, W)
with torch.no_grad():
    m(x)                                             # sizes the shared workspace, packs the weights
    ws = _lib.shared_workspace(nbytes, dev)
    outs = []
    for pat in (0x00, 0xFF, 0x00, 0x3C, 0x7F):
        ws.fill_(pat)
        torch.cuda.synchronize()
        outs.append(m(x).clone())
torch.cuda.synchronize()
for i, (pat, o) in enumerate(zip((0x00, 0xFF, 0x00, 0x3C, 0x7F), outs)):
    d = (o - outs[0]).abs()
    bad = int((~torch.isfinite(o)).sum())
    print(f"pattern 0x{pat:02X}: non-finite {bad}, differs from the zeroed-workspace run in {int((d > 0).sum())} elements, max {float(torch.nan_to_num(d).max()):.3e}")
