"""Run-to-run reproducibility of the full-frame bf16 forward (eager and CUDA graph): prints how many outputs differ and where."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
dev = torch.device("cuda", 0)
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2848, 4256)
size = sys.argv[3] if len(sys.argv) > 3 else "S"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
m = rf.RawFormer(model_size=size, precision="bf16")
m.load_state_dict(T.make_state_dict(m, seed=77, scale=1.0)); m = m.to(dev).eval()
x = torch.from_numpy(T.gen_input("rand", (B, 1, H, W), H + W)).to(dev)
from bayer_low_light_image_enhancement_b200 import _lib
print("workspace GB", _lib.load().rf_rawformer_workspace_bytes(m.dim, 1, 0, B, H, W) / 1e9)
with torch.no_grad():
    outs = [m(x).clone() for _ in range(5)]
    if not os.environ.get("REPRO_NO_GRAPH"):
        m.enable_cuda_graphs()
        outs += [m(x).clone() for _ in range(5)]
torch.cuda.synchronize()
for i, o in enumerate(outs[1:], 1):
    d = (o - outs[0]).abs()
    nz = int((d > 0).sum())
    if nz:
        if i == 1:
            rowm = d.amax(dim=(0, 1, 3))            # max over batch, channel, column -> per output row
            big = (rowm > 0.25 * rowm.max()).nonzero().flatten()
            print("rows with large differences:", int(big.min()), "..", int(big.max()), "count", int(big.numel()), "; per-row max (every 178th row):", [f"{float(v):.1e}" for v in rowm[::178]])
            colm = d.amax(dim=(0, 1, 2))
            print("per-col max (every 266th col):", [f"{float(v):.1e}" for v in colm[::266]])
        idx = (d > 0).nonzero()
        print(f"run {i} ({'graph' if i >= 5 else 'eager'}): {nz} elements differ, max abs {float(d.max()):.3e}, rows {int(idx[:,2].min())}..{int(idx[:,2].max())}, cols {int(idx[:,3].min())}..{int(idx[:,3].max())}")
    else:
        print(f"run {i} ({'graph' if i >= 5 else 'eager'}): identical")
