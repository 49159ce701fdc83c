"""Dev: one full-frame S forward with the per-phase counters of k_lnconv (RAWFORMER_B200_LNCONV_DBG=1)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
dev = torch.device("cuda", 0)
m = rf.RawFormer(model_size="S", precision="bf16")
m.load_state_dict(T.make_state_dict(m, seed=1234)); m = m.to(dev).eval()
x = torch.rand(1, 1, 2848, 4256, device=dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    print("---- second forward", file=sys.stderr, flush=True)
    m(x); torch.cuda.synchronize()
