import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T
import bayer_low_light_image_enhancement_b200 as rf
dev = torch.device("cuda", 0)
C, hw = int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3]))
att = rf.Attention(C, 8, True)
att.load_state_dict(T.make_state_dict(att, seed=C, scale=1.5))
att = att.to(dev); att.precision = "bf16"
x = torch.randn(1, C, *hw, device=dev)
y = att(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
