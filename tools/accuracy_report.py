"""bf16-mode accuracy against the reference goldens (tests/golden): PSNR per whole-model case, relative to the output range.
Run on the GPU box; prints one line per case and a JSON summary."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rf_testlib as T

def psnr(a, b, rng):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * np.log10(rng ** 2 / mse)

dev = torch.device("cuda", 0)
res = {}
for case in T.MODEL_CASES:
    name, variant, dim, H, W, kind, seed, scale, b = case
    ref = T.load_golden(name)["out"]
    rng = max(float(ref.max() - ref.min()), 1e-6)
    row = {}
    for prec in ("fp32", "bf16"):
        m = T.build_model(variant, dim, precision=prec)
        m.load_state_dict(T.make_state_dict(m, seed=1234 + seed, scale=scale), strict=True)
        m = m.to(dev).eval()
        with torch.no_grad():
            out = m(torch.from_numpy(T.gen_input(kind, (b, 1, H, W), seed)).to(dev)).float().cpu().numpy()
        row[prec] = round(psnr(out, ref, rng), 2)
        row[prec + "_maxabs"] = float(np.abs(out - ref).max())
    res[name] = row
    print(name, row)
print(json.dumps(res))
