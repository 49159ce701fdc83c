"""Golden vectors for the TrueColor head / tail by EXECUTING THE REFERENCE CLASSES
(``EnhancedBayerProcessor`` / ``CameraAwareColorCorrection`` of TrueColorRawFormer.py and BayerTORGBColorMultiLvl.py) on CPU.

    python tests/golden/make_golden_truecolor.py       # needs /root/reference; writes tests/golden/truecolor.npz

Inputs and weights are reproducible from seeds (tests/rf_testlib.py), so only the reference outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rf_testlib as T  # noqa: E402
from make_golden import _stub, load_ref  # noqa: E402

CASES = T.TRUECOLOR_CASES


@torch.no_grad()
def main():
    torch.manual_seed(0)
    _stub("ptflops", get_model_complexity_info=lambda *a, **k: (None, None))
    mods = {0: load_ref("TrueColorRawFormer.py", "ref_truecolor"), 1: load_ref("BayerTORGBColorMultiLvl.py", "ref_truecolor_ml")}
    out = {}
    for name, variant, kind, shape, seed, scale in CASES:
        ref_mod = mods[variant]
        ours = T.build_truecolor(kind, variant)
        ref = (ref_mod.EnhancedBayerProcessor() if kind.startswith("head") else ref_mod.CameraAwareColorCorrection()).eval()
        assert list(ref.state_dict()) == list(ours.state_dict()), (list(ref.state_dict()), list(ours.state_dict()))
        assert all(tuple(a.shape) == tuple(b.shape) for a, b in zip(ref.state_dict().values(), ours.state_dict().values()))
        # default initialisation of the non-random parameters must agree (wb_gains, color_matrix, gamma)
        for k, v in ref.state_dict().items():
            if k in ("wb_gains", "color_matrix", "gamma", "gamma_param", "y_weights"):
                assert torch.equal(v, ours.state_dict()[k]), k
        sd = T.make_truecolor_state_dict(ours, seed, scale)
        ref.load_state_dict(sd, strict=True)
        x = torch.from_numpy(T.truecolor_input(kind, shape, seed))
        res = ref(x)
        res = res if isinstance(res, tuple) else (res,)
        for i, r in enumerate(res):
            out[f"{name}.out{i}"] = r.numpy()
        print(name, [tuple(r.shape) for r in res], [float(r.abs().max()) for r in res])
    path = os.path.join(HERE, "truecolor.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
