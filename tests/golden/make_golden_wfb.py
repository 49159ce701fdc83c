"""Golden vectors for the WFB "WMB" block pieces by EXECUTING THE REFERENCE CLASSES on CPU:
``FEB`` / ``ProcessBlock`` / ``FFAB`` of RawFomer_WFB_FFAB/blocks.py and ``Illumination_Estimator`` of
RawFomer_WFB_FFAB/model.py (timm / mamba_ssm / ptflops are import-only there and stubbed, see make_golden.py).

    python tests/golden/make_golden_wfb.py       # needs /root/reference; writes tests/golden/wfb.npz

Inputs and weights are reproducible from seeds (tests/rf_testlib.py), so only the reference outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rf_testlib as T  # noqa: E402
from make_golden import load_reference_modules  # noqa: E402


@torch.no_grad()
def main():
    torch.manual_seed(0)
    _, _, _, blocks, model = load_reference_modules()
    assert blocks is not None and model is not None
    out = {}
    for name, kind, c, shape, seed, scale in T.WFB_CASES:
        ours = T.build_wfb(kind, c)
        ref = {"feb": lambda: blocks.FEB(c), "pb": lambda: blocks.ProcessBlock(c), "ffab": lambda: blocks.FFAB(c),
               "illu": lambda: model.Illumination_Estimator(c, n_fea_in=c + 1, n_fea_out=c)}[kind]().eval()
        assert list(ref.state_dict()) == list(ours.state_dict()), (list(ref.state_dict()), list(ours.state_dict()))
        assert all(tuple(a.shape) == tuple(b.shape) for a, b in zip(ref.state_dict().values(), ours.state_dict().values()))
        sd = T.make_state_dict(ours, seed=seed, scale=scale)
        ref.load_state_dict(sd, strict=True)
        x = torch.from_numpy(T.wfb_input(kind, shape, seed))
        res = ref(x)
        res = res if isinstance(res, tuple) else (res,)
        for i, r in enumerate(res):
            out[f"{name}.out{i}"] = r.numpy()
        print(name, [tuple(r.shape) for r in res], [float(r.abs().max()) for r in res])
    path = os.path.join(HERE, "wfb.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
