"""TrueColor head / tail: ``EnhancedBayerProcessor`` and ``CameraAwareColorCorrection`` of
TrueColorRawFormer.py (variant 0) and BayerTORGBColorMultiLvl.py (variant 1).

CPU: the numpy oracle against goldens produced by executing the reference classes (tests/golden/make_golden_truecolor.py).
GPU: the CUDA kernels, through the module mirrors and the C ABI, against the same goldens (fp32, max-abs <= 2e-5) and against
the oracle on a larger ragged frame."""
import numpy as np
import pytest
import torch

import rf_testlib as T
from oracle import rawformer_oracle as O

GOLD = T.load_golden("truecolor")
TOL = 2e-5


def _oracle(kind, variant, sd, x):
    sdn = T.sd_numpy(sd)
    if kind.startswith("head"):
        return O.enhanced_bayer_processor(sdn, x, variant)
    return (O.camera_aware_color_correction(sdn, x, variant),)


@pytest.mark.parametrize("case", T.TRUECOLOR_CASES, ids=[c[0] for c in T.TRUECOLOR_CASES])
def test_oracle_matches_reference(case):
    name, variant, kind, shape, seed, scale = case
    sd = T.make_truecolor_state_dict(T.build_truecolor(kind, variant), seed, scale)
    outs = _oracle(kind, variant, sd, T.truecolor_input(kind, shape, seed))
    for i, o in enumerate(outs):
        ref = GOLD[f"{name}.out{i}"]
        assert o.shape == ref.shape and o.dtype == np.float32
        err = float(np.abs(o.astype(np.float64) - ref).max())
        assert err <= TOL, f"{name} out{i}: max-abs {err:.3e}"


def test_state_dict_names_match_reference_layout():
    import bayer_low_light_image_enhancement_b200 as rf

    head = rf.truecolor.EnhancedBayerProcessor()
    assert [k for k in head.state_dict()][:3] == ["wb_gains", "color_matrix", "y_weights"]
    assert tuple(head.demosaic_refine[0].weight.shape) == (32, 4, 3, 3) and tuple(head.demosaic_refine[2].weight.shape) == (4, 32, 3, 3)
    ml = rf.truecolor.multilevel.EnhancedBayerProcessor()
    assert tuple(ml.demosaic_refine[0].weight.shape) == (32, 3, 3, 3) and ml.wb_gains.tolist() == pytest.approx([1.8, 1.0, 1.0, 1.6])
    assert "gamma" in rf.truecolor.CameraAwareColorCorrection().state_dict()
    assert "gamma_param" in rf.truecolor.multilevel.CameraAwareColorCorrection().state_dict()
    with pytest.raises(RuntimeError):
        head(torch.zeros(1, 4, 8, 8))          # CPU tensor: there is no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("case", T.TRUECOLOR_CASES, ids=[c[0] for c in T.TRUECOLOR_CASES])
def test_cuda_matches_reference(case):
    name, variant, kind, shape, seed, scale = case
    dev = torch.device("cuda", 0)
    m = T.build_truecolor(kind, variant)
    m.load_state_dict(T.make_truecolor_state_dict(m, seed, scale), strict=True)
    m = m.to(dev).eval()
    with torch.no_grad():
        outs = m(torch.from_numpy(T.truecolor_input(kind, shape, seed)).to(dev))
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        ref = GOLD[f"{name}.out{i}"]
        got = o.float().cpu().numpy()
        assert got.shape == ref.shape and np.isfinite(got).all()
        err = float(np.abs(got.astype(np.float64) - ref).max())
        assert err <= TOL, f"{name} out{i}: max-abs {err:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_cuda_matches_oracle_ragged_frame(variant):
    """A frame that is not a multiple of the thread-block size in either direction, batch 2: head then tail."""
    dev = torch.device("cuda", 0)
    head, tail = T.build_truecolor("head", variant), T.build_truecolor("tail", variant)
    sdh, sdt = T.make_truecolor_state_dict(head, 50 + variant, 1.5), T.make_truecolor_state_dict(tail, 60 + variant, 2.0)
    head.load_state_dict(sdh, strict=True)
    tail.load_state_dict(sdt, strict=True)
    x = T.gen_input("rand", (2, 4, 77, 131), 70 + variant)
    img = T.gen_input("rand", (2, 3, 77, 131), 80 + variant) * 1.4 - 0.2
    with torch.no_grad():
        got_h = head.to(dev).eval()(torch.from_numpy(x).to(dev))
        got_t = tail.to(dev).eval()(torch.from_numpy(img).to(dev))
    ref_h = O.enhanced_bayer_processor(T.sd_numpy(sdh), x, variant)
    ref_t = O.camera_aware_color_correction(T.sd_numpy(sdt), img, variant)
    for i, (g, r) in enumerate(zip(got_h, ref_h)):
        err = float(np.abs(g.float().cpu().numpy().astype(np.float64) - r).max())
        assert err <= TOL, f"head variant {variant} out{i}: max-abs {err:.3e}"
    err = float(np.abs(got_t.float().cpu().numpy().astype(np.float64) - ref_t).max())
    assert err <= TOL, f"tail variant {variant}: max-abs {err:.3e}"
    assert float(got_t.min()) >= 0.0 and float(got_t.max()) <= 1.0
