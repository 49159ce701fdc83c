"""WFB "WMB" block pieces: ``FEB`` / ``ProcessBlock`` / ``FFAB`` (RawFomer_WFB_FFAB/blocks.py:11-92) and
``Illumination_Estimator`` (RawFomer_WFB_FFAB/model.py:174-200).

CPU: the numpy oracle against goldens produced by executing the reference classes (tests/golden/make_golden_wfb.py).
GPU: the CUDA kernels, through the module mirrors and the C ABI, against the same goldens; the dense-DFT rfft2 / irfft2 against
numpy's FFT on the sizes of a SID Sony frame's LL band (712 x 1064 = 2^3*89 x 2^3*7*19) and on odd sizes; round trips.

Tolerance: fp32 max-abs 2e-4 on outputs of magnitude <= 10 (the phase of a bin goes through a learned 1x1 stack and back through
cos / sin, so an fp32-rounding difference in atan2 is amplified by the weights; FFAB chains seven such blocks)."""
import numpy as np
import pytest
import torch

import rf_testlib as T
from oracle import rawformer_oracle as O

GOLD = T.load_golden("wfb")
TOL = 2e-4


def _oracle(kind, sd, x):
    sdn = T.sd_numpy(sd)
    if kind == "feb":
        return (O.feb(sdn, x),)
    if kind == "pb":
        return (O.process_block(sdn, x),)
    if kind == "ffab":
        return (O.ffab(sdn, x),)
    return O.illumination_estimator(sdn, x)


@pytest.mark.parametrize("case", T.WFB_CASES, ids=[c[0] for c in T.WFB_CASES])
def test_oracle_matches_reference(case):
    name, kind, c, shape, seed, scale = case
    sd = T.make_state_dict(T.build_wfb(kind, c), seed=seed, scale=scale)
    outs = _oracle(kind, sd, T.wfb_input(kind, shape, seed))
    for i, o in enumerate(outs):
        ref = GOLD[f"{name}.out{i}"]
        assert o.shape == ref.shape and o.dtype == np.float32
        err = float(np.abs(o.astype(np.float64) - ref).max())
        assert err <= TOL, f"{name} out{i}: max-abs {err:.3e}"


def test_state_dict_layout_and_no_fallback():
    import bayer_low_light_image_enhancement_b200 as rf

    f = rf.FFAB(8)
    keys = list(f.state_dict())
    assert keys[:2] == ["conv0.0.weight", "conv0.0.bias"] and "conv4.0.frequency_process.process2.2.bias" in keys
    assert tuple(f.conv4[0].cat.weight.shape) == (16, 16, 1, 1) and tuple(f.convout[1].weight.shape) == (8, 16, 1, 1)
    ill = rf.Illumination_Estimator(16, n_fea_in=17, n_fea_out=16)
    assert tuple(ill.depth_conv.weight.shape) == (16, 1, 5, 5)
    with pytest.raises(RuntimeError):
        f(torch.zeros(1, 8, 4, 4))              # CPU tensor: there is no fallback
    w = rf.WMB(16)
    assert "ffab.conv1.cat.weight" in w.state_dict() and "illu.depth_conv.weight" in w.state_dict()
    with pytest.raises(NotImplementedError):
        w(torch.zeros(1, 16, 8, 8))             # the Mamba branch is third party: it must be supplied


@pytest.mark.gpu
@pytest.mark.parametrize("case", T.WFB_CASES, ids=[c[0] for c in T.WFB_CASES])
def test_cuda_matches_reference(case):
    name, kind, c, shape, seed, scale = case
    dev = torch.device("cuda", 0)
    m = T.build_wfb(kind, c)
    m.load_state_dict(T.make_state_dict(m, seed=seed, scale=scale), strict=True)
    m = m.to(dev).eval()
    with torch.no_grad():
        outs = m(torch.from_numpy(T.wfb_input(kind, shape, seed)).to(dev))
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        ref = GOLD[f"{name}.out{i}"]
        o = o.cpu().numpy()
        assert o.shape == ref.shape and o.dtype == np.float32
        err = float(np.abs(o.astype(np.float64) - ref).max())
        assert err <= TOL, f"{name} out{i}: max-abs {err:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 2, 712, 1064), (2, 3, 45, 77), (1, 1, 1, 7), (3, 1, 16, 2)], ids=str)
def test_rfft2_irfft2_against_numpy(shape):
    from bayer_low_light_image_enhancement_b200 import wfb

    dev = torch.device("cuda", 0)
    x = T.gen_input("randn", shape, 77)
    spec = wfb.rfft2_ortho(torch.from_numpy(x).to(dev))
    ref = np.fft.rfft2(x.astype(np.float64), norm="ortho")
    got = spec.cpu().numpy()
    scale = float(np.abs(ref).max())
    assert got.shape == shape[:2] + (2, shape[2], shape[3] // 2 + 1)
    assert np.abs(got[:, :, 0] - ref.real).max() <= 2e-5 * scale and np.abs(got[:, :, 1] - ref.imag).max() <= 2e-5 * scale
    back = wfb.irfft2_ortho(spec, shape[3]).cpu().numpy()
    assert np.abs(back - x).max() <= 2e-5 * max(1.0, float(np.abs(x).max()))
    # a spectrum that is NOT the transform of a real image (what FEB feeds the inverse, blocks.py:32-35): same result as a
    # complex inverse over the rows' axis followed by a complex-to-real transform that drops the DC / Nyquist imaginary parts
    spec2 = spec.clone()
    spec2[:, :, 1, :, 0] = 3.0
    if shape[3] % 2 == 0:
        spec2[:, :, 1, :, -1] = -2.0
    spec2[:, :, 0] *= 1.5
    s2 = spec2.cpu().numpy().astype(np.float64)
    ref2 = np.fft.irfft2(s2[:, :, 0] + 1j * s2[:, :, 1], s=shape[2:], norm="ortho")
    back2 = wfb.irfft2_ortho(spec2, shape[3]).cpu().numpy()
    assert np.abs(back2 - ref2).max() <= 2e-5 * max(1.0, float(np.abs(ref2).max()))


@pytest.mark.gpu
def test_feb_ragged_band_against_oracle():
    """One FEB at a ragged, non-power-of-two band (178 x 266 = the LL band of a 1/4-scale Sony frame) against the oracle."""
    dev = torch.device("cuda", 0)
    m = T.build_wfb("feb", 8)
    sd = T.make_state_dict(m, seed=61, scale=1.0)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    x = T.gen_input("randn", (1, 8, 178, 266), 62)
    with torch.no_grad():
        got = m(torch.from_numpy(x).to(dev)).cpu().numpy()
    ref = O.feb(T.sd_numpy(sd), x)
    assert float(np.abs(got - ref).max()) <= 5e-4
