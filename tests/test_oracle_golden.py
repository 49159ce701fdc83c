"""Pins the numpy oracle (oracle/rawformer_oracle.py) against golden vectors produced by executing the
reference itself (tests/golden/make_golden.py) and against the README's own DWT->IDWT known answer.
CPU only."""
import numpy as np
import pytest

import rf_testlib as T
from oracle import rawformer_oracle as O

OPS = T.load_golden("ops")


def _sd(module, seed, scale):
    return T.sd_numpy(T.make_state_dict(module, seed=seed, scale=scale))


@pytest.mark.parametrize("case", T.MODEL_CASES, ids=[c[0] for c in T.MODEL_CASES])
def test_whole_model(case):
    name, variant, dim, H, W, kind, seed, scale, b = case
    sd = _sd(T.build_model(variant, dim), 1234 + seed, scale)
    x = T.gen_input(kind, (b, 1, H, W), seed)
    out = O.rawformer_forward(sd, x) if variant == "flca" else O.rawformer_ml_forward(sd, x)
    ref = T.load_golden(name)["out"]
    tol = 2e-5 * max(1.0, float(np.abs(ref).max()))
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= tol


@pytest.mark.parametrize("case", T.BLOCK_CASES, ids=[c[0] for c in T.BLOCK_CASES])
def test_block_and_parts(case):
    name, variant, C, (hf, wf), (hy, wy), seed, scale, b = case
    g = T.load_golden(name)
    sd = _sd(T.build_block(variant, C), seed, scale)
    feat, x_ds = T.block_inputs(case)
    y, cr, cb = O.bayer_luma_chroma(x_ds)
    for a, r in ((y, g["y"]), (cr, g["cr"]), (cb, g["cb"])):
        assert np.abs(a - r).max() <= 1e-6
    lv = 2 if variant == "ml" else 0
    fl = O.flca_pyramid if variant == "ml" else O.flca
    checks = {
        "flca": fl(T.sub_sd(sd, "FLCA."), feat, y, cr, cb),
        "ln": O.layernorm_channels(feat, sd["Transformer.norm1.body.weight"], sd["Transformer.norm1.body.bias"]),
        "attn": O.attention(T.sub_sd(sd, "Transformer.attn."), feat),
        "ffn": O.conv_ffn(T.sub_sd(sd, "Transformer.ffn."), feat),
        "trans": O.transformer_block(T.sub_sd(sd, "Transformer."), feat),
        "out": O.conv_transformer(sd, feat, y, cr, cb, 8, lv),
    }
    for k, v in checks.items():
        tol = 3e-5 * max(1.0, float(np.abs(g[k]).max()))
        assert np.abs(v - g[k]).max() <= tol, k


def test_index_ops_bit_exact():
    assert np.array_equal(O.downshuffle(T.gen_input("int", (2, 3, 8, 12), 1), 2), OPS["downshuffle_r2"])
    assert np.array_equal(O.downshuffle(T.gen_input("int", (1, 1, 32, 48), 2), 2), OPS["downshuffle_raw"])
    assert np.array_equal(O.downshuffle(T.gen_input("int", (1, 2, 12, 12), 2), 4), OPS["downshuffle_r4"])
    assert np.array_equal(O.pixelshuffle(T.gen_input("int", (2, 12, 5, 7), 3), 2), OPS["pixelshuffle_r2"])
    assert np.array_equal(O.pixelshuffle(T.gen_input("int", (1, 18, 4, 6), 3), 3), OPS["pixelshuffle_r3"])


@pytest.mark.parametrize("tag,shape", [("even", (2, 3, 10, 16)), ("odd", (1, 2, 9, 13)), ("oddh", (1, 1, 7, 8)),
                                       ("oddw", (1, 1, 8, 7))])
def test_haar_dwt(tag, shape):
    x = T.gen_input("randn", shape, 4)
    LL, (LH, HL, HH) = O.haar_dwt(x)
    got = np.stack([LL, LH, HL, HH], 0)
    assert got.shape == OPS[f"haar_{tag}"].shape
    assert np.abs(got - OPS[f"haar_{tag}"]).max() <= 5e-7  # association order only (hard parts)
    assert np.array_equal(O.haar_filt().reshape(4, 1, 2, 2), OPS["haar_filt"])
    assert float(O.haar_filt()[0, 0, 0]).hex() == "0x1.fffffe0000000p-2"


def test_luma_chroma():
    got = np.stack(O.bayer_luma_chroma(T.gen_input("rand", (2, 4, 10, 14), 5)), 0)
    assert np.abs(got - OPS["luma_rand"]).max() <= 1e-6
    z = np.stack(O.bayer_luma_chroma(np.zeros((1, 4, 6, 8), np.float32)), 0)
    assert np.array_equal(z, OPS["luma_zeros"])


def test_readme_custom_dwt_known_answer():
    x = OPS["readme_x"]
    d = O.custom_dwt(x)
    assert np.abs(d - OPS["readme_dwt"]).max() <= 1e-6
    rec = O.custom_idwt(d)
    assert np.abs(rec - OPS["readme_rec"]).max() <= 1e-6
    mse = float(np.mean((x.astype(np.float64) - rec) ** 2))
    assert abs(mse - 0.36616483330726624) < 1e-6  # README.md:148-170 self-check value (non-zero by construction)
    assert abs(float(OPS["readme_mse"]) - 0.36616483330726624) < 1e-9


def test_custom_dwt_variants_bit_exact():
    xi = T.gen_input("int", (2, 5, 6, 10), 6)
    xs = T.gen_input("int", (2, 8, 3, 5), 7)
    kh = [[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]
    assert np.array_equal(O.custom_dwt(xi), OPS["cdwt_int"])
    assert np.array_equal(O.custom_idwt(xs), OPS["cidwt_int"])
    assert np.array_equal(O.custom_dwt(xi, kernel=kh, norm=False), OPS["cdwt_haar_nonorm"])
    assert np.array_equal(O.custom_idwt(xs, kernel=kh, norm=False), OPS["cidwt_haar_nonorm"])
    assert np.array_equal(O.custom_dwt(xi, kernel=kh, use_custom=False), OPS["cdwt_nocustom"])


def test_wfb_dwt_iwt():
    assert np.array_equal(O.dwt_init(T.gen_input("int", (2, 3, 6, 8), 8)), OPS["dwt_init"])
    assert np.array_equal(O.iwt_init(T.gen_input("int", (8, 3, 3, 4), 9)), OPS["iwt_init"])
    x = T.gen_input("randn", (1, 2, 8, 8), 9)
    assert np.abs(O.iwt_init(O.dwt_init(x)) - OPS["iwt_roundtrip"]).max() <= 1e-6


def test_downsample_layernorm():
    import torch

    class _DS(torch.nn.Module):  # key layout of the reference Downsample: body.0.weight
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Sequential(torch.nn.Conv2d(32, 16, 3, padding=1, bias=False))

    sd = T.sd_numpy(T.make_state_dict(_DS(), seed=20, scale=1.5))
    got = O.downsample(sd["body.0.weight"], T.gen_input("randn", (2, 32, 8, 12), 21))
    assert np.abs(got - OPS["downsample_c32"]).max() <= 2e-5

    class _LN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.LayerNorm(48)

    sd = T.sd_numpy(T.make_state_dict(_LN(), seed=22))
    x = T.gen_input("randn", (2, 48, 5, 7), 23) * 3 + 1
    got = O.layernorm_channels(x, sd["body.weight"], sd["body.bias"])
    assert np.abs(got - OPS["layernorm_c48"]).max() <= 1e-5


def test_wfb_feedforward_and_layernorms():
    keys = [str(k) for k in OPS["wfb_ffn_keys"]]
    sd = {k: OPS["wfb_ffn_sd." + k] for k in keys}
    got = O.feedforward_gated(sd, T.gen_input("randn", (2, 32, 6, 9), 25))
    assert np.abs(got - OPS["wfb_ffn_c32"]).max() <= 3e-5
    xx = T.gen_input("randn", (2, 35, 32), 26) * 2 + 0.5
    wv = np.random.default_rng(27).uniform(0.5, 1.5, 32).astype(np.float32)
    bv = np.random.default_rng(28).uniform(-0.2, 0.2, 32).astype(np.float32)
    x4 = np.ascontiguousarray(xx.transpose(0, 2, 1))[:, :, :, None]  # [B, C, N, 1]
    bf = O.layernorm_biasfree(x4, wv)[:, :, :, 0].transpose(0, 2, 1)
    wb = O.layernorm_withbias(x4, wv, bv)[:, :, :, 0].transpose(0, 2, 1)
    assert np.abs(bf - OPS["wfb_ln_biasfree"]).max() <= 1e-5
    assert np.abs(wb - OPS["wfb_ln_withbias"]).max() <= 1e-5


def test_ml_tail_and_bilinear():
    xo = T.gen_input("rand", (2, 3, 16, 24), 30)
    xp = T.gen_input("rand", (2, 4, 8, 12), 31)
    assert np.abs(O.color_anchor_correction_rgb(xo, xp, 0.12) - OPS["ml_color_anchor"]).max() <= 1e-6
    g = T.gen_input("rand", (1, 1, 12, 20), 32)
    for tag, size in (("x2", (24, 40)), ("d2", (6, 10)), ("d4", (3, 5)), ("odd", (9, 14)), ("same", (12, 20))):
        assert np.abs(O.bilinear_resize(g, size) - OPS[f"bilinear_{tag}"]).max() <= 1e-6, tag


@pytest.mark.parametrize("case", T.MODEL_CASES, ids=[c[0] for c in T.MODEL_CASES])
def test_torch_port(case):
    """The functional-PyTorch CPU port used as bench.py's cpu_baseline / reference arm, against the same goldens."""
    import torch

    from oracle import rawformer_torch as P

    name, variant, dim, H, W, kind, seed, scale, b = case
    sd = T.make_state_dict(T.build_model(variant, dim), 1234 + seed, scale)
    x = torch.from_numpy(T.gen_input(kind, (b, 1, H, W), seed))
    out = P.rawformer_forward(sd, x, variant).numpy()
    ref = T.load_golden(name)["out"]
    assert np.abs(out - ref).max() <= 2e-5 * max(1.0, float(np.abs(ref).max()))


def test_preprocess_u16_oracle_matches_reference_statements():
    """RAW normalisation: the oracle against tests/golden/pre.npz, which make_golden_pre.py produced by
    executing the reference loader's own statements (WFB/load_dataset.py:88-89, correctdataloader.py:103)."""
    g = T.load_golden("pre")
    raw = g["raw"]
    assert raw.dtype == np.uint16 and raw.min() == 0 and raw.max() == 65535
    for ap in (100, 300):
        for clamp in (False, True):
            ref = g[f"out_ap{ap}_{'clamp' if clamp else 'noclamp'}"]
            got = O.preprocess_u16(raw, 512.0, 16383.0, float(ap), clamp)
            assert got.dtype == np.float32 and np.array_equal(got, ref), (ap, clamp)
    assert float(g["out_ap300_noclamp"].max()) == 300.0 and float(g["out_ap300_clamp"].max()) == 1.0
