"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the C ABI via the
package's nn.Module mirrors, against (a) golden vectors produced by the real reference and (b) the numpy oracle
on freshly seeded inputs.  Tolerances: bit-exact for index/permute work; fp32 mode max-abs <= 1e-4 (scaled by the
output magnitude when the stress weights push it beyond 1); bf16 mode by PSNR (stated per test)."""
import numpy as np
import pytest
import torch

import rf_testlib as T

pytestmark = pytest.mark.gpu

OPS = T.load_golden("ops")


def dev():
    return torch.device("cuda", 0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev())


def npy(t):
    return t.detach().float().cpu().numpy()


def report(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    i = np.unravel_index(np.argmax(d), d.shape) if d.size else ()
    return (f"{name}: max-abs {d.max() if d.size else 0:.3e} at {i} (got {got[i] if d.size else 0:.6f}, ref "
            f"{ref[i] if d.size else 0:.6f}), mean-abs {d.mean() if d.size else 0:.3e}, ref range [{ref.min():.4f},{ref.max():.4f}]")


def assert_close(name, got, ref, tol=1e-4, scaled=True):
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{name}: non-finite values in the CUDA output"
    lim = tol * (max(1.0, float(np.abs(ref).max())) if scaled else 1.0)
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) if ref.size else 0.0
    assert err <= lim, report(name, got, ref) + f"  > tol {lim:.3e}"


def psnr(a, b, data_range):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)


@pytest.fixture(scope="module")
def rf():
    import bayer_low_light_image_enhancement_b200 as pkg

    pkg.set_default_precision("fp32")
    return pkg


# ---------------------------------------------------------------------------------------------------------
# index / wavelet operators: bit-exact
# ---------------------------------------------------------------------------------------------------------
def test_downshuffle_pixelshuffle_bit_exact(rf):
    assert np.array_equal(npy(rf.downshuffle(cu(T.gen_input("int", (2, 3, 8, 12), 1)), 2)), OPS["downshuffle_r2"])
    assert np.array_equal(npy(rf.downshuffle(cu(T.gen_input("int", (1, 1, 32, 48), 2)), 2)), OPS["downshuffle_raw"])
    assert np.array_equal(npy(rf.downshuffle(cu(T.gen_input("int", (1, 2, 12, 12), 2)), 4)), OPS["downshuffle_r4"])
    assert np.array_equal(npy(rf.PixelShuffle(2)(cu(T.gen_input("int", (2, 12, 5, 7), 3)))), OPS["pixelshuffle_r2"])
    assert np.array_equal(npy(rf.PixelShuffle(3)(cu(T.gen_input("int", (1, 18, 4, 6), 3)))), OPS["pixelshuffle_r3"])
    # vector fast paths (W % 8 == 0) against the oracle, incl. a full-frame round trip
    from oracle import rawformer_oracle as O

    x = T.gen_input("randn", (2, 3, 16, 64), 40)
    assert np.array_equal(npy(rf.downshuffle(cu(x), 2)), O.downshuffle(x, 2))
    y = T.gen_input("randn", (2, 8, 6, 16), 41)
    assert np.array_equal(npy(rf.PixelShuffle(2)(cu(y))), O.pixelshuffle(y, 2))
    big = torch.rand(1, 1, 2848, 4256, device=dev())
    assert torch.equal(rf.PixelShuffle(2)(rf.downshuffle(big, 2)), big)


def test_empty_inputs(rf):
    assert tuple(rf.downshuffle(torch.zeros(0, 1, 8, 8, device=dev()), 2).shape) == (0, 4, 4, 4)
    assert tuple(rf.CustomDWT()(torch.zeros(0, 3, 8, 8, device=dev())).shape) == (0, 12, 4, 4)
    assert tuple(rf.dwt_init(torch.zeros(0, 3, 8, 8, device=dev())).shape) == (0, 3, 4, 4)


def test_custom_dwt_idwt(rf):
    xi = T.gen_input("int", (2, 5, 6, 10), 6)
    xs = T.gen_input("int", (2, 8, 3, 5), 7)
    kh = [[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]
    assert np.array_equal(npy(rf.CustomDWT()(cu(xi))), OPS["cdwt_int"])
    assert np.array_equal(npy(rf.CustomIDWT()(cu(xs))), OPS["cidwt_int"])
    assert np.array_equal(npy(rf.CustomDWT(kernel=kh, norm=False)(cu(xi))), OPS["cdwt_haar_nonorm"])
    assert np.array_equal(npy(rf.CustomIDWT(kernel=kh, norm=False)(cu(xs))), OPS["cidwt_haar_nonorm"])
    assert np.array_equal(npy(rf.CustomDWT(kernel=kh, use_custom=False)(cu(xi))), OPS["cdwt_nocustom"])
    # README self-check (README.md:148-170): shapes and the non-zero reconstruction MSE
    x = OPS["readme_x"]
    d = rf.CustomDWT()(cu(x))
    rec = rf.CustomIDWT()(d)
    assert_close("readme_dwt", npy(d), OPS["readme_dwt"], 1e-6)
    assert_close("readme_rec", npy(rec), OPS["readme_rec"], 1e-6)
    mse = float(torch.nn.functional.mse_loss(cu(x), rec).item())
    assert abs(mse - 0.36616483330726624) < 1e-6
    # vector fast path on integer-valued data is exact; full-size shape of the north star ([1,32,1424,2128])
    from oracle import rawformer_oracle as O

    xv = T.gen_input("int", (2, 3, 8, 32), 42)
    assert np.array_equal(npy(rf.CustomDWT()(cu(xv))), O.custom_dwt(xv))
    sv = T.gen_input("int", (1, 8, 6, 16), 43)
    assert np.array_equal(npy(rf.CustomIDWT()(cu(sv))), O.custom_idwt(sv))
    big = torch.randint(-8, 9, (1, 32, 1424, 2128), device=dev()).float()
    db = rf.CustomDWT(kernel=kh, norm=True)(big)  # orthogonal Haar/2: exact inverse on integer data
    assert tuple(db.shape) == (1, 128, 712, 1064)
    assert torch.equal(rf.CustomIDWT(kernel=kh, norm=True)(db), big)


def test_wfb_dwt_iwt(rf):
    assert np.array_equal(npy(rf.dwt_init(cu(T.gen_input("int", (2, 3, 6, 8), 8)))), OPS["dwt_init"])
    assert np.array_equal(npy(rf.iwt_init(cu(T.gen_input("int", (8, 3, 3, 4), 9)))), OPS["iwt_init"])
    x = T.gen_input("randn", (1, 2, 8, 8), 9)
    assert_close("iwt_roundtrip", npy(rf.IWT()(rf.DWT()(cu(x)))), OPS["iwt_roundtrip"], 1e-6)
    big = torch.randint(-8, 9, (2, 4, 712, 1064), device=dev()).float()
    assert torch.equal(rf.iwt_init(rf.dwt_init(big)), big)


@pytest.mark.parametrize("tag,shape", [("even", (2, 3, 10, 16)), ("odd", (1, 2, 9, 13)), ("oddh", (1, 1, 7, 8)),
                                       ("oddw", (1, 1, 8, 7))])
def test_haar_dwt(rf, tag, shape):
    LL, (LH, HL, HH) = rf.HaarDWT().to(dev())(cu(T.gen_input("randn", shape, 4)))
    got = np.stack([npy(LL), npy(LH), npy(HL), npy(HH)], 0)
    assert_close(f"haar_{tag}", got, OPS[f"haar_{tag}"], 5e-7)


def test_luma_chroma_layernorm_downsample(rf):
    got = np.stack([npy(t) for t in rf.BayerLumaChroma().to(dev())(cu(T.gen_input("rand", (2, 4, 10, 14), 5)))], 0)
    assert_close("luma_rand", got, OPS["luma_rand"], 1e-6)
    z = np.stack([npy(t) for t in rf.BayerLumaChroma().to(dev())(torch.zeros(1, 4, 6, 8, device=dev()))], 0)
    assert np.array_equal(z, OPS["luma_zeros"])
    ln = rf.LayerNorm(48)
    ln.load_state_dict(T.make_state_dict(ln, seed=22))
    x = T.gen_input("randn", (2, 48, 5, 7), 23) * 3 + 1
    assert_close("layernorm", npy(ln.to(dev())(cu(x))), OPS["layernorm_c48"], 1e-5)
    ds = rf.Downsample(32)
    ds.load_state_dict(T.make_state_dict(ds, seed=20, scale=1.5))
    assert_close("downsample", npy(ds.to(dev())(cu(T.gen_input("randn", (2, 32, 8, 12), 21)))), OPS["downsample_c32"], 1e-4)


# ---------------------------------------------------------------------------------------------------------
# Conv_Transformer ("WaveTransformBlock") and its parts, fp32 mode, against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", T.BLOCK_CASES, ids=[c[0] for c in T.BLOCK_CASES])
def test_block_and_parts_fp32(rf, case):
    name, variant, C, (hf, wf), (hy, wy), seed, scale, b = case
    g = T.load_golden(name)
    blk = T.build_block(variant, C)
    blk.load_state_dict(T.make_state_dict(blk, seed=seed, scale=scale), strict=True)
    blk = blk.to(dev()).eval()
    feat, x_ds = T.block_inputs(case)
    ft = cu(feat)
    y, cr, cb = rf.BayerLumaChroma().to(dev())(cu(x_ds))
    assert_close("y", npy(y), g["y"], 1e-6)
    assert_close("cr", npy(cr), g["cr"], 1e-6)
    with torch.no_grad():
        assert_close("ln", npy(blk.Transformer.norm1(ft)), g["ln"], 1e-5)
        assert_close("ffn", npy(blk.Transformer.ffn(ft)), g["ffn"])
        assert_close("attn", npy(blk.Transformer.attn(ft)), g["attn"])
        assert_close("flca", npy(blk.FLCA(ft, y, cr, cb)), g["flca"])
        assert_close("trans", npy(blk.Transformer(ft)), g["trans"])
        assert_close("out", npy(blk(ft, y, cr, cb)), g["out"])


@pytest.mark.parametrize("case", T.BLOCK_CASES[:3] + T.BLOCK_CASES[5:6], ids=lambda c: c[0])
def test_block_bf16(rf, case):
    name, variant, C, (hf, wf), (hy, wy), seed, scale, b = case
    g = T.load_golden(name)
    blk = T.build_block(variant, C)
    blk.load_state_dict(T.make_state_dict(blk, seed=seed, scale=scale), strict=True)
    blk = blk.to(dev()).eval()
    blk.precision = "bf16"
    feat, x_ds = T.block_inputs(case)
    y, cr, cb = rf.BayerLumaChroma().to(dev())(cu(x_ds))
    out = npy(blk(cu(feat), y, cr, cb))
    rng = float(g["out"].max() - g["out"].min())
    p = psnr(out, g["out"], rng)
    assert p >= 38.0, f"bf16 block PSNR vs reference {p:.1f} dB (range {rng:.3f}); " + report("out", out, g["out"])


# ---------------------------------------------------------------------------------------------------------
# whole models
# ---------------------------------------------------------------------------------------------------------
def _model(case, precision):
    name, variant, dim, H, W, kind, seed, scale, b = case
    m = T.build_model(variant, dim, precision=precision)
    m.load_state_dict(T.make_state_dict(m, seed=1234 + seed, scale=scale), strict=True)
    return m.to(dev()).eval(), cu(T.gen_input(kind, (b, 1, H, W), seed))


@pytest.mark.parametrize("case", T.MODEL_CASES, ids=[c[0] for c in T.MODEL_CASES])
def test_whole_model_fp32(case):
    m, x = _model(case, "fp32")
    with torch.no_grad():
        out = npy(m(x))
    assert_close(case[0], out, T.load_golden(case[0])["out"], 1e-4)


@pytest.mark.parametrize("case", T.MODEL_CASES, ids=[c[0] for c in T.MODEL_CASES])
def test_whole_model_bf16(case):
    """bf16 mode: PSNR(ours, ref) relative to the output's own range, and the north-star criterion
    |PSNR(ours,GT) - PSNR(ref,GT)| <= 0.05 dB against a seeded synthetic ground truth."""
    m, x = _model(case, "bf16")
    with torch.no_grad():
        out = npy(m(x))
    ref = T.load_golden(case[0])["out"]
    assert np.isfinite(out).all()
    rng = max(float(ref.max() - ref.min()), 1e-6)
    p = psnr(out, ref, rng)
    assert p >= 35.0, f"PSNR(ours, ref) = {p:.1f} dB; " + report(case[0], out, ref)
    gt = np.clip(ref + np.random.default_rng(5).normal(0, 0.05 * rng, ref.shape), ref.min(), ref.max())
    d = abs(psnr(out, gt, rng) - psnr(ref, gt, rng))
    assert d <= 0.05, f"|PSNR(ours,GT) - PSNR(ref,GT)| = {d:.3f} dB"


def test_baseline_config0_packed_512(rf):
    """BASELINE configs[0]: RawFormer-S, batch 1, packed 512x512x4 crop (raw 1024x1024) against the
    reference's fp32 CPU forward (oracle/rawformer_torch.py, pinned to the reference by tests/test_oracle_golden.py):
    fp32 engine max-abs <= 1e-4; bf16 engine PSNR >= 35 dB and |PSNR(ours,GT) - PSNR(ref,GT)| <= 0.05 dB."""
    from oracle import rawformer_torch as P

    m = rf.RawFormer(model_size="S", precision="fp32")
    sd = T.make_state_dict(m, seed=1234, scale=1.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev()).eval()
    x = torch.rand(1, 1, 1024, 1024, generator=torch.Generator().manual_seed(0))
    ref = P.rawformer_forward(sd, x, "flca").numpy()
    with torch.no_grad():
        out32 = npy(m(x.to(dev())))
    assert_close("config0 fp32", out32, ref, 1e-4)
    m16 = rf.RawFormer(model_size="S", precision="bf16")
    m16.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out16 = npy(m16.to(dev()).eval()(x.to(dev())))
    rng = max(float(ref.max() - ref.min()), 1e-6)
    p = psnr(out16, ref, rng)
    assert p >= 35.0, f"config0 bf16 PSNR(ours, ref) = {p:.1f} dB; " + report("config0 bf16", out16, ref)
    gt = np.clip(ref + np.random.default_rng(5).normal(0, 0.05 * rng, ref.shape), ref.min(), ref.max())
    d = abs(psnr(out16, gt, rng) - psnr(ref, gt, rng))
    assert d <= 0.05, f"|PSNR(ours,GT) - PSNR(ref,GT)| = {d:.3f} dB"
    print(report("config0 fp32", out32, ref), f"; bf16 PSNR {p:.1f} dB")


def test_model_size_sugar_and_state_dict(rf):
    for size, dim in rf.MODEL_SIZES.items():
        m = rf.RawFormer(model_size=size)
        assert m.dim == dim
    m = rf.RawFormer(dim=32)
    sd = {"module." + k: v for k, v in T.make_state_dict(m, seed=7).items()}
    m.load_state_dict({k[len("module."):]: v for k, v in sd.items()}, strict=True)  # test.py:89-91
    assert len(m.state_dict()) == 246
    assert len(rf.multilevel.RawFormer(dim=32).state_dict()) == 310


def test_model_batch_and_repack(rf):
    """Batch entries are independent (per-image statistics); re-packing follows in-place parameter updates."""
    case = T.MODEL_CASES[0]
    m, x = _model(case, "fp32")
    x2 = torch.cat([x, torch.flip(x, dims=[3])], 0)
    with torch.no_grad():
        o2 = m(x2)
        o0 = m(x)
        o1 = m(torch.flip(x, dims=[3]).contiguous())
    assert torch.allclose(o2[0], o0[0], atol=2e-6)
    assert torch.allclose(o2[1], o1[0], atol=2e-6)
    with torch.no_grad():
        m.conv_out.bias.add_(0.5)
        o3 = m(x)
    assert float((o3 - o0).abs().max()) > 0.05


def test_errors(rf):
    m = rf.RawFormer(dim=32).to(dev())
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 40, 64, device=dev()))  # H not a multiple of 16
    with pytest.raises(RuntimeError):
        rf.RawFormer(dim=32)(torch.zeros(1, 1, 32, 32))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        rf.CustomDWT()(torch.zeros(1, 1, 5, 6, device=dev()))
    with pytest.raises(ValueError):
        rf.RawFormer(dim=36)


def test_full_frame_properties(rf):
    """BASELINE sizes (SID Sony 2848x4256): no NaN/Inf, fp32-vs-bf16 agreement, launch accounting."""
    from bayer_low_light_image_enhancement_b200 import _lib

    torch.manual_seed(0)
    x = torch.rand(1, 1, 2848, 4256, device=dev())
    m32 = rf.RawFormer(model_size="S", precision="fp32")
    sd = T.make_state_dict(m32, seed=99, scale=1.5)
    m32.load_state_dict(sd)
    m32 = m32.to(dev()).eval()
    m16 = rf.RawFormer(model_size="S", precision="bf16")
    m16.load_state_dict(sd)
    m16 = m16.to(dev()).eval()
    lib = _lib.load()
    lib.rf_reset_launch_count()
    with torch.no_grad():
        o32 = m32(x)
        n_launch = lib.rf_launch_count()
        o16 = m16(x)
    assert n_launch > 50
    assert tuple(o32.shape) == (1, 3, 2848, 4256)
    assert torch.isfinite(o32).all() and torch.isfinite(o16).all()
    rng = float(o32.max() - o32.min())
    mse = float(((o32 - o16) ** 2).mean())
    p = 10 * np.log10(rng ** 2 / max(mse, 1e-30))
    assert p >= 35.0, f"full-frame fp32 vs bf16 PSNR {p:.1f} dB"
    # crop consistency away from the borders is NOT expected (global statistics) -- but the top-left 64x64 of the
    # frame must equal the oracle-checked small-model path in its index work: out[2y+i,2x+j] channel layout.
    assert float(o32.abs().max()) < 1e4


# ---------------------------------------------------------------------------------------------------------
# callers either side of the forward and the WFB gated FFN
# ---------------------------------------------------------------------------------------------------------
def test_postprocess_preprocess(rf):
    pred = torch.randn(2, 3, 20, 28, device=dev()) * 0.8 + 0.4
    got = rf.postprocess_u8(pred).cpu().numpy()
    ref = (torch.clamp(pred, 0, 1).cpu().numpy().transpose(0, 2, 3, 1) * 255).astype(np.uint8)  # test.py:117-118
    assert np.array_equal(got, ref)
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 16384, size=(2, 16, 24)).astype(np.uint16)
    for ap in (100, 300):
        x = np.clip(raw.astype(np.float32), 512, 16383)                       # WFB/load_dataset.py:88-89
        x = (x - 512) / (16383 - 512 + 1e-6) * ap
        x = np.minimum(x, 1.0).astype(np.float32)                              # correctdataloader.py:103
        got = rf.preprocess_u16(torch.from_numpy(raw.view(np.int16)).to(dev()).view(torch.uint16), 512.0, 16383.0, float(ap))
        assert_close("preprocess", got.cpu().numpy()[:, 0], x, 1e-6, scaled=False)


def test_wfb_feedforward_gated(rf):
    keys = [str(k) for k in OPS["wfb_ffn_keys"]]
    sd = {k: torch.from_numpy(OPS["wfb_ffn_sd." + k]) for k in keys}
    ff = rf.FeedForward(32, 2.66, False)
    ff.load_state_dict(sd, strict=True)
    ff = ff.to(dev()).eval()
    out = ff(cu(T.gen_input("randn", (2, 32, 6, 9), 25)))
    assert_close("wfb_ffn", npy(out), OPS["wfb_ffn_c32"], 1e-4)


# ---------------------------------------------------------------------------------------------------------
# bf16 kernels at sizes that span several tiles / chunks (odd sizes, ragged edges): bf16 mode vs fp32 mode of the
# same sub-module (the fp32 mode is itself checked against the reference goldens above)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,hf,wf,b", [(32, 70, 90, 1), (64, 37, 50, 2), (48, 41, 100, 1), (128, 19, 45, 1),
                                       (256, 18, 34, 1)])
def test_bf16_submodules_multitile(rf, C, hf, wf, b):
    blk = T.build_block("flca", C)
    blk.load_state_dict(T.make_state_dict(blk, seed=50 + C, scale=1.5), strict=True)
    blk = blk.to(dev()).eval()
    feat = cu(T.gen_input("randn", (b, C, hf, wf), 60 + C))
    x_ds = cu(T.gen_input("rand", (b, 4, hf, wf), 61 + C))
    y, cr, cb = rf.BayerLumaChroma().to(dev())(x_ds)

    def run(precision):
        for m in blk.modules():
            if hasattr(m, "precision"):
                m.precision = precision
        with torch.no_grad():
            return {
                "ffn": npy(blk.Transformer.ffn(feat)),
                "attn": npy(blk.Transformer.attn(feat)),
                "flca": npy(blk.FLCA(feat, y, cr, cb)),
                "trans": npy(blk.Transformer(feat)),
                "out": npy(blk(feat, y, cr, cb)),
            }

    ref, got = run("fp32"), run("bf16")
    for k in ref:
        assert np.isfinite(got[k]).all(), k
        rng = float(ref[k].max() - ref[k].min())
        p = psnr(got[k], ref[k], rng)
        assert p >= 40.0, f"{k} (C={C}, {hf}x{wf}): bf16 vs fp32 PSNR {p:.1f} dB; " + report(k, got[k], ref[k])


@pytest.mark.parametrize("C,hw", [(32, 448), (64, 320), (128, 224), (256, 160)])
def test_bf16_pipeline_wraparound(rf, C, hw):
    """Images large enough that every persistent CTA walks several tiles (ring / staging-buffer / TMEM-buffer reuse in
    the tcgen05 GEMM, the depthwise and the im2col kernels): bf16 mode against fp32 mode of the same block."""
    blk = T.build_block("flca", C)
    blk.load_state_dict(T.make_state_dict(blk, seed=70 + C, scale=1.5), strict=True)
    blk = blk.to(dev()).eval()
    g = torch.Generator(device="cpu").manual_seed(C)
    feat = torch.randn(1, C, hw, hw + 16, generator=g).to(dev())
    x_ds = torch.rand(1, 4, hw, hw + 16, generator=g).to(dev())
    y, cr, cb = rf.BayerLumaChroma().to(dev())(x_ds)
    outs = {}
    for prec in ("fp32", "bf16"):
        for m in blk.modules():
            if hasattr(m, "precision"):
                m.precision = prec
        with torch.no_grad():
            outs[prec] = npy(blk(feat, y, cr, cb))
    ref, got = outs["fp32"], outs["bf16"]
    assert np.isfinite(got).all()
    p = psnr(got, ref, float(ref.max() - ref.min()))
    assert p >= 40.0, f"C={C}: bf16 vs fp32 PSNR {p:.1f} dB; " + report("out", got, ref)


def test_frame_pipeline(rf):
    """FramePipeline (overlapped H2D / forward / D2H) returns what the direct call returns, frame by frame and in order.
    (The forward's float atomics make two runs differ in the last bf16 bits, so the comparison is by PSNR, and every
    pipelined result must match ITS frame's direct result far better than any other frame's.)"""
    m = rf.RawFormer(dim=32, precision="bf16")
    m.load_state_dict(T.make_state_dict(m, seed=5, scale=1.5))
    m = m.to(dev()).eval()
    frames = [torch.rand(1, 1, 96, 160, generator=torch.Generator().manual_seed(i)).pin_memory() for i in range(5)]
    outs = [torch.empty(1, 3, 96, 160).pin_memory() for _ in frames]
    pipe = rf.FramePipeline(m, depth=2)
    for x, o in zip(frames, outs):
        pipe.submit(x, o)
    pipe.flush()
    with torch.no_grad():
        direct = [npy(m(x.to(dev()))) for x in frames]
    for i, o in enumerate(outs):
        o = o.numpy()
        rng = float(direct[i].max() - direct[i].min())
        own = psnr(o, direct[i], rng)
        others = max(psnr(o, direct[j], rng) for j in range(len(frames)) if j != i)
        assert own >= 45.0 and own > others + 15.0, f"frame {i}: PSNR vs own direct result {own:.1f} dB, best other {others:.1f} dB"


def test_cuda_graph_replay(rf):
    """enable_cuda_graphs(): the captured forward reproduces the eager forward, replays follow new input data in the
    captured buffer, and a second input buffer gets its own graph."""
    m = rf.RawFormer(dim=32, precision="bf16")
    m.load_state_dict(T.make_state_dict(m, seed=6, scale=1.5))
    m = m.to(dev()).eval()
    xa = torch.rand(1, 1, 96, 160, device=dev())
    xb = torch.rand(1, 1, 96, 160, device=dev())
    with torch.no_grad():
        ea, eb = npy(m(xa)), npy(m(xb))
        m.enable_cuda_graphs()
        ga = npy(m(xa))                      # capture + replay
        ga2 = npy(m(xa))                     # replay
        xa.copy_(xb)                         # new data in the captured buffer
        gab = npy(m(xa))
        gb = npy(m(xb))                      # second buffer: second graph
        m.enable_cuda_graphs(False)
    rng = float(ea.max() - ea.min())
    for name, got, ref in (("first", ga, ea), ("replay", ga2, ea), ("new data", gab, eb), ("second buffer", gb, eb)):
        p = psnr(got, ref, rng)
        assert p >= 45.0, f"{name}: graph vs eager PSNR {p:.1f} dB"
    assert psnr(gab, ea, rng) < 40.0         # (the replay really used the new data)
