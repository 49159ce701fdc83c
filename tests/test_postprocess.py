"""Caller side of the forward (test.py:17-40,111-123): uint8 conversion, Bayer channel order, R/B
auto-correction and PSNR.

CPU: the numpy oracle against golden vectors produced by the reference's own functions (tests/golden/make_golden_post.py).
GPU: the device operators against the oracle, bit-exact (integer / byte work)."""
import math

import numpy as np
import pytest
import torch

import rf_testlib as T
from oracle import rawformer_oracle as O

PATTERNS = ("RGGB", "BGGR", "GBRG", "GRBG", "rggb", "XXXX")


def _golden():
    g = T.load_golden("post")
    names = sorted(k[3:] for k in g if k.startswith("in_"))
    return g, names


def test_oracle_matches_reference_golden():
    g, names = _golden()
    assert names == ["blue", "equal", "green", "red", "tiny"]
    swapped = 0
    for n in names:
        img = g["in_" + n]
        for pat in PATTERNS:
            ours = O.auto_correct_rb(O.correct_bayer_channels(img.copy(), pat))
            assert np.array_equal(ours, g[f"out_{n}_{pat}"]), (n, pat)
            swapped += int(not np.array_equal(g[f"out_{n}_{pat}"], O.correct_bayer_channels(img.copy(), pat)))
    assert swapped > 0                       # the data-dependent branch is exercised both ways
    # equal means: strict '<' -> no swap
    assert np.array_equal(O.auto_correct_rb(g["in_equal"]), g["in_equal"])


def test_oracle_u8_conversion_and_psnr_known_answers():
    pred = np.array([[[[-0.5, 0.0, 0.5]], [[1.0, 1.5, 0.999]], [[0.25, 0.75, 1e-9]]]], np.float32)      # [1,3,1,3]
    out = O.postprocess_u8(pred)
    assert out.shape == (1, 1, 3, 3) and out.dtype == np.uint8
    assert out[0, 0].tolist() == [[0, 255, 63], [0, 255, 191], [127, 254, 0]]      # truncation, not rounding
    a = np.zeros((4, 4, 3), np.uint8)
    b = a.copy()
    assert O.psnr_u8(a, b) == math.inf
    b[0, 0, 0] = 48                                                                  # mse = 48^2 / 48 = 48
    assert abs(O.psnr_u8(a, b) - 10.0 * math.log10(255.0 ** 2 / 48.0)) < 1e-12


def test_oracle_ssim_known_answers():
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(24, 31, 3), dtype=np.uint8)
    assert O.ssim_u8(a, a) == 1.0                                   # identical images: every map value is exactly 1
    flat0, flat1 = np.full((9, 9, 3), 100, np.uint8), np.full((9, 9, 3), 120, np.uint8)
    # constant images: variances are 0, S = (2*100*120 + C1) / (100^2 + 120^2 + C1) everywhere
    c1 = (0.01 * 255) ** 2
    assert abs(O.ssim_u8(flat0, flat1) - (2 * 100 * 120 + c1) / (100 ** 2 + 120 ** 2 + c1)) < 1e-12
    noisy = np.clip(a.astype(np.int16) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    s = O.ssim_u8(a, noisy)
    assert 0.5 < s < 1.0 and abs(s - O.ssim_u8(noisy, a)) < 1e-12   # symmetric
    with pytest.raises(ValueError):
        O.ssim_u8(a[:5], a[:5])


def test_bayer_downshuffle_validation():
    """dataloader.py:7-43: r must be 2, the pattern one of the four Bayer layouts (checked before any device work)."""
    import bayer_low_light_image_enhancement_b200 as rf

    x = torch.zeros(1, 1, 4, 4)
    with pytest.raises(ValueError):
        rf.bayer_downshuffle(x, np.array([[0, 1], [1, 2]]), r=4)
    with pytest.raises(ValueError):
        rf.bayer_downshuffle(x, np.array([[0, 0], [1, 2]]))
    with pytest.raises(ValueError):
        rf.bayer_downshuffle(torch.zeros(1, 3, 4, 4), np.array([[0, 1], [1, 2]]))
    with pytest.raises(RuntimeError):                      # valid request on a CPU tensor: no fallback
        rf.bayer_downshuffle(x, np.array([[0, 1], [1, 2]]))


@pytest.mark.gpu
def test_bayer_downshuffle_device():
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    x = torch.from_numpy(T.gen_input("int", (2, 1, 24, 40), 4)).to(dev)
    # the reference's arithmetic (dataloader.py:35-43): phases in positional order, whatever the pattern
    ref = torch.cat([x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]], dim=1)
    for pat in ([[0, 1], [1, 2]], [[2, 1], [1, 0]], [[1, 0], [2, 1]], [[1, 2], [0, 1]]):
        assert torch.equal(rf.bayer_downshuffle(x, np.array(pat)), ref)
    assert torch.equal(rf.bayer_downshuffle(x, [[0, 1], [1, 2]]), rf.downshuffle(x, 2))


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_corrections_bit_exact():
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    g, names = _golden()
    # (a) uint8 images through correct_rgb_u8 against the reference goldens, batched so that one call takes both branches
    for pat in PATTERNS:
        batch = np.stack([g["in_" + n] for n in ("red", "blue", "equal", "green")])
        got = rf.correct_rgb_u8(torch.from_numpy(batch).to(dev), pat).cpu().numpy()
        for i, n in enumerate(("red", "blue", "equal", "green")):
            assert np.array_equal(got[i], g[f"out_{n}_{pat}"]), (n, pat)
        got = rf.correct_rgb_u8(torch.from_numpy(batch).to(dev), pat, auto_rb=False).cpu().numpy()
        assert np.array_equal(got, O.correct_bayer_channels(batch, pat))
    tiny = torch.from_numpy(g["in_tiny"][None]).to(dev)
    assert np.array_equal(rf.correct_rgb_u8(tiny, "BGGR").cpu().numpy()[0], g["out_tiny_BGGR"])
    # (b) float predictions through postprocess_rgb_u8 against the oracle chain, odd sizes, values outside [0,1]
    rng = np.random.default_rng(7)
    for shape in ((3, 3, 37, 53), (1, 3, 128, 256), (2, 3, 1, 1)):
        pred = (rng.standard_normal(shape) * 0.6 + 0.45).astype(np.float32)
        pred[0, 2] += 0.3                                     # image 0: blue-heavy -> swapped
        for pat in ("RGGB", "GBRG"):
            ref = np.stack([O.auto_correct_rb(O.correct_bayer_channels(im, pat)) for im in O.postprocess_u8(pred)])
            got = rf.postprocess_rgb_u8(torch.from_numpy(pred).to(dev), pat).cpu().numpy()
            assert got.dtype == np.uint8 and np.array_equal(got, ref), (shape, pat)
        assert np.array_equal(rf.postprocess_rgb_u8(torch.from_numpy(pred).to(dev), "RGGB", auto_rb=False).cpu().numpy(),
                              O.postprocess_u8(pred))
        assert np.array_equal(rf.postprocess_u8(torch.from_numpy(pred).to(dev)).cpu().numpy(), O.postprocess_u8(pred))
    with pytest.raises(ValueError):
        rf.correct_rgb_u8(torch.zeros(1, 4, 4, 3, device=dev))       # not uint8


@pytest.mark.gpu
def test_device_psnr_exact():
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, size=(3, 211, 173, 3), dtype=np.uint8)
    b = a.copy()
    b[0] = rng.integers(0, 256, size=b[0].shape, dtype=np.uint8)      # unrelated images
    b[1, ::7, ::5, 1] ^= 3                                             # small sparse error; image 2 identical
    got = rf.psnr_u8(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    ref = [O.psnr_u8(a[i], b[i]) for i in range(3)]
    assert got[2] == math.inf and ref[2] == math.inf
    assert got[0] == ref[0] and got[1] == ref[1], (got, ref)          # exact integer error sums -> identical doubles
    # full-frame size: sum of squares well beyond 32 bits
    big_a = torch.full((1, 2848, 4256, 3), 255, dtype=torch.uint8, device=dev)
    big_b = torch.zeros_like(big_a)
    assert rf.psnr_u8(big_a, big_b) == [10.0 * math.log10(255.0 ** 2 / 255.0 ** 2)]


@pytest.mark.gpu
def test_device_ssim():
    """Windowed integer sums are exact; the map is evaluated in double like skimage: agreement to ~1e-12 (scipy's separable
    uniform_filter rounds the window means differently in the last bits)."""
    import bayer_low_light_image_enhancement_b200 as rf

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(13)
    for h, w in ((7, 7), (40, 53), (129, 200)):
        a = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
        b = a.copy()
        b[0] = np.clip(a[0].astype(np.int16) + rng.integers(-25, 26, size=a[0].shape), 0, 255).astype(np.uint8)
        b[1] = rng.integers(0, 256, size=a[1].shape, dtype=np.uint8)
        got = rf.ssim_u8(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
        ref = [O.ssim_u8(a[i], b[i]) for i in range(3)]
        assert abs(got[2] - 1.0) <= 1e-12          # identical images (the device may contract a*b+c into FMAs)
        for g_, r_ in zip(got, ref):
            assert abs(g_ - r_) <= 1e-10, (h, w, got, ref)
    with pytest.raises(ValueError):
        rf.ssim_u8(torch.zeros(1, 6, 9, 3, dtype=torch.uint8, device=dev), torch.zeros(1, 6, 9, 3, dtype=torch.uint8, device=dev))
    # the reference's evaluation loop (test.py:111-124) end to end on the device against the oracle chain
    pred = (rng.standard_normal((1, 3, 64, 96)) * 0.4 + 0.5).astype(np.float32)
    gt = (rng.random((1, 64, 96, 3)) * 255).astype(np.uint8)
    p_dev = rf.postprocess_rgb_u8(torch.from_numpy(pred).to(dev), "RGGB")
    g_dev = rf.correct_rgb_u8(torch.from_numpy(gt).to(dev), "RGGB")
    p_ref = O.auto_correct_rb(O.correct_bayer_channels(O.postprocess_u8(pred)[0], "RGGB"))
    g_ref = O.auto_correct_rb(O.correct_bayer_channels(gt[0], "RGGB"))
    assert rf.psnr_u8(p_dev, g_dev)[0] == O.psnr_u8(p_ref, g_ref)
    assert abs(rf.ssim_u8(p_dev, g_dev)[0] - O.ssim_u8(p_ref, g_ref)) <= 1e-10


# ---------------------------------------------------------------------------------------------------------
# size-independent properties of the restatement (CPU)
# ---------------------------------------------------------------------------------------------------------
def test_oracle_properties():
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from hypothesis.extra import numpy as hnp

    imgs = hnp.arrays(np.uint8, st.tuples(st.integers(7, 12), st.integers(7, 12), st.just(3)))

    @settings(max_examples=40, deadline=None)
    @given(imgs, st.sampled_from(["RGGB", "BGGR", "GBRG", "GRBG"]))
    def check(img, pat):
        out = O.auto_correct_rb(O.correct_bayer_channels(img, pat))
        # a permutation of the channels: same multiset of planes, and red is never darker than blue afterwards
        assert sorted(out[..., c].tobytes() for c in range(3)) == sorted(img[..., c].tobytes() for c in range(3))
        assert out[..., 0].mean() >= out[..., 2].mean()
        assert np.array_equal(O.auto_correct_rb(out), out)                       # idempotent
        swap = O.correct_bayer_channels(O.correct_bayer_channels(img, "BGGR"), "BGGR")
        assert np.array_equal(swap, img)                                         # the three swaps are involutions
        other = np.roll(img, 1, axis=0)
        assert O.psnr_u8(img, other) == O.psnr_u8(other, img)
        s = O.ssim_u8(img, other)
        assert -1.0 - 1e-12 <= s <= 1.0 + 1e-12 and abs(s - O.ssim_u8(other, img)) < 1e-12

    check()
