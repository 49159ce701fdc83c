"""Full-frame parity at the BENCHMARKED sizes (BASELINE configs 2, 3, 4, 5): one SID Sony frame, raw 2848x4256, through the
CUDA path against the reference's fp32 CPU forward (oracle/rawformer_torch.py, the functional-PyTorch port pinned to the
reference-generated goldens by tests/test_oracle_golden.py).  The CPU forward of a full frame takes 5-20 s per model on
the GPU box's host cores.

bf16 mode (the benchmarked mode): PSNR(ours, ref) >= 55 dB relative to the output's own range AND the north-star
criterion |PSNR(ours, GT) - PSNR(ref, GT)| <= 0.05 dB against a seeded synthetic ground truth.
fp32 mode (RawFormer-S): max-abs <= 1e-4.
"""
import numpy as np
import pytest
import torch

import rf_testlib as T

pytestmark = pytest.mark.gpu

H_RAW, W_RAW = 2848, 4256
SIZES = {"S": 32, "B": 48, "L": 64}


def psnr(a, b, data_range):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)


def _frame():
    return torch.rand(1, 1, H_RAW, W_RAW, generator=torch.Generator().manual_seed(0))


def _reference(variant, dim, x):
    from oracle import rawformer_torch as P

    sd = T.make_state_dict(T.build_model(variant, dim), seed=1234, scale=1.0)
    with torch.no_grad():
        return sd, P.rawformer_forward(sd, x, variant).numpy()


@pytest.mark.parametrize("variant,size", [("flca", "S"), ("flca", "B"), ("flca", "L"), ("ml", "S")],
                         ids=["config2_S", "config3_B", "config4_L", "config5_ml_S"])
def test_full_frame_bf16_vs_reference(variant, size):
    dev = torch.device("cuda", 0)
    dim = SIZES[size]
    x = _frame()
    sd, ref = _reference(variant, dim, x)
    m = T.build_model(variant, dim, precision="bf16")
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(x.to(dev)).float().cpu().numpy()
    del m
    torch.cuda.empty_cache()
    assert out.shape == ref.shape == (1, 3, H_RAW, W_RAW)
    assert np.isfinite(out).all()
    rng = max(float(ref.max() - ref.min()), 1e-6)
    p = psnr(out, ref, rng)
    gt = np.clip(ref + np.random.default_rng(5).normal(0, 0.05 * rng, ref.shape).astype(np.float32), ref.min(), ref.max())
    d = abs(psnr(out, gt, rng) - psnr(ref, gt, rng))
    print(f"full frame {variant}-{size} bf16: PSNR(ours, ref) {p:.2f} dB, |dPSNR vs GT| {d:.4f} dB, max-abs "
          f"{np.abs(out - ref).max():.3e} (range {rng:.4f})")
    assert p >= 55.0, f"{variant}-{size}: PSNR(ours, ref) = {p:.1f} dB"
    assert d <= 0.05, f"{variant}-{size}: |PSNR(ours,GT) - PSNR(ref,GT)| = {d:.3f} dB"


def test_full_frame_fp32_vs_reference():
    dev = torch.device("cuda", 0)
    x = _frame()
    sd, ref = _reference("flca", 32, x)
    m = T.build_model("flca", 32, precision="fp32")
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(x.to(dev)).float().cpu().numpy()
    err = float(np.abs(out.astype(np.float64) - ref.astype(np.float64)).max())
    print(f"full frame flca-S fp32: max-abs {err:.3e}, ref range [{ref.min():.4f}, {ref.max():.4f}]")
    assert err <= 1e-4 * max(1.0, float(np.abs(ref).max()))
