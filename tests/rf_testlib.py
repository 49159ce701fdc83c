"""Shared helpers for the test-suite: deterministic inputs and weights that need neither the reference
tree nor a GPU, plus the list of golden cases (``tests/golden/make_golden.py`` writes them by executing the
real reference; the tests read them back).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def gen_input(kind: str, shape, seed: int = 0) -> np.ndarray:
    """Synthetic inputs of SURVEY 8d: 'rand' U[0,1); 'dark' = SID-like dark frame through the reference
    normalisation (WFB/load_dataset.py:81-90) and clamp; 'randn'; 'zeros'; 'const'; 'hot' (single hot pixel);
    'int' = small integer-valued floats (exact sums, for bit-exact index tests)."""
    rng = np.random.default_rng([seed, 7919])
    if kind == "rand":
        return rng.random(shape, dtype=np.float32)
    if kind == "randn":
        return rng.standard_normal(shape).astype(np.float32)
    if kind == "int":
        return rng.integers(-8, 9, size=shape).astype(np.float32)
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    if kind == "const":
        return np.full(shape, 0.25, np.float32)
    if kind == "hot":
        x = np.zeros(shape, np.float32)
        x[..., shape[-2] // 2, shape[-1] // 3] = 1.0
        return x
    if kind == "dark":
        raw = 512.0 + rng.poisson(20.0, size=shape) + rng.normal(0.0, 3.0, size=shape)
        raw = np.clip(np.round(raw), 0, 16383)
        ap = 100.0 if seed % 2 == 0 else 300.0
        x = (np.clip(raw, 512, 16383) - 512.0) / (16383.0 - 512.0 + 1e-6) * ap
        return np.clip(x, 0.0, 1.0).astype(np.float32)
    raise ValueError(kind)


_KEEP = ("filt", "r_w", "g_w", "b_w", "running_mean", "running_var", "num_batches_tracked")


def make_state_dict(module: torch.nn.Module, seed: int = 1234, scale: float = 1.0) -> dict:
    """Deterministic, construction-order-independent weights for any module of this package (or the reference:
    the key names are the same).  Conv weights ~ U(-b,b), b = scale/sqrt(fan_in); LayerNorm affine, attention
    temperature and the FLCA balances are randomised so that no parameter sits at its neutral default."""
    out = {}
    for idx, (k, v) in enumerate(module.state_dict().items()):
        leaf = k.rsplit(".", 1)[-1]
        rng = np.random.default_rng([seed, idx])
        shape = tuple(v.shape)
        if leaf in _KEEP or k.split(".")[-1] in _KEEP:
            if leaf == "running_var":
                val = rng.uniform(0.5, 1.5, shape)
            elif leaf == "running_mean":
                val = rng.uniform(-0.2, 0.2, shape)
            else:
                out[k] = v.detach().clone()
                continue
        elif leaf == "temperature":
            val = rng.uniform(0.5, 2.0, shape)
        elif leaf in ("alpha", "beta", "gamma"):
            val = rng.uniform(0.5, 1.5, shape)
        elif ".norm" in k or k.startswith("norm") or ".bn." in k or ".body." in k and v.dim() == 1:
            val = rng.uniform(0.5, 1.5, shape) if leaf == "weight" else rng.uniform(-0.2, 0.2, shape)
        elif leaf == "weight" and v.dim() >= 2:
            fan_in = int(np.prod(shape[1:]))
            if "up" in k.split(".")[0] and v.dim() == 4 and shape[2] == 2:  # ConvTranspose2d [Ci,Co,2,2]
                fan_in = shape[0]
            b = scale / np.sqrt(max(fan_in, 1))
            val = rng.uniform(-b, b, shape)
        elif leaf == "weight":  # 1-d weight of an unknown norm
            val = rng.uniform(0.5, 1.5, shape)
        else:  # bias
            val = rng.uniform(-0.1, 0.1, shape) * scale
        out[k] = torch.from_numpy(np.asarray(val, np.float32).reshape(shape)).clone()
    return out


def sd_numpy(sd: dict, dtype=np.float32) -> dict:
    return {k: np.asarray(v.detach().cpu().numpy(), dtype) for k, v in sd.items()}


def sub_sd(sd: dict, prefix: str) -> dict:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ---------------------------------------------------------------------------------------------------
# golden cases.  Every case is reproducible from these few numbers; only the reference OUTPUT is stored.
# ---------------------------------------------------------------------------------------------------
MODEL_CASES = [
    # name, variant, dim, H, W (raw), input kind, seed, weight scale, batch
    ("flca_s_64x64_rand", "flca", 32, 64, 64, "rand", 0, 1.0, 1),
    ("flca_s_96x160_dark", "flca", 32, 96, 160, "dark", 1, 2.0, 2),
    ("flca_b_64x96_rand", "flca", 48, 64, 96, "rand", 2, 1.5, 1),
    ("flca_l_64x64_dark", "flca", 64, 64, 64, "dark", 3, 1.0, 1),
    ("flca_s_64x64_zeros", "flca", 32, 64, 64, "zeros", 0, 1.0, 1),
    ("flca_s_64x64_hot", "flca", 32, 64, 64, "hot", 0, 2.0, 1),
    ("flca_s_64x64_const", "flca", 32, 64, 64, "const", 0, 1.0, 1),
    ("ml_s_64x64_rand", "ml", 32, 64, 64, "rand", 0, 1.0, 1),
    ("ml_b_96x64_dark", "ml", 48, 96, 64, "dark", 1, 1.5, 2),
    ("ml_l_64x64_rand", "ml", 64, 64, 64, "rand", 2, 1.0, 1),
]

# sub-module cases: name, kind, channels C, feature (Hf,Wf), guidance (Hy,Wy), seed, scale, batch
BLOCK_CASES = [
    ("block_flca_c32_s0", "flca", 32, (16, 24), (16, 24), 10, 1.5, 2),
    ("block_flca_c64_s1", "flca", 64, (12, 20), (24, 40), 11, 1.5, 1),
    ("block_flca_c96_s2", "flca", 96, (6, 10), (24, 40), 12, 1.5, 1),
    ("block_flca_c256_s3", "flca", 256, (3, 5), (24, 40), 13, 1.0, 1),
    ("block_flca_c48_odd", "flca", 48, (9, 14), (19, 27), 14, 1.5, 1),
    ("block_ml_c32_s0", "ml", 32, (16, 24), (16, 24), 15, 1.5, 2),
    ("block_ml_c64_s1", "ml", 64, (12, 20), (24, 40), 16, 1.5, 1),
    ("block_ml_c128_s3", "ml", 128, (3, 5), (24, 40), 17, 1.0, 1),
]


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, name + ".npz")


def load_golden(name: str) -> dict:
    with np.load(golden_path(name)) as z:
        return {k: z[k] for k in z.files}


def build_model(variant: str, dim: int, precision=None):
    import bayer_low_light_image_enhancement_b200 as rf

    cls = rf.RawFormer if variant == "flca" else rf.multilevel.RawFormer
    return cls(dim=dim, precision=precision)


def build_block(variant: str, C: int):
    import bayer_low_light_image_enhancement_b200 as rf

    return rf.Conv_Transformer(C) if variant == "flca" else rf.multilevel.Conv_Transformer(C)


def block_inputs(case):
    name, variant, C, (hf, wf), (hy, wy), seed, scale, b = case
    feat = gen_input("randn", (b, C, hf, wf), seed)
    x_ds = gen_input("rand", (b, 4, hy, wy), seed + 100)
    return feat, x_ds


# ---------------------------------------------------------------------------------------------------
# TrueColor head / tail: name, variant (0 TrueColorRawFormer.py, 1 BayerTORGBColorMultiLvl.py), kind,
# input shape, seed, weight scale
# ---------------------------------------------------------------------------------------------------
TRUECOLOR_CASES = [
    ("tc_head_v0", 0, "head", (2, 4, 18, 26), 40, 1.5),
    ("tc_head_v1", 1, "head", (2, 4, 18, 26), 41, 1.5),
    ("tc_tail_v0", 0, "tail", (2, 3, 20, 28), 42, 2.0),
    ("tc_tail_v1", 1, "tail", (2, 3, 20, 28), 43, 2.0),
    ("tc_head_v0_zeros", 0, "head0", (1, 4, 8, 8), 44, 1.0),
]


def build_truecolor(kind, variant):
    import bayer_low_light_image_enhancement_b200 as rf

    ns = rf.truecolor if variant == 0 else rf.truecolor.multilevel
    return ns.EnhancedBayerProcessor() if kind.startswith("head") else ns.CameraAwareColorCorrection()


def make_truecolor_state_dict(module, seed, scale):
    """make_state_dict plus non-neutral white balance / colour matrix / gamma."""
    sd = make_state_dict(module, seed=seed, scale=scale)
    rng = np.random.default_rng([seed, 99])
    for k in sd:
        if k == "wb_gains":
            sd[k] = torch.from_numpy(rng.uniform(0.7, 2.0, 4).astype(np.float32))
        elif k == "color_matrix":
            m = np.eye(3, 4, dtype=np.float32) + rng.uniform(-0.2, 0.2, (3, 4)).astype(np.float32)
            sd[k] = torch.from_numpy(m)
        elif k in ("gamma", "gamma_param"):
            sd[k] = torch.tensor(float(rng.uniform(1.6, 2.6)), dtype=torch.float32)
    return sd


def truecolor_input(kind, shape, seed):
    if kind == "head0":
        return np.zeros(shape, np.float32)
    x = gen_input("rand", shape, seed)
    if kind == "tail":                      # exercise both clamps of the tail: values below 0 and above 1
        x = x * 1.4 - 0.2
    return x.astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# WFB "WMB" block pieces: name, kind, channels, input shape, seed, weight scale
# ---------------------------------------------------------------------------------------------------
WFB_CASES = [
    ("wfb_feb_c8_even", "feb", 8, (2, 8, 12, 20), 50, 1.5),
    ("wfb_feb_c4_odd", "feb", 4, (1, 4, 9, 15), 51, 1.5),
    ("wfb_pb_c8", "pb", 8, (1, 8, 10, 14), 52, 1.5),
    ("wfb_ffab_c8", "ffab", 8, (2, 8, 12, 18), 53, 1.0),
    ("wfb_ffab_c32", "ffab", 32, (1, 32, 16, 24), 54, 1.0),
    ("wfb_illu_c16", "illu", 16, (2, 16, 11, 13), 55, 1.5),
]


def build_wfb(kind, c):
    import bayer_low_light_image_enhancement_b200 as rf

    if kind == "feb":
        return rf.FEB(c)
    if kind == "pb":
        return rf.ProcessBlock(c)
    if kind == "ffab":
        return rf.FFAB(c)
    return rf.Illumination_Estimator(c, n_fea_in=c + 1, n_fea_out=c)


def wfb_input(kind, shape, seed):
    x = gen_input("randn", shape, seed)
    if kind == "feb":
        x = x * 4.0            # a few values beyond the +-10 input clamp
    return x.astype(np.float32)
