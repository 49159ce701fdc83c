"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE ITSELF (CPU, fp32).

Run in the build container only (needs /root/reference, which never travels to the GPU box):

    python tests/golden/make_golden.py

Inputs and weights are reproducible from seeds (tests/rf_testlib.py), so only the reference outputs are
stored.  The reference modules are imported by file path with stub modules for imports their forward never
touches (ptflops, imageio, timm, mamba_ssm) -- SURVEY appendix C.  Nothing is copied from the reference tree.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import rf_testlib as T  # noqa: E402

REF = os.environ.get("RF_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    _stub("ptflops", get_model_complexity_info=lambda *a, **k: (None, None))
    _stub("imageio")
    flca = load_ref("FrequencyawareLumaChromaAttentionRAWFormer.py", "ref_flca_rf")
    ml = load_ref("MultiLvlFrequencyawareLumaChromaAttentionRAWFormer.py", "ref_ml_rf")
    readme = {"__name__": "readme"}
    text = open(os.path.join(REF, "README.md")).read()
    exec(text.split("### DWT and IDWT")[1].split("```")[1], readme)
    # WFB blocks / model (timm + mamba_ssm are import-only for the pieces we use)
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.vision_transformer", VisionTransformer=object, _cfg=None)
    _stub("timm.models.registry", register_model=None)
    _stub("timm.models.layers", trunc_normal_=None, DropPath=None, to_2tuple=None)
    _stub("mamba_ssm", Mamba=object)
    wfb_blocks = wfb_model = None
    try:
        sys.path.insert(0, os.path.join(REF, "RawFomer_WFB_FFAB"))
        wfb_blocks = load_ref("RawFomer_WFB_FFAB/blocks.py", "blocks")
        wfb_model = load_ref("RawFomer_WFB_FFAB/model.py", "ref_wfb_model")
    except Exception as e:  # pragma: no cover
        print("WFB import failed:", repr(e))
    return flca, ml, readme, wfb_blocks, wfb_model


def save(name, **arrays):
    path = T.golden_path(name)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  wrote {os.path.relpath(path, T.ROOT)}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@torch.no_grad()
def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    flca_rf, ml_rf, readme, wfb_blocks, wfb_model = load_reference_modules()

    # ---------------- whole models ----------------
    for case in T.MODEL_CASES:
        name, variant, dim, H, W, kind, seed, scale, b = case
        print(name)
        ours = T.build_model(variant, dim)
        sd = T.make_state_dict(ours, seed=1234 + seed, scale=scale)
        ref = (flca_rf if variant == "flca" else ml_rf).RawFormer(dim=dim).eval()
        ref_keys = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        our_keys = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert ref_keys == our_keys, "state_dict mismatch between the reference and the mirror modules"
        assert list(ref.state_dict()) == list(ours.state_dict()), "state_dict key ORDER differs"
        ref.load_state_dict(sd, strict=True)
        x = T.gen_input(kind, (b, 1, H, W), seed)
        out = ref(t(x)).numpy()
        out64 = ref.double()(t(x).double()).numpy()
        print(f"    out range [{out.min():.4f}, {out.max():.4f}]  fp32-vs-fp64 max-abs {np.abs(out - out64).max():.2e}")
        save(name, out=out.astype(np.float32), err64=np.float64(np.abs(out - out64).max()))

    # default-initialisation equality: same seed, same construction order => identical parameters
    torch.manual_seed(1234)
    ref = flca_rf.RawFormer(dim=32)
    torch.manual_seed(1234)
    ours = T.build_model("flca", 32)
    same = all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), ours.state_dict().values()))
    print("default-init equality with the reference under manual_seed(1234):", same)
    assert same

    # ---------------- Conv_Transformer blocks and their parts ----------------
    for case in T.BLOCK_CASES:
        name, variant, C, (hf, wf), (hy, wy), seed, scale, b = case
        print(name)
        ours = T.build_block(variant, C)
        sd = T.make_state_dict(ours, seed=seed, scale=scale)
        mod = flca_rf if variant == "flca" else ml_rf
        ref = mod.Conv_Transformer(C).eval()
        assert list(ref.state_dict()) == list(ours.state_dict())
        ref.load_state_dict(sd, strict=True)
        feat, x_ds = T.block_inputs(case)
        y, cr, cb = mod.BayerLumaChroma()(t(x_ds))
        ft = t(feat)
        res = dict(
            out=ref(ft, y, cr, cb), flca=ref.FLCA(ft, y, cr, cb), trans=ref.Transformer(ft),
            attn=ref.Transformer.attn(ft), ffn=ref.Transformer.ffn(ft), ln=ref.Transformer.norm1(ft),
            y=y, cr=cr, cb=cb)
        save(name, **{k: v.numpy() for k, v in res.items()})

    # ---------------- small operators ----------------
    print("ops")
    ops = {}
    x = t(T.gen_input("int", (2, 3, 8, 12), 1))
    ops["downshuffle_r2"] = flca_rf.downshuffle(x, 2).numpy()
    x1 = t(T.gen_input("int", (1, 1, 32, 48), 2))
    ops["downshuffle_raw"] = flca_rf.downshuffle(x1, 2).numpy()
    x4 = t(T.gen_input("int", (1, 2, 12, 12), 2))
    ops["downshuffle_r4"] = flca_rf.downshuffle(x4, 4).numpy()
    ops["pixelshuffle_r2"] = torch.nn.PixelShuffle(2)(t(T.gen_input("int", (2, 12, 5, 7), 3))).numpy()
    ops["pixelshuffle_r3"] = torch.nn.PixelShuffle(3)(t(T.gen_input("int", (1, 18, 4, 6), 3))).numpy()
    haar = flca_rf.HaarDWT()
    for tag, shape in (("even", (2, 3, 10, 16)), ("odd", (1, 2, 9, 13)), ("oddh", (1, 1, 7, 8)), ("oddw", (1, 1, 8, 7))):
        LL, (LH, HL, HH) = haar(t(T.gen_input("randn", shape, 4)))
        ops[f"haar_{tag}"] = torch.stack([LL, LH, HL, HH], 0).numpy()
    ops["haar_filt"] = haar.filt.numpy()
    xl = t(T.gen_input("rand", (2, 4, 10, 14), 5))
    ops["luma_rand"] = torch.stack(flca_rf.BayerLumaChroma()(xl), 0).numpy()
    ops["luma_zeros"] = torch.stack(flca_rf.BayerLumaChroma()(torch.zeros(1, 4, 6, 8)), 0).numpy()
    # README CustomDWT / CustomIDWT, incl. the README's own self-check
    torch.manual_seed(0)
    xr = torch.randn(1, 3, 64, 64)
    k = [[1, 1, 1, 1], [1, -1, 1, 1], [1, 1, -1, 1], [1, 1, 1, -1]]
    dwt, idwt = readme["CustomDWT"](kernel=k), readme["CustomIDWT"](kernel=k)
    xd = dwt(xr)
    rec = idwt(xd)
    ops["readme_x"] = xr.numpy()
    ops["readme_dwt"] = xd.numpy()
    ops["readme_rec"] = rec.numpy()
    ops["readme_mse"] = np.float64(F.mse_loss(xr, rec).item())
    print("    README reconstruction MSE:", ops["readme_mse"])
    xi = t(T.gen_input("int", (2, 5, 6, 10), 6))
    ops["cdwt_int"] = readme["CustomDWT"]()(xi).numpy()
    ops["cidwt_int"] = readme["CustomIDWT"]()(t(T.gen_input("int", (2, 8, 3, 5), 7))).numpy()
    kh = [[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]
    ops["cdwt_haar_nonorm"] = readme["CustomDWT"](kernel=kh, norm=False)(xi).numpy()
    ops["cidwt_haar_nonorm"] = readme["CustomIDWT"](kernel=kh, norm=False)(t(T.gen_input("int", (2, 8, 3, 5), 7))).numpy()
    ops["cdwt_nocustom"] = readme["CustomDWT"](kernel=kh, use_custom=False)(xi).numpy()
    if wfb_blocks is not None:
        xw = t(T.gen_input("int", (2, 3, 6, 8), 8))
        ops["dwt_init"] = wfb_blocks.dwt_init(xw).numpy()
        ops["iwt_init"] = wfb_blocks.iwt_init(t(T.gen_input("int", (8, 3, 3, 4), 9))).numpy()
        ops["iwt_roundtrip"] = wfb_blocks.iwt_init(wfb_blocks.dwt_init(t(T.gen_input("randn", (1, 2, 8, 8), 9)))).numpy()
    # Downsample / LayerNorm
    ds = flca_rf.Downsample(32).eval()
    sd = T.make_state_dict(ds, seed=20, scale=1.5)
    ds.load_state_dict(sd)
    ops["downsample_c32"] = ds(t(T.gen_input("randn", (2, 32, 8, 12), 21))).numpy()
    ln = flca_rf.LayerNorm(48).eval()
    ln.load_state_dict(T.make_state_dict(ln, seed=22))
    ops["layernorm_c48"] = ln(t(T.gen_input("randn", (2, 48, 5, 7), 23) * 3 + 1)).numpy()
    if wfb_model is not None:
        ff = wfb_model.FeedForward(32, 2.66, False).eval()
        ffsd = T.make_state_dict(ff, seed=24, scale=1.5)
        ff.load_state_dict(ffsd)
        ops["wfb_ffn_keys"] = np.array(list(ffsd.keys()))
        for kk, vv in ffsd.items():
            ops["wfb_ffn_sd." + kk] = vv.numpy()
        ops["wfb_ffn_c32"] = ff(t(T.gen_input("randn", (2, 32, 6, 9), 25))).numpy()
        lnb = wfb_model.BiasFree_LayerNorm(32)
        lnw = wfb_model.WithBias_LayerNorm(32)
        xx = t(T.gen_input("randn", (2, 35, 32), 26) * 2 + 0.5)
        wv = t(np.random.default_rng(27).uniform(0.5, 1.5, 32).astype(np.float32))
        bv = t(np.random.default_rng(28).uniform(-0.2, 0.2, 32).astype(np.float32))
        lnb.weight.data.copy_(wv)
        lnw.weight.data.copy_(wv)
        lnw.bias.data.copy_(bv)
        ops["wfb_ln_biasfree"] = lnb(xx).numpy()
        ops["wfb_ln_withbias"] = lnw(xx).numpy()
    # ML tail helpers
    xo = t(T.gen_input("rand", (2, 3, 16, 24), 30))
    xp = t(T.gen_input("rand", (2, 4, 8, 12), 31))
    ops["ml_color_anchor"] = ml_rf.color_anchor_correction_rgb(xo, xp, alpha=0.12).numpy()
    # bilinear resize semantics (F.interpolate align_corners=False) on guidance-like maps
    g = t(T.gen_input("rand", (1, 1, 12, 20), 32))
    for tag, size in (("x2", (24, 40)), ("d2", (6, 10)), ("d4", (3, 5)), ("odd", (9, 14)), ("same", (12, 20))):
        ops[f"bilinear_{tag}"] = F.interpolate(g, size=size, mode="bilinear", align_corners=False).numpy()
    save("ops", **ops)


if __name__ == "__main__":
    main()
