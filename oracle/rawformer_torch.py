"""CPU port of the reference forward in functional PyTorch (ATen/MKL-DNN kernels) -- the CPU *baseline*.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rules as ``rawformer_oracle.py``): imported by ``tests/`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (and its opt-in ``--torch-eager`` baseline leg, which runs
the same operator sequence eagerly on the GPU as the "existing Blackwell path" to compare with); never by the product
package.

Why a second restatement: the reference itself is pure Python/PyTorch (no native code to compile into
``oracle/_ref``) and cannot travel to the GPU box, while the numpy oracle is written for clarity, not speed.  This
port issues the same ATen operators in the same order as the reference modules, so timing it on the box's host
cores is a faithful stand-in for "the reference's CPU forward".  Parity status: **pinned** -- checked against the
reference-generated golden vectors in ``tests/test_oracle_golden.py::test_torch_port``.

Citations: FLCA_RF = FrequencyawareLumaChromaAttentionRAWFormer.py, ML_RF = MultiLvlFrequencyaware...RAWFormer.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _p(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def downshuffle(x, r=2):
    """FLCA_RF.py:18-33."""
    return F.pixel_unshuffle(x, r)


def haar_dwt(x, filt):
    """FLCA_RF.py:56-73."""
    B, C, H, W = x.shape
    if (H & 1) or (W & 1):
        x = F.pad(x, (0, W & 1, 0, H & 1), mode="reflect")
    y = F.conv2d(x, filt.repeat(C, 1, 1, 1), stride=2, groups=C)
    y = y.view(B, C, 4, y.shape[-2], y.shape[-1])
    return y[:, :, 0], (y[:, :, 1], y[:, :, 2], y[:, :, 3])


def luma_chroma(x_ds, sd):
    """FLCA_RF.py:87-97."""
    r, g, b = x_ds[:, 0:1], 0.5 * (x_ds[:, 1:2] + x_ds[:, 2:3]), x_ds[:, 3:4]
    y = sd["luma_chroma.r_w"] * r + sd["luma_chroma.g_w"] * g + sd["luma_chroma.b_w"] * b
    y = y / y.amax(dim=(2, 3), keepdim=True).clamp_min(1e-6)
    return y, r - y, b - y


def _resize(t, size):
    return F.interpolate(t, size=size, mode="bilinear", align_corners=False)


def _se(sd, x):
    m = F.adaptive_avg_pool2d(x, 1)
    h = F.relu(F.conv2d(m, sd["se.1.weight"], sd["se.1.bias"]))
    return torch.sigmoid(F.conv2d(h, sd["se.3.weight"], sd["se.3.bias"]))


def flca(sd, feat, y, cr, cb):
    """FLCA_RF.py:136-162."""
    size = feat.shape[-2:]
    LL, (LH, HL, HH) = haar_dwt(y, sd["dwt.filt"])
    hi = torch.sqrt(LH.pow(2) + HL.pow(2) + HH.pow(2) + 1e-8)
    a_low = torch.sigmoid(F.conv2d(_resize(LL, size), sd["low_attn.0.weight"], padding=1))
    a_high = torch.tanh(F.conv2d(_resize(hi, size), sd["high_attn.0.weight"], padding=1))
    a_chr = torch.sigmoid(F.conv2d(torch.cat([_resize(cr, size), _resize(cb, size)], 1), sd["chroma_attn.0.weight"], padding=1))
    x = feat * (1 + sd["alpha"] * a_low + sd["beta"] * a_high + sd["gamma"] * a_chr)
    return x * _se(sd, x)


def flca_pyramid(sd, feat, y, cr, cb, levels=2):
    """ML_RF.py:132-183."""
    size = feat.shape[-2:]
    x = feat
    cur = y
    lows, highs = [], []
    for _ in range(levels):
        LL, (LH, HL, HH) = haar_dwt(cur, sd["dwt.filt"])
        lows.append(LL)
        highs.append(torch.sqrt(LH.pow(2) + HL.pow(2) + HH.pow(2) + 1e-8))
        cur = LL

    def res_proj(z):
        z = F.relu(F.conv2d(z, sd["res_proj.0.weight"], sd["res_proj.0.bias"]))
        return F.conv2d(z, sd["res_proj.2.weight"], sd["res_proj.2.bias"])

    for l in range(levels):
        y_low, y_high = _resize(lows[l], size), _resize(highs[l], size)
        a_low = torch.sigmoid(F.conv2d(y_low, sd[f"low_attn.{l}.0.weight"], padding=1))
        a_high = torch.tanh(F.conv2d(y_high, sd[f"high_attn.{l}.0.weight"], padding=1))
        g_in = torch.cat([F.adaptive_avg_pool2d(y_low, 1), F.adaptive_avg_pool2d(y_high, 1)], 1)
        gates = torch.sigmoid(F.conv2d(g_in, sd[f"freq_gate_head.{l}.weight"], sd[f"freq_gate_head.{l}.bias"]))
        x = x + torch.tanh(res_proj(x * (gates[:, 0:1] * a_low + gates[:, 1:2] * a_high))) * 0.2
    cr_r, cb_r = _resize(cr, size), _resize(cb, size)
    a_chr = torch.sigmoid(F.conv2d(torch.cat([cr_r, cb_r], 1), sd["chroma_attn.0.weight"], padding=1))
    mag = torch.sqrt(cr_r.pow(2) + cb_r.pow(2) + 1e-8)
    gamma = torch.sigmoid(F.conv2d(F.adaptive_avg_pool2d(mag, 1), sd["chroma_gate.weight"], sd["chroma_gate.bias"]))
    x = x + torch.tanh(res_proj(x * (gamma * a_chr))) * 0.2
    return x * _se(sd, x)


def layernorm(x, w, b):
    """FLCA_RF.py:185-187."""
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, 1e-5).permute(0, 3, 1, 2)


def attention(sd, x, heads=8):
    """FLCA_RF.py:221-235."""
    b, c, h, w = x.shape
    qkv = F.conv2d(F.conv2d(x, sd["qkv.weight"], sd["qkv.bias"]), sd["qkv_dwconv.weight"], sd["qkv_dwconv.bias"],
                   padding=1, groups=3 * c)
    q, k, v = (t.reshape(b, heads, c // heads, h * w) for t in qkv.chunk(3, dim=1))
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    attn = ((q @ k.transpose(-2, -1)) * sd["temperature"]).softmax(dim=-1)
    out = (attn @ v).reshape(b, c, h, w)
    return F.conv2d(out, sd["project_out.weight"], sd["project_out.bias"])


def conv_ffn(sd, x):
    """FLCA_RF.py:204-209."""
    x = F.conv2d(x, sd["pointwise1.weight"], sd["pointwise1.bias"])
    x = F.conv2d(x, sd["depthwise.weight"], sd["depthwise.bias"], padding=1, groups=x.shape[1])
    return F.conv2d(F.gelu(x), sd["pointwise2.weight"], sd["pointwise2.bias"])


def conv_transformer(sd, feat, y, cr, cb, levels=0):
    """FLCA_RF.py:272-278 / ML_RF.py:252-258."""
    f = flca_pyramid(_p(sd, "FLCA."), feat, y, cr, cb, levels) if levels else flca(_p(sd, "FLCA."), feat, y, cr, cb)
    t = _p(sd, "Transformer.")
    x = feat + attention(_p(t, "attn."), layernorm(feat, t["norm1.body.weight"], t["norm1.body.bias"]))
    x = x + conv_ffn(_p(t, "ffn."), layernorm(x, t["norm2.body.weight"], t["norm2.body.bias"]))
    x = F.conv2d(torch.cat([f, x], 1), sd["channel_reduce.weight"], sd["channel_reduce.bias"])
    return F.leaky_relu(F.conv2d(x, sd["Conv_out.weight"], sd["Conv_out.bias"], padding=1), 0.2)


@torch.no_grad()
def rawformer_forward(sd, x, variant="flca"):
    """FLCA_RF.py:330-370 (variant 'ml': ML_RF.py:356-416)."""
    ml = variant == "ml"
    lv = 2 if ml else 0
    x_ds = downshuffle(x, 2)
    y, cr, cb = luma_chroma(x_ds, sd)
    ct = lambda i, f: conv_transformer(_p(sd, f"conv_tran{i}."), f, y, cr, cb, lv)
    dk = (lambda n: f"down{n}.0.weight") if ml else (lambda n: f"down{n}.body.0.weight")
    down = lambda n, f: downshuffle(F.conv2d(f, sd[dk(n)], padding=1), 2)
    c1 = ct(1, F.conv2d(x_ds, sd["embedding.weight"], sd["embedding.bias"], padding=1))
    c2 = ct(2, down(1, c1))
    c3 = ct(3, down(2, c2))
    c4 = ct(4, down(3, c3))

    def up(n, lo, skip):
        u = F.conv_transpose2d(lo, sd[f"up{n}.weight"], sd[f"up{n}.bias"], stride=2)
        return F.conv2d(torch.cat([u, skip], 1), sd[f"channel_reduce{n}.weight"], sd[f"channel_reduce{n}.bias"])

    c5 = ct(5, up(1, c4, c3))
    c6 = ct(6, up(2, c5, c2))
    c7 = ct(7, up(3, c6, c1))
    out = F.pixel_shuffle(F.leaky_relu(F.conv2d(c7, sd["conv_out.weight"], sd["conv_out.bias"], padding=1), 0.2), 2)
    if ml:
        rgb = torch.cat([x_ds[:, 0:1], 0.5 * (x_ds[:, 1:2] + x_ds[:, 2:3]), x_ds[:, 3:4]], 1)
        full = _resize(rgb, out.shape[-2:])
        out = out + 0.12 * (full.mean(dim=(2, 3), keepdim=True) - out.mean(dim=(2, 3), keepdim=True))
        cur = y
        for _ in range(2):
            cur, _hi = haar_dwt(cur, sd["haar.filt"])
        out_y = 0.299 * out[:, 0:1] + 0.587 * out[:, 1:2] + 0.114 * out[:, 2:3]
        out = out + (_resize(cur, out.shape[-2:]) - out_y) * 0.03
    return out
