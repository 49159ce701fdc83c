"""CPU oracle for the RawFormer inference hot path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*: a plain numpy restatement of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``bayer_low_light_image_enhancement_b200``) never imports anything from ``oracle/`` and fails loudly
when its CUDA library is missing.

Parity status: **pinned**.  Every function below is checked in ``tests/test_oracle_golden.py``
against golden vectors produced by executing the reference modules themselves
(``tests/golden/make_golden.py`` imports ``/root/reference`` and writes ``tests/golden/*.npz``), and
against the one numeric known-answer the reference ships (README DWT->IDWT MSE 0.36616483,
``README.md:148-170``).

Reference shorthands used in the citations:
  FLCA_RF = FrequencyawareLumaChromaAttentionRAWFormer.py
  ML_RF   = MultiLvlFrequencyawareLumaChromaAttentionRAWFormer.py
  WFB     = RawFomer_WFB_FFAB/

All tensors are numpy arrays in the reference's own layout (NCHW); weights are a flat dict keyed by
the reference's ``state_dict`` names.  ``dtype`` of the inputs decides the arithmetic (float32 to
mimic the reference, float64 for a tie-breaker "truth").
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf as _erf

# ----------------------------------------------------------------------------------------------
# elementary layers
# ----------------------------------------------------------------------------------------------


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gelu_erf(x):
    """nn.GELU() default = exact erf form (FLCA_RF.py:194,202)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def leaky_relu(x, slope=0.2):
    """nn.LeakyReLU(0.2) (FLCA_RF.py:270,297)."""
    return np.where(x >= 0, x, x * x.dtype.type(slope))


def conv1x1(x, w, b=None):
    """nn.Conv2d(kernel_size=1).  x [B,Ci,H,W], w [Co,Ci,1,1]."""
    B, Ci, H, W = x.shape
    w2 = w.reshape(w.shape[0], Ci).astype(x.dtype)
    y = np.matmul(w2[None], x.reshape(B, Ci, H * W)).reshape(B, -1, H, W)
    if b is not None:
        y = y + b.astype(x.dtype)[None, :, None, None]
    return y


def conv3x3(x, w, b=None):
    """Dense nn.Conv2d(kernel_size=3, stride=1, padding=1) with zero padding."""
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    xp = np.zeros((B, Ci, H + 2, W + 2), x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    y = np.zeros((B, Co, H * W), x.dtype)
    w = w.astype(x.dtype)
    for dy in range(3):
        for dx in range(3):
            sl = np.ascontiguousarray(xp[:, :, dy:dy + H, dx:dx + W]).reshape(B, Ci, H * W)
            y += np.matmul(w[None, :, :, dy, dx], sl)
    y = y.reshape(B, Co, H, W)
    if b is not None:
        y = y + b.astype(x.dtype)[None, :, None, None]
    return y


def dwconv3x3(x, w, b=None):
    """Depthwise nn.Conv2d(C, C, 3, padding=1, groups=C).  w [C,1,3,3]."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    w = w.astype(x.dtype)
    y = np.zeros_like(x)
    for dy in range(3):
        for dx in range(3):
            y += xp[:, :, dy:dy + H, dx:dx + W] * w[None, :, 0, dy, dx, None, None]
    if b is not None:
        y = y + b.astype(x.dtype)[None, :, None, None]
    return y


def conv_transpose2x2(x, w, b):
    """nn.ConvTranspose2d(Ci, Co, 2, stride=2): out[co,2y+i,2x+j] = sum_ci x[ci,y,x] W[ci,co,i,j] + b
    (FLCA_RF.py:315,319,323)."""
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    out = np.zeros((B, Co, 2 * H, 2 * W), x.dtype)
    xf = x.reshape(B, Ci, H * W)
    w = w.astype(x.dtype)
    for i in range(2):
        for j in range(2):
            wt = w[:, :, i, j].T  # [Co,Ci]
            out[:, :, i::2, j::2] = np.matmul(wt[None], xf).reshape(B, Co, H, W)
    return out + b.astype(x.dtype)[None, :, None, None]


def layernorm_channels(x, g, b, eps=1e-5):
    """LayerNorm.forward (FLCA_RF.py:180-187): nn.LayerNorm(C) per pixel, biased variance."""
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    xn = (x - mu) / np.sqrt(var + x.dtype.type(eps))
    return xn * g.astype(x.dtype)[None, :, None, None] + b.astype(x.dtype)[None, :, None, None]


def _axis_taps(n_in, n_out, dtype):
    # F.interpolate(mode='bilinear', align_corners=False) source index rule
    scale = n_in / n_out
    dst = np.arange(n_out, dtype=np.float64)
    src = np.maximum((dst + 0.5) * scale - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = (src - i0).astype(dtype)
    return i0, i1, lam


def bilinear_resize(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) (FLCA_RF.py:145-148)."""
    B, C, H, W = x.shape
    Ho, Wo = size
    if (Ho, Wo) == (H, W):
        return x.copy()
    y0, y1, ly = _axis_taps(H, Ho, x.dtype)
    x0, x1, lx = _axis_taps(W, Wo, x.dtype)
    ly = ly[None, None, :, None]
    lx = lx[None, None, None, :]
    one = x.dtype.type(1)
    top = x[:, :, y0][:, :, :, x0] * (one - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (one - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (one - ly) + bot * ly


# ----------------------------------------------------------------------------------------------
# index / shuffle work (bit-exact)
# ----------------------------------------------------------------------------------------------


def downshuffle(x, r=2):
    """downshuffle (FLCA_RF.py:18-33) == pixel-unshuffle; out channel = c*r*r + r*i + j."""
    B, C, H, W = x.shape
    h, w = H // r, W // r
    x = x[:, :, :h * r, :w * r].reshape(B, C, h, r, w, r)
    return np.ascontiguousarray(x.transpose(0, 1, 3, 5, 2, 4)).reshape(B, C * r * r, h, w)


def pixelshuffle(x, r=2):
    """nn.PixelShuffle(2) (FLCA_RF.py:328,369): out[c,r*y+i,r*x+j] = in[c*r*r + r*i + j, y, x]."""
    B, C, H, W = x.shape
    c = C // (r * r)
    x = x.reshape(B, c, r, r, H, W).transpose(0, 1, 4, 2, 5, 3)
    return np.ascontiguousarray(x).reshape(B, c, H * r, W * r)


# ----------------------------------------------------------------------------------------------
# wavelets
# ----------------------------------------------------------------------------------------------


def haar_filt(dtype=np.float32):
    """HaarDWT.__init__ (FLCA_RF.py:47-54): outer products of fl32(1/sqrt 2) vectors, rounded in fp32
    (so the coefficient is 0x1.fffffep-2, not 0.5)."""
    s = np.float32(1.0) / np.float32(math.sqrt(2.0))
    h = np.array([s, s], np.float32)
    g = np.array([s, -s], np.float32)
    f = np.stack([np.outer(h, h), np.outer(h, g), np.outer(g, h), np.outer(g, g)]).astype(np.float32)
    return f.astype(dtype)  # [4,2,2]


def haar_dwt(x, filt=None):
    """HaarDWT.forward (FLCA_RF.py:56-73).  Returns LL,(LH,HL,HH), each [B,C,ceil(H/2),ceil(W/2)].
    Reflect-pads right/bottom when a dimension is odd (FLCA_RF.py:63-66)."""
    B, C, H, W = x.shape
    if (H & 1) or (W & 1):
        x = np.pad(x, ((0, 0), (0, 0), (0, H & 1), (0, W & 1)), mode="reflect")
    f = haar_filt(x.dtype) if filt is None else np.asarray(filt).reshape(4, 2, 2).astype(x.dtype)
    a = x[:, :, 0::2, 0::2]
    b = x[:, :, 0::2, 1::2]
    c = x[:, :, 1::2, 0::2]
    d = x[:, :, 1::2, 1::2]
    outs = []
    for n in range(4):
        outs.append(f[n, 0, 0] * a + f[n, 0, 1] * b + f[n, 1, 0] * c + f[n, 1, 1] * d)
    return outs[0], (outs[1], outs[2], outs[3])


_README_K = ((1, 1, 1, 1), (1, -1, 1, 1), (1, 1, -1, 1), (1, 1, 1, -1))


def _custom_kernel(kernel, use_custom, norm, dtype):
    k = np.array(kernel if (use_custom and kernel is not None) else _README_K, np.float32).reshape(4, 4)
    if norm:
        k = k / np.float32(2.0)
    return k.astype(dtype)


def custom_dwt(x, kernel=None, use_custom=True, norm=True):
    """CustomDWT.forward (README.md:111-117; weights README.md:95-109).
    sub-band n = sum_t K[n,t]*[a,b,c,d][t]; output channel = n*C + c (sub-band major)."""
    B, C, H, W = x.shape
    K = _custom_kernel(kernel, use_custom, norm, x.dtype)
    taps = (x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2])
    out = np.zeros((B, 4, C, H // 2, W // 2), x.dtype)
    for n in range(4):
        for t in range(4):
            out[:, n] += K[n, t] * taps[t][:, :, :H // 2, :W // 2]
    return out.reshape(B, 4 * C, H // 2, W // 2)


def custom_idwt(x, kernel=None, use_custom=True, norm=True):
    """CustomIDWT.forward (README.md:139-144): conv_transpose2d with the same 4x1x2x2 weight,
    i.e. [a,b,c,d][t] = sum_n K[n,t]*sub[n]."""
    B, C4, H, W = x.shape
    C = C4 // 4
    K = _custom_kernel(kernel, use_custom, norm, x.dtype)
    sub = x.reshape(B, 4, C, H, W)
    out = np.zeros((B, C, 2 * H, 2 * W), x.dtype)
    for t, (i, j) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        acc = np.zeros((B, C, H, W), x.dtype)
        for n in range(4):
            acc += K[n, t] * sub[:, n]
        out[:, :, i::2, j::2] = acc
    return out


def dwt_init(x):
    """dwt_init (WFB/blocks.py:102-115): sub-bands concatenated on the BATCH axis, order LL,HL,LH,HH."""
    x01 = x[:, :, 0::2, :] / 2
    x02 = x[:, :, 1::2, :] / 2
    x1, x2, x3, x4 = x01[:, :, :, 0::2], x02[:, :, :, 0::2], x01[:, :, :, 1::2], x02[:, :, :, 1::2]
    return np.concatenate((x1 + x2 + x3 + x4, -x1 - x2 + x3 + x4, -x1 + x2 - x3 + x4, x1 - x2 - x3 + x4), 0)


def iwt_init(x):
    """iwt_init (WFB/blocks.py:119-136): exact inverse of dwt_init; always float32 output."""
    B4, C, H, W = x.shape
    B = B4 // 4
    x1, x2, x3, x4 = (x[k * B:(k + 1) * B] / 2 for k in range(4))
    h = np.zeros((B, C, 2 * H, 2 * W), np.float32)
    h[:, :, 0::2, 0::2] = x1 - x2 - x3 + x4
    h[:, :, 1::2, 0::2] = x1 - x2 + x3 - x4
    h[:, :, 0::2, 1::2] = x1 + x2 - x3 - x4
    h[:, :, 1::2, 1::2] = x1 + x2 + x3 + x4
    return h


# ----------------------------------------------------------------------------------------------
# guidance
# ----------------------------------------------------------------------------------------------


def bayer_luma_chroma(x_ds, eps=1e-6):
    """BayerLumaChroma.forward (FLCA_RF.py:87-97).  x_ds [B,4,h,w] = (R,G1,G2,B).
    Note cr/cb subtract the *normalised* y from the *raw* r,b."""
    t = x_ds.dtype.type
    r = x_ds[:, 0:1]
    g = t(0.5) * (x_ds[:, 1:2] + x_ds[:, 2:3])
    b = x_ds[:, 3:4]
    y = t(np.float32(0.299)) * r + t(np.float32(0.587)) * g + t(np.float32(0.114)) * b
    y = y / np.maximum(y.max(axis=(2, 3), keepdims=True), t(eps))
    return y, r - y, b - y


def _sd(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def flca(sd, feat, y, cr, cb, eps=1e-8):
    """FLCA.forward (FLCA_RF.py:136-162)."""
    Hf, Wf = feat.shape[-2:]
    t = feat.dtype.type
    LL, (LH, HL, HH) = haar_dwt(y, sd.get("dwt.filt"))
    y_high = np.sqrt(LH ** 2 + HL ** 2 + HH ** 2 + t(eps))
    y_low = bilinear_resize(LL, (Hf, Wf))
    y_high = bilinear_resize(y_high, (Hf, Wf))
    cr = bilinear_resize(cr, (Hf, Wf))
    cb = bilinear_resize(cb, (Hf, Wf))
    a_low = sigmoid(conv3x3(y_low, sd["low_attn.0.weight"]))
    a_high = np.tanh(conv3x3(y_high, sd["high_attn.0.weight"]))
    a_chr = sigmoid(conv3x3(np.concatenate([cr, cb], 1), sd["chroma_attn.0.weight"]))
    spatial = 1 + t(sd["alpha"]) * a_low + t(sd["beta"]) * a_high + t(sd["gamma"]) * a_chr
    x = feat * spatial
    return x * _se(sd, x)


def _se(sd, x):
    # SE (FLCA_RF.py:124-130,160-161): avgpool -> 1x1 -> relu -> 1x1 -> sigmoid
    m = x.mean(axis=(2, 3), keepdims=True)
    h = np.maximum(conv1x1(m, sd["se.1.weight"], sd["se.1.bias"]), 0)
    return sigmoid(conv1x1(h, sd["se.3.weight"], sd["se.3.bias"]))


def flca_pyramid(sd, feat, y, cr, cb, levels=2, eps=1e-8, max_residual_scale=0.2):
    """FLCA_Pyramid.forward (ML_RF.py:132-183)."""
    Hf, Wf = feat.shape[-2:]
    t = feat.dtype.type
    x = feat
    lows, highs = [], []
    cur = y
    for _ in range(levels):  # _pyramid_y (ML_RF.py:122-130)
        LL, (LH, HL, HH) = haar_dwt(cur, sd.get("dwt.filt"))
        lows.append(LL)
        highs.append(np.sqrt(LH ** 2 + HL ** 2 + HH ** 2 + t(eps)))
        cur = LL

    def res_proj(z):
        z = np.maximum(conv1x1(z, sd["res_proj.0.weight"], sd["res_proj.0.bias"]), 0)
        return conv1x1(z, sd["res_proj.2.weight"], sd["res_proj.2.bias"])

    for l in range(levels):
        y_low = bilinear_resize(lows[l], (Hf, Wf))
        y_high = bilinear_resize(highs[l], (Hf, Wf))
        a_low = sigmoid(conv3x3(y_low, sd[f"low_attn.{l}.0.weight"]))
        a_high = np.tanh(conv3x3(y_high, sd[f"high_attn.{l}.0.weight"]))
        g_in = np.concatenate([y_low.mean(axis=(2, 3), keepdims=True), y_high.mean(axis=(2, 3), keepdims=True)], 1)
        gates = sigmoid(conv1x1(g_in, sd[f"freq_gate_head.{l}.weight"], sd[f"freq_gate_head.{l}.bias"]))
        spatial = gates[:, 0:1] * a_low + gates[:, 1:2] * a_high
        x = x + np.tanh(res_proj(x * spatial)) * t(max_residual_scale)
    cr_r = bilinear_resize(cr, (Hf, Wf))
    cb_r = bilinear_resize(cb, (Hf, Wf))
    a_chr = sigmoid(conv3x3(np.concatenate([cr_r, cb_r], 1), sd["chroma_attn.0.weight"]))
    chr_mag = np.sqrt(cr_r ** 2 + cb_r ** 2 + t(eps))
    gamma = sigmoid(conv1x1(chr_mag.mean(axis=(2, 3), keepdims=True), sd["chroma_gate.weight"], sd["chroma_gate.bias"]))
    x = x + np.tanh(res_proj(x * (gamma * a_chr))) * t(max_residual_scale)
    return x * _se(sd, x)


# ----------------------------------------------------------------------------------------------
# transformer branch
# ----------------------------------------------------------------------------------------------


def attention(sd, x, num_heads=8):
    """Attention.forward (FLCA_RF.py:221-235): transposed (channel) attention."""
    B, C, H, W = x.shape
    t = x.dtype.type
    qkv = dwconv3x3(conv1x1(x, sd["qkv.weight"], sd["qkv.bias"]), sd["qkv_dwconv.weight"], sd["qkv_dwconv.bias"])
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, num_heads, C // num_heads, H * W) for i in range(3))
    q = q / np.maximum(np.sqrt((q * q).sum(-1, keepdims=True)), t(1e-12))  # F.normalize, eps is a max
    k = k / np.maximum(np.sqrt((k * k).sum(-1, keepdims=True)), t(1e-12))
    attn = np.matmul(q, k.transpose(0, 1, 3, 2)) * sd["temperature"].astype(x.dtype)[None]
    attn = attn - attn.max(-1, keepdims=True)
    attn = np.exp(attn)
    attn = attn / attn.sum(-1, keepdims=True)
    out = np.matmul(attn, v).reshape(B, C, H, W)
    return conv1x1(out, sd["project_out.weight"], sd["project_out.bias"])


def conv_ffn(sd, x):
    """conv_ffn.forward (FLCA_RF.py:204-209)."""
    x = conv1x1(x, sd["pointwise1.weight"], sd["pointwise1.bias"])
    x = dwconv3x3(x, sd["depthwise.weight"], sd["depthwise.bias"])
    x = gelu_erf(x)
    return conv1x1(x, sd["pointwise2.weight"], sd["pointwise2.bias"])


def transformer_block(sd, x, num_heads=8):
    """TransformerBlock.forward (FLCA_RF.py:251-254)."""
    x = x + attention(_sd(sd, "attn."), layernorm_channels(x, sd["norm1.body.weight"], sd["norm1.body.bias"]), num_heads)
    x = x + conv_ffn(_sd(sd, "ffn."), layernorm_channels(x, sd["norm2.body.weight"], sd["norm2.body.bias"]))
    return x


def conv_transformer(sd, feat, y, cr, cb, num_heads=8, pyramid_levels=0):
    """Conv_Transformer.forward (FLCA_RF.py:272-278; ML_RF.py:252-258)."""
    if pyramid_levels:
        f = flca_pyramid(_sd(sd, "FLCA."), feat, y, cr, cb, levels=pyramid_levels)
    else:
        f = flca(_sd(sd, "FLCA."), feat, y, cr, cb)
    tr = transformer_block(_sd(sd, "Transformer."), feat, num_heads)
    x = conv1x1(np.concatenate([f, tr], 1), sd["channel_reduce.weight"], sd["channel_reduce.bias"])
    return leaky_relu(conv3x3(x, sd["Conv_out.weight"], sd["Conv_out.bias"]))


def downsample(w, x):
    """Downsample.forward (FLCA_RF.py:176-177): bias-free 3x3 C->C/2 then pixel-unshuffle."""
    return downshuffle(conv3x3(x, w), 2)


# ----------------------------------------------------------------------------------------------
# whole models
# ----------------------------------------------------------------------------------------------


def _unet(sd, x_ds, y, cr, cb, down_keys, pyramid_levels):
    ct = lambda i, f: conv_transformer(_sd(sd, f"conv_tran{i}."), f, y, cr, cb, 8, pyramid_levels)
    x0 = conv3x3(x_ds, sd["embedding.weight"], sd["embedding.bias"])
    c1 = ct(1, x0)
    c2 = ct(2, downsample(sd[down_keys[0]], c1))
    c3 = ct(3, downsample(sd[down_keys[1]], c2))
    c4 = ct(4, downsample(sd[down_keys[2]], c3))

    def up(i, lo, skip):
        u = conv_transpose2x2(lo, sd[f"up{i}.weight"], sd[f"up{i}.bias"])
        return conv1x1(np.concatenate([u, skip], 1), sd[f"channel_reduce{i}.weight"], sd[f"channel_reduce{i}.bias"])

    c5 = ct(5, up(1, c4, c3))
    c6 = ct(6, up(2, c5, c2))
    c7 = ct(7, up(3, c6, c1))
    return pixelshuffle(leaky_relu(conv3x3(c7, sd["conv_out.weight"], sd["conv_out.bias"])), 2)


def rawformer_forward(sd, x):
    """RawFormer.forward (FLCA_RF.py:330-370).  x [B,1,H,W] -> [B,3,H,W]."""
    x_ds = downshuffle(x, 2)
    y, cr, cb = bayer_luma_chroma(x_ds)
    return _unet(sd, x_ds, y, cr, cb, [f"down{i}.body.0.weight" for i in (1, 2, 3)], 0)


def color_anchor_correction_rgb(out_rgb, x_ds, alpha=0.12):
    """color_anchor_correction_rgb (ML_RF.py:270-288)."""
    t = out_rgb.dtype.type
    rgb = np.concatenate([x_ds[:, 0:1], t(0.5) * (x_ds[:, 1:2] + x_ds[:, 2:3]), x_ds[:, 3:4]], 1)
    full = bilinear_resize(rgb, out_rgb.shape[-2:])
    return out_rgb + t(alpha) * (full.mean(axis=(2, 3), keepdims=True) - out_rgb.mean(axis=(2, 3), keepdims=True))


def rawformer_ml_forward(sd, x, flca_levels=2):
    """ML_RF RawFormer.forward (ML_RF.py:356-416)."""
    t = x.dtype.type
    x_ds = downshuffle(x, 2)
    y, cr, cb = bayer_luma_chroma(x_ds)
    cur = y
    for _ in range(2):  # LL anchor, always two levels (ML_RF.py:361-367)
        cur, _hi = haar_dwt(cur, sd.get("haar.filt"))
    out = _unet(sd, x_ds, y, cr, cb, [f"down{i}.0.weight" for i in (1, 2, 3)], flca_levels)
    out = color_anchor_correction_rgb(out, x_ds, 0.12)
    ll_up = bilinear_resize(cur, out.shape[-2:])
    out_y = t(0.299) * out[:, 0:1] + t(0.587) * out[:, 1:2] + t(0.114) * out[:, 2:3]
    return out + (ll_up - out_y) * t(0.03)


# ----------------------------------------------------------------------------------------------
# WFB ("WMB") building blocks on the hot-path list
# ----------------------------------------------------------------------------------------------


def _bn_eval(x, sd, prefix, eps=1e-5):
    g, b = sd[prefix + "weight"], sd[prefix + "bias"]
    m, v = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    s = (g / np.sqrt(v + eps)).astype(x.dtype)
    return x * s[None, :, None, None] + (b - m * g / np.sqrt(v + eps)).astype(x.dtype)[None, :, None, None]


def feedforward_gated(sd, x):
    """FeedForward.forward (WFB/model.py:58-65), eval-mode BatchNorm (Conv2d_BN, WFB/model.py:17-25)."""
    identity = x
    x = conv1x1(x, sd["project_in.weight"], sd.get("project_in.bias"))
    r1 = _bn_eval(dwconv3x3(x, sd["rep_conv1.c.weight"]), sd, "rep_conv1.bn.")
    r2 = _bn_eval(x * sd["rep_conv2.c.weight"].astype(x.dtype)[None, :, 0, 0, 0, None, None], sd, "rep_conv2.bn.")
    x1 = x + r1 + r2
    x2 = dwconv3x3(x, sd["dwconv.weight"], sd.get("dwconv.bias"))
    x = gelu_erf(x2) * x1 + gelu_erf(x1) * x2
    return conv1x1(x, sd["project_out.weight"], sd.get("project_out.bias")) + identity


def layernorm_withbias(x, w, b):
    """WithBias_LayerNorm (WFB/model.py:106-122) over channels per pixel."""
    return layernorm_channels(x, w, b, 1e-5)


def layernorm_biasfree(x, w):
    """BiasFree_LayerNorm (WFB/model.py:89-103): divides by sqrt(var+eps) WITHOUT subtracting the mean."""
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    return x / np.sqrt(var + x.dtype.type(1e-5)) * w.astype(x.dtype)[None, :, None, None]


# ----------------------------------------------------------------------------------------------
# TrueColor head / tail.  variant 0 = TrueColorRawFormer.py, 1 = BayerTORGBColorMultiLvl.py
# ----------------------------------------------------------------------------------------------
def softplus(x):
    """F.softplus / nn.Softplus defaults (beta 1, threshold 20)."""
    x = np.asarray(x)
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20)))).astype(x.dtype)


def enhanced_bayer_processor(sd, x, variant=0, eps=1e-6):
    """EnhancedBayerProcessor.forward: TrueColorRawFormer.py:109-142 (variant 0), BayerTORGBColorMultiLvl.py:103-136
    (variant 1).  x [B,4,H,W] -> (y, cr, cb, rgb) with rgb = rgb_linear (variant 0) or the residually refined linear RGB."""
    dt = x.dtype
    M = sd["color_matrix"].astype(dt)
    yw = sd["y_weights"].astype(dt)

    def chroma(r, g, b, y):
        c = conv3x3(np.concatenate([r, g, b, y], 1), sd["chroma_extractor.0.weight"], sd["chroma_extractor.0.bias"])
        c = np.tanh(conv3x3(np.maximum(c, 0), sd["chroma_extractor.2.weight"], sd["chroma_extractor.2.bias"]))
        return c[:, 0:1], c[:, 1:2]

    def linear(r, g, b):
        rgb = np.concatenate([r, g, b], 1)
        lin = np.einsum("ij,bjhw->bihw", M[:, :3], rgb) + M[:, 3][None, :, None, None]
        y = (lin * yw[None, :, None, None]).sum(1, keepdims=True)
        y = y / np.maximum(y.max(axis=(2, 3), keepdims=True), dt.type(eps))
        return lin.astype(dt), y.astype(dt)

    if variant == 0:
        wb = x * sd["wb_gains"].astype(dt)[None, :, None, None]
        h = np.maximum(conv3x3(wb, sd["demosaic_refine.0.weight"], sd["demosaic_refine.0.bias"]), 0)
        refined = softplus(conv3x3(h, sd["demosaic_refine.2.weight"], sd["demosaic_refine.2.bias"]))
        r, g, b = refined[:, 0:1], dt.type(0.5) * (refined[:, 1:2] + refined[:, 2:3]), refined[:, 3:4]
        lin, y = linear(r, g, b)
        cr, cb = chroma(r, g, b, y)
        return y, cr, cb, lin
    gains = softplus(sd["wb_gains"].astype(dt)) + dt.type(1e-6)
    wb = x * gains[None, :, None, None]
    r, g, b = wb[:, 0:1], dt.type(0.5) * (wb[:, 1:2] + wb[:, 2:3]), wb[:, 3:4]
    lin, y = linear(r, g, b)
    cr, cb = chroma(r, g, b, y)
    h = gelu_erf(conv3x3(lin, sd["demosaic_refine.0.weight"], sd["demosaic_refine.0.bias"]))
    refined = lin + conv3x3(h, sd["demosaic_refine.2.weight"], sd["demosaic_refine.2.bias"])
    return y, cr, cb, refined.astype(dt)


def camera_aware_color_correction(sd, x, variant=0):
    """CameraAwareColorCorrection.forward: TrueColorRawFormer.py:170-185 (variant 0), BayerTORGBColorMultiLvl.py:164-181
    (variant 1).  x [B,3,H,W] -> [B,3,H,W] in [0,1]."""
    dt = x.dtype
    gamma = sd["gamma"].astype(dt) if variant == 0 else softplus(sd["gamma_param"].astype(dt)) + dt.type(1e-6)
    v = np.power(np.clip(x, 0, 1), dt.type(1.0) / gamma).astype(dt)
    t = conv1x1(np.maximum(conv1x1(v, sd["color_transform.0.weight"], sd["color_transform.0.bias"]), 0),
                sd["color_transform.2.weight"], sd["color_transform.2.bias"])
    out = []
    for i in range(t.shape[1]):
        ch = t[:, i:i + 1]
        m = sigmoid(conv1x1(np.maximum(conv1x1(ch, sd["tone_curve.0.weight"], sd["tone_curve.0.bias"]), 0),
                            sd["tone_curve.2.weight"], sd["tone_curve.2.bias"]))
        out.append(m if variant == 0 else np.clip(ch * (dt.type(0.8) + dt.type(0.4) * m), 0, 1))
    return np.clip(np.concatenate(out, 1), 0, 1).astype(dt)


# ----------------------------------------------------------------------------------------------
# caller side of the forward: uint8 conversion, Bayer channel order, R/B auto-correction, PSNR  
# ----------------------------------------------------------------------------------------------
# Pinned: ``tests/golden/make_golden_post.py`` executes the reference's own ``correct_bayer_channels`` /
# ``auto_correct_rb`` (cut out of test.py, whose top-level imports need skimage / imageio) and stores their outputs.
# ``psnr_u8`` restates ``skimage.metrics.peak_signal_noise_ratio`` (scikit-image: version not pinned by the reference and
# not installed here) from its published definition -- parity UNPINNED for that one function.
BAYER_PERMS = {"BGGR": (2, 1, 0), "GBRG": (1, 0, 2), "GRBG": (0, 2, 1)}


def preprocess_u16(raw, black=512.0, white=16383.0, ratio=100.0, clamp=True):
    """RAW normalisation of the short exposure, WFB/load_dataset.py:88-89: ``np.clip(raw.astype(float32), 512, 16383)`` then
    ``(x - 512) / (16383 - 512 + 1e-6) * ap`` -- float32 throughout (the Python-float divisor is cast to float32 by numpy's
    scalar promotion) -- and, with ``clamp``, ``np.minimum(x, 1.0)`` (correctdataloader.py:103).  raw: uint16 array.
    Pinned by tests/golden/pre.npz (tests/golden/make_golden_pre.py executes the reference's own statements)."""
    x = np.clip(np.asarray(raw).astype(np.float32), np.float32(black), np.float32(white))
    x = (x - np.float32(black)) / np.float32(float(white) - float(black) + 1e-6) * np.float32(ratio)
    return np.minimum(x, np.float32(1.0)) if clamp else x


def postprocess_u8(pred):
    """test.py:117-118: clamp(pred, 0, 1)[b] -> HWC -> * 255 -> astype(uint8).  pred [B,3,H,W] -> [B,H,W,3]."""
    x = np.clip(np.asarray(pred, np.float32), 0.0, 1.0).transpose(0, 2, 3, 1)
    return (x * np.float32(255.0)).astype(np.uint8)


def correct_bayer_channels(rgb, pattern="RGGB"):
    """test.py:17-27: channel order by Bayer pattern; unknown patterns (and RGGB) are left alone."""
    perm = BAYER_PERMS.get(str(pattern).upper())
    return rgb if perm is None else rgb[..., list(perm)]


def auto_correct_rb(rgb):
    """test.py:29-38: swap R and B when the red mean is below the blue mean (ONE image, HWC)."""
    if rgb[..., 0].mean() < rgb[..., 2].mean():
        rgb = rgb[..., [2, 1, 0]]
    return rgb


def psnr_u8(a, b):
    """skimage.metrics.peak_signal_noise_ratio(a, b) for uint8 images as called in test.py:123: data_range = 255 (from the
    dtype), mse = mean((a - b)^2) in float64, 10 * log10(data_range^2 / mse)."""
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2, dtype=np.float64)
    return math.inf if err == 0 else 10.0 * math.log10(255.0 ** 2 / err)


def ssim_u8(a, b, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity(a, b, channel_axis=-1) for uint8 HWC images as called in test.py:124
    (scikit-image's published algorithm, defaults: uniform 7x7 window, use_sample_covariance=True, data_range from the
    dtype = 255; float64 arithmetic; mean of the SSIM map cropped by (win_size - 1) // 2, then mean over channels).
    Parity UNPINNED (scikit-image is not installed here); the windowed means use scipy.ndimage.uniform_filter like skimage."""
    from scipy.ndimage import uniform_filter

    if min(a.shape[0], a.shape[1]) < win_size:
        raise ValueError("win_size exceeds image extent")
    R = 255.0
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    pad = (win_size - 1) // 2
    vals = []
    for ch in range(a.shape[-1]):
        x, y = a[..., ch].astype(np.float64), b[..., ch].astype(np.float64)
        ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
        uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(vals))


# ----------------------------------------------------------------------------------------------
# WFB "WMB" block pieces: FEB / ProcessBlock / FFAB (WFB/blocks.py:11-92),
# Illumination_Estimator (WFB/model.py:174-200).  Pinned to tests/golden/wfb.npz (reference classes executed on CPU).
# ----------------------------------------------------------------------------------------------
def feb(sd, x):
    """FEB.forward, WFB/blocks.py:23-39."""
    dt = x.dtype
    H, W = x.shape[-2:]
    x = np.clip(x, -10, 10).astype(dt)                                              # :25
    pre = conv1x1(x, sd["fpre.weight"], sd["fpre.bias"])
    x_freq = np.fft.rfft2(pre, norm="ortho")                                        # :27
    # The four self-conjugate bins of a real input (k in {0, H/2}, f in {0, W/2}) are real.  torch's CPU FFT (MKL) returns
    # imag = +0 there (checked against the goldens); pocketfft leaves a rounding residue of either sign, which would flip the
    # phase of a negative real bin between +pi and -pi.
    ks = [0] + ([H // 2] if H % 2 == 0 else [])
    fs = [0] + ([W // 2] if W % 2 == 0 else [])
    for k in ks:
        for f in fs:
            x_freq[..., k, f] = x_freq[..., k, f].real
    mag = (np.abs(x_freq) + 1e-6).astype(dt)                                        # :28
    pha = np.angle(x_freq).astype(dt)                                               # :29

    def process(t, p):
        t = conv1x1(t, sd[p + ".0.weight"], sd[p + ".0.bias"])
        t = leaky_relu(t, 0.1)
        return conv1x1(t, sd[p + ".2.weight"], sd[p + ".2.bias"])

    mag = np.clip(process(mag, "process1"), 0, 1e4).astype(dt)                      # :30
    pha = process(pha, "process2")                                                  # :31
    z = (mag * np.cos(pha)) + 1j * (mag * np.sin(pha))                              # :32-34
    y = np.fft.irfft2(z, s=(H, W), norm="ortho").astype(dt)                         # :35
    return np.clip(y + x, -10, 10).astype(dt)                                       # :36-37


def process_block(sd, x):
    """ProcessBlock.forward, WFB/blocks.py:48-56."""
    xf = feb(_sd(sd, "frequency_process."), x)
    return conv1x1(xf, sd["cat.weight"], sd["cat.bias"]) + x


def ffab(sd, x):
    """FFAB.forward, WFB/blocks.py:83-92."""
    def tail(p, a, b):
        t = process_block(_sd(sd, p + ".0."), np.concatenate((a, b), 1))
        return conv1x1(t, sd[p + ".1.weight"], sd[p + ".1.bias"])

    x = process_block(_sd(sd, "conv0.1."), conv1x1(x, sd["conv0.0.weight"], sd["conv0.0.bias"]))
    x1 = process_block(_sd(sd, "conv1."), x)
    x2 = process_block(_sd(sd, "conv2."), x1)
    x3 = process_block(_sd(sd, "conv3."), x2)
    x4 = tail("conv4", x2, x3)
    x5 = tail("conv5", x1, x4)
    return tail("convout", x, x5)


def dwconv5x5(x, w, b=None):
    """nn.Conv2d(C, C, 5, padding=2, groups=C).  w [C,1,5,5]."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 4, W + 4), x.dtype)
    xp[:, :, 2:-2, 2:-2] = x
    y = np.zeros_like(x)
    w = w.astype(x.dtype)
    for dy in range(5):
        for dx in range(5):
            y += xp[:, :, dy:dy + H, dx:dx + W] * w[None, :, 0, dy, dx, None, None]
    if b is not None:
        y = y + b.astype(x.dtype)[None, :, None, None]
    return y


def illumination_estimator(sd, img):
    """Illumination_Estimator.forward, WFB/model.py:185-200 -> (illu_fea, illu_map)."""
    mean_c = img.mean(axis=1, keepdims=True).astype(img.dtype)
    inp = np.concatenate((img, mean_c), 1)
    x1 = conv1x1(inp, sd["conv1.weight"], sd["conv1.bias"])
    illu_fea = dwconv5x5(x1, sd["depth_conv.weight"], sd["depth_conv.bias"])
    illu_map = conv1x1(illu_fea, sd["conv2.weight"], sd["conv2.bias"])
    return illu_fea, illu_map
