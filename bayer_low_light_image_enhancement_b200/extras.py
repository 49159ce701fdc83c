"""Operators either side of the forward and the WFB gated-GELU FFN, as C-ABI calls.

* ``postprocess_u8``   -- test.py:117-118: clamp(pred,0,1) -> *255 -> uint8 (truncation) -> HWC.
* ``preprocess_u16``   -- WFB/load_dataset.py:88-89 + correctdataloader.py:103 on a uint16 Bayer frame.
* ``FeedForward``      -- RawFomer_WFB_FFAB/model.py:42-87 (eval mode): the two Conv2d_BN branches and the identity are folded
  into one depthwise 3x3 exactly like the reference's own ``fuse()``; hidden = int(dim*factor) is zero-padded to a multiple of 8.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream_ptr
from .modules import _Op


def postprocess_u8(pred: torch.Tensor) -> torch.Tensor:
    """[B,3,H,W] float -> [B,H,W,3] uint8, the arithmetic of test.py:117-118 on the device."""
    pred = _Op._prep(pred, "pred", 3)
    b, _, h, w = pred.shape
    out = torch.empty(b, h, w, 3, dtype=torch.uint8, device=pred.device)
    if out.numel():
        check(_lib.load().rf_postprocess_u8(ptr(pred), ptr(out), b, h, w, stream_ptr(pred.device)), "rf_postprocess_u8")
    return out


BAYER_PERMS = {"RGGB": (0, 1, 2), "BGGR": (2, 1, 0), "GBRG": (1, 0, 2), "GRBG": (0, 2, 1)}   # test.py:17-27


def _perm(pattern):
    p = BAYER_PERMS.get(str(pattern).upper(), (0, 1, 2))      # unknown patterns are left alone, like the reference
    return (C.c_int * 3)(*p)


def postprocess_rgb_u8(pred: torch.Tensor, pattern: str = "RGGB", auto_rb: bool = True) -> torch.Tensor:
    """test.py:117-120 on the device: clamp -> *255 -> uint8 -> HWC -> ``correct_bayer_channels(pattern)`` ->
    ``auto_correct_rb`` (per image).  [B,3,H,W] float -> [B,H,W,3] uint8."""
    pred = _Op._prep(pred, "pred", 3)
    b, _, h, w = pred.shape
    out = torch.empty(b, h, w, 3, dtype=torch.uint8, device=pred.device)
    if out.numel():
        ws = torch.empty(2 * b, dtype=torch.int64, device=pred.device)
        check(_lib.load().rf_postprocess_rgb_u8(ptr(pred), ptr(out), _perm(pattern), int(bool(auto_rb)), b, h, w, ptr(ws),
                                                ws.numel() * 8, stream_ptr(pred.device)), "rf_postprocess_rgb_u8")
    return out


def correct_rgb_u8(img: torch.Tensor, pattern: str = "RGGB", auto_rb: bool = True) -> torch.Tensor:
    """``auto_correct_rb(correct_bayer_channels(img, pattern))`` (test.py:111-113) for a uint8 [B,H,W,3] image; returns a
    corrected copy."""
    if img.dim() != 4 or img.shape[-1] != 3 or img.dtype != torch.uint8:
        raise ValueError("img must be a uint8 tensor [B,H,W,3]")
    _lib.init_device(img.device)
    out = img.contiguous().clone()
    b, h, w, _ = out.shape
    if out.numel():
        ws = torch.empty(2 * b, dtype=torch.int64, device=out.device)
        check(_lib.load().rf_correct_rgb_u8(ptr(out), _perm(pattern), int(bool(auto_rb)), b, h, w, ptr(ws), ws.numel() * 8,
                                            stream_ptr(out.device)), "rf_correct_rgb_u8")
    return out


def psnr_u8(a: torch.Tensor, b: torch.Tensor):
    """``skimage.metrics.peak_signal_noise_ratio(a, b)`` for uint8 images (test.py:123), one value per image of the batch:
    the squared error is summed exactly (uint64) on the device; 10*log10(255^2 / mse) in float64 on the host."""
    import math

    if a.shape != b.shape or a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.dim() < 2:
        raise ValueError("a and b must be uint8 tensors of the same shape [B,...]")
    _lib.init_device(a.device)
    a, b = a.contiguous(), b.contiguous()
    nb = a.shape[0]
    n = a[0].numel() if nb else 0
    sse = torch.zeros(max(nb, 1), dtype=torch.int64, device=a.device)
    if nb:
        check(_lib.load().rf_sse_u8(ptr(a), ptr(b), ptr(sse), nb, n, stream_ptr(a.device)), "rf_sse_u8")
    out = []
    for v in sse[:nb].cpu().tolist():
        mse = v / n if n else 0.0
        out.append(math.inf if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse))
    return out


def ssim_u8(a: torch.Tensor, b: torch.Tensor):
    """``skimage.metrics.structural_similarity(a, b, channel_axis=-1)`` for uint8 [B,H,W,3] images (test.py:124), one value
    per image: 7x7 uniform window, data_range 255, sample covariance, border of 3 cropped.  H, W >= 7."""
    if a.shape != b.shape or a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.dim() != 4 or a.shape[-1] != 3:
        raise ValueError("a and b must be uint8 tensors of the same shape [B,H,W,3]")
    nb, h, w, _ = a.shape
    if h < 7 or w < 7:
        raise ValueError("win_size exceeds image extent: H and W must be at least 7")
    _lib.init_device(a.device)
    a, b = a.contiguous(), b.contiguous()
    acc = torch.zeros(max(nb, 1), dtype=torch.float64, device=a.device)
    if nb:
        check(_lib.load().rf_ssim_u8(ptr(a), ptr(b), ptr(acc), nb, h, w, stream_ptr(a.device)), "rf_ssim_u8")
    n = 3.0 * (h - 6) * (w - 6)
    return [v / n for v in acc[:nb].cpu().tolist()]


def preprocess_u16(raw: torch.Tensor, black: float = 512.0, white: float = 16383.0, ratio: float = 100.0,
                   clamp: bool = True, out: torch.Tensor = None) -> torch.Tensor:
    """uint16 Bayer [B,H,W] -> float32 [B,1,H,W]: clip to [black, white], subtract the black level, scale by
    ratio / (white - black + 1e-6) (WFB/load_dataset.py:88-89); ``clamp`` adds min(., 1) (correctdataloader.py:103).
    ``out`` (optional): a float32 [B,1,H,W] tensor to write into."""
    if raw.dim() != 3:
        raise ValueError("raw must be [B,H,W]")
    _lib.init_device(raw.device)
    if raw.dtype not in (torch.uint16, torch.int16):
        raise ValueError("raw must be a 16-bit integer tensor")
    raw = raw.contiguous()
    b, h, w = raw.shape
    if out is None:
        out = torch.empty(b, 1, h, w, dtype=torch.float32, device=raw.device)
    elif tuple(out.shape) != (b, 1, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != raw.device:
        raise ValueError("out must be a contiguous float32 [B,1,H,W] tensor on the device of raw")
    if out.numel():
        check(_lib.load().rf_preprocess_u16(ptr(raw), ptr(out), float(black), float(white), float(ratio), int(bool(clamp)),
                                            b, h, w, stream_ptr(raw.device)), "rf_preprocess_u16")
    return out


class _RowsLayerNorm(_Op):
    """Shared body of the two hand-written LayerNorms of the WFB variant: normalises the LAST axis of x."""

    _mode = 0

    def __init__(self, normalized_shape):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        normalized_shape = torch.Size(normalized_shape)
        assert len(normalized_shape) == 1
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.normalized_shape = normalized_shape

    def forward(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() < 1 or x.shape[-1] != self.normalized_shape[0]:
            raise ValueError(f"last axis of x must have {self.normalized_shape[0]} elements")
        _lib.init_device(x.device)
        xin = f32c(x.detach())
        out = torch.empty_like(xin)
        c = xin.shape[-1]
        rows = xin.numel() // c if c else 0
        bias = getattr(self, "bias", None)
        wt = f32c(self.weight.detach())
        bs = None if bias is None else f32c(bias.detach())
        if rows:
            check(_lib.load().rf_layernorm_rows(ptr(xin), ptr(wt), ptr(bs), ptr(out), 1e-5, self._mode, rows, c,
                                                stream_ptr(x.device)), "rf_layernorm_rows")
        return out


class BiasFree_LayerNorm(_RowsLayerNorm):
    """x / sqrt(var(x) + 1e-5) * weight over the last axis -- the mean is NOT subtracted from x.
    Reference: RawFomer_WFB_FFAB/model.py:89-103."""

    _mode = 1


class WithBias_LayerNorm(_RowsLayerNorm):
    """(x - mean) / sqrt(var + 1e-5) * weight + bias over the last axis.  Reference: RawFomer_WFB_FFAB/model.py:106-122."""

    _mode = 0

    def __init__(self, normalized_shape):
        super().__init__(normalized_shape)
        self.bias = nn.Parameter(torch.zeros(self.normalized_shape))


class WFBLayerNorm(_Op):
    """``LayerNorm(dim, LayerNorm_type)`` of the WFB variant on NCHW tensors ('BiasFree' or anything else = with bias).
    Reference: RawFomer_WFB_FFAB/model.py:125-135 (to_3d -> body -> to_4d)."""

    def __init__(self, dim, LayerNorm_type):
        super().__init__()
        self.body = BiasFree_LayerNorm(dim) if LayerNorm_type == "BiasFree" else WithBias_LayerNorm(dim)

    def forward(self, x):
        x = self._prep(x, "x", self.body.normalized_shape[0])
        b, c, h, w = x.shape
        out = torch.empty_like(x)
        wt = f32c(self.body.weight.detach())
        bias = getattr(self.body, "bias", None)
        bs = None if bias is None else f32c(bias.detach())
        if out.numel():
            check(_lib.load().rf_layernorm(ptr(x), ptr(wt), ptr(bs), ptr(out), 1e-5, self.body._mode, b, c, h, w,
                                           stream_ptr(x.device)), "rf_layernorm")
        return out


class _Conv2dBN(nn.Sequential):
    """Parameter container with the reference's names ('c', 'bn').  Reference: WFB/model.py:17-25."""

    def __init__(self, ch, ks):
        super().__init__()
        self.add_module("c", nn.Conv2d(ch, ch, ks, 1, ks // 2, groups=ch, bias=False))
        self.add_module("bn", nn.BatchNorm2d(ch))


class FeedForward(_Op):
    """Gated-GELU FFN of the WFB variant, inference (eval-mode BatchNorm).  Reference: WFB/model.py:42-65."""

    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.dim, self.hidden = dim, hidden
        self.rep_conv1 = _Conv2dBN(hidden, 3)
        self.rep_conv2 = _Conv2dBN(hidden, 1)
        self.project_in = nn.Conv2d(dim, hidden, 1, bias=bias)
        self.dwconv = nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden, bias=bias)
        self.project_out = nn.Conv2d(hidden, dim, 1, bias=bias)

    @torch.no_grad()
    def _folded(self, device):
        """x1 = x + BN(dw3(x)) + BN(dw1(x)) as one depthwise 3x3 + bias (the reference's fuse(), WFB/model.py:67-87)."""
        def bn_fold(seq):
            s = seq.bn.weight / torch.sqrt(seq.bn.running_var + seq.bn.eps)
            return seq.c.weight * s[:, None, None, None], seq.bn.bias - seq.bn.running_mean * s

        w3, b3 = bn_fold(self.rep_conv1)
        w1, b1 = bn_fold(self.rep_conv2)
        wa = w3.clone()
        wa[:, :, 1, 1] += w1[:, :, 0, 0] + 1.0
        ba = b3 + b1
        hp = (self.hidden + 7) // 8 * 8
        pad = hp - self.hidden

        def padn(t, dim=0):
            if pad == 0 or t is None:
                return None if t is None else f32c(t.to(device))
            shape = list(t.shape)
            shape[dim] = pad
            return f32c(torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim).to(device))

        return dict(
            hp=hp, w_in=padn(self.project_in.weight), b_in=padn(self.project_in.bias), wa=padn(wa), ba=padn(ba),
            wb=padn(self.dwconv.weight), bb=padn(self.dwconv.bias), w_out=padn(self.project_out.weight, 1),
            b_out=None if self.project_out.bias is None else f32c(self.project_out.bias.to(device)))

    def forward(self, x):
        if self.training:
            raise RuntimeError("FeedForward kernels implement inference (call .eval()): BatchNorm uses running statistics")
        x = self._prep(x, "x", self.dim)
        b, c, h, w = x.shape
        f = self._folded(x.device)
        hp = f["hp"]
        lib = _lib.load()
        es = 2 if self._dtype() == _lib.RF_BF16 else 4
        nbytes = (b * h * w * (2 * c + 2 * hp) + 2 * hp * c) * es + 4 * (24 * hp + 2 * c) + (1 << 16)
        ws = _lib.shared_workspace(nbytes, x.device)
        out = torch.empty_like(x)
        check(lib.rf_feedforward_gated(ptr(f["w_in"]), ptr(f["b_in"]), ptr(f["wa"]), ptr(f["ba"]), ptr(f["wb"]), ptr(f["bb"]),
                                       ptr(f["w_out"]), ptr(f["b_out"]), c, hp, self._dtype(), ptr(x), ptr(out), b, h, w,
                                       ptr(ws), ws.numel(), stream_ptr(x.device)), "rf_feedforward_gated")
        return out
