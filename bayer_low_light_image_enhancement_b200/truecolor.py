"""TrueColor head / tail: the learned colour front end (``EnhancedBayerProcessor``) and tone-mapping tail
(``CameraAwareColorCorrection``) that ``TrueColorRawFormer.py`` and ``BayerTORGBColorMultiLvl.py`` put around the U-Net body.

Same constructor signatures, parameter / buffer names and shapes as the reference classes (their ``state_dict`` loads with
``strict=True``); every ``forward`` is a few C-ABI calls into fp32 CUDA kernels (``csrc/rf_truecolor.cu``).  The classes at
module level mirror ``TrueColorRawFormer.py:79-185``; ``multilevel`` holds the ``BayerTORGBColorMultiLvl.py:72-181`` variants
(softplus-positive white balance and gamma, GELU residual demosaic refinement on linear RGB, multiplicative tone curve).
"""
from __future__ import annotations

import ctypes as C
import math
import types

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream_ptr
from .modules import _Op

ACT = {"none": 0, "relu": 1, "softplus": 2, "tanh": 3, "gelu": 4}


def _conv3(x, conv: nn.Conv2d, act: str, in_scale=None, resid=None):
    b, ci, h, w = x.shape
    co = conv.out_channels
    out = torch.empty(b, co, h, w, dtype=torch.float32, device=x.device)
    wt = f32c(conv.weight.detach())
    bs = None if conv.bias is None else f32c(conv.bias.detach())
    if out.numel():
        check(_lib.load().rf_conv3x3_small(ptr(x), ptr(in_scale), ptr(wt), ptr(bs), ptr(resid), ptr(out), ci, co, ACT[act], b, h, w,
                                           stream_ptr(x.device)), "rf_conv3x3_small")
    return out


def _softplus_host(v: float) -> float:
    """F.softplus (beta 1, threshold 20) of one float32 value, evaluated in float32 like the reference's tensor op."""
    t = torch.tensor(v, dtype=torch.float32)
    return float(torch.nn.functional.softplus(t))


class EnhancedBayerProcessor(_Op):
    """x [B,4,H,W] (R, G1, G2, B planes) -> y [B,1,H,W], cr, cb [B,1,H,W], rgb [B,3,H,W].
    Reference: TrueColorRawFormer.py:79-142."""

    _variant = 0

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps
        self._build()
        self.register_buffer("y_weights", torch.tensor([0.2126, 0.7152, 0.0722], dtype=torch.float32))

    def _build(self):
        self.wb_gains = nn.Parameter(torch.tensor([1.0, 1.0, 1.0, 1.0], dtype=torch.float32))
        self.color_matrix = nn.Parameter(torch.eye(3, 4, dtype=torch.float32))
        self.demosaic_refine = nn.Sequential(nn.Conv2d(4, 32, 3, padding=1), nn.ReLU(inplace=True),
                                             nn.Conv2d(32, 4, 3, padding=1), nn.Softplus())
        self.chroma_extractor = nn.Sequential(nn.Conv2d(4, 16, 3, padding=1), nn.ReLU(inplace=True),
                                              nn.Conv2d(16, 2, 3, padding=1), nn.Tanh())

    def _mix(self, planes, gains, apply_gains):
        b, _, h, w = planes.shape
        dev = planes.device
        rgb_linear = torch.empty(b, 3, h, w, dtype=torch.float32, device=dev)
        chroma_in = torch.empty(b, 4, h, w, dtype=torch.float32, device=dev)
        y = torch.empty(b, 1, h, w, dtype=torch.float32, device=dev)
        ymax = torch.empty(max(b, 1), dtype=torch.float32, device=dev)
        cm = [float(v) for v in self.color_matrix.detach().float().cpu().reshape(-1).tolist()]
        if len(cm) != 12:
            raise ValueError("color_matrix must be [3,4]")
        yw = [float(v) for v in self.y_weights.detach().float().cpu().tolist()]
        if planes.numel():
            check(_lib.load().rf_truecolor_mix(ptr(planes), (C.c_float * 4)(*gains), int(apply_gains), (C.c_float * 12)(*cm),
                                               (C.c_float * 3)(*yw), float(self.eps), ptr(rgb_linear), ptr(chroma_in), ptr(y),
                                               ptr(ymax), b, h, w, stream_ptr(dev)), "rf_truecolor_mix")
        return rgb_linear, chroma_in, y

    def _chroma(self, chroma_in):
        c = _conv3(_conv3(chroma_in, self.chroma_extractor[0], "relu"), self.chroma_extractor[2], "tanh")
        return c[:, 0:1], c[:, 1:2]

    def forward(self, x):
        x = self._prep(x, "x", 4)
        gains = [float(v) for v in self.wb_gains.detach().float().cpu().tolist()]
        g_dev = f32c(self.wb_gains.detach())
        # white balance is applied on load of the first convolution; Softplus keeps the refined planes positive
        refined = _conv3(_conv3(x, self.demosaic_refine[0], "relu", in_scale=g_dev), self.demosaic_refine[2], "softplus")
        rgb_linear, chroma_in, y = self._mix(refined, gains, apply_gains=False)
        cr, cb = self._chroma(chroma_in)
        return y, cr, cb, rgb_linear


class CameraAwareColorCorrection(_Op):
    """x [B,3,H,W] -> clamp(x,0,1)^(1/gamma) -> 1x1 MLP 3-64-3 -> per-channel tone curve (1-32-1, sigmoid) -> [0,1].
    Reference: TrueColorRawFormer.py:148-185."""

    _variant = 0

    def __init__(self, out_channels=3):
        super().__init__()
        if out_channels != 3:
            raise NotImplementedError("kernels implement the reference configuration: 3 output channels")
        self._make_gamma()
        self.color_transform = nn.Sequential(nn.Conv2d(out_channels, 64, 1), nn.ReLU(inplace=True), nn.Conv2d(64, out_channels, 1))
        self.tone_curve = nn.Sequential(nn.Conv2d(1, 32, 1), nn.ReLU(inplace=True), nn.Conv2d(32, 1, 1), nn.Sigmoid())

    def _make_gamma(self):
        self.gamma = nn.Parameter(torch.tensor(2.2, dtype=torch.float32))

    def _gamma_value(self) -> float:
        return float(self.gamma.detach().float().cpu())

    def forward(self, x):
        x = self._prep(x, "x", 3)
        b, _, h, w = x.shape
        out = torch.empty_like(x)
        ct0, ct2, tc0, tc2 = self.color_transform[0], self.color_transform[2], self.tone_curve[0], self.tone_curve[2]
        t = [f32c(p.detach()) for p in (ct0.weight, ct0.bias, ct2.weight, ct2.bias, tc0.weight, tc0.bias, tc2.weight, tc2.bias)]
        if out.numel():
            check(_lib.load().rf_color_correction(ptr(x), self._gamma_value(), self._variant, *(ptr(v) for v in t), ptr(out), b, h, w,
                                                  stream_ptr(x.device)), "rf_color_correction")
        return out


class _EnhancedBayerProcessorML(EnhancedBayerProcessor):
    """Reference: BayerTORGBColorMultiLvl.py:72-136 -- positive white balance through softplus, colour matrix on the
    linearly demosaiced planes, chroma from (r, g, b, y), residual GELU refinement of the linear RGB."""

    _variant = 1

    def _build(self):
        self.wb_gains = nn.Parameter(torch.tensor([1.8, 1.0, 1.0, 1.6], dtype=torch.float32))
        self.color_matrix = nn.Parameter(torch.cat([torch.eye(3, 3, dtype=torch.float32), torch.zeros(3, 1, dtype=torch.float32)], dim=1))
        self.demosaic_refine = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.GELU(), nn.Conv2d(32, 3, 3, padding=1))
        self.chroma_extractor = nn.Sequential(nn.Conv2d(4, 16, 3, padding=1), nn.ReLU(inplace=True),
                                              nn.Conv2d(16, 2, 3, padding=1), nn.Tanh())

    def forward(self, x):
        x = self._prep(x, "x", 4)
        # gains = softplus(wb_gains) + 1e-6 (float32 arithmetic, like the reference's tensor ops)
        gains = [float(torch.tensor(_softplus_host(float(v)), dtype=torch.float32) + torch.tensor(1e-6, dtype=torch.float32))
                 for v in self.wb_gains.detach().float().cpu().tolist()]
        rgb_linear, chroma_in, y = self._mix(x, gains, apply_gains=True)
        cr, cb = self._chroma(chroma_in)
        refined = _conv3(_conv3(rgb_linear, self.demosaic_refine[0], "gelu"), self.demosaic_refine[2], "none", resid=rgb_linear)
        return y, cr, cb, refined


class _CameraAwareColorCorrectionML(CameraAwareColorCorrection):
    """Reference: BayerTORGBColorMultiLvl.py:141-181 -- gamma = softplus(gamma_param) + 1e-6; the tone curve scales each
    channel by 0.8 + 0.4 * sigmoid(.) instead of replacing it."""

    _variant = 1

    def _make_gamma(self):
        self.gamma_param = nn.Parameter(torch.tensor(2.2, dtype=torch.float32))

    def _gamma_value(self) -> float:
        g = torch.nn.functional.softplus(self.gamma_param.detach().float().cpu()) + 1e-6
        return float(g)


multilevel = types.SimpleNamespace(EnhancedBayerProcessor=_EnhancedBayerProcessorML,
                                   CameraAwareColorCorrection=_CameraAwareColorCorrectionML)
_EnhancedBayerProcessorML.__name__ = "EnhancedBayerProcessor"
_CameraAwareColorCorrectionML.__name__ = "CameraAwareColorCorrection"
