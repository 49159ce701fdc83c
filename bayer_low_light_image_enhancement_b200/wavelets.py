"""Wavelet operators of the reference, as C-ABI calls into vectorised gather/scatter kernels.

* ``CustomDWT`` / ``CustomIDWT`` -- exist only in the reference README (README.md:92-144).  The default 4x4
  matrix is NOT a Haar basis and IDWT(DWT(x)) != x (README self-check MSE 0.36616483); that behaviour is
  reproduced, not fixed.
* ``dwt_init`` / ``iwt_init`` / ``DWT`` / ``IWT`` -- batch-concatenated Haar (RawFomer_WFB_FFAB/blocks.py:102-154).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream_ptr
from .modules import _Op

_README_KERNEL = [[1, 1, 1, 1], [1, -1, 1, 1], [1, 1, -1, 1], [1, 1, 1, -1]]


class _CustomWavelet(_Op):
    def __init__(self, kernel=None, use_custom=True, norm=True):
        super().__init__()
        k = torch.tensor(kernel if (use_custom and kernel is not None) else _README_KERNEL, dtype=torch.float32)
        k = k.view(4, 1, 2, 2)
        if norm:
            k = k / 2.0
        self.register_buffer("weight", k)

    def _k16(self):
        # the 4x4 matrix travels as a kernel argument: read it back from the device ONCE per weight version (a .cpu() per
        # call would synchronise the stream on every forward)
        key = (self.weight.data_ptr(), self.weight._version)
        if getattr(self, "_k16_key", None) != key:
            vals = self.weight.detach().float().cpu().reshape(16).tolist()
            self._k16_val, self._k16_key = (C.c_float * 16)(*vals), key
        return self._k16_val


class CustomDWT(_CustomWavelet):
    """[B,C,H,W] -> [B,4C,H/2,W/2], output channel = sub_band*C + c.  Reference: README.md:92-117."""

    def forward(self, x):
        x = self._prep(x)
        b, c, h, w = x.shape
        if h % 2 or w % 2:
            raise ValueError("CustomDWT needs even H and W (the reference's view() fails otherwise)")
        out = torch.empty(b, 4 * c, h // 2, w // 2, dtype=torch.float32, device=x.device)
        if out.numel():
            check(_lib.load().rf_custom_dwt(ptr(x), ptr(out), self._k16(), b, c, h, w, stream_ptr(x.device)),
                  "rf_custom_dwt")
        return out


class CustomIDWT(_CustomWavelet):
    """[B,4C,H,W] -> [B,C,2H,2W] (conv_transpose2d with the same 4x1x2x2 weight).  Reference: README.md:120-144."""

    def forward(self, x):
        x = self._prep(x)
        b, c4, h, w = x.shape
        if c4 % 4:
            raise ValueError("CustomIDWT needs a channel count divisible by 4")
        out = torch.empty(b, c4 // 4, 2 * h, 2 * w, dtype=torch.float32, device=x.device)
        if out.numel():
            check(_lib.load().rf_custom_idwt(ptr(x), ptr(out), self._k16(), b, c4 // 4, h, w, stream_ptr(x.device)),
                  "rf_custom_idwt")
        return out


def dwt_init(x):
    """[B,C,H,W] -> [4B,C,H/2,W/2] with LL,HL,LH,HH stacked on the batch axis.  Reference: WFB/blocks.py:102-115."""
    x = _Op._prep(x)
    b, c, h, w = x.shape
    if h % 2 or w % 2:
        raise ValueError("dwt_init needs even H and W")
    out = torch.empty(4 * b, c, h // 2, w // 2, dtype=torch.float32, device=x.device)
    if out.numel():
        check(_lib.load().rf_dwt_init(ptr(x), ptr(out), b, c, h, w, stream_ptr(x.device)), "rf_dwt_init")
    return out


def iwt_init(x):
    """[4B,C,H,W] -> [B,C,2H,2W], always float32.  Reference: WFB/blocks.py:119-136."""
    x = _Op._prep(x)
    b4, c, h, w = x.shape
    if b4 % 4:
        raise ValueError("iwt_init needs a batch divisible by 4")
    out = torch.empty(b4 // 4, c, 2 * h, 2 * w, dtype=torch.float32, device=x.device)
    if out.numel():
        check(_lib.load().rf_iwt_init(ptr(x), ptr(out), b4 // 4, c, h, w, stream_ptr(x.device)), "rf_iwt_init")
    return out


class DWT(nn.Module):
    """Reference: WFB/blocks.py:139-145."""

    def __init__(self):
        super().__init__()
        self.requires_grad = False

    def forward(self, x):
        return dwt_init(x)


class IWT(nn.Module):
    """Reference: WFB/blocks.py:148-154."""

    def __init__(self):
        super().__init__()
        self.requires_grad = False

    def forward(self, x):
        return iwt_init(x)
