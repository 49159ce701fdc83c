"""Multi-level variant (BASELINE config 5): ``MultiLvlFrequencyawareLumaChromaAttentionRAWFormer.py`` (ML_RF).

Same U-Net skeleton as ``modules.RawFormer`` with ``FLCA_Pyramid`` (two-level luma pyramid, gated and
tanh-limited residual steps) instead of ``FLCA``, ``downN = nn.Sequential(conv)`` (state_dict key
``downN.0.weight``), and a tail that adds the colour-anchor mean correction and the LL-anchor luminance nudge.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .modules import Conv_Transformer as _BaseBlock
from .modules import FLCA as _BaseFLCA
from .modules import HaarDWT, _conv, _downsample_call, _fill
from .modules import RawFormer as _BaseRawFormer


class FLCA_Pyramid(_BaseFLCA):
    """Reference: ML_RF.py:86-183 (levels = 2, max_residual_scale = 0.2 are what the kernels implement)."""

    variant = _lib.RF_VARIANT_ML

    def __init__(self, channels, levels=2, r_ratio=8, eps=1e-8, max_residual_scale=0.2):
        nn.Module.__init__(self)
        if levels != 2 or abs(max_residual_scale - 0.2) > 1e-12 or abs(eps - 1e-8) > 1e-20:
            raise NotImplementedError("kernels implement levels=2, max_residual_scale=0.2, eps=1e-8 (reference defaults)")
        self.channels = channels
        self.levels = levels
        self.eps = eps
        self.max_residual_scale = float(max_residual_scale)
        self.dwt = HaarDWT()
        self.low_attn = nn.ModuleList(
            [nn.Sequential(_conv(1, channels, 3, bias=False), nn.Sigmoid()) for _ in range(levels)])
        self.high_attn = nn.ModuleList(
            [nn.Sequential(_conv(1, channels, 3, bias=False), nn.Tanh()) for _ in range(levels)])
        self.freq_gate_head = nn.ModuleList([_conv(2, 2, 1) for _ in range(levels)])
        self.chroma_attn = nn.Sequential(_conv(2, channels, 3, bias=False), nn.Sigmoid())
        self.chroma_gate = _conv(1, 1, 1)
        hidden = max(8, channels // r_ratio)
        self.se = nn.Sequential(nn.AdaptiveAvgPool2d(1), _conv(channels, hidden, 1), nn.ReLU(inplace=True),
                                _conv(hidden, channels, 1), nn.Sigmoid())
        self.res_proj = nn.Sequential(_conv(channels, channels, 1), nn.ReLU(inplace=True), _conv(channels, channels, 1))

    def _weights(self, bw, keep):
        _fill(bw, keep, flca_chroma_w=self.chroma_attn[0].weight, flca_se_w1=self.se[1].weight,
              flca_se_b1=self.se[1].bias, flca_se_w2=self.se[3].weight, flca_se_b2=self.se[3].bias,
              flca_filt=self.dwt.filt, pyr_cgate_w=self.chroma_gate.weight, pyr_cgate_b=self.chroma_gate.bias,
              pyr_res_w0=self.res_proj[0].weight, pyr_res_b0=self.res_proj[0].bias, pyr_res_w2=self.res_proj[2].weight,
              pyr_res_b2=self.res_proj[2].bias)
        for l in range(2):
            for name, t in (("pyr_low_w", self.low_attn[l][0].weight), ("pyr_high_w", self.high_attn[l][0].weight),
                            ("pyr_gate_w", self.freq_gate_head[l].weight), ("pyr_gate_b", self.freq_gate_head[l].bias)):
                t = _lib.f32c(t.detach())
                keep.append(t)
                getattr(bw, name)[l] = t.data_ptr()


class Conv_Transformer(_BaseBlock):
    """Reference: ML_RF.py:245-258."""

    variant = _lib.RF_VARIANT_ML

    def __init__(self, in_channel, num_heads=8, ffn_expansion_factor=2, flca_levels=2):
        self._levels = flca_levels
        super().__init__(in_channel, num_heads, ffn_expansion_factor)

    def _make_flca(self, c):
        return FLCA_Pyramid(c, levels=self._levels)


class RawFormer(_BaseRawFormer):
    """Reference: ML_RF.py:313-416."""

    variant = _lib.RF_VARIANT_ML

    def __init__(self, inp_channels=1, out_channels=3, dim=48, num_heads=[8, 8, 8, 8], ffn_expansion_factor=2,
                 flca_levels=2, model_size=None, precision=None):
        self._levels = flca_levels
        super().__init__(inp_channels, out_channels, dim, num_heads, ffn_expansion_factor, model_size, precision)

    def _block(self, c):
        return Conv_Transformer(c, 8, 2, self._levels)

    def _down(self, c):
        return nn.Sequential(_conv(c, c // 2, 3, bias=False))

    def _down_weight(self, n):
        return getattr(self, f"down{n}")[0].weight

    def _extra_init(self):
        self.haar = HaarDWT()

    @staticmethod
    def simple_demosaic_from_packed(x_ds):
        """(R, (G1+G2)/2, B) at packed resolution.  Reference: ML_RF.py:348-354 (host-side helper, torch ops)."""
        return torch.cat([x_ds[:, 0:1], 0.5 * (x_ds[:, 1:2] + x_ds[:, 2:3]), x_ds[:, 3:4]], dim=1)
