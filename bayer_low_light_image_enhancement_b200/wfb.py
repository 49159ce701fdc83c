"""WFB "WMB" block pieces: ``Illumination_Estimator`` (RawFomer_WFB_FFAB/model.py:174-200) and the rFFT
amplitude / phase blocks ``FEB`` / ``ProcessBlock`` / ``FFAB`` (RawFomer_WFB_FFAB/blocks.py:11-92).

Same constructor signatures, parameter names and shapes as the reference classes (their ``state_dict`` loads with
``strict=True``); every ``forward`` is a sequence of C-ABI calls into fp32 CUDA kernels (``csrc/rf_wfb.cu``): 1x1 convolutions
with the block's clamps / LeakyReLU / residual in the epilogue, the 2-d real FFT as dense DFT matrix products (any H x W),
magnitude / phase and back, the 5x5 depthwise convolution.  ``WMB`` (model.py:203-245) is the reference's composition of these
with the batch-concat DWT / IWT, the WithBias LayerNorm and the gated-GELU FeedForward of this package; its high-band branch
``WM`` is ``mamba_ssm.Mamba`` (third party, version unpinned, absent from the reference tree) and has to be supplied by the
caller (``WMB(dim, mb=module)``) -- without it the block refuses to run (no fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream_ptr
from .modules import _Op

_PLANS: dict = {}


def _plan(h: int, w: int, device: torch.device) -> torch.Tensor:
    """Twiddle matrices of the H x W transform on `device` (made once, on the device, in double precision)."""
    key = (h, w, device.index if device.index is not None else torch.cuda.current_device())
    p = _PLANS.get(key)
    if p is None:
        lib = _lib.load()
        n = int(lib.rf_dft2_plan_floats(h, w))
        p = torch.empty(n, dtype=torch.float32, device=device)
        check(lib.rf_dft2_plan_init(ptr(p), h, w, stream_ptr(device)), "rf_dft2_plan_init")
        if len(_PLANS) >= 8:
            _PLANS.clear()
        _PLANS[key] = p
    return p


def rfft2_ortho(x: torch.Tensor) -> torch.Tensor:
    """torch.fft.rfft2(x, norm='ortho') of a real [B,C,H,W] tensor -> [B,C,2,H,W//2+1] (real plane, imaginary plane)."""
    x = _Op._prep(x)
    b, c, h, w = x.shape
    spec = torch.empty(b, c, 2, h, w // 2 + 1, dtype=torch.float32, device=x.device)
    if spec.numel():
        tmp = torch.empty_like(spec)
        check(_lib.load().rf_rfft2_ortho(ptr(x), ptr(_plan(h, w, x.device)), ptr(spec), ptr(tmp), b * c, h, w, stream_ptr(x.device)),
              "rf_rfft2_ortho")
    return spec


def irfft2_ortho(spec: torch.Tensor, w: int) -> torch.Tensor:
    """torch.fft.irfft2(re + i im, s=(H, w), norm='ortho') of spec [B,C,2,H,w//2+1] -> real [B,C,H,w]."""
    if spec.dim() != 5 or spec.shape[2] != 2 or spec.shape[4] != w // 2 + 1:
        raise ValueError("spec must be [B,C,2,H,W//2+1]")
    _lib.init_device(spec.device)
    spec = f32c(spec.detach())
    b, c, _, h, _ = spec.shape
    out = torch.empty(b, c, h, w, dtype=torch.float32, device=spec.device)
    if out.numel():
        tmp = torch.empty_like(spec)
        check(_lib.load().rf_irfft2_ortho(ptr(spec), ptr(_plan(h, w, spec.device)), ptr(out), ptr(tmp), b * c, h, w,
                                          stream_ptr(spec.device)), "rf_irfft2_ortho")
    return out


def _conv1x1(x, conv: nn.Conv2d, x2=None, in_clamp=0.0, act=0, out_clamp=None, resid=None):
    """conv (1x1) of cat(x, x2) with the block's input clamp / LeakyReLU(0.1) / output clamp / residual in the epilogue.
    x: [B,C,...] with any trailing shape (planes of P elements)."""
    b, ci = x.shape[0], x.shape[1]
    ci2 = 0 if x2 is None else x2.shape[1]
    co = conv.out_channels
    if conv.in_channels != ci + ci2 or conv.kernel_size != (1, 1):
        raise ValueError("conv does not match the input channels")
    p = x[0, 0].numel() if x.numel() else 0
    out = torch.empty((b, co) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    wt = f32c(conv.weight.detach()).reshape(co, ci + ci2)
    bs = None if conv.bias is None else f32c(conv.bias.detach())
    lo, hi = (0.0, 0.0) if out_clamp is None else out_clamp
    if out.numel():
        check(_lib.load().rf_conv1x1_nchw(ptr(x), ptr(x2), ptr(wt), ptr(bs), ptr(resid), ptr(out), ci, ci2, co, float(in_clamp),
                                          int(act), float(lo), float(hi), b, p, stream_ptr(x.device)), "rf_conv1x1_nchw")
    return out


class FEB(_Op):
    """Frequency enhancement block: clamp -> 1x1 -> rFFT2 -> two 1x1-LeakyReLU-1x1 stacks on magnitude and phase -> irFFT2 ->
    + clamped input -> clamp.  Reference: RawFomer_WFB_FFAB/blocks.py:11-39."""

    def __init__(self, nc):
        super().__init__()
        self.fpre = nn.Conv2d(nc, nc, 1, 1, 0)
        self.process1 = nn.Sequential(nn.Conv2d(nc, nc, 1, 1, 0), nn.LeakyReLU(0.1, inplace=True), nn.Conv2d(nc, nc, 1, 1, 0))
        self.process2 = nn.Sequential(nn.Conv2d(nc, nc, 1, 1, 0), nn.LeakyReLU(0.1, inplace=True), nn.Conv2d(nc, nc, 1, 1, 0))

    def forward(self, x):
        xin = self._prep(x)
        b, c, h, w = xin.shape
        lib = _lib.load()
        st = stream_ptr(xin.device)
        pre = _conv1x1(xin, self.fpre, in_clamp=10.0)
        spec = rfft2_ortho(pre)
        wf = w // 2 + 1
        mag = torch.empty(b, c, h, wf, dtype=torch.float32, device=xin.device)
        pha = torch.empty_like(mag)
        if mag.numel():
            check(lib.rf_spec_abs_angle(ptr(spec), ptr(mag), ptr(pha), b * c, h, w, st), "rf_spec_abs_angle")
        mag = _conv1x1(_conv1x1(mag, self.process1[0], act=1), self.process1[2], out_clamp=(0.0, 1e4))
        pha = _conv1x1(_conv1x1(pha, self.process2[0], act=1), self.process2[2])
        if mag.numel():
            check(lib.rf_spec_polar(ptr(mag), ptr(pha), ptr(spec), b * c, h, w, st), "rf_spec_polar")
        y = irfft2_ortho(spec, w)
        out = torch.empty_like(xin)
        if out.numel():
            check(lib.rf_add_clamp(ptr(y), ptr(xin), ptr(out), 10.0, out.numel(), st), "rf_add_clamp")
        return out


class ProcessBlock(_Op):
    """x + cat(FEB(x)).  Reference: RawFomer_WFB_FFAB/blocks.py:41-56."""

    def __init__(self, in_nc):
        super().__init__()
        self.spatial_process = nn.Identity()
        self.frequency_process = FEB(in_nc)
        self.cat = nn.Conv2d(in_nc, in_nc, 1, 1, 0)

    def forward(self, x):
        x = self._prep(x)
        return _conv1x1(self.frequency_process(x), self.cat, resid=x)


class FFAB(_Op):
    """Seven ProcessBlocks with dense skip concatenations.  Reference: RawFomer_WFB_FFAB/blocks.py:60-92."""

    def __init__(self, nc):
        super().__init__()
        self.conv0 = nn.Sequential(nn.Conv2d(nc, nc, 1, 1, 0), ProcessBlock(nc))
        self.conv1 = ProcessBlock(nc)
        self.conv2 = ProcessBlock(nc)
        self.conv3 = ProcessBlock(nc)
        self.conv4 = nn.Sequential(ProcessBlock(nc * 2), nn.Conv2d(nc * 2, nc, 1, 1, 0))
        self.conv5 = nn.Sequential(ProcessBlock(nc * 2), nn.Conv2d(nc * 2, nc, 1, 1, 0))
        self.convout = nn.Sequential(ProcessBlock(nc * 2), nn.Conv2d(nc * 2, nc, 1, 1, 0))

    @staticmethod
    def _tail(seq, a, b):
        return _conv1x1(seq[0](torch.cat((a, b), 1)), seq[1])

    def forward(self, x):
        x = self._prep(x)
        x = self.conv0[1](_conv1x1(x, self.conv0[0]))
        x1 = self.conv1(x)
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self._tail(self.conv4, x2, x3)
        x5 = self._tail(self.conv5, x1, x4)
        return self._tail(self.convout, x, x5)


class Illumination_Estimator(_Op):
    """illu_fea = depthwise5x5(conv1(cat(img, mean_c(img)))), illu_map = conv2(illu_fea).
    Reference: RawFomer_WFB_FFAB/model.py:174-200."""

    def __init__(self, n_fea_middle, n_fea_in=4, n_fea_out=3):
        super().__init__()
        self.conv1 = nn.Conv2d(n_fea_in, n_fea_middle, kernel_size=1, bias=True)
        self.depth_conv = nn.Conv2d(n_fea_middle, n_fea_middle, kernel_size=5, padding=2, bias=True, groups=n_fea_middle)
        self.conv2 = nn.Conv2d(n_fea_middle, n_fea_out, kernel_size=1, bias=True)

    def forward(self, img):
        img = self._prep(img, "img", self.conv1.in_channels - 1)
        b, c, h, w = img.shape
        lib = _lib.load()
        st = stream_ptr(img.device)
        mean_c = torch.empty(b, 1, h, w, dtype=torch.float32, device=img.device)
        mid = self.depth_conv.out_channels
        illu_fea = torch.empty(b, mid, h, w, dtype=torch.float32, device=img.device)
        if img.numel():
            check(lib.rf_channel_mean(ptr(img), ptr(mean_c), b, c, h * w, st), "rf_channel_mean")
        x1 = _conv1x1(img, self.conv1, x2=mean_c)
        if illu_fea.numel():
            dw = f32c(self.depth_conv.weight.detach())
            db = None if self.depth_conv.bias is None else f32c(self.depth_conv.bias.detach())
            check(lib.rf_dwconv5x5_nchw(ptr(x1), ptr(dw), ptr(db), ptr(illu_fea), b, mid, h, w, st), "rf_dwconv5x5_nchw")
        illu_map = _conv1x1(illu_fea, self.conv2)
        return illu_fea, illu_map


def data_transform(x):
    """model.py:10-11."""
    return 2 * x - 1.0


def inverse_data_transform(x):
    """model.py:14-15."""
    return torch.clamp((x + 1.0) / 2.0, 0.0, 1.0)


class WMB(_Op):
    """Wavelet block: norm1 -> 2x-1 -> DWT -> {LL: Illumination_Estimator -> FFAB; high bands: WM} -> IWT -> clamp((y+1)/2) ->
    x + . -> x + FeedForward(norm2(x)).  Reference: RawFomer_WFB_FFAB/model.py:203-245.

    ``mb`` is the high-band branch (the reference's ``WM``: two 3x3 convolutions around ``mamba_ssm.Mamba``).  mamba_ssm is a
    third-party CUDA package the reference does not pin or vendor, so this package does not restate it: pass a module with
    WM's call signature ([3B,C,h,w] -> [3B,C,h,w]); without one ``forward`` raises."""

    def __init__(self, dim, num_heads=1, ffn_expansion_factor=2.66, bias=True, LayerNorm_type="WithBias", mb=None):
        super().__init__()
        from .extras import FeedForward, WFBLayerNorm
        from .wavelets import DWT, IWT

        self.DWT = DWT()
        self.IWT = IWT()
        self.norm1 = WFBLayerNorm(dim, LayerNorm_type)
        self.illu = Illumination_Estimator(dim, n_fea_in=dim + 1, n_fea_out=dim)
        self.ffab = FFAB(dim)
        self.norm2 = WFBLayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)
        if mb is not None:
            self.mb = mb

    def forward(self, input_):
        if not hasattr(self, "mb"):
            raise NotImplementedError("WMB needs the high-band branch WM (mamba_ssm.Mamba, third party): pass mb=module")
        x = self._prep(input_)
        n = x.shape[0]
        x = data_transform(self.norm1(x))
        input_dwt = self.DWT(x)
        input_ll, input_high = input_dwt[:n, ...], input_dwt[n:, ...]
        input_ll, _ = self.illu(input_ll.contiguous())
        input_ll = self.ffab(input_ll)
        input_high = self.mb(input_high.contiguous())
        output = inverse_data_transform(self.IWT(torch.cat((input_ll, input_high), dim=0)))
        x = x + output
        return x + self.ffn(self.norm2(x))
