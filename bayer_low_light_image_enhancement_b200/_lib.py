"""ctypes binding of ``librawformer_b200.so`` (the C ABI declared in ``include/rawformer_b200.h``).

The library is built in-tree by ``build.py`` (nvcc, sm_100a).  There is no CPU path: every operator in
this package raises if the shared library is missing or the tensor is not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "librawformer_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

RF_F32, RF_BF16 = 0, 1
RF_VARIANT_FLCA, RF_VARIANT_ML = 0, 1

_fp = C.c_void_p  # device pointers travel as void*


class BlockWeights(C.Structure):
    """Mirror of ``rf_block_weights`` (include/rawformer_b200.h)."""

    _names = (
        "flca_low_w flca_high_w flca_chroma_w flca_se_w1 flca_se_b1 flca_se_w2 flca_se_b2 flca_alpha flca_beta "
        "flca_gamma flca_filt norm1_w norm1_b temperature qkv_w qkv_b qkv_dw_w qkv_dw_b proj_w proj_b norm2_w "
        "norm2_b pw1_w pw1_b ffn_dw_w ffn_dw_b pw2_w pw2_b reduce_w reduce_b convout_w convout_b"
    ).split()
    _fields_ = (
        [(n, _fp) for n in _names]
        + [("pyr_low_w", _fp * 2), ("pyr_high_w", _fp * 2), ("pyr_gate_w", _fp * 2), ("pyr_gate_b", _fp * 2)]
        + [(n, _fp) for n in ("pyr_cgate_w", "pyr_cgate_b", "pyr_res_w0", "pyr_res_b0", "pyr_res_w2", "pyr_res_b2")]
    )


class ModelWeights(C.Structure):
    """Mirror of ``rf_model_weights``."""

    _fields_ = [
        ("embedding_w", _fp),
        ("embedding_b", _fp),
        ("blocks", BlockWeights * 7),
        ("down_w", _fp * 3),
        ("up_w", _fp * 3),
        ("up_b", _fp * 3),
        ("reduce_w", _fp * 3),
        ("reduce_b", _fp * 3),
        ("conv_out_w", _fp),
        ("conv_out_b", _fp),
        ("rgb_w_host", C.c_float * 3),
    ]


class Band(C.Structure):
    """Mirror of ``rf_band`` (row-tiled single frame)."""

    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("row0", C.c_int), ("rows", C.c_int), ("comm", _fp * 8),
                ("epoch", C.c_uint)]


_lib = None
_lock = threading.Lock()
_inited_devices = set()

_i, _sz, _f = C.c_int, C.c_size_t, C.c_float
_BWp, _MWp, _BDp = C.POINTER(BlockWeights), C.POINTER(ModelWeights), C.POINTER(Band)
_ubp = C.POINTER(C.c_ubyte)

# name -> (restype, argtypes).  Keep in sync with include/rawformer_b200.h (tests/test_abi.py checks the symbols).
_SIGS = {
    "rf_strerror": (C.c_char_p, [_i]),
    "rf_version": (_i, []),
    "rf_last_cuda_error": (_i, []),
    "rf_init": (_i, [_i]),
    "rf_set_tcgen05": (_i, [_i]),
    "rf_launch_count": (C.c_longlong, []),
    "rf_reset_launch_count": (None, []),
    "rf_downshuffle": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _fp]),
    "rf_pixelshuffle": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _fp]),
    "rf_custom_dwt": (_i, [_fp, _fp, C.POINTER(_f), _i, _i, _i, _i, _fp]),
    "rf_custom_idwt": (_i, [_fp, _fp, C.POINTER(_f), _i, _i, _i, _i, _fp]),
    "rf_haar_dwt": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp]),
    "rf_dwt_init": (_i, [_fp, _fp, _i, _i, _i, _i, _fp]),
    "rf_iwt_init": (_i, [_fp, _fp, _i, _i, _i, _i, _fp]),
    "rf_luma_chroma": (_i, [_fp, _fp, _fp, _fp, C.POINTER(_f), _f, _i, _i, _i, _fp, _sz, _fp]),
    "rf_layernorm": (_i, [_fp, _fp, _fp, _fp, _f, _i, _i, _i, _i, _i, _fp]),
    "rf_layernorm_rows": (_i, [_fp, _fp, _fp, _fp, _f, _i, C.c_longlong, _i, _fp]),
    "rf_block_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "rf_flca_forward": (_i, [_BWp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _sz, _fp]),
    "rf_attention_forward": (_i, [_BWp, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_conv_ffn_forward": (_i, [_BWp, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_transformer_block_forward": (_i, [_BWp, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_conv_transformer_forward": (_i, [_BWp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _sz, _fp]),
    "rf_downsample_forward": (_i, [_fp, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_feedforward_gated": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_model_packed_bytes": (_sz, [_i, _i, _i]),
    "rf_model_pack": (_i, [_MWp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_rawformer_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "rf_rawformer_forward": (_i, [_fp, _i, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp]),
    "rf_rawformer_forward_profiled": (
        _i,
        [_fp, _i, _i, _i, _fp, _fp, _i, _i, _i, _fp, _sz, _fp, C.POINTER(_f), C.POINTER(_i), _i, C.POINTER(_i)],
    ),
    "rf_band_comm_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "rf_band_comm_alloc": (_i, [_sz, C.POINTER(_fp), _ubp]),
    "rf_band_comm_open": (_i, [_ubp, C.POINTER(_fp)]),
    "rf_band_comm_close": (_i, [_fp]),
    "rf_band_comm_free": (_i, [_fp]),
    "rf_band_comm_status": (_i, [_fp, C.POINTER(_i), _fp]),
    "rf_band_comm_reset": (_i, [_fp, _fp]),
    "rf_band_out_rows": (_i, [_BDp, C.POINTER(_i), C.POINTER(_i)]),
    "rf_rawformer_band_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _BDp]),
    "rf_rawformer_forward_band": (_i, [_fp, _i, _i, _i, _fp, _fp, _i, _i, _BDp, _fp, _sz, _fp]),
    "rf_rawformer_forward_band_profiled": (
        _i, [_fp, _i, _i, _i, _fp, _fp, _i, _i, _BDp, _fp, _sz, _fp, C.POINTER(_f), C.POINTER(_i), _i, C.POINTER(_i)]),
    "rf_kernel_name": (C.c_char_p, [_i]),
    "rf_profiled_launch_info": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rf_conv3x3_small": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _fp]),
    "rf_truecolor_mix": (_i, [_fp, C.POINTER(_f), _i, C.POINTER(_f), C.POINTER(_f), _f, _fp, _fp, _fp, _fp, _i, _i, _i, _fp]),
    "rf_color_correction": (_i, [_fp, _f, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp]),
    "rf_postprocess_u8": (_i, [_fp, _fp, _i, _i, _i, _fp]),
    "rf_preprocess_u16": (_i, [_fp, _fp, _f, _f, _f, _i, _i, _i, _i, _fp]),
    "rf_postprocess_rgb_u8": (_i, [_fp, _fp, C.POINTER(_i), _i, _i, _i, _i, _fp, _sz, _fp]),
    "rf_correct_rgb_u8": (_i, [_fp, C.POINTER(_i), _i, _i, _i, _i, _fp, _sz, _fp]),
    "rf_sse_u8": (_i, [_fp, _fp, _fp, _i, C.c_longlong, _fp]),
    "rf_ssim_u8": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp]),
    "rf_dft2_plan_floats": (_sz, [_i, _i]),
    "rf_dft2_plan_init": (_i, [_fp, _i, _i, _fp]),
    "rf_rfft2_ortho": (_i, [_fp, _fp, _fp, _fp, C.c_longlong, _i, _i, _fp]),
    "rf_irfft2_ortho": (_i, [_fp, _fp, _fp, _fp, C.c_longlong, _i, _i, _fp]),
    "rf_spec_abs_angle": (_i, [_fp, _fp, _fp, C.c_longlong, _i, _i, _fp]),
    "rf_spec_polar": (_i, [_fp, _fp, _fp, C.c_longlong, _i, _i, _fp]),
    "rf_conv1x1_nchw": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _i, _f, _f, _i, C.c_longlong, _fp]),
    "rf_add_clamp": (_i, [_fp, _fp, _fp, _f, C.c_longlong, _fp]),
    "rf_channel_mean": (_i, [_fp, _fp, _i, _i, C.c_longlong, _fp]),
    "rf_dwconv5x5_nchw": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp]),
}


def exported_symbols():
    return sorted(_SIGS)


def load():
    """Load the in-tree shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bayer_low_light_image_enhancement_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def strerror(code: int) -> str:
    return load().rf_strerror(int(code)).decode()


def check(code: int, what: str = ""):
    if code != 0:
        lib = load()
        extra = ""
        if code == -4:
            extra = f" (cuda error {lib.rf_last_cuda_error()})"
        raise RuntimeError(f"rawformer_b200 {what}: {strerror(code)}{extra}")


def init_device(device: torch.device):
    """rf_init for the tensor's device (sm_100 check).  Raises on CPU tensors: there is no CPU path."""
    if device.type != "cuda":
        raise RuntimeError("rawformer_b200 operators need CUDA tensors on a B200 (sm_100a); there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        check(load().rf_init(idx), "rf_init")
        _inited_devices.add(idx)
    return idx


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous float32 view/copy on the same device."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def dtype_code(precision: str) -> int:
    p = str(precision).lower()
    if p in ("fp32", "f32", "float32"):
        return RF_F32
    if p in ("bf16", "bfloat16"):
        return RF_BF16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


class Workspace:
    """Caller-owned scratch memory (the C ABI never allocates): one buffer per (device, stream), grown on demand.

    Stream-ordered work of ONE stream may share a buffer; two streams never do (forwards on different streams would race
    on it).  A CUDA graph must NOT capture a buffer of this pool -- a later, larger request replaces it and the graph would
    replay into freed memory; graph owners allocate their own workspace and keep it with the graph
    (``RawFormer._forward_graph``, ``RowTiledRawFormer``)."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        stream = torch.cuda.current_stream(device)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the shared workspace must not be captured into a CUDA graph: allocate a workspace that "
                               "lives as long as the graph")
        key = (device.type, device.index, stream.cuda_stream)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None:
                b.record_stream(stream)          # work already enqueued on this stream may still use the old buffer
            # zero-filled once: a scratch buffer straight from the allocator may hold any bit pattern (open issue
            # "workspace history")
            b = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


_shared_ws = Workspace()


def shared_workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return _shared_ws.get(nbytes, device)
