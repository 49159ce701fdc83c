"""B200-native RawFormer inference hot path (packed-Bayer 4-channel U-Net forward).

Public surface = the reference's own module names:

    from bayer_low_light_image_enhancement_b200 import RawFormer
    model = RawFormer(dim=32)            # or RawFormer(model_size='S'), precision='bf16'
    model.load_state_dict(reference_state_dict, strict=True)
    rgb = model.cuda().eval()(raw)        # raw [B,1,H,W] -> [B,3,H,W]

``multilevel`` holds the ``MultiLvlFrequencyawareLumaChromaAttentionRAWFormer.py`` variant (config 5).
All compute runs in ``librawformer_b200.so`` (hand-written sm_100a CUDA behind a C ABI); there is no fallback.
"""
from . import modules_ml as multilevel
from ._lib import LIB_PATH, exported_symbols
from .modules import (MODEL_SIZES, Attention, BayerLumaChroma, Conv_Transformer, Downsample, FLCA, HaarDWT, LayerNorm,
                      PixelShuffle, RawFormer, TransformerBlock, WaveTransformBlock, bayer_downshuffle, conv_ffn, downshuffle,
                      get_default_precision, set_default_precision)
from .modules_ml import FLCA_Pyramid
from .modules_ml import RawFormer as RawFormerMultiLevel
from .extras import (BiasFree_LayerNorm, FeedForward, WFBLayerNorm, WithBias_LayerNorm, correct_rgb_u8, postprocess_rgb_u8, postprocess_u8, preprocess_u16, psnr_u8,
                     ssim_u8)
from . import truecolor
from . import wfb
from .wfb import FEB, FFAB, WMB, Illumination_Estimator, ProcessBlock
from .pipeline import FramePipeline
from .rowtiled import LocalBands, RowTiledRawFormer, plan_bands
from .wavelets import DWT, IWT, CustomDWT, CustomIDWT, dwt_init, iwt_init

__all__ = [
    "RawFormer", "RawFormerMultiLevel", "Conv_Transformer", "WaveTransformBlock", "FLCA", "FLCA_Pyramid", "HaarDWT",
    "BayerLumaChroma", "Attention", "conv_ffn", "TransformerBlock", "LayerNorm", "Downsample", "PixelShuffle",
    "downshuffle", "bayer_downshuffle", "CustomDWT", "CustomIDWT", "DWT", "IWT", "dwt_init", "iwt_init", "multilevel", "MODEL_SIZES",
    "FeedForward", "WithBias_LayerNorm", "BiasFree_LayerNorm", "WFBLayerNorm", "postprocess_u8", "postprocess_rgb_u8", "correct_rgb_u8", "psnr_u8", "ssim_u8", "preprocess_u16", "truecolor", "FramePipeline", "RowTiledRawFormer", "LocalBands", "plan_bands", "set_default_precision", "get_default_precision", "LIB_PATH", "exported_symbols",
]
