/*
 * rawformer_b200.h — C ABI of the B200-native RawFormer inference hot path.
 *
 * The reference (Gaurav14cs17/Bayer_Low_light_Image_Enhancement) is pure PyTorch and has no FFI; the
 * entry points below are what a binding for its hot path would call.  Each one cites the reference
 * interface it replaces (file:line relative to the reference tree; FLCA_RF = Frequencyaware-
 * LumaChromaAttentionRAWFormer.py, ML_RF = MultiLvlFrequencyawareLumaChromaAttentionRAWFormer.py,
 * WFB = RawFomer_WFB_FFAB/).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the parameter name ends in `_host`.
 *  - Tensors at this boundary use the reference's own layout: contiguous NCHW float32.
 *  - The caller owns every buffer, including the workspace; nothing is allocated or freed here.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*); calls are re-entrant and
 *    stream-ordered, and may be captured into a CUDA graph.
 *  - Return value: RF_OK (0) or a negative rf_status; nothing is thrown across the ABI.
 *  - sm_100a only: rf_init() fails with RF_ERR_ARCH on any other device.  There is no CPU fallback.
 *  - `dtype` selects the INTERNAL precision: RF_F32 (fp32 activations, fp32 FFMA; parity mode,
 *    max-abs <= 1e-4 vs the reference) or RF_BF16 (bf16 activations and tensor-core operands, fp32
 *    accumulation and statistics).
 */
#ifndef RAWFORMER_B200_H
#define RAWFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rf_status {
  RF_OK = 0,
  RF_ERR_BAD_SHAPE = -1,   /* shape rule violated (H,W multiples of 16, dim % 8, even DWT sizes ...) */
  RF_ERR_BAD_ARG = -2,     /* null / misaligned pointer, unknown enum */
  RF_ERR_ARCH = -3,        /* device is not sm_100 */
  RF_ERR_CUDA = -4,        /* a CUDA runtime call or launch failed (see rf_last_cuda_error) */
  RF_ERR_WORKSPACE = -5,   /* workspace too small */
  RF_ERR_UNSUPPORTED = -6  /* valid request this build does not implement */
} rf_status;

typedef enum rf_dtype { RF_F32 = 0, RF_BF16 = 1 } rf_dtype;

/* Model variant: FLCA_RF.py::RawFormer or ML_RF.py::RawFormer (flca_levels = 2). */
typedef enum rf_variant { RF_VARIANT_FLCA = 0, RF_VARIANT_ML = 1 } rf_variant;

const char* rf_strerror(int status);
int rf_version(void);
/* Last CUDA error code observed by this library on the calling thread (0 if none). */
int rf_last_cuda_error(void);
/* Checks that `device` is an sm_100 part and caches its properties.  Must succeed before any other call. */
int rf_init(int device);
/* Number of kernels this library launched on the calling thread since the last rf_reset_launch_count(). */
/* bf16 mode: 1 (default) = contractions run on the tcgen05/TMEM/TMA kernels, 0 = on the CUDA-core FFMA kernels
 * (A/B testing of the two device paths; both are CUDA).  enable < 0 only queries.  Returns the previous setting. */
int rf_set_tcgen05(int enable);
long long rf_launch_count(void);
void rf_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Index / wavelet operators (bit-exact index work).  NCHW float32 in and out.
 * ---------------------------------------------------------------------------------------------- */

/* downshuffle(var, r)  — FLCA_RF.py:18-33.  in [B,C,H,W] -> out [B,C*r*r,H/r,W/r], channel c*r*r+r*i+j. */
int rf_downshuffle(const float* in, float* out, int B, int C, int H, int W, int r, void* stream);
/* nn.PixelShuffle(r) — FLCA_RF.py:328,369.  in [B,C*r*r,H,W] -> out [B,C,H*r,W*r]. */
int rf_pixelshuffle(const float* in, float* out, int B, int C_out, int H, int W, int r, void* stream);
/* CustomDWT.forward — README.md:111-117.  k16_host = the 4x4 matrix (after the /2 of norm=True),
 * row n = sub-band, column t = tap (a,b,c,d = TL,TR,BL,BR).  in [B,C,H,W] -> out [B,4C,H/2,W/2] (n*C+c). */
int rf_custom_dwt(const float* in, float* out, const float* k16_host, int B, int C, int H, int W, void* stream);
/* CustomIDWT.forward — README.md:139-144.  in [B,4C,H,W] -> out [B,C,2H,2W]. */
int rf_custom_idwt(const float* in, float* out, const float* k16_host, int B, int C, int H, int W, void* stream);
/* HaarDWT.forward — FLCA_RF.py:56-73.  filt = device [4,1,2,2] buffer (`dwt.filt`).  in [B,C,H,W];
 * LL,LH,HL,HH each [B,C,ceil(H/2),ceil(W/2)] (reflect pad right/bottom for odd sizes). */
int rf_haar_dwt(const float* in, const float* filt, float* LL, float* LH, float* HL, float* HH,
                int B, int C, int H, int W, void* stream);
/* dwt_init — WFB/blocks.py:102-115.  in [B,C,H,W] -> out [4B,C,H/2,W/2] (LL,HL,LH,HH on the batch axis). */
int rf_dwt_init(const float* in, float* out, int B, int C, int H, int W, void* stream);
/* iwt_init — WFB/blocks.py:119-136.  in [4B,C,H,W] -> out [B,C,2H,2W]. */
int rf_iwt_init(const float* in, float* out, int B, int C, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Guidance and normalisation
 * ---------------------------------------------------------------------------------------------- */

/* BayerLumaChroma.forward — FLCA_RF.py:87-97.  x_ds [B,4,h,w] -> y,cr,cb [B,1,h,w].
 * rgb_w_host = {r_w,g_w,b_w}.  workspace: 4*B bytes. */
int rf_luma_chroma(const float* x_ds, float* y, float* cr, float* cb, const float* rgb_w_host, float eps,
                   int B, int h, int w, void* workspace, size_t workspace_bytes, void* stream);
/* LayerNorm.forward — FLCA_RF.py:185-187 (mode 0, nn.LayerNorm over C per pixel);
 * WithBias_LayerNorm / BiasFree_LayerNorm — WFB/model.py:89-122 (mode 0 / mode 1: no mean subtraction, no bias). */
int rf_layernorm(const float* in, const float* weight, const float* bias, float* out, float eps, int mode,
                 int B, int C, int H, int W, void* stream);
/* The same two normalisations on channels-last rows [rows, C] (the 'b (h w) c' tensors WithBias_LayerNorm /
 * BiasFree_LayerNorm.forward receive, WFB/model.py:102-103,119-122).  C % 8 == 0; bias may be NULL for mode 1. */
int rf_layernorm_rows(const float* in, const float* weight, const float* bias, float* out, float eps, int mode,
                      long long rows, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv_Transformer ("WaveTransformBlock") and its parts.
 * Weights are passed in PyTorch's native layouts (float32, device); absent parts may be NULL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rf_block_weights {
  /* FLCA — FLCA_RF.py:103-134 */
  const float* flca_low_w;     /* low_attn.0.weight    [C,1,3,3] */
  const float* flca_high_w;    /* high_attn.0.weight   [C,1,3,3] */
  const float* flca_chroma_w;  /* chroma_attn.0.weight [C,2,3,3] */
  const float* flca_se_w1;     /* se.1.weight [hid,C,1,1], hid = max(8, C/8) */
  const float* flca_se_b1;     /* se.1.bias   [hid] */
  const float* flca_se_w2;     /* se.3.weight [C,hid,1,1] */
  const float* flca_se_b2;     /* se.3.bias   [C] */
  const float* flca_alpha;     /* 0-d */
  const float* flca_beta;      /* 0-d */
  const float* flca_gamma;     /* 0-d */
  const float* flca_filt;      /* dwt.filt [4,1,2,2] */
  /* TransformerBlock — FLCA_RF.py:238-254 */
  const float* norm1_w;        /* norm1.body.weight [C] */
  const float* norm1_b;
  const float* temperature;    /* attn.temperature [8,1,1] */
  const float* qkv_w;          /* attn.qkv.weight [3C,C,1,1] */
  const float* qkv_b;
  const float* qkv_dw_w;       /* attn.qkv_dwconv.weight [3C,1,3,3] */
  const float* qkv_dw_b;
  const float* proj_w;         /* attn.project_out.weight [C,C,1,1] */
  const float* proj_b;
  const float* norm2_w;
  const float* norm2_b;
  const float* pw1_w;          /* ffn.pointwise1.weight [2C,C,1,1] */
  const float* pw1_b;
  const float* ffn_dw_w;       /* ffn.depthwise.weight [2C,1,3,3] */
  const float* ffn_dw_b;
  const float* pw2_w;          /* ffn.pointwise2.weight [C,2C,1,1] */
  const float* pw2_b;
  /* Conv_Transformer tail — FLCA_RF.py:268-277 */
  const float* reduce_w;       /* channel_reduce.weight [C,2C,1,1] */
  const float* reduce_b;
  const float* convout_w;      /* Conv_out.weight [C,C,3,3] */
  const float* convout_b;
  /* FLCA_Pyramid extras — ML_RF.py:86-120 (NULL for RF_VARIANT_FLCA) */
  const float* pyr_low_w[2];   /* low_attn.{l}.0.weight  [C,1,3,3] */
  const float* pyr_high_w[2];  /* high_attn.{l}.0.weight [C,1,3,3] */
  const float* pyr_gate_w[2];  /* freq_gate_head.{l}.weight [2,2,1,1] */
  const float* pyr_gate_b[2];  /* freq_gate_head.{l}.bias [2] */
  const float* pyr_cgate_w;    /* chroma_gate.weight [1,1,1,1] */
  const float* pyr_cgate_b;    /* chroma_gate.bias [1] */
  const float* pyr_res_w0;     /* res_proj.0.weight [C,C,1,1] */
  const float* pyr_res_b0;
  const float* pyr_res_w2;     /* res_proj.2.weight [C,C,1,1] */
  const float* pyr_res_b2;
} rf_block_weights;

size_t rf_block_workspace_bytes(int C, int dtype, int B, int Hf, int Wf, int Hy, int Wy);

/* FLCA.forward(feat, y, cr, cb) — FLCA_RF.py:136-162 (variant ML: FLCA_Pyramid.forward, ML_RF.py:132-183, levels = 2).
 * feat [B,C,Hf,Wf]; y,cr,cb [B,1,Hy,Wy]. */
int rf_flca_forward(const rf_block_weights* w, int C, int dtype, int variant, const float* feat, const float* y,
                    const float* cr, const float* cb, float* out, int B, int Hf, int Wf, int Hy, int Wy, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Attention.forward(x) — FLCA_RF.py:221-235 (num_heads = 8). */
int rf_attention_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                         void* workspace, size_t workspace_bytes, void* stream);
/* conv_ffn.forward(x) — FLCA_RF.py:204-209 (hidden = 2C). */
int rf_conv_ffn_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                        void* workspace, size_t workspace_bytes, void* stream);
/* TransformerBlock.forward(x) — FLCA_RF.py:251-254. */
int rf_transformer_block_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H,
                                 int W, void* workspace, size_t workspace_bytes, void* stream);
/* Conv_Transformer.forward(feat, y, cr, cb) — FLCA_RF.py:272-278 (variant ML: ML_RF.py:252-258). */
int rf_conv_transformer_forward(const rf_block_weights* w, int C, int dtype, int variant, const float* feat,
                                const float* y, const float* cr, const float* cb, float* out, int B, int Hf, int Wf,
                                int Hy, int Wy, void* workspace, size_t workspace_bytes, void* stream);
/* Downsample.forward(x) — FLCA_RF.py:176-177.  conv_w = body.0.weight [C/2,C,3,3]; out [B,2C,H/2,W/2]. */
int rf_downsample_forward(const float* conv_w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* stream);

/* FeedForward.forward(x) — WFB/model.py:42-65 (gated-GELU FFN) with the two eval-mode Conv2d_BN branches and the
 * identity already folded into ONE depthwise 3x3 + bias (the reference's own fuse(), WFB/model.py:67-87):
 *   t = project_in(x); x1 = dwA(t); x2 = dwB(t); y = project_out(gelu(x2)*x1 + gelu(x1)*x2) + x.
 * project_in_w [hidden,C,1,1], dwA_w/dwB_w [hidden,1,3,3], project_out_w [C,hidden,1,1]; biases may be NULL.
 * hidden % 8 == 0 is required by the 16-byte vector kernels (pad int(2.66*C) up on the host with zero weights). */
int rf_feedforward_gated(const float* project_in_w, const float* project_in_b, const float* dwA_w, const float* dwA_b,
                         const float* dwB_w, const float* dwB_b, const float* project_out_w, const float* project_out_b,
                         int C, int hidden, int dtype, const float* x, float* out, int B, int H, int W, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Callers either side of the forward
 * ---------------------------------------------------------------------------------------------- */
/* test.py:117-118: clamp(pred,0,1) -> *255 -> uint8 (truncation) -> HWC.  in [B,3,H,W] f32 -> out [B,H,W,3] u8. */
int rf_postprocess_u8(const float* in, unsigned char* out, int B, int H, int W, void* stream);
/* test.py:17-40 + 117-120 (prediction side): the conversion above, then correct_bayer_channels -- out channel k =
 * channel perm_host[k] ({0,1,2} RGGB, {2,1,0} BGGR, {1,0,2} GBRG, {0,2,1} GRBG) -- and, if auto_rb, auto_correct_rb: R and B
 * of an image are swapped when its red mean is below its blue mean (decided on exact integer sums).
 * workspace (auto_rb only): 16*B bytes, 8-byte aligned. */
int rf_postprocess_rgb_u8(const float* in, unsigned char* out, const int* perm_host, int auto_rb, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* stream);
/* test.py:111-113 (ground-truth side): the same two corrections on a uint8 [B,H,W,3] image, in place. */
int rf_correct_rgb_u8(unsigned char* img, const int* perm_host, int auto_rb, int B, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream);
/* test.py:123, skimage.metrics.peak_signal_noise_ratio on uint8 images: sse[b] = sum (a-b)^2 over the n_per_image bytes of
 * image b (exact, uint64, device); PSNR = 10*log10(255^2 * n_per_image / sse). */
int rf_sse_u8(const unsigned char* a, const unsigned char* b, unsigned long long* sse, int B, long long n_per_image,
              void* stream);
/* test.py:124, skimage.metrics.structural_similarity(a, b, channel_axis=-1) on uint8 [B,H,W,3] images (7x7 uniform window,
 * data_range 255, K1 0.01, K2 0.03, sample covariance, border of 3 cropped): sum_out[b] (device, double) = sum of the SSIM
 * map over the valid pixels and the 3 channels; SSIM = sum_out[b] / (3*(H-6)*(W-6)).  H, W >= 7. */
int rf_ssim_u8(const unsigned char* a, const unsigned char* b, double* sum_out, int B, int H, int W, void* stream);
/* WFB/load_dataset.py:88-89: clip(raw,black,white) -> (x-black)/(white-black+1e-6)*ratio (fp32 arithmetic, as numpy
 * evaluates it); clamp != 0 adds min(.,1) (correctdataloader.py:103).  raw [B,H,W] u16 -> out [B,1,H,W] f32; both
 * pointers 16-byte aligned. */
int rf_preprocess_u16(const unsigned short* raw, float* out, float black, float white, float ratio, int clamp, int B, int H,
                      int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TrueColor head / tail -- TrueColorRawFormer.py ("variant 0") and BayerTORGBColorMultiLvl.py
 * ("variant 1") put a learned colour front end and a tone-mapping tail around the same U-Net body.  fp32 NCHW.
 * ---------------------------------------------------------------------------------------------- */
/* nn.Conv2d(Cin, Cout, 3, padding=1) for a handful of channels (the demosaic_refine / chroma_extractor stacks of
 * EnhancedBayerProcessor, TrueColorRawFormer.py:90-103): out = act(conv(in * in_scale) + bias) + resid.
 * in [B,Cin,H,W], weight [Cout,Cin,3,3] (PyTorch layout), in_scale [Cin] / bias [Cout] / resid [B,Cout,H,W] may be NULL.
 * Cin <= 64, Cout in {2,3,4,16,32}; act: 0 none, 1 ReLU, 2 Softplus(beta 1, threshold 20), 3 tanh, 4 GELU(erf). */
int rf_conv3x3_small(const float* in, const float* in_scale, const float* weight, const float* bias, const float* resid,
                     float* out, int Cin, int Cout, int act, int B, int H, int W, void* stream);
/* The per-pixel part of EnhancedBayerProcessor.forward (TrueColorRawFormer.py:120-137; BayerTORGBColorMultiLvl.py:107-127):
 * planes [B,4,H,W] = (R,G1,G2,B) [times gains_host[4] if apply_gains] -> r, g = (G1+G2)/2, b -> rgb_linear = M rgb + bias
 * (color_matrix_host [3][4]) -> y = <rgb_linear, y_weights_host> / max(amax_hw, eps).  Writes rgb_linear [B,3,H,W],
 * chroma_in [B,4,H,W] = (r, g, b, y) (the chroma_extractor's input) and y [B,1,H,W].  ymax_ws: B floats of scratch. */
int rf_truecolor_mix(const float* planes, const float* gains_host, int apply_gains, const float* color_matrix_host,
                     const float* y_weights_host, float eps, float* rgb_linear, float* chroma_in, float* y, float* ymax_ws,
                     int B, int H, int W, void* stream);
/* CameraAwareColorCorrection.forward -- TrueColorRawFormer.py:170-185 (variant 0: gamma = the parameter; tone curve output
 * IS the pixel), BayerTORGBColorMultiLvl.py:164-181 (variant 1: gamma = softplus(gamma_param) + 1e-6, computed by the caller;
 * tone curve modulates by 0.8 + 0.4 * sigmoid).  x, out [B,3,H,W]; w1 [64,3], b1 [64], w2 [3,64], b2 [3] (color_transform),
 * w3 [32], b3 [32], w4 [32], b4 [1] (tone_curve), all device fp32. */
int rf_color_correction(const float* x, float gamma, int variant, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* w3, const float* b3, const float* w4, const float* b4, float* out, int B,
                        int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * WFB "WMB" block pieces -- RawFomer_WFB_FFAB/blocks.py:11-92 (FEB / ProcessBlock / FFAB: rFFT amplitude /
 * phase blocks) and RawFomer_WFB_FFAB/model.py:174-200 (Illumination_Estimator).  fp32 NCHW; the module mirrors in wfb.py
 * compose these calls exactly as the reference's forward does.  (WM = mamba_ssm.Mamba is third-party and not built.)
 * ---------------------------------------------------------------------------------------------- */
/* torch.fft.rfft2 / irfft2 (norm='ortho') of [BC,H,W] real planes, any H, W, as dense DFT matrix products in fp32.
 * plan: rf_dft2_plan_floats(H, W) floats, filled once per (H, W) by rf_dft2_plan_init (twiddles made in double precision
 * on the device).  spec / tmp: [BC][2][H][W/2+1] floats (real plane, imaginary plane); tmp is scratch of the same size. */
size_t rf_dft2_plan_floats(int H, int W);
int rf_dft2_plan_init(float* plan, int H, int W, void* stream);
int rf_rfft2_ortho(const float* x, const float* plan, float* spec, float* tmp, long long BC, int H, int W, void* stream);
/* ignores the imaginary parts of the DC / Nyquist bins like torch.fft.irfft2(s=(H, W)) (blocks.py:35) */
int rf_irfft2_ortho(const float* spec, const float* plan, float* out, float* tmp, long long BC, int H, int W, void* stream);
/* blocks.py:28-29: mag = |z| + 1e-6, pha = angle(z), both [BC][H*(W/2+1)]; the four self-conjugate bins take imag = +0 */
int rf_spec_abs_angle(const float* spec, float* mag, float* pha, long long BC, int H, int W, void* stream);
/* blocks.py:32-34: spec = (mag cos pha, mag sin pha) */
int rf_spec_polar(const float* mag, const float* pha, float* spec, long long BC, int H, int W, void* stream);
/* nn.Conv2d(Cin + Cin2, Cout, 1) on [B,C,P] planes (weight [Cout][Cin + Cin2]; in2 = the second half of a torch.cat, may
 * be NULL with Cin2 = 0): out = clamp(act(W clamp(in, +-in_clamp) + bias), out_lo, out_hi) + resid.  in_clamp <= 0: no
 * input clamp; act: 0 none, 1 LeakyReLU(0.1); out_lo >= out_hi: no output clamp; bias / resid may be NULL. */
int rf_conv1x1_nchw(const float* in, const float* in2, const float* weight, const float* bias, const float* resid, float* out,
                    int Cin, int Cin2, int Cout, float in_clamp, int act, float out_lo, float out_hi, int B, long long P,
                    void* stream);
/* blocks.py:24,36-37: out = clamp(a + clamp(b, -lim, lim), -lim, lim), n elements */
int rf_add_clamp(const float* a, const float* b, float* out, float lim, long long n, void* stream);
/* model.py:192: mean over the channels, in [B,C,P] -> out [B,1,P] */
int rf_channel_mean(const float* in, float* out, int B, int C, long long P, void* stream);
/* model.py:181-182: nn.Conv2d(C, C, 5, padding=2, groups=C), weight [C,1,5,5] */
int rf_dwconv5x5_nchw(const float* in, const float* weight, const float* bias, float* out, int B, int C, int H, int W,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole model — RawFormer.forward, FLCA_RF.py:330-370 (variant ML: ML_RF.py:356-416)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rf_model_weights {
  const float* embedding_w;        /* embedding.weight [d,4,3,3] */
  const float* embedding_b;
  rf_block_weights blocks[7];      /* conv_tran1..7 */
  const float* down_w[3];          /* down{n}.body.0.weight (ML: down{n}.0.weight) [C/2,C,3,3] */
  const float* up_w[3];            /* up{n}.weight [2C,C,2,2] (ConvTranspose2d layout: in,out,kh,kw) */
  const float* up_b[3];
  const float* reduce_w[3];        /* channel_reduce{n}.weight [C,2C,1,1] */
  const float* reduce_b[3];
  const float* conv_out_w;         /* conv_out.weight [12,d,3,3] */
  const float* conv_out_b;
  float rgb_w_host[3];             /* luma_chroma.{r_w,g_w,b_w} (host values) */
} rf_model_weights;

/* Bytes of the packed (kernel-layout) parameter blob for a model of base width `dim`. */
size_t rf_model_packed_bytes(int dim, int dtype, int variant);
/* Re-packs PyTorch-layout weights into the kernel layout (device-side, stream-ordered). */
int rf_model_pack(const rf_model_weights* w, int dim, int dtype, int variant, void* packed, size_t packed_bytes,
                  void* stream);
/* Workspace needed by rf_rawformer_forward for raw frames [B,1,H,W]. */
size_t rf_rawformer_workspace_bytes(int dim, int dtype, int variant, int B, int H, int W);
/* raw [B,1,H,W] float32 -> out [B,3,H,W] float32.  H, W multiples of 16; dim % 8 == 0. */
int rf_rawformer_forward(const void* packed, int dim, int dtype, int variant, const float* raw, float* out, int B,
                         int H, int W, void* workspace, size_t workspace_bytes, void* stream);
/* Same forward with a cudaEvent pair around every launch; synchronises the stream.  Fills up to `cap` entries of
 * kernel_ms_host / kernel_id_host (RF_K_* ids below) and returns the launch count in *n_host. */
int rf_rawformer_forward_profiled(const void* packed, int dim, int dtype, int variant, const float* raw, float* out,
                                  int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream,
                                  float* kernel_ms_host, int* kernel_id_host, int cap, int* n_host);
/* ------------------------------------------------------------------------------------------------
 * Row-tiled single frame — RawFormer.forward (FLCA_RF.py:330-370) of ONE frame split into bands of whole rows over
 * the GPUs of a box (BASELINE config 4).  One process per GPU; rank r owns raw rows [row0, row0 + rows) (multiples of
 * 16, at least 64).  Per Conv_Transformer the ranks exchange RF_BAND_HALO rows of the block input with their band
 * neighbours and all-reduce, in one exchange, the three per-image reductions of the block (squeeze-excite channel sums
 * FLCA_RF.py:160, |q|^2,|k|^2 and the per-head Gram FLCA_RF.py:228-230) through peer-mapped memory: every rank owns a "comm region"
 * (flags + mailboxes) that its peers write with plain stores over NVLink; no NCCL call is on the data path.
 * RF_BF16 / RF_VARIANT_FLCA only.
 * ---------------------------------------------------------------------------------------------- */
#define RF_BAND_MAX_RANKS 8
#define RF_BAND_HALO 4           /* packed rows at every U-Net stage */
#define RF_IPC_HANDLE_BYTES 64   /* sizeof(cudaIpcMemHandle_t) */

typedef struct rf_band {
  int rank, nranks;
  int row0, rows;                       /* interior raw rows of this rank */
  void* comm[RF_BAND_MAX_RANKS];        /* comm region of every rank as mapped in THIS process (comm[rank] = own) */
  unsigned epoch;                       /* non-zero: a real forward (the frame counter itself lives in the comm region
                                           and is advanced on the device, so a forward can be replayed as a CUDA
                                           graph); every rank must run the same number of real forwards.
                                           0 = rehearsal: same launches, but no rank signals or waits (run it once
                                           per rank before the first frame so that every kernel is loaded) */
} rf_band;

/* Bytes of one rank's comm region (identical on all ranks). */
size_t rf_band_comm_bytes(int dim, int dtype, int variant, int H, int W, int nranks);
/* cudaMalloc + zero a comm region on the current device and export its IPC handle (handle_host may be NULL). */
int rf_band_comm_alloc(size_t bytes, void** ptr_host, unsigned char* handle_host);
/* Map a peer's comm region into this process (cudaIpcOpenMemHandle); rf_band_comm_close undoes it. */
int rf_band_comm_open(const unsigned char* handle_host, void** ptr_host);
int rf_band_comm_close(void* ptr);
int rf_band_comm_free(void* ptr);
/* Sticky error word of the local comm region: 0, or 1 + the index of the sync point whose wait timed out
 * (a peer never arrived).  Synchronises `stream`. */
int rf_band_comm_status(const void* comm_own, int* err_host, void* stream);
/* Clear the header of the LOCAL comm region (error word, frame counter, arrival counters) after a failed frame.  Every
 * rank must call it between two barriers of the group (no forward in flight anywhere), then restart its epoch count.
 * Synchronises `stream`. */
int rf_band_comm_reset(void* comm_own, void* stream);
/* Rows of the band image the forward writes: out is [1,3,2*(ht+rows/2+hb),W] where ht/hb = RF_BAND_HALO for a
 * neighbour above/below, else 0; the interior starts at raw row 2*ht of it. */
int rf_band_out_rows(const rf_band* band, int* out_rows_host, int* interior_row0_host);
size_t rf_rawformer_band_workspace_bytes(int dim, int dtype, int variant, int H, int W, const rf_band* band);
/* raw [1,1,H,W] float32 = the WHOLE frame (replicated on every rank: the 1-channel guidance is computed locally);
 * out = this rank's band image of the result (see rf_band_out_rows).  dtype RF_BF16 only (RF_ERR_UNSUPPORTED otherwise);
 * variant RF_VARIANT_FLCA (FLCA_RF.py) or RF_VARIANT_ML (ML_RF.py: its colour anchor adds one sync point). */
int rf_rawformer_forward_band(const void* packed, int dim, int dtype, int variant, const float* raw, float* out, int H,
                              int W, const rf_band* band, void* workspace, size_t workspace_bytes, void* stream);
/* Same with a cudaEvent pair around every launch (see rf_rawformer_forward_profiled); the times of the band_halo /
 * band_allreduce launches include the wait for the peers.  Every rank must make the same call. */
int rf_rawformer_forward_band_profiled(const void* packed, int dim, int dtype, int variant, const float* raw, float* out,
                                       int H, int W, const rf_band* band, void* workspace, size_t workspace_bytes,
                                       void* stream, float* kernel_ms_host, int* kernel_id_host, int cap, int* n_host);

/* Human-readable name of a kernel id reported by rf_rawformer_forward_profiled. */
const char* rf_kernel_name(int kernel_id);
/* Algorithmic (compulsory) bytes and FLOPs the launch `index` of the last profiled forward moved/did. */
int rf_profiled_launch_info(int index, double* bytes_host, double* flops_host);

#ifdef __cplusplus
}
#endif
#endif /* RAWFORMER_B200_H */
