// Internal launcher interface between the kernel translation units and the orchestration (rf_model.cu).
// All tensors are device pointers; activations are NHWC of the context's dtype unless stated otherwise.
#pragma once
#include "rf_common.cuh"

namespace rf {

// ---- kernel-layout parameters of one Conv_Transformer (pointers into the packed blob) -----------------
struct PackedBlock {
  int C = 0, hid = 0;
  // FLCA (variant FLCA): [9 taps][4 maps: LL, |high|, cr, cb][C]; (variant ML): [9][6: LL1,hi1,LL2,hi2,cr,cb][C]
  float* flca_w = nullptr;
  float* abg = nullptr;        // alpha, beta, gamma (FLCA only)
  float* se_w1 = nullptr;      // [hid][C]
  float* se_b1 = nullptr;
  float* se_w2 = nullptr;      // [C][hid]
  float* se_b2 = nullptr;
  float* ln1_g = nullptr; float* ln1_b = nullptr;
  void* qkv_w = nullptr;       // T [3C][C]
  float* qkv_b = nullptr;
  void* qkv_wf = nullptr;      // T [3C][C]: qkv_w * diag(ln1_g)   (LayerNorm folded, bf16 mode)
  float* qkv_cs = nullptr;     // [3C] row sums of qkv_wf
  float* qkv_bf = nullptr;     // [3C] qkv_b + qkv_w * ln1_b
  float* qkv_dw_w = nullptr;   // [9][3C]
  float* qkv_dw_b = nullptr;
  float* temperature = nullptr;  // [8]
  float* proj_w = nullptr;     // fp32 master [C][C]
  float* proj_b = nullptr;
  float* ln2_g = nullptr; float* ln2_b = nullptr;
  void* pw1_w = nullptr;       // T [2C][C]
  float* pw1_b = nullptr;
  void* pw1_wf = nullptr;      // T [2C][C]: pw1_w * diag(ln2_g)
  float* pw1_cs = nullptr;
  float* pw1_bf = nullptr;
  float* ffn_dw_w = nullptr;   // [9][2C]
  float* ffn_dw_b = nullptr;
  void* pw2_w = nullptr;       // T [C][2C]
  float* pw2_b = nullptr;
  float* red_w = nullptr;      // fp32 master [C][2C]
  float* red_b = nullptr;
  void* convout_w = nullptr;   // T [C][9][C]  (k = tap*C + ci)
  float* convout_b = nullptr;
  // norm -> 1x1 -> depthwise 3x3 as one dense 3x3 conv (rf_lnconv.cu; bf16 mode, C <= 64 only, else null)
  void* ffn_cw = nullptr;      // T [2C][9][C] = ffn_dw[n][tap] * pw1[n][c] * ln2_g[c]
  float* ffn_bt = nullptr;     // [9][2C] bias by border state
  void* qkv_cw = nullptr;      // T [3C][9][C]
  float* qkv_bt = nullptr;     // [9][3C]
  float* cat_p2 = nullptr;     // fp32 [C][9][2C] = Conv_out o channel_reduce (C = 32 only)
  float* cat_bt = nullptr;     // [9][C]
  // ML extras
  float* gate_w = nullptr;     // [2 levels][2][2]
  float* gate_b = nullptr;     // [2][2]
  float* cgate = nullptr;      // [2] = weight, bias
  void* res_w0 = nullptr;      // T [C][C]
  float* res_b0 = nullptr;
  void* res_w2 = nullptr;      // T [C][C]
  float* res_b2 = nullptr;
};

struct PackedModel {
  int dim = 0;
  float* rgb_w = nullptr;      // [3]
  float* embed_w = nullptr;    // [9][4][d]
  float* embed_b = nullptr;
  PackedBlock blocks[7];
  void* down_w[3] = {};        // T [C/2][9][C]
  void* up_w[3] = {};          // T [4*Co][Ci], row = (2i+j)*Co + co
  float* up_b[3] = {};         // [4*Co]
  void* red_w[3] = {};         // T [C][2C]
  float* red_b[3] = {};
  float* head_w = nullptr;     // [9][d][12]
  float* head_b = nullptr;     // [12]
  void* head_wt = nullptr;     // T [16][9][d] (rows 12..15 zero): implicit-GEMM form of conv_out for the tensor cores
  float* head_b16 = nullptr;   // [16] (12 used)
  float* haar = nullptr;       // [16] analysis filter (a,b,c,d taps of LL,LH,HL,HH)
};

// ---- generic "pixel-row" GEMM / implicit conv (rf_gemm.cu) ----------------------------------------------
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_TANH_RES = 3 /* Y = R + 0.2*tanh(acc+bias) */ };
enum { AMODE_ROWS = 0, AMODE_CONV3 = 1 };
enum { OMODE_ROWS = 0, OMODE_CONVT = 1, OMODE_UNSHUFFLE = 2, OMODE_ATOMIC_F32 = 3 /* split-K: atomicAdd into fp32 Y */,
       OMODE_HEAD = 4 /* N = 16 (12 used): LeakyReLU(0.2) + PixelShuffle(2) into fp32 NCHW [B,3,2H,2W] (FLCA_RF.py:368-369) */ };

struct GemmP {
  const void* A1 = nullptr; const void* A2 = nullptr;  // [B][M][K1], [B][M][K2] (A2 optional: concatenated K)
  const void* Wt = nullptr;                             // [N][K1+K2] (K contiguous); per image if w_img != 0
  const float* bias = nullptr;                          // [N]
  const void* R = nullptr;                              // residual [B][M][N] (OMODE_ROWS only)
  void* Y = nullptr;
  i64 lda1 = 0, lda2 = 0, ldr = 0, ldy = 0;             // row pitches in elements
  i64 w_img = 0;                                        // elements between per-image weights (0 = shared)
  i64 a1_img = 0;                                       // elements between images of A1 (0 = lda1*M)
  i64 ldw = 0;                                          // row pitch of Wt in elements (0 = K1+K2)
  int ksplit = 1;                                       // split-K factor (OMODE_ATOMIC_F32 only)
  int M = 0, N = 0, K1 = 0, K2 = 0, B = 1;
  int act = ACT_NONE, amode = AMODE_ROWS, omode = OMODE_ROWS;
  int H = 0, W = 0;                                     // image size of the rows (m = y*W + x) for CONV3/CONVT/UNSHUFFLE
  int kernel_id = RF_K_MISC;
  // LayerNorm folded into the contraction (bf16 mode, FLCA_RF.py:183-187 + the 1x1 conv that follows): Wt / bias are the
  // folded weights W*diag(g) and b + W*beta, ln_cs[n] = sum_k Wt[n][k], and
  //   Y = rstd[m] * (acc - mean[m] * ln_cs[n]) + bias[n],  mean/rstd from ln_stats[row][ln_npart] = partial (sum, sumsq)
  const float* ln_stats = nullptr;
  const float* ln_cs = nullptr;
  int ln_npart = 0, ln_C = 0;
  float ln_eps = 0.f;
  // optional (OMODE_ROWS): partial (sum, sum of squares) of the stored (rounded) output rows -> stats_out[row][npart]
  // (float2 each); launch_gemm returns npart (0 when not requested)
  float* stats_out = nullptr;
};
int launch_gemm(Ctx& ctx, const GemmP& p);
// (sum, sum of squares) over the C channels of every pixel row -> stats[row] (float2); the LayerNorm statistics of the
// folded form above when no producer epilogue emitted them
void launch_row_stats(Ctx& ctx, const void* x, float* stats, i64 rows, int C);
// fold LayerNorm(gamma, beta) into the 1x1 conv W [N][K] (fp32, PyTorch layout): Wf = T(W*gamma), cs = rowsum(Wf),
// bf = bias + W*beta
void launch_fold_ln(Ctx& ctx, const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* cs,
                    float* bf, int N, int K);

// ---- layout / index kernels (rf_index_ops.cu) --------------------------------------------------------------
void launch_nchw_to_nhwc(Ctx& ctx, const float* in, void* out, int B, int C, i64 HW);   // fp32 NCHW -> T NHWC
void launch_nhwc_to_nchw(Ctx& ctx, const void* in, float* out, int B, int C, i64 HW);   // T NHWC -> fp32 NCHW
void launch_fill_f32(Ctx& ctx, float* p, float v, i64 n);
// n zero-initialised floats: from the pre-cleared region of the workspace when there is room (no launch), else an arena
// block cleared by a fill kernel
float* zeroed_f32(Ctx& ctx, size_t n);
// generic strided 3-d copy with conversion: dst[doff + a*da + b*db + c*dc] = src[a*sa + b*sb + c*sc]
void launch_pack3(Ctx& ctx, const float* src, void* dst, int dst_dtype, int A, int Bn, int Cn, i64 sa, i64 sb, i64 sc,
                  i64 da, i64 db, i64 dc, i64 doff);

// ---- guidance (rf_guidance.cu) ---------------------------------------------------------------------------------
// raw [B,1,H,W] -> x_ds [B,h,w,4] fp32, y_raw [B,h,w], ymax[B] (must be pre-filled with -inf)
void launch_pack_luma(Ctx& ctx, const float* raw, float* x_ds, float* y_raw, float* ymax, const float* rgb_w, int B,
                      int H, int W);
// y = y_raw / max(ymax, eps); cr = r - y; cb = b - y     (planar [B,h,w] each)
void launch_luma_finalize(Ctx& ctx, const float* x_ds, const float* y_raw, const float* ymax, float eps, float* y,
                          float* cr, float* cb, int B, int h, int w);
// Haar analysis of a 1-channel map + high-band magnitude: y [B,Hy,Wy] -> LL, yh [B,ceil(Hy/2),ceil(Wy/2)]
void launch_dwt_high(Ctx& ctx, const float* y, const float* filt16, float* LL, float* yh, int B, int Hy, int Wy);
// bilinear-resized guidance at one stage: G [B,Hf,Wf,NG]; NG = 4 (LL,yh,cr,cb) or 8 (LL1,yh1,LL2,yh2,cr,cb,chr_mag,0)
// means (optional, ML): [B][8] sums of the NG maps over the stage (divide by Hf*Wf on use); must be zeroed.
void launch_guidance_stage(Ctx& ctx, const float* LL1, const float* yh1, int H1, int W1, const float* LL2,
                           const float* yh2, int H2, int W2, const float* cr, const float* cb, int Hy, int Wy, float* G,
                           int NG, float* sums, int B, int Hf, int Wf, void* G16a = nullptr, void* G16b = nullptr,
                           int y_begin = 0, int y_rows = -1 /* rows of G to produce; default: all */);

// ---- FLCA (rf_flca.cu) ---------------------------------------------------------------------------------------------
int flca_num_partials(int C, int B, i64 P);
// xmod = feat * (1 + a*sig(conv(LL)) + b*tanh(conv(yh)) + g*sig(conv(cr,cb))); partial [B][nblk][C] channel sums
// G16 (optional): the bf16 [hi|lo] form of G for the tensor-core path (see launch_split_bf16x8)
// returns the number of partial slots per image the kernel used (stride of `partial` = that number; <= nblk)
int launch_flca_mod(Ctx& ctx, const void* feat, const float* G, const void* G16, const float* w36, const float* abg,
                    void* xmod, float* partial, int nblk, int B, int Hf, int Wf, int C);
// tensor-core im2col forms (rf_im2col_tc.cu, bf16 only); false when the shape is not supported
bool im2col_tc_supported(const Ctx& ctx, int C);
// G16 / x16: the 4 fp32 maps of every pixel as [hi x4 | lo x4] bf16 (16 bytes per pixel), made by launch_split_bf16x8
// (g4 points at the first of 4 consecutive floats of pixel 0; pixels are stride_floats apart: 4, or 8 for the ML guidance)
void launch_split_bf16x8(Ctx& ctx, const float* g4, void* out16, i64 npix, int stride_floats = 4);
bool launch_pyr_spatial_tc(Ctx& ctx, const void* x, const void* G16, const float* w54, const float* gates, void* xs, int mode,
                           int level, int B, int Hf, int Wf, int C);
bool launch_flca_mod_tc(Ctx& ctx, const void* feat, const void* G16, const float* w36, const float* abg, void* xmod,
                        float* partial, int B, int Hf, int Wf, int C, int* used_slots = nullptr);
bool launch_embed_tc(Ctx& ctx, const void* x16, const float* w, const float* b, void* out, int B, int h, int w_, int d);
// ML: xs = x * (ga * sig(conv(mapA)) + gb * tanh(conv(mapB)))  [mode 0, level l] or xs = x * gc*sig(conv(cr,cb)) [mode 1]
void launch_pyr_spatial(Ctx& ctx, const void* x, const float* G8, const float* w54, const float* gates, void* xs, int mode,
                        int level, int B, int Hf, int Wf, int C);
// gates[b][0..5] = (alpha_0, beta_0, alpha_1, beta_1, gamma, 0) from the stage sums
void launch_pyr_gates(Ctx& ctx, const float* sums, i64 P, const float* gate_w, const float* gate_b, const float* cgate,
                      float* gates, int B);
// per-channel sums of an NHWC tensor -> partial [B][nblk][C]
void launch_channel_sums(Ctx& ctx, const void* x, float* partial, int nblk, int B, i64 P, int C);
// s = sigmoid(W2 relu(W1 mean + b1) + b2)  -> scale [B][C]
void launch_se_finalize(Ctx& ctx, const float* partial, int nblk, i64 P, const float* w1, const float* b1,
                        const float* w2, const float* b2, float* scale, int B, int C, int hid);
// the squeeze-excite MLP and the fold wred[b][n][k] = red_w[n][k] * (k < C ? scale[b][k] : 1) (T output) in one launch
// (the scale is also written to scale [B][C])
void launch_se_fold(Ctx& ctx, const float* partial, int nblk, i64 P, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* scale, const float* red_w, void* wred, int B, int C, int hid);
// out = x * scale[b][c]
void launch_scale_channels(Ctx& ctx, const void* x, const float* scale, void* out, int B, i64 P, int C);

// ---- transformer branch (rf_attn.cu) -----------------------------------------------------------------------------
void launch_layernorm(Ctx& ctx, const void* x, const float* g, const float* b, void* out, float eps, int mode, i64 rows,
                      int C);
// depthwise 3x3 on qkv_pre [B,H,W,3C]; writes v [B,H,W,C]; accumulates stats[b] = {gram [8][c][c], qn2 [C], kn2 [C]}
void launch_dwqkv_gram(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* v, float* stats, int B,
                       int H, int W, int C);
inline i64 attn_stats_floats(int C) { return (i64)C * C + 2 * C; }
// variant for the tensor-core Gram: dw(qkv_pre) split into qk [.,2C] (q | k) and v [.,C], both dense NHWC;
// sumsq[b][2C] = squared norms of q,k (must be zeroed)
// sq_part (optional): [B][num_sms][2C] pre-zeroed per-CTA partial squared norms; when the kernel used them (return value =
// number of slots per image > 0) sumsq is left untouched and the caller sums the slots in order (launch_attn_reduce)
int launch_dwqkv_nhwc(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq,
                      int B, int H, int W, int C, float* sq_part = nullptr);
// TMA-staged bf16 depthwise kernels (rf_dw_tma.cu); false when the shape is not supported
bool launch_dwconv_tma(Ctx& ctx, const void* in, const float* dw_w, const float* dw_b, void* out, int gelu, int B, int H, int W,
                       int Cn);
bool launch_dwqkv_tma(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq,
                      int B, int H, int W, int C, float* sq_part = nullptr, int* nslots = nullptr);
bool launch_dwconv_tma_sub(Ctx& ctx, const void* in, int in_pitch, const float* dw_w, int wpitch, const float* dw_b, void* out,
                           int B, int H, int W, int Cn);
// q|k depthwise + Gram + squared norms of ONE image in one kernel (rf_qk_gram.cu): q|k never reach HBM.  Writes per-CTA
// partial slots gram_part [slot][C][C/8] and sq_part [slot][2C] (slot < the returned count <= slot_cap); 0 = unsupported.
bool qk_gram_supported(const Ctx& ctx, int C);
int launch_dwqk_gram(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, float* gram_part, float* sq_part,
                     int H, int W, int C, int slot_cap);
// Gram of one image (bf16 NHWC qk [P][2C]) as per-split partials: part[z][C][C/8] = the per-head diagonal blocks of
// q^T k over pixel slice z (plain stores, z < returned split count <= gram_max_splits()); 0 if the tcgen05 path is unavailable.
// The caller sums the slices in order (launch_attn_reduce) -> bit-reproducible, unlike atomic accumulation.
int launch_gram_tcgen05(Ctx& ctx, const void* qk, float* part, int C, i64 P);
int gram_max_splits();
// stats[C*C (diagonal blocks)] = sum_z gram_part[z]; norms[2C] = sum_slot sq_part[slot]  (fixed order); one image
void launch_attn_reduce(Ctx& ctx, const float* gram_part, int nsplit, const float* sq_part, int nslots, float* stats, float* norms,
                        int C);
bool tcgen05_enabled();
// Mw[b] = proj_w * blockdiag(softmax(gram / (|q||k|) * temperature))   (T [B][C][C])
// norms (optional): [B][2C] squared norms of q,k kept outside `stats` (else they are read from stats[b][C*C ..])
void launch_attn_finalize(Ctx& ctx, const float* stats, const float* temperature, const float* proj_w, void* Mw, int B,
                          int C, const float* norms = nullptr);
// depthwise 3x3 (+bias) with optional exact GELU: in/out [B,H,W,Cn]
void launch_dwconv(Ctx& ctx, const void* in, const float* dw_w, const float* dw_b, void* out, int gelu, int B, int H,
                   int W, int Cn, int kernel_id);
// fused conv_ffn (rf_ffn_fused.cu): out = x + pointwise2(gelu(depthwise(pointwise1(norm2(x))))) in ONE kernel, norm2 folded
// into W1f / cs / b1; stats = [rows][npart] (sum, sumsq) of x's rows.  false when the shape is not supported.
bool ffn_fused_supported(const Ctx& ctx, int C, int W);
bool launch_ffn_fused(Ctx& ctx, const void* x, const void* W1f, const float* cs, const float* b1, const float* stats, int npart,
                      const float* dw_w, const float* dw_b, const void* W2, const float* b2, void* out, int B, int H, int W,
                      int C);
// norm -> 1x1 conv -> depthwise 3x3 as ONE dense 3x3 convolution on the tensor cores (rf_lnconv.cu, bf16, C = 32):
// pack-time weights Weff [N][9][K] and the border-state bias table [9][N] from the PyTorch-layout fp32 parameters
void launch_pack_lnconv(Ctx& ctx, const float* W, const float* gamma, const float* beta, const float* bias, const float* dw,
                        const float* dwb, void* cw, float* btab, int N, int K);
bool lnconv_supported(const Ctx& ctx, int C, int H, int W);
// out = x + conv_ffn(norm2(x)); stats = [rows] (sum, sumsq) of x's rows
bool launch_lnconv_ffn(Ctx& ctx, const void* x, const float* stats, const void* cw, const float* btab, const void* W2,
                       const float* b2, void* out, int B, int H, int W, int C);
// ONE image: v [H,W,C] + per-CTA partial slots of the Gram / squared norms of q, k (as launch_dwqk_gram); returns the slots
int launch_lnconv_qkv(Ctx& ctx, const void* x, const float* stats, const void* cw, const float* btab, void* v, float* gram_part,
                      float* sq_part, int H, int W, int C, int slot_cap);
// Conv_out (C -> C 3x3 + bias + LeakyReLU 0.2) through the same pipeline; w = T [C][9][C]
bool launch_lnconv_conv3(Ctx& ctx, const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int C);
// project_out fused in front of the FFN (C = 32, one image): out = x1 + conv_ffn(norm2(x1)), x1 = x + Mw v + proj_b
bool lnconv_proj_supported(const Ctx& ctx, int C, int H, int W);
bool launch_lnconv_ffn_proj(Ctx& ctx, const void* x, const void* v, const void* Mw, const float* proj_b, const void* cw,
                            const float* btab, const void* W2, const float* b2, void* out, int H, int W, int C);
// channel_reduce folded into Conv_out (C = 32): one dense 3x3 conv of the pair (xmod | x2) with per-image weights
void launch_pack_cat(Ctx& ctx, const float* wout, const float* bout, const float* wred, const float* bred, float* p2, float* bt,
                     int C);
void launch_cat_scale(Ctx& ctx, const float* p2, const float* scale, void* weff, int B, int C);
bool lnconv_cat_supported(const Ctx& ctx, int C, int H, int W);
bool launch_lnconv_cat(Ctx& ctx, const void* xmod, const void* x2, const void* weff, const float* btab, void* out, int H, int W,
                       int C);
// embedding 3x3 4->d from x_ds (fp32 [B,h,w,4]) ; head 3x3 d->12 + lrelu + pixel-shuffle to fp32 NCHW [B,3,2h,2w]
void launch_embed(Ctx& ctx, const float* x_ds, const void* x16, const float* w, const float* b, void* out, int B, int h,
                  int w_, int d);
void launch_head(Ctx& ctx, const void* in, const float* w, const float* b, float* out, int B, int h, int w_, int d);
// ML tail: out += 0.12*(mean(up(in_rgb)) - mean(out)); out += 0.03*(up8(LL2) - Y(out))
// out[b * ostride + i] = sum over the slots s (in a fixed order) of part[(b * nslots + s) * width + i], i < width <= 8: the ordered
// second stage of the per-CTA partial sums of the multi-level variant's guidance means and colour-anchor sums
void launch_sum_slots(Ctx& ctx, const float* part, int nslots, int width, float* out, int ostride, int B);
// (out == nullptr: the input sums only)
void launch_tail_stats(Ctx& ctx, const float* out, const float* x_ds, float* sums, int B, int h, int w_);
// sums3[ch] += sum over rows [row0, row0 + rows) of out [3][rows_total][Wo]  (row-tiled forward: the band's interior rows)
void launch_out_sums(Ctx& ctx, const float* out, int rows_total, int row0, int rows, int Wo, float* sums3);
// y_off / rows (row-tiled forward): out holds rows [y_off, y_off + rows) of the 2h x 2w_ frame; rows < 0: the whole frame
void launch_tail_apply(Ctx& ctx, float* out, const float* sums, const float* LL2, int H2, int W2, int B, int h, int w_,
                       int y_off = 0, int rows = -1);

// ---- row-tiled forward: cross-GPU steps over peer-mapped comm regions (rf_band.cu); ctx.band must be set -------------
// fetch the BAND_HALO halo rows of the band image x [ht + rows_in + hb][W][C] from the band neighbours' interiors
void band_halo_exchange(Ctx& ctx, void* x, int W, int C);
// v[0..n) summed over the ranks, in place (one sync point)
void band_allreduce_small(Ctx& ctx, float* v, int n);
// advance the frame counter of the local comm region (first launch of a real forward)
void band_begin(Ctx& ctx);
// one all-reduce (sum over ranks) per Conv_Transformer: the attention statistics of C channels (diagonal Gram blocks +
// squared norms) and the squeeze-excite channel sums se [se_slots][C] (-> row 0 holds the frame's sums)
// norms: the [2C] squared norms of q,k (nullptr: inside stats at C*C)
void band_allreduce(Ctx& ctx, float* stats, int C, float* se, int se_slots, float* norms = nullptr);

}  // namespace rf
