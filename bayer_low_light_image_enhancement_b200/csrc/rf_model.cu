// Orchestration of the RawFormer forward (FLCA_RF.py:330-370, ML_RF.py:356-416) and of its sub-modules, plus the
// C-ABI entry points that take PyTorch-layout weights / NCHW fp32 tensors (include/rawformer_b200.h).
//
// Everything here is host code that enqueues kernels on the caller's stream and bump-allocates from the caller's
// workspace.  The same code runs in "dry" mode to answer the *_workspace_bytes queries.
#include <string.h>

#include <stdlib.h>

#include "rf_kernels.cuh"

namespace rf {

int launch_gemm_tcgen05(Ctx& ctx, const GemmP& p);
int profile_begin(cudaStream_t stream, int cap);
int profile_end(float* ms_host, int* ids_host, int cap_out, int* n_host);

// ---------------------------------------------------------------------------------------------
// packed-parameter layout (a pure function of dim/dtype/variant, so pack and forward agree)
// ---------------------------------------------------------------------------------------------
struct Layout {
  char* base;
  size_t off = 0;
  explicit Layout(void* b) : base((char*)b) {}
  void* take(size_t bytes) {
    off = align_up(off, 256);
    char* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f32(size_t n) { return (float*)take(n * 4); }
  void* elems(size_t n, int dtype) { return take(n * esize(dtype)); }
};

static PackedBlock layout_block(Layout& L, int C, int dtype, int variant) {
  PackedBlock pb;
  pb.C = C;
  pb.hid = C / 8 > 8 ? C / 8 : 8;
  const size_t c = C;
  pb.flca_w = L.f32(9 * (variant == RF_VARIANT_ML ? 6 : 4) * c);
  pb.abg = L.f32(4);
  pb.se_w1 = L.f32(pb.hid * c); pb.se_b1 = L.f32(pb.hid);
  pb.se_w2 = L.f32(c * pb.hid); pb.se_b2 = L.f32(c);
  pb.ln1_g = L.f32(c); pb.ln1_b = L.f32(c);
  pb.qkv_w = L.elems(3 * c * c, dtype); pb.qkv_b = L.f32(3 * c);
  pb.qkv_wf = L.elems(3 * c * c, dtype); pb.qkv_cs = L.f32(3 * c); pb.qkv_bf = L.f32(3 * c);
  pb.qkv_dw_w = L.f32(27 * c); pb.qkv_dw_b = L.f32(3 * c);
  pb.temperature = L.f32(8);
  pb.proj_w = L.f32(c * c); pb.proj_b = L.f32(c);
  pb.ln2_g = L.f32(c); pb.ln2_b = L.f32(c);
  pb.pw1_w = L.elems(2 * c * c, dtype); pb.pw1_b = L.f32(2 * c);
  pb.pw1_wf = L.elems(2 * c * c, dtype); pb.pw1_cs = L.f32(2 * c); pb.pw1_bf = L.f32(2 * c);
  pb.ffn_dw_w = L.f32(18 * c); pb.ffn_dw_b = L.f32(2 * c);
  pb.pw2_w = L.elems(2 * c * c, dtype); pb.pw2_b = L.f32(c);
  pb.red_w = L.f32(2 * c * c); pb.red_b = L.f32(c);
  pb.convout_w = L.elems(9 * c * c, dtype); pb.convout_b = L.f32(c);
  if (dtype == RF_BF16 && C <= 64) {
    pb.ffn_cw = L.elems(2 * c * 9 * c, dtype); pb.ffn_bt = L.f32(9 * 2 * c);
    pb.qkv_cw = L.elems(3 * c * 9 * c, dtype); pb.qkv_bt = L.f32(9 * 3 * c);
    if (C == 32) { pb.cat_p2 = L.f32(c * 9 * 2 * c); pb.cat_bt = L.f32(9 * c); }
  }
  if (variant == RF_VARIANT_ML) {
    pb.gate_w = L.f32(8); pb.gate_b = L.f32(4); pb.cgate = L.f32(2);
    pb.res_w0 = L.elems(c * c, dtype); pb.res_b0 = L.f32(c);
    pb.res_w2 = L.elems(c * c, dtype); pb.res_b2 = L.f32(c);
  }
  return pb;
}

static const int kBlockStage[7] = {0, 1, 2, 3, 2, 1, 0};

static PackedModel layout_model(Layout& L, int dim, int dtype, int variant) {
  PackedModel pm;
  pm.dim = dim;
  const size_t d = dim;
  pm.rgb_w = L.f32(4);
  pm.haar = L.f32(16);
  pm.embed_w = L.f32(36 * d); pm.embed_b = L.f32(d);
  for (int i = 0; i < 7; ++i) pm.blocks[i] = layout_block(L, dim << kBlockStage[i], dtype, variant);
  for (int n = 0; n < 3; ++n) {
    const size_t C = d << n;           // down{n+1}: C -> C/2 (then unshuffle -> 2C)
    pm.down_w[n] = L.elems((C / 2) * 9 * C, dtype);
    const size_t Co = d << (2 - n);    // up{n+1}: 2Co -> Co
    pm.up_w[n] = L.elems(4 * Co * 2 * Co, dtype);
    pm.up_b[n] = L.f32(4 * Co);
    pm.red_w[n] = L.elems(Co * 2 * Co, dtype);
    pm.red_b[n] = L.f32(Co);
  }
  pm.head_w = L.f32(9 * d * 12); pm.head_b = L.f32(12);
  pm.head_wt = L.elems(16 * 9 * d, dtype); pm.head_b16 = L.f32(16);
  return pm;
}

// ---------------------------------------------------------------------------------------------
// packing from PyTorch layouts
// ---------------------------------------------------------------------------------------------
static void copy_f32(Ctx& ctx, const float* src, float* dst, i64 n) {
  if (src) launch_pack3(ctx, src, dst, RF_F32, 1, 1, (int)n, 0, 0, 1, 0, 0, 1, 0);
}
static void copy_T(Ctx& ctx, const float* src, void* dst, i64 n) {
  if (src) launch_pack3(ctx, src, dst, ctx.dtype, 1, 1, (int)n, 0, 0, 1, 0, 0, 1, 0);
}
// depthwise / 1-input-channel 3x3 [Cn,Cin,3,3] (input channel `ch`) -> dst[(t*NW + g)*Cn + c]
static void pack_taps(Ctx& ctx, const float* src, int Cin, int ch, float* dst, int NW, int g, int Cn) {
  if (src) launch_pack3(ctx, src + ch * 9, dst, RF_F32, 1, 9, Cn, 0, 1, (i64)Cin * 9, 0, (i64)NW * Cn, 1, (i64)g * Cn);
}
// dense 3x3 [N,Ci,3,3] -> T [N][9][Ci]
static void pack_conv3(Ctx& ctx, const float* src, void* dst, int N, int Ci) {
  if (src) launch_pack3(ctx, src, dst, ctx.dtype, N, 9, Ci, (i64)Ci * 9, 1, 9, (i64)9 * Ci, Ci, 1, 0);
}

static void pack_block(Ctx& ctx, const rf_block_weights& w, const PackedBlock& pb, int variant) {
  const int C = pb.C, hid = pb.hid;
  if (variant == RF_VARIANT_ML) {
    for (int l = 0; l < 2; ++l) {
      pack_taps(ctx, w.pyr_low_w[l], 1, 0, pb.flca_w, 6, 2 * l, C);
      pack_taps(ctx, w.pyr_high_w[l], 1, 0, pb.flca_w, 6, 2 * l + 1, C);
      copy_f32(ctx, w.pyr_gate_w[l], pb.gate_w + 4 * l, 4);
      copy_f32(ctx, w.pyr_gate_b[l], pb.gate_b + 2 * l, 2);
    }
    pack_taps(ctx, w.flca_chroma_w, 2, 0, pb.flca_w, 6, 4, C);
    pack_taps(ctx, w.flca_chroma_w, 2, 1, pb.flca_w, 6, 5, C);
    copy_f32(ctx, w.pyr_cgate_w, pb.cgate, 1);
    copy_f32(ctx, w.pyr_cgate_b, pb.cgate + 1, 1);
    copy_T(ctx, w.pyr_res_w0, pb.res_w0, (i64)C * C);
    copy_f32(ctx, w.pyr_res_b0, pb.res_b0, C);
    copy_T(ctx, w.pyr_res_w2, pb.res_w2, (i64)C * C);
    copy_f32(ctx, w.pyr_res_b2, pb.res_b2, C);
  } else {
    pack_taps(ctx, w.flca_low_w, 1, 0, pb.flca_w, 4, 0, C);
    pack_taps(ctx, w.flca_high_w, 1, 0, pb.flca_w, 4, 1, C);
    pack_taps(ctx, w.flca_chroma_w, 2, 0, pb.flca_w, 4, 2, C);
    pack_taps(ctx, w.flca_chroma_w, 2, 1, pb.flca_w, 4, 3, C);
    copy_f32(ctx, w.flca_alpha, pb.abg, 1);
    copy_f32(ctx, w.flca_beta, pb.abg + 1, 1);
    copy_f32(ctx, w.flca_gamma, pb.abg + 2, 1);
  }
  copy_f32(ctx, w.flca_se_w1, pb.se_w1, (i64)hid * C);
  copy_f32(ctx, w.flca_se_b1, pb.se_b1, hid);
  copy_f32(ctx, w.flca_se_w2, pb.se_w2, (i64)C * hid);
  copy_f32(ctx, w.flca_se_b2, pb.se_b2, C);
  copy_f32(ctx, w.norm1_w, pb.ln1_g, C);
  copy_f32(ctx, w.norm1_b, pb.ln1_b, C);
  copy_T(ctx, w.qkv_w, pb.qkv_w, (i64)3 * C * C);
  copy_f32(ctx, w.qkv_b, pb.qkv_b, 3 * C);
  launch_fold_ln(ctx, w.qkv_w, w.norm1_w, w.norm1_b, w.qkv_b, pb.qkv_wf, pb.qkv_cs, pb.qkv_bf, 3 * C, C);
  pack_taps(ctx, w.qkv_dw_w, 1, 0, pb.qkv_dw_w, 1, 0, 3 * C);
  copy_f32(ctx, w.qkv_dw_b, pb.qkv_dw_b, 3 * C);
  copy_f32(ctx, w.temperature, pb.temperature, 8);
  copy_f32(ctx, w.proj_w, pb.proj_w, (i64)C * C);
  copy_f32(ctx, w.proj_b, pb.proj_b, C);
  copy_f32(ctx, w.norm2_w, pb.ln2_g, C);
  copy_f32(ctx, w.norm2_b, pb.ln2_b, C);
  copy_T(ctx, w.pw1_w, pb.pw1_w, (i64)2 * C * C);
  copy_f32(ctx, w.pw1_b, pb.pw1_b, 2 * C);
  launch_fold_ln(ctx, w.pw1_w, w.norm2_w, w.norm2_b, w.pw1_b, pb.pw1_wf, pb.pw1_cs, pb.pw1_bf, 2 * C, C);
  pack_taps(ctx, w.ffn_dw_w, 1, 0, pb.ffn_dw_w, 1, 0, 2 * C);
  copy_f32(ctx, w.ffn_dw_b, pb.ffn_dw_b, 2 * C);
  copy_T(ctx, w.pw2_w, pb.pw2_w, (i64)2 * C * C);
  copy_f32(ctx, w.pw2_b, pb.pw2_b, C);
  copy_f32(ctx, w.reduce_w, pb.red_w, (i64)2 * C * C);
  copy_f32(ctx, w.reduce_b, pb.red_b, C);
  pack_conv3(ctx, w.convout_w, pb.convout_w, C, C);
  copy_f32(ctx, w.convout_b, pb.convout_b, C);
  if (pb.ffn_cw != nullptr) {
    launch_pack_lnconv(ctx, w.pw1_w, w.norm2_w, w.norm2_b, w.pw1_b, w.ffn_dw_w, w.ffn_dw_b, pb.ffn_cw, pb.ffn_bt, 2 * C, C);
    launch_pack_lnconv(ctx, w.qkv_w, w.norm1_w, w.norm1_b, w.qkv_b, w.qkv_dw_w, w.qkv_dw_b, pb.qkv_cw, pb.qkv_bt, 3 * C, C);
    if (pb.cat_p2 != nullptr) launch_pack_cat(ctx, w.convout_w, w.convout_b, w.reduce_w, w.reduce_b, pb.cat_p2, pb.cat_bt, C);
  }
}

// ---------------------------------------------------------------------------------------------
// building blocks on NHWC activations
// ---------------------------------------------------------------------------------------------
struct Stage {           // guidance of one U-Net stage
  int H = 0, W = 0;
  float* G = nullptr;    // [B,H,W,NG]
  void* G16 = nullptr;   // bf16 mode: [B,H,W] x 16 B = [hi x4 | lo x4] bf16 of G's maps 0..3 (tensor-core FLCA kernels)
  void* G16b = nullptr;  // ML variant: the same for maps 4..7 (cr, cb, mag, 0)
  float* sums = nullptr; // ML: [B][8] sums of the guidance maps
};

static GemmP gemm_rows(const void* A, int K, const void* Wt, const float* bias, void* Y, int N, int B, i64 M, int kid) {
  GemmP p;
  p.A1 = A; p.K1 = K; p.lda1 = K;
  p.Wt = Wt; p.bias = bias; p.Y = Y; p.ldy = N;
  p.M = (int)M; p.N = N; p.B = B; p.kernel_id = kid;
  return p;
}

// FLCA branch: returns xmod (un-scaled) and the SE scale [B][C]
// With `se_out`, the squeeze-excite MLP is left to the caller (conv_transformer fuses it with the channel_reduce fold):
// se_out = {partial sums [B][nblk][C], nblk}.
struct SePartial { float* partial = nullptr; int nblk = 0; };
static void flca_branch(Ctx& ctx, const PackedBlock& pb, int variant, const void* feat, const Stage& sg, int B, void** xmod_out,
                        float** scale_out, SePartial* se_out = nullptr) {
  const int C = pb.C, H = sg.H, W = sg.W;
  const i64 P = (i64)H * W;
  Arena& A = ctx.arena;
  int nblk = flca_num_partials(C, B, P);            // capacity; the modulation kernel reports the slots it used
  float* partial = zeroed_f32(ctx, (size_t)B * nblk * C);
  float* scale = A.get<float>((size_t)B * C);
  void* xmod = nullptr;
  if (variant == RF_VARIANT_FLCA) {
    xmod = A.elems((size_t)B * P * C, ctx.dtype);
    nblk = launch_flca_mod(ctx, feat, sg.G, sg.G16, pb.flca_w, pb.abg, xmod, partial, nblk, B, H, W, C);
  } else {
    float* gates = A.get<float>((size_t)B * 6);
    // (row-tiled forward: the guidance maps and their sums are the whole frame's, replicated on every rank)
    launch_pyr_gates(ctx, sg.sums, ctx.band != nullptr ? ctx.band->P_full : P, pb.gate_w, pb.gate_b, pb.cgate, gates, B);
    void* xa = A.elems((size_t)B * P * C, ctx.dtype);
    void* xb = A.elems((size_t)B * P * C, ctx.dtype);
    const size_t mk = A.mark();
    void* xs = A.elems((size_t)B * P * C, ctx.dtype);
    void* t1 = A.elems((size_t)B * P * C, ctx.dtype);
    const void* cur = feat;
    void* nxt = xa;
    for (int step = 0; step < 3; ++step) {
      const int smode = step < 2 ? 0 : 1, slevel = step < 2 ? step : 0;
      bool done = false;
      if (sg.G16 != nullptr && im2col_tc_supported(ctx, C)) {
        if (ctx.dry) {
          done = true;
        } else {
          const double px = (double)B * H * W;
          ScopedLaunch sl(RF_K_PYR_SPATIAL, px * C * 4.0 + px * 16.0, px * C * 2.0 * (smode == 0 ? 18 : 18));
          done = launch_pyr_spatial_tc(ctx, cur, smode == 0 ? sg.G16 : sg.G16b, pb.flca_w, gates, xs, smode, slevel, B, H, W, C);
          if (!done) recorder().last_cuda_error = (int)cudaErrorNotSupported;
        }
      }
      if (!done && sg.G != nullptr) launch_pyr_spatial(ctx, cur, sg.G, pb.flca_w, gates, xs, smode, slevel, B, H, W, C);
      GemmP g1 = gemm_rows(xs, C, pb.res_w0, pb.res_b0, t1, C, B, P, RF_K_GEMM_PYR_RES1);
      g1.act = ACT_RELU;
      launch_gemm(ctx, g1);
      GemmP g2 = gemm_rows(t1, C, pb.res_w2, pb.res_b2, nxt, C, B, P, RF_K_GEMM_PYR_RES2);
      g2.act = ACT_TANH_RES;
      g2.R = cur; g2.ldr = C;
      launch_gemm(ctx, g2);
      cur = nxt;
      nxt = (nxt == xa) ? xb : xa;
    }
    A.release(mk);
    xmod = const_cast<void*>(cur);  // == xa after three steps
    if (ctx.band != nullptr)        // row-tiled forward (B = 1): the band's interior rows only, summed over the ranks later
      launch_channel_sums(ctx, (const char*)xmod + (size_t)ctx.band->ht * W * C * esize(ctx.dtype), partial, nblk, B,
                          (i64)ctx.band->rows_in * W, C);
    else
      launch_channel_sums(ctx, xmod, partial, nblk, B, P, C);
  }
  if (ctx.band != nullptr) {
    // row-tiled forward: the channel sums of the whole frame arrive with the block's attention statistics (ONE exchange
    // per block, attention()); conv_transformer runs the squeeze-excite MLP after the transformer branch
    ctx.band->se_partial = partial;
    ctx.band->se_slots = nblk;
  }
  if (se_out != nullptr) {
    se_out->partial = partial;
    se_out->nblk = nblk;
  } else {
    launch_se_finalize(ctx, partial, nblk, P, pb.se_w1, pb.se_b1, pb.se_w2, pb.se_b2, scale, B, C, pb.hid);
  }
  *xmod_out = xmod;
  *scale_out = scale;
}

// LayerNorm statistics of the input rows when the norm is folded into the first 1x1 conv of a branch (bf16 mode)
struct LnFold {
  const float* stats = nullptr;   // [rows][npart] float2 (sum, sumsq)
  int npart = 0;
};

// out = (resid ? resid : 0) + Attention(xin).  xin is the normalised input, or -- with `ln` -- the raw block input
// whose LayerNorm (norm1) is folded into the qkv projection.  stats_out (optional): row statistics of `out`.
// keep (optional): project_out is left to the caller (fused in front of the FFN, rf_lnconv.cu): the v tensor and the per-image
// weights Mw stay allocated in the arena (the caller releases them) and no GEMM is launched
struct AttnKeep {
  const void* v = nullptr;
  void* Mw = nullptr;
};
static int attention(Ctx& ctx, const PackedBlock& pb, const void* xin, const void* resid, void* out, int B, int H, int W,
                     const LnFold* ln = nullptr, float* stats_out = nullptr, AttnKeep* keep = nullptr) {
  const int C = pb.C;
  const i64 P = (i64)H * W;
  Arena& A = ctx.arena;
  const size_t mk = A.mark();
  // norm1 -> qkv -> qkv_dwconv as one dense 3x3 conv on the tensor cores with the Gram / norms / v epilogue (rf_lnconv.cu)
  const bool conv_qkv = ln != nullptr && ln->npart == 1 && pb.qkv_cw != nullptr && lnconv_supported(ctx, C, H, W) &&
                        (ln->stats != nullptr || C == 32 || ctx.dry);
  void* qkv = conv_qkv ? nullptr : A.elems((size_t)B * P * 3 * C, ctx.dtype);
  if (!conv_qkv) {
    GemmP gq = gemm_rows(xin, C, ln ? pb.qkv_wf : pb.qkv_w, ln ? pb.qkv_bf : pb.qkv_b, qkv, 3 * C, B, P, RF_K_GEMM_QKV);
    if (ln) { gq.ln_stats = ln->stats; gq.ln_npart = ln->npart; gq.ln_cs = pb.qkv_cs; gq.ln_C = C; gq.ln_eps = 1e-5f; }
    launch_gemm(ctx, gq);
  }
  const i64 nst = attn_stats_floats(C);
  float* stats = zeroed_f32(ctx, (size_t)B * nst);
  const void* v = nullptr;
  const float* norms = nullptr;
  i64 ldv = C;
  if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {
    // depthwise pass writes q|k [P][2C] and v [P][C] (dense NHWC, coalesced) and reduces |q|^2,|k|^2; the Gram is a
    // split-K tcgen05 kernel with MN-major operands reading q,k straight from that tensor.
    void* vbuf = A.elems((size_t)B * P * C, ctx.dtype);
    // squared norms and Gram: per-CTA / per-split partial slots (plain stores) + ONE ordered reduction per image, so two
    // runs of the forward are bit-identical (no float atomics)
    float* sumsq = zeroed_f32(ctx, (size_t)B * 2 * C);      // (zeroed for the atomically accumulating fallback kernel)
    const int sq_cap = num_sms();
    const int gsplit_cap = gram_max_splits();
    float* gram_part = A.get<float>((size_t)gsplit_cap * C * (C / 8));
    // row-tiled forward: the Gram and the norms run over the band's interior rows only, then are summed over the ranks
    const i64 row0 = ctx.band != nullptr ? (i64)ctx.band->ht * W : 0;
    const i64 Pg = ctx.band != nullptr ? (i64)ctx.band->rows_in * W : P;
    if (conv_qkv) {
      float* sq_part = A.get<float>((size_t)sq_cap * 2 * C);
      if (!ctx.dry) {
        for (int b = 0; b < B; ++b) {
          const int ns = launch_lnconv_qkv(ctx, (const char*)xin + (size_t)b * P * C * 2,
                                           ln->stats ? ln->stats + (size_t)b * P * 2 : nullptr, pb.qkv_cw,
                                           pb.qkv_bt, (char*)vbuf + (size_t)b * P * C * 2, gram_part, sq_part, H, W, C, sq_cap);
          if (ns <= 0) recorder().last_cuda_error = (int)cudaErrorNotSupported;
          launch_attn_reduce(ctx, gram_part, ns, sq_part, ns, stats + b * nst, sumsq + (size_t)b * 2 * C, C);
        }
      }
    } else if (qk_gram_supported(ctx, C)) {
      // q|k depthwise + Gram + norms in one kernel per image (q|k never reach HBM); v by the plain depthwise kernel
      float* sq_part = A.get<float>((size_t)sq_cap * 2 * C);
      if (!ctx.dry) {
        for (int b = 0; b < B; ++b) {
          const int ns = launch_dwqk_gram(ctx, (const char*)qkv + (size_t)b * P * 3 * C * 2, pb.qkv_dw_w, pb.qkv_dw_b, gram_part,
                                          sq_part, H, W, C, sq_cap);
          if (ns <= 0) recorder().last_cuda_error = (int)cudaErrorNotSupported;
          launch_attn_reduce(ctx, gram_part, ns, sq_part, ns, stats + b * nst, sumsq + (size_t)b * 2 * C, C);
        }
        const double px = (double)B * P;
        ScopedLaunch sl(RF_K_DW_QKV_GRAM, px * C * 2.0 * 2.0, 18.0 * px * C);
        if (!launch_dwconv_tma_sub(ctx, (const char*)qkv + (size_t)2 * C * 2, 3 * C, pb.qkv_dw_w + 2 * C, 3 * C,
                                   pb.qkv_dw_b + 2 * C, vbuf, B, H, W, C))
          recorder().last_cuda_error = (int)cudaErrorNotSupported;
      }
    } else {
      void* qk = A.elems((size_t)B * P * 2 * C, ctx.dtype);
      float* sq_part = zeroed_f32(ctx, (size_t)B * sq_cap * 2 * C);
      const int nslots = launch_dwqkv_nhwc(ctx, qkv, pb.qkv_dw_w, pb.qkv_dw_b, qk, vbuf, sumsq, B, H, W, C, sq_part);
      if (!ctx.dry) {
        for (int b = 0; b < B; ++b) {
          const int nsplit = launch_gram_tcgen05(ctx, (const char*)qk + ((size_t)b * P + row0) * 2 * C * 2, gram_part, C, Pg);
          if (nsplit <= 0 || nsplit > gsplit_cap) recorder().last_cuda_error = (int)cudaErrorNotSupported;
          // (the depthwise kernel lays the slots out as [b][slot < nslots][2C])
          launch_attn_reduce(ctx, gram_part, nsplit, sq_part + (size_t)b * nslots * 2 * C, nslots, stats + b * nst,
                             sumsq + (size_t)b * 2 * C, C);
        }
      }
    }
    if (ctx.band != nullptr) band_allreduce(ctx, stats, C, ctx.band->se_partial, ctx.band->se_slots, sumsq);
    norms = sumsq;     // the squared norms stay where the depthwise kernel left them
    v = vbuf;
  } else {
    if (ctx.band != nullptr) recorder().last_cuda_error = (int)cudaErrorNotSupported;
    void* vbuf = A.elems((size_t)B * P * C, ctx.dtype);
    launch_dwqkv_gram(ctx, qkv, pb.qkv_dw_w, pb.qkv_dw_b, vbuf, stats, B, H, W, C);
    v = vbuf;
  }
  void* Mw = A.elems((size_t)B * C * C, ctx.dtype);
  launch_attn_finalize(ctx, stats, pb.temperature, pb.proj_w, Mw, B, C, norms);
  if (keep != nullptr) {
    keep->v = v;
    keep->Mw = Mw;
    return 0;
  }
  GemmP g = gemm_rows(v, C, Mw, pb.proj_b, out, C, B, P, RF_K_GEMM_PROJ);
  g.lda1 = ldv;
  g.w_img = (i64)C * C;
  g.R = resid; g.ldr = C;
  g.stats_out = stats_out;
  const int np = launch_gemm(ctx, g);
  A.release(mk);
  return np;
}

// out = (resid ? resid : 0) + conv_ffn(xin); with `ln`, xin is the raw input and norm2 is folded into pointwise1
static void ffn(Ctx& ctx, const PackedBlock& pb, const void* xin, const void* resid, void* out, int B, int H, int W,
                const LnFold* ln = nullptr) {
  const int C = pb.C;
  const i64 P = (i64)H * W;
  Arena& A = ctx.arena;
  // one kernel for the whole FFN (hidden tensor in shared / tensor memory) where the shape allows
  // norm2 -> pointwise1 -> depthwise as one dense 3x3 conv on the tensor cores, GELU + pointwise2 + residual in its
  // epilogues (rf_lnconv.cu)
  if (ln != nullptr && xin == resid && ln->npart == 1 && pb.ffn_cw != nullptr && lnconv_supported(ctx, C, H, W)) {
    if (ctx.dry) return;
    if (!launch_lnconv_ffn(ctx, xin, ln->stats, pb.ffn_cw, pb.ffn_bt, pb.pw2_w, pb.pw2_b, out, B, H, W, C))
      recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return;
  }
  if (ln != nullptr && xin == resid && ffn_fused_supported(ctx, C, W) && ln->npart == 1) {
    if (ctx.dry) return;
    if (launch_ffn_fused(ctx, xin, pb.pw1_wf, pb.pw1_cs, pb.pw1_bf, ln->stats, ln->npart, pb.ffn_dw_w, pb.ffn_dw_b, pb.pw2_w,
                         pb.pw2_b, out, B, H, W, C))
      return;
    recorder().last_cuda_error = (int)cudaErrorNotSupported;   // (the dry run sized the arena without the hidden tensors)
    return;
  }
  const size_t mk = A.mark();
  void* hpre = A.elems((size_t)B * P * 2 * C, ctx.dtype);
  {
    GemmP g1 = gemm_rows(xin, C, ln ? pb.pw1_wf : pb.pw1_w, ln ? pb.pw1_bf : pb.pw1_b, hpre, 2 * C, B, P, RF_K_GEMM_PW1);
    if (ln) { g1.ln_stats = ln->stats; g1.ln_npart = ln->npart; g1.ln_cs = pb.pw1_cs; g1.ln_C = C; g1.ln_eps = 1e-5f; }
    launch_gemm(ctx, g1);
  }
  void* h = A.elems((size_t)B * P * 2 * C, ctx.dtype);
  launch_dwconv(ctx, hpre, pb.ffn_dw_w, pb.ffn_dw_b, h, 1, B, H, W, 2 * C, RF_K_DW_GELU);
  GemmP g = gemm_rows(h, 2 * C, pb.pw2_w, pb.pw2_b, out, C, B, P, RF_K_GEMM_PW2);
  g.R = resid; g.ldr = C;
  launch_gemm(ctx, g);
  A.release(mk);
}

// pre (optional): LayerNorm statistics of `feat` already emitted by the kernel that produced it
static void transformer(Ctx& ctx, const PackedBlock& pb, const void* feat, void* out, int B, int H, int W,
                        const LnFold* pre = nullptr) {
  const int C = pb.C;
  const i64 P = (i64)H * W;
  Arena& A = ctx.arena;
  const size_t mk = A.mark();
  if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {
    // bf16 mode: both LayerNorms are folded into the 1x1 conv that follows them (W*diag(g), per-row mean/rstd applied in
    // the GEMM epilogue), so the normalised tensors never exist.  norm1 statistics: one read pass over the block input;
    // norm2 statistics: emitted by the project_out GEMM epilogue that produces x1.
    LnFold l1;
    if (pre != nullptr && pre->stats != nullptr && pre->npart > 0) {
      l1 = *pre;
    } else if (C == 32 && pb.qkv_cw != nullptr && lnconv_supported(ctx, C, H, W)) {
      l1.stats = nullptr; l1.npart = 1;       // the dense-conv qkv kernel computes norm1's statistics itself (no pass over feat)
    } else {
      float* st1 = A.get<float>((size_t)B * P * 2);
      launch_row_stats(ctx, feat, st1, B * P, C);
      l1.stats = st1; l1.npart = 1;
    }
    // project_out fused in front of the FFN (C = 32): x1 = feat + Mw v + b exists per halo patch inside the kernel only (neither
    // x1 nor its statistics tensor is allocated)
    const bool fuse_proj = ctx.dtype == RF_BF16 && pb.ffn_cw != nullptr && pb.qkv_cw != nullptr && lnconv_proj_supported(ctx, C, H, W);
    if (fuse_proj) {
      AttnKeep keep;
      attention(ctx, pb, feat, feat, nullptr, B, H, W, &l1, nullptr, &keep);
      if (!ctx.dry) {
        for (int b = 0; b < B; ++b) {
          const size_t img = (size_t)b * P * C * 2;
          if (!launch_lnconv_ffn_proj(ctx, (const char*)feat + img, (const char*)keep.v + img, (const char*)keep.Mw + (size_t)b * C * C * 2,
                                      pb.proj_b, pb.ffn_cw, pb.ffn_bt, pb.pw2_w, pb.pw2_b, (char*)out + img, H, W, C))
            recorder().last_cuda_error = (int)cudaErrorNotSupported;
        }
      }
      A.release(mk);
      return;
    }
    void* x1 = A.elems((size_t)B * P * C, ctx.dtype);
    float* st2 = A.get<float>((size_t)B * P * 2 * 2);     // up to two N tiles of partials
    LnFold l2;
    l2.stats = st2;
    l2.npart = attention(ctx, pb, feat, feat, x1, B, H, W, &l1, st2);
    ffn(ctx, pb, x1, x1, out, B, H, W, &l2);
  } else {
    void* x1 = A.elems((size_t)B * P * C, ctx.dtype);
    void* ln = A.elems((size_t)B * P * C, ctx.dtype);
    launch_layernorm(ctx, feat, pb.ln1_g, pb.ln1_b, ln, 1e-5f, 0, B * P, C);
    attention(ctx, pb, ln, feat, x1, B, H, W);
    launch_layernorm(ctx, x1, pb.ln2_g, pb.ln2_b, ln, 1e-5f, 0, B * P, C);
    ffn(ctx, pb, ln, x1, out, B, H, W);
  }
  A.release(mk);
}

static void conv3x3(Ctx& ctx, const void* in, const void* w, const float* bias, void* out, int Cin, int N, int act, int omode,
                    int B, int H, int W, int kid) {
  GemmP p;
  p.A1 = in; p.K1 = 9 * Cin; p.lda1 = Cin; p.amode = AMODE_CONV3;
  p.Wt = w; p.bias = bias; p.Y = out;
  p.ldy = omode == OMODE_UNSHUFFLE ? 4 * N : N;
  p.M = H * W; p.N = N; p.B = B; p.H = H; p.W = W; p.act = act; p.omode = omode; p.kernel_id = kid;
  launch_gemm(ctx, p);
}

static void conv_transformer(Ctx& ctx, const PackedBlock& pb, int variant, const void* feat, const Stage& sg, void* out,
                             int B, const LnFold* pre = nullptr) {
  const int C = pb.C, H = sg.H, W = sg.W;
  const i64 P = (i64)H * W;
  Arena& A = ctx.arena;
  const size_t mk = A.mark();
  void* xmod;
  float* scale;
  SePartial se;
  flca_branch(ctx, pb, variant, feat, sg, B, &xmod, &scale, &se);
  void* wred = A.elems((size_t)B * C * 2 * C, ctx.dtype);
  // channel_reduce folded into Conv_out (C = 32, rf_lnconv.cu): the squeeze-excite scale goes into per-image 3x3 weights
  const bool cat_conv = ctx.dtype == RF_BF16 && pb.cat_p2 != nullptr && lnconv_cat_supported(ctx, C, H, W);
  void* weff = cat_conv ? A.elems((size_t)B * C * 9 * 2 * C, ctx.dtype) : nullptr;
  // squeeze-excite MLP + fold of its scale into channel_reduce: one launch.  Row-tiled forward: after the transformer
  // branch, whose all-reduce brings the channel sums of the whole frame (into row 0 of the partial sums)
  // (whole-frame forward: on the side stream, next to the transformer branch -- only channel_reduce needs its result)
  cudaStream_t side = nullptr;
  bool forked = false;
  if (ctx.band == nullptr) {
    forked = side_fork(ctx, &side);
    cudaStream_t main_stream = ctx.stream;
    if (forked) ctx.stream = side;
    launch_se_fold(ctx, se.partial, se.nblk, P, pb.se_w1, pb.se_b1, pb.se_w2, pb.se_b2, scale, pb.red_w, wred, B, C, pb.hid);
    if (cat_conv) launch_cat_scale(ctx, pb.cat_p2, scale, weff, B, C);
    ctx.stream = main_stream;
  }
  void* x2 = A.elems((size_t)B * P * C, ctx.dtype);
  transformer(ctx, pb, feat, x2, B, H, W, pre);
  if (forked) side_join(ctx, side);
  if (ctx.band != nullptr) {
    launch_se_fold(ctx, se.partial, 1, ctx.band->P_full, pb.se_w1, pb.se_b1, pb.se_w2, pb.se_b2, scale, pb.red_w, wred, B, C,
                   pb.hid);
    if (cat_conv) launch_cat_scale(ctx, pb.cat_p2, scale, weff, B, C);
  }
  if (cat_conv) {
    if (!ctx.dry) {
      for (int b = 0; b < B; ++b) {
        const size_t img = (size_t)b * P * C * 2;
        if (!launch_lnconv_cat(ctx, (const char*)xmod + img, (const char*)x2 + img, (const char*)weff + (size_t)b * C * 9 * 2 * C * 2,
                               pb.cat_bt, (char*)out + img, H, W, C))
          recorder().last_cuda_error = (int)cudaErrorNotSupported;
      }
    }
    A.release(mk);
    return;
  }
  void* xr = A.elems((size_t)B * P * C, ctx.dtype);
  GemmP g = gemm_rows(xmod, C, wred, pb.red_b, xr, C, B, P, RF_K_GEMM_CAT_REDUCE);
  g.A2 = x2; g.K2 = C; g.lda2 = C;
  g.w_img = (i64)C * 2 * C;
  launch_gemm(ctx, g);
  if (!(ctx.dtype == RF_BF16 && pb.convout_b != nullptr && lnconv_supported(ctx, C, H, W) &&
        (ctx.dry || launch_lnconv_conv3(ctx, xr, pb.convout_w, pb.convout_b, out, B, H, W, C))))
    conv3x3(ctx, xr, pb.convout_w, pb.convout_b, out, C, C, ACT_LRELU, OMODE_ROWS, B, H, W, RF_K_CONV3X3_OUT);
  A.release(mk);
}

// guidance of one stage from planar y-derived maps
static void make_stage(Ctx& ctx, int variant, Stage& sg, int Hf, int Wf, const float* LL1, const float* yh1, int H1, int W1,
                       const float* LL2, const float* yh2, int H2, int W2, const float* cr, const float* cb, int Hy, int Wy,
                       int B, int y_begin = 0, int y_rows = -1, int C_stage = 0) {
  const int NG = variant == RF_VARIANT_ML ? 8 : 4;
  sg.H = Hf; sg.W = Wf;
  // the fp32 maps are read by the CUDA-core kernels only: where the stage's FLCA kernels run on the tensor cores (they read
  // the [hi | lo] bf16 pixels) the maps are neither allocated nor written (half of this pass's stores)
  const bool tc_only = C_stage > 0 && ctx.dtype == RF_BF16 && tcgen05_enabled() && im2col_tc_supported(ctx, C_stage);
  sg.G = tc_only ? nullptr : ctx.arena.get<float>((size_t)B * Hf * Wf * NG);
  sg.sums = nullptr;
  if (variant == RF_VARIANT_ML) {
    sg.sums = ctx.arena.get<float>((size_t)B * 8);
    launch_fill_f32(ctx, sg.sums, 0.f, (i64)B * 8);
  }
  sg.G16 = nullptr; sg.G16b = nullptr;
  if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {   // the [hi | lo] bf16 pixels of the tensor-core FLCA kernels, same pass
    sg.G16 = ctx.arena.alloc((size_t)B * Hf * Wf * 16);
    if (variant == RF_VARIANT_ML) sg.G16b = ctx.arena.alloc((size_t)B * Hf * Wf * 16);
  }
  launch_guidance_stage(ctx, LL1, yh1, H1, W1, LL2, yh2, H2, W2, cr, cb, Hy, Wy, sg.G, NG, sg.sums, B, Hf, Wf, sg.G16, sg.G16b,
                        y_begin, y_rows);
}

struct GuidanceMaps {
  float *LL1 = nullptr, *yh1 = nullptr, *LL2 = nullptr, *yh2 = nullptr;
  int H1 = 0, W1 = 0, H2 = 0, W2 = 0;
};
static GuidanceMaps make_pyramid(Ctx& ctx, int variant, const float* y, const float* filt, int B, int Hy, int Wy) {
  GuidanceMaps g;
  g.H1 = (Hy + 1) / 2; g.W1 = (Wy + 1) / 2;
  g.LL1 = ctx.arena.get<float>((size_t)B * g.H1 * g.W1);
  g.yh1 = ctx.arena.get<float>((size_t)B * g.H1 * g.W1);
  launch_dwt_high(ctx, y, filt, g.LL1, g.yh1, B, Hy, Wy);
  if (variant == RF_VARIANT_ML) {
    g.H2 = (g.H1 + 1) / 2; g.W2 = (g.W1 + 1) / 2;
    g.LL2 = ctx.arena.get<float>((size_t)B * g.H2 * g.W2);
    g.yh2 = ctx.arena.get<float>((size_t)B * g.H2 * g.W2);
    launch_dwt_high(ctx, g.LL1, filt, g.LL2, g.yh2, B, g.H1, g.W1);
  }
  return g;
}

// ---------------------------------------------------------------------------------------------
// whole model
// ---------------------------------------------------------------------------------------------
float* zeroed_f32(Ctx& ctx, size_t n) {
  const size_t bytes = align_up(n * sizeof(float), 256);
  if (ctx.zero_off + bytes <= ctx.zero_cap) {
    float* p = ctx.zero_base ? reinterpret_cast<float*>(ctx.zero_base + ctx.zero_off) : nullptr;
    ctx.zero_off += bytes;
    return p;
  }
  float* p = ctx.arena.get<float>(n);
  launch_fill_f32(ctx, p, 0.f, (i64)n);
  return p;
}

static int model_forward(Ctx& ctx, const PackedModel& pm, int variant, const float* raw, float* out, int B, int H, int W) {
  const int d = pm.dim;
  const int h = H / 2, w = W / 2;
  Arena& A = ctx.arena;
  {
    // all per-block accumulators of the frame, cleared by one memset
    size_t zb = 0;
    for (int i = 0; i < 7; ++i) {
      const size_t C = (size_t)d << kBlockStage[i];
      zb += align_up(B * (C * C + 2 * C) * 4, 256) + align_up((size_t)B * num_sms() * 2 * C * 4, 256) +
            align_up((size_t)B * flca_num_partials((int)C, B, 0) * C * 4, 256);
    }
    ctx.zero_cap = zb;
    ctx.zero_off = 0;
    ctx.zero_base = (char*)A.alloc(zb);
    if (!ctx.dry) {
      if (!ctx.fits()) return RF_ERR_WORKSPACE;
      RF_CUDA(cudaMemsetAsync(ctx.zero_base, 0, zb, ctx.stream));
    }
  }
  const i64 P0 = (i64)h * w;
  float* x_ds = A.get<float>((size_t)B * P0 * 4);
  float* y_raw = A.get<float>((size_t)B * P0);
  float* ymax = A.get<float>((size_t)B);
  float* y = A.get<float>((size_t)B * P0);
  float* cr = A.get<float>((size_t)B * P0);
  float* cb = A.get<float>((size_t)B * P0);
  launch_fill_f32(ctx, ymax, -INFINITY, B);
  launch_pack_luma(ctx, raw, x_ds, y_raw, ymax, pm.rgb_w, B, H, W);
  launch_luma_finalize(ctx, x_ds, y_raw, ymax, 1e-6f, y, cr, cb, B, h, w);
  GuidanceMaps gm = make_pyramid(ctx, variant, y, pm.haar, B, h, w);
  Stage st[4];
  // (the guidance of stages 1-3 on a side stream next to the embedding and block 0: measured, no gain -- 5.451 vs 5.435 ms)
  for (int s = 0; s < 4; ++s)
    make_stage(ctx, variant, st[s], h >> s, w >> s, gm.LL1, gm.yh1, gm.H1, gm.W1, gm.LL2, gm.yh2, gm.H2, gm.W2, cr, cb, h, w,
               B, 0, -1, d << s);
  auto feat_buf = [&](int s) { return A.elems((size_t)B * (P0 >> (2 * s)) * ((size_t)d << s), ctx.dtype); };
  void* x0 = feat_buf(0);
  void* x16 = nullptr;
  if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {
    x16 = A.alloc((size_t)B * P0 * 16);
    launch_split_bf16x8(ctx, x_ds, x16, (i64)B * P0);
  }
  launch_embed(ctx, x_ds, x16, pm.embed_w, pm.embed_b, x0, B, h, w, d);
  void* enc[4];
  void* cur = x0;
  for (int s = 0; s < 4; ++s) {
    enc[s] = feat_buf(s);
    conv_transformer(ctx, pm.blocks[s], variant, cur, st[s], enc[s], B);
    if (s < 3) {
      void* pooled = feat_buf(s + 1);
      const int C = d << s;
      conv3x3(ctx, enc[s], pm.down_w[s], nullptr, pooled, C, C / 2, ACT_NONE, OMODE_UNSHUFFLE, B, h >> s, w >> s,
              RF_K_DOWN_CONV);
      cur = pooled;
    }
  }
  cur = enc[3];
  for (int n = 0; n < 3; ++n) {
    const int s = 2 - n;               // output stage
    const int Co = d << s, Ci = 2 * Co;
    const int Hs = h >> s, Ws = w >> s;
    const i64 Ps = (i64)Hs * Ws;
    void* up = feat_buf(s);
    GemmP g = gemm_rows(cur, Ci, pm.up_w[n], pm.up_b[n], up, 4 * Co, B, Ps / 4, RF_K_UP_CONVT);
    g.omode = OMODE_CONVT; g.H = Hs / 2; g.W = Ws / 2; g.ldy = Co;
    launch_gemm(ctx, g);
    void* fused = feat_buf(s);
    GemmP r = gemm_rows(up, Co, pm.red_w[n], pm.red_b[n], fused, Co, B, Ps, RF_K_SKIP_REDUCE);
    r.A2 = enc[s]; r.K2 = Co; r.lda2 = Co;
    // the skip-fusion GEMM's epilogue also emits the LayerNorm statistics of its rows (= norm1 input of the next block)
    LnFold pre;
    if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {
      float* stf = A.get<float>((size_t)B * Ps * 2 * 2);
      r.stats_out = stf;
      pre.stats = stf;
    }
    pre.npart = launch_gemm(ctx, r);
    void* dec = feat_buf(s);
    conv_transformer(ctx, pm.blocks[4 + n], variant, fused, st[s], dec, B, pre.stats ? &pre : nullptr);
    cur = dec;
  }
  bool head_done = false;
  if (ctx.dtype == RF_BF16 && tcgen05_enabled()) {
    // conv_out + LeakyReLU + PixelShuffle as an implicit GEMM (N padded to 16) with a scatter epilogue
    GemmP hp;
    hp.A1 = cur; hp.K1 = 9 * d; hp.lda1 = d; hp.amode = AMODE_CONV3;
    hp.Wt = pm.head_wt; hp.bias = pm.head_b16; hp.Y = out; hp.ldy = 0;
    hp.M = h * w; hp.N = 16; hp.B = B; hp.H = h; hp.W = w; hp.omode = OMODE_HEAD; hp.kernel_id = RF_K_HEAD;
    head_done = ctx.dry || launch_gemm_tcgen05(ctx, hp) >= 0;
  }
  if (!head_done) launch_head(ctx, cur, pm.head_w, pm.head_b, out, B, h, w, d);
  if (variant == RF_VARIANT_ML) {
    float* sums = A.get<float>((size_t)B * 8);
    launch_fill_f32(ctx, sums, 0.f, (i64)B * 8);
    launch_tail_stats(ctx, out, x_ds, sums, B, h, w);
    launch_tail_apply(ctx, out, sums, gm.LL2, gm.H2, gm.W2, B, h, w);
  }
  return RF_OK;
}

static int check_model_args(int dim, int dtype, int variant, int B, int H, int W);

// ---------------------------------------------------------------------------------------------
// row-tiled single frame (BASELINE config 4): this rank's band of ONE frame
// ---------------------------------------------------------------------------------------------
// Geometry: the frame is cut in units of 16 raw rows (= 8 >> s packed rows at stage s), so every stage has whole rows.
// Every activation is a band image [ht + n_s + hb][w_s][C_s] (ht, hb = BAND_HALO towards a neighbour, 0 at the frame
// border).  Validity of the halo rows along one Conv_Transformer (FLCA_RF.py:272-278), starting from 4 exchanged rows:
//   qkv 1x1 (4) -> dw3x3 (3) -> x1 (3) -> pointwise1 (3) -> dw3x3+GELU (2) -> x2 (2) -> channel_reduce (2) -> Conv_out (1)
// and Downsample's / conv_out's 3x3 consumes the last one.  The 1-channel guidance (y, cr, cb, Haar pyramid: FLCA_RF.py:
// 87-97,140-148) is computed for the WHOLE frame on every rank (replicate, 12 MB), so FLCA needs no halo.
static int model_forward_band(Ctx& ctx, const PackedModel& pm, int variant, const float* raw, float* out, int H, int W,
                              const rf_band& rb) {
  Band& bd = *ctx.band;
  const int d = pm.dim, B = 1;
  const int h = H / 2, w = W / 2;
  Arena& A = ctx.arena;
  const int u0 = rb.row0 / 16, nu = rb.rows / 16;
  bd.ht = rb.rank > 0 ? BAND_HALO : 0;
  bd.hb = rb.rank + 1 < rb.nranks ? BAND_HALO : 0;
  const int ht = bd.ht, hb = bd.hb;
  auto rows_in = [&](int s) { return nu * (8 >> s); };           // interior rows at stage s
  auto rows_img = [&](int s) { return ht + rows_in(s) + hb; };   // rows of the band image
  auto first_row = [&](int s) { return u0 * (8 >> s) - ht; };    // frame row of band-image row 0
  auto set_stage = [&](int s) {
    bd.rows_in = rows_in(s);
    bd.P_full = (i64)(h >> s) * (w >> s);
  };
  {
    size_t zb = 0;
    for (int i = 0; i < 7; ++i) {
      const size_t C = (size_t)d << kBlockStage[i];
      zb += align_up((C * C + 2 * C) * 4, 256) + align_up((size_t)num_sms() * 2 * C * 4, 256) +
            align_up((size_t)flca_num_partials((int)C, 1, 0) * C * 4, 256);
    }
    ctx.zero_cap = zb;
    ctx.zero_off = 0;
    ctx.zero_base = (char*)A.alloc(zb);
    if (!ctx.dry) {
      if (!ctx.fits()) return RF_ERR_WORKSPACE;
      RF_CUDA(cudaMemsetAsync(ctx.zero_base, 0, zb, ctx.stream));
    }
  }
  band_begin(ctx);
  // ---- whole-frame guidance (replicated) --------------------------------------------------------------------------
  const i64 P0 = (i64)h * w;
  float* x_ds = A.get<float>((size_t)P0 * 4);
  float* y_raw = A.get<float>((size_t)P0);
  float* ymax = A.get<float>(1);
  float* y = A.get<float>((size_t)P0);
  float* cr = A.get<float>((size_t)P0);
  float* cb = A.get<float>((size_t)P0);
  Band* const band = ctx.band;
  ctx.band = nullptr;                                   // the guidance kernels see the whole frame
  launch_fill_f32(ctx, ymax, -INFINITY, 1);
  launch_pack_luma(ctx, raw, x_ds, y_raw, ymax, pm.rgb_w, B, H, W);
  launch_luma_finalize(ctx, x_ds, y_raw, ymax, 1e-6f, y, cr, cb, B, h, w);
  GuidanceMaps gm = make_pyramid(ctx, variant, y, pm.haar, B, h, w);
  Stage st[4];
  for (int s = 0; s < 4; ++s) {
    // (the maps are laid out for the whole frame, but only the band's rows are produced -- except for the multi-level
    // variant, whose gates need the maps' whole-frame sums, ML_RF.py:151-156: every rank makes the whole maps)
    const bool ml = variant == RF_VARIANT_ML;
    make_stage(ctx, variant, st[s], h >> s, w >> s, gm.LL1, gm.yh1, gm.H1, gm.W1, gm.LL2, gm.yh2, gm.H2, gm.W2, cr, cb, h, w, B,
               ml ? 0 : first_row(s), ml ? -1 : rows_img(s), d << s);
    // the band's view of the stage guidance: rows [first_row, first_row + rows_img)
    const size_t px0 = (size_t)first_row(s) * (w >> s);
    if (st[s].G) st[s].G += px0 * (ml ? 8 : 4);
    if (st[s].G16) st[s].G16 = (char*)st[s].G16 + px0 * 16;
    if (st[s].G16b) st[s].G16b = (char*)st[s].G16b + px0 * 16;
    st[s].H = rows_img(s);
  }
  auto feat_buf = [&](int s) { return A.elems((size_t)rows_img(s) * (w >> s) * ((size_t)d << s), ctx.dtype); };
  auto row_off = [&](void* p, int rows, int s) {        // p + `rows` rows of a stage-s band image
    return (void*)((char*)p + (size_t)rows * (w >> s) * ((size_t)d << s) * esize(ctx.dtype));
  };
  void* x0 = feat_buf(0);
  const size_t px0 = (size_t)first_row(0) * w;
  void* x16 = A.alloc((size_t)rows_img(0) * w * 16);
  launch_split_bf16x8(ctx, x_ds + px0 * 4, x16, (i64)rows_img(0) * w);
  launch_embed(ctx, x_ds + px0 * 4, x16, pm.embed_w, pm.embed_b, x0, B, rows_img(0), w, d);
  ctx.band = band;

  // ---- encoder ---------------------------------------------------------------------------------------------------------
  void* enc[4];
  void* cur = x0;
  for (int s = 0; s < 4; ++s) {
    set_stage(s);
    enc[s] = feat_buf(s);
    band_halo_exchange(ctx, cur, w >> s, d << s);
    conv_transformer(ctx, pm.blocks[s], variant, cur, st[s], enc[s], B);
    if (s < 3) {
      // Downsample over the whole band image: its ht/2 + n_{s+1} + hb/2 output rows land ht/2 rows into the next
      // stage's band image (the outer rows are filled by that block's halo exchange)
      void* pooled = feat_buf(s + 1);
      const int C = d << s;
      conv3x3(ctx, enc[s], pm.down_w[s], nullptr, row_off(pooled, ht / 2, s + 1), C, C / 2, ACT_NONE, OMODE_UNSHUFFLE, B,
              rows_img(s), w >> s, RF_K_DOWN_CONV);
      cur = pooled;
    }
  }
  // ---- decoder ---------------------------------------------------------------------------------------------------------
  cur = enc[3];
  for (int n = 0; n < 3; ++n) {
    const int s = 2 - n;
    const int Co = d << s, Ci = 2 * Co;
    const int Ws = w >> s;
    set_stage(s);
    // ConvTranspose2d of the coarse band image minus its outer ht/2 (hb/2) rows = exactly the fine band image
    const int rows_c = rows_img(s + 1) - ht / 2 - hb / 2;
    void* up = feat_buf(s);
    GemmP g = gemm_rows(row_off(cur, ht / 2, s + 1), Ci, pm.up_w[n], pm.up_b[n], up, 4 * Co, B, (i64)rows_c * (Ws / 2),
                        RF_K_UP_CONVT);
    g.omode = OMODE_CONVT; g.H = rows_c; g.W = Ws / 2; g.ldy = Co;
    launch_gemm(ctx, g);
    void* fused = feat_buf(s);
    GemmP r = gemm_rows(up, Co, pm.red_w[n], pm.red_b[n], fused, Co, B, (i64)rows_img(s) * Ws, RF_K_SKIP_REDUCE);
    r.A2 = enc[s]; r.K2 = Co; r.lda2 = Co;
    launch_gemm(ctx, r);
    band_halo_exchange(ctx, fused, Ws, Co);
    void* dec = feat_buf(s);
    conv_transformer(ctx, pm.blocks[4 + n], variant, fused, st[s], dec, B);
    cur = dec;
  }
  // ---- head: conv_out + LeakyReLU + PixelShuffle over the band image (the caller keeps the interior rows) ----------
  GemmP hp;
  hp.A1 = cur; hp.K1 = 9 * d; hp.lda1 = d; hp.amode = AMODE_CONV3;
  hp.Wt = pm.head_wt; hp.bias = pm.head_b16; hp.Y = out; hp.ldy = 0;
  hp.M = rows_img(0) * w; hp.N = 16; hp.B = B; hp.H = rows_img(0); hp.W = w; hp.omode = OMODE_HEAD; hp.kernel_id = RF_K_HEAD;
  if (!ctx.dry && launch_gemm_tcgen05(ctx, hp) < 0) return RF_ERR_UNSUPPORTED;
  if (variant == RF_VARIANT_ML) {
    // colour anchor + LL nudge (ML_RF.py:270-288): the input means are the whole frame's (every rank has the packed frame), the
    // three output-channel sums run over the band's interior rows and are summed over the ranks (one more sync point)
    float* sums = A.get<float>(8);
    launch_fill_f32(ctx, sums, 0.f, 8);
    launch_tail_stats(ctx, nullptr, x_ds, sums, 1, h, w);
    launch_out_sums(ctx, out, 2 * rows_img(0), 2 * ht, 2 * rows_in(0), 2 * w, sums + 3);
    band_allreduce_small(ctx, sums + 3, 3);
    launch_tail_apply(ctx, out, sums, gm.LL2, gm.H2, gm.W2, 1, h, w, 2 * first_row(0), 2 * rows_img(0));
  }
  return RF_OK;
}

static int check_band_args(int dim, int dtype, int variant, int H, int W, const rf_band* band) {
  if (!band) return RF_ERR_BAD_ARG;
  RF_TRY(check_model_args(dim, dtype, variant, 1, H, W));
  if (dtype != RF_BF16) return RF_ERR_UNSUPPORTED;                // (the fp32 parity engine has no band mode)
  if (dim % 32 && dim % 48) return RF_ERR_UNSUPPORTED;            // tensor-core FLCA / embed kernels (dim 32/48/64)
  if (band->nranks < 1 || band->nranks > RF_BAND_MAX_RANKS || band->rank < 0 || band->rank >= band->nranks)
    return RF_ERR_BAD_ARG;
  if (band->row0 % 16 || band->rows % 16 || band->rows < 16 * BAND_HALO) return RF_ERR_BAD_SHAPE;
  if (band->row0 < 0 || band->row0 + band->rows > H) return RF_ERR_BAD_SHAPE;
  if ((band->rank == 0) != (band->row0 == 0)) return RF_ERR_BAD_SHAPE;
  if ((band->rank == band->nranks - 1) != (band->row0 + band->rows == H)) return RF_ERR_BAD_SHAPE;
  return RF_OK;
}

static void band_from_abi(const rf_band& rb, Band& bd, bool dry) {
  bd.rank = rb.rank; bd.nranks = rb.nranks; bd.epoch = rb.epoch;
  for (int r = 0; r < rb.nranks; ++r) bd.comm[r] = dry ? nullptr : (char*)rb.comm[r];
}

static int check_model_args(int dim, int dtype, int variant, int B, int H, int W) {
  if (dtype != RF_F32 && dtype != RF_BF16) return RF_ERR_BAD_ARG;
  if (variant != RF_VARIANT_FLCA && variant != RF_VARIANT_ML) return RF_ERR_BAD_ARG;
  if (dim <= 0 || dim % 8) return RF_ERR_BAD_SHAPE;
  if (dim > 64) return RF_ERR_UNSUPPORTED;  // C <= 512 (Gram register tiling, TMEM columns)
  if (B <= 0 || H <= 0 || W <= 0 || H % 16 || W % 16) return RF_ERR_BAD_SHAPE;
  return RF_OK;
}

static int check_block_args(int C, int dtype, int B, int H, int W) {
  if (dtype != RF_F32 && dtype != RF_BF16) return RF_ERR_BAD_ARG;
  if (C <= 0 || C % 8) return RF_ERR_BAD_SHAPE;
  if (C > 512) return RF_ERR_UNSUPPORTED;
  if (B <= 0 || H <= 0 || W <= 0) return RF_ERR_BAD_SHAPE;
  return RF_OK;
}

static int finish(Ctx& ctx) {
  if (!ctx.fits()) return RF_ERR_WORKSPACE;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e);
  if (recorder().last_cuda_error != 0) return RF_ERR_CUDA;
  return RF_OK;
}

static Ctx make_ctx(void* ws, size_t ws_bytes, void* stream, int dtype, bool dry) {
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  ctx.arena.base = dry ? nullptr : (char*)ws;
  ctx.arena.cap = ws_bytes;
  if (!dry && ws != nullptr && (uintptr_t)ws % 256) {  // keep every arena block 256-byte aligned
    size_t adj = 256 - (uintptr_t)ws % 256;
    ctx.arena.base += adj;
    ctx.arena.cap = ws_bytes > adj ? ws_bytes - adj : 0;
  }
  ctx.dry = dry;
  ctx.dtype = dtype;
  if (!dry) recorder().last_cuda_error = 0;
  return ctx;
}

// sub-module entry: pack the given weights into the arena, convert NCHW->NHWC, build the stage guidance
struct BlockEnv {
  PackedBlock pb;
  Stage sg;
  void* x = nullptr;     // NHWC input
  void* outT = nullptr;  // NHWC output
};

template <typename F>
static int run_block_entry(const rf_block_weights* w, int C, int dtype, int variant, const float* feat, const float* y,
                           const float* cr, const float* cb, float* out, int B, int Hf, int Wf, int Hy, int Wy, void* ws,
                           size_t ws_bytes, void* stream, bool dry, size_t* peak, F body) {
  Ctx ctx = make_ctx(ws, ws_bytes, stream, dtype, dry);
  Arena& A = ctx.arena;
  BlockEnv env;
  {
    Layout L(dry ? nullptr : A.base);
    env.pb = layout_block(L, C, dtype, variant);
    A.alloc(L.off);
  }
  const i64 P = (i64)Hf * Wf;
  env.x = A.elems((size_t)B * P * C, dtype);
  env.outT = A.elems((size_t)B * P * C, dtype);
  if (!dry && !ctx.fits()) return RF_ERR_WORKSPACE;
  if (!dry) {
    // the workspace must hold the plan; re-check after the body via finish()
    pack_block(ctx, *w, env.pb, variant);
    launch_nchw_to_nhwc(ctx, feat, env.x, B, C, P);
  }
  if (y != nullptr || dry) {
    if (Hy > 0 && Wy > 0) {
      GuidanceMaps gm = make_pyramid(ctx, variant, y, w ? w->flca_filt : nullptr, B, Hy, Wy);
      make_stage(ctx, variant, env.sg, Hf, Wf, gm.LL1, gm.yh1, gm.H1, gm.W1, gm.LL2, gm.yh2, gm.H2, gm.W2, cr, cb, Hy, Wy, B, 0, -1,
                 C);
    }
  }
  env.sg.H = Hf; env.sg.W = Wf;
  if (!dry && !ctx.fits()) return RF_ERR_WORKSPACE;
  body(ctx, env);
  if (!dry && !ctx.fits()) return RF_ERR_WORKSPACE;
  launch_nhwc_to_nchw(ctx, env.outT, out, B, C, P);
  if (peak) *peak = A.peak;
  return dry ? RF_OK : finish(ctx);
}

}  // namespace rf

using namespace rf;

extern "C" {

size_t rf_block_workspace_bytes(int C, int dtype, int B, int Hf, int Wf, int Hy, int Wy) {
  if (check_block_args(C, dtype, B, Hf, Wf) != RF_OK) return 0;
  size_t best = 0;
  for (int variant = 0; variant < 2; ++variant) {
    size_t peak = 0;
    run_block_entry(nullptr, C, dtype, variant, nullptr, nullptr, nullptr, nullptr, nullptr, B, Hf, Wf, Hy, Wy, nullptr, 0,
                    nullptr, true, &peak, [&](Ctx& ctx, BlockEnv& env) {
                      conv_transformer(ctx, env.pb, variant, env.x, env.sg, env.outT, B);
                    });
    if (peak > best) best = peak;
  }
  return best + 4096;
}

#define RF_BLOCK_PRECHECK(need_guidance)                                              \
  if (!w || !out || !workspace) return RF_ERR_BAD_ARG;                                \
  {                                                                                   \
    int _s = check_block_args(C, dtype, B, H_, W_);                                   \
    if (_s != RF_OK) return _s;                                                       \
  }

int rf_flca_forward(const rf_block_weights* w, int C, int dtype, int variant, const float* feat, const float* y,
                    const float* cr, const float* cb, float* out, int B, int Hf, int Wf, int Hy, int Wy, void* workspace,
                    size_t workspace_bytes, void* stream) {
  const int H_ = Hf, W_ = Wf;
  RF_BLOCK_PRECHECK(true)
  if (!feat || !y || !cr || !cb || Hy <= 0 || Wy <= 0) return RF_ERR_BAD_ARG;
  if (variant != RF_VARIANT_FLCA && variant != RF_VARIANT_ML) return RF_ERR_BAD_ARG;
  if (!w->flca_filt) return RF_ERR_BAD_ARG;
  return run_block_entry(w, C, dtype, variant, feat, y, cr, cb, out, B, Hf, Wf, Hy, Wy, workspace, workspace_bytes, stream,
                         false, nullptr, [&](Ctx& ctx, BlockEnv& env) {
                           void* xmod;
                           float* scale;
                           flca_branch(ctx, env.pb, variant, env.x, env.sg, B, &xmod, &scale);
                           launch_scale_channels(ctx, xmod, scale, env.outT, B, (i64)Hf * Wf, C);
                         });
}

int rf_attention_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                         void* workspace, size_t workspace_bytes, void* stream) {
  const int H_ = H, W_ = W;
  RF_BLOCK_PRECHECK(false)
  if (!x) return RF_ERR_BAD_ARG;
  return run_block_entry(w, C, dtype, RF_VARIANT_FLCA, x, nullptr, nullptr, nullptr, out, B, H, W, 0, 0, workspace,
                         workspace_bytes, stream, false, nullptr,
                         [&](Ctx& ctx, BlockEnv& env) { attention(ctx, env.pb, env.x, nullptr, env.outT, B, H, W); });
}

int rf_conv_ffn_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                        void* workspace, size_t workspace_bytes, void* stream) {
  const int H_ = H, W_ = W;
  RF_BLOCK_PRECHECK(false)
  if (!x) return RF_ERR_BAD_ARG;
  return run_block_entry(w, C, dtype, RF_VARIANT_FLCA, x, nullptr, nullptr, nullptr, out, B, H, W, 0, 0, workspace,
                         workspace_bytes, stream, false, nullptr,
                         [&](Ctx& ctx, BlockEnv& env) { ffn(ctx, env.pb, env.x, nullptr, env.outT, B, H, W); });
}

int rf_transformer_block_forward(const rf_block_weights* w, int C, int dtype, const float* x, float* out, int B, int H,
                                 int W, void* workspace, size_t workspace_bytes, void* stream) {
  const int H_ = H, W_ = W;
  RF_BLOCK_PRECHECK(false)
  if (!x) return RF_ERR_BAD_ARG;
  return run_block_entry(w, C, dtype, RF_VARIANT_FLCA, x, nullptr, nullptr, nullptr, out, B, H, W, 0, 0, workspace,
                         workspace_bytes, stream, false, nullptr,
                         [&](Ctx& ctx, BlockEnv& env) { transformer(ctx, env.pb, env.x, env.outT, B, H, W); });
}

int rf_conv_transformer_forward(const rf_block_weights* w, int C, int dtype, int variant, const float* feat, const float* y,
                                const float* cr, const float* cb, float* out, int B, int Hf, int Wf, int Hy, int Wy,
                                void* workspace, size_t workspace_bytes, void* stream) {
  const int H_ = Hf, W_ = Wf;
  RF_BLOCK_PRECHECK(true)
  if (!feat || !y || !cr || !cb || Hy <= 0 || Wy <= 0) return RF_ERR_BAD_ARG;
  if (variant != RF_VARIANT_FLCA && variant != RF_VARIANT_ML) return RF_ERR_BAD_ARG;
  if (!w->flca_filt) return RF_ERR_BAD_ARG;
  return run_block_entry(w, C, dtype, variant, feat, y, cr, cb, out, B, Hf, Wf, Hy, Wy, workspace, workspace_bytes, stream,
                         false, nullptr, [&](Ctx& ctx, BlockEnv& env) {
                           conv_transformer(ctx, env.pb, variant, env.x, env.sg, env.outT, B);
                         });
}

int rf_downsample_forward(const float* conv_w, int C, int dtype, const float* x, float* out, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!conv_w || !x || !out || !workspace) return RF_ERR_BAD_ARG;
  RF_TRY(check_block_args(C, dtype, B, H, W));
  if ((H & 1) || (W & 1) || (C % 16)) return RF_ERR_BAD_SHAPE;
  Ctx ctx = make_ctx(workspace, workspace_bytes, stream, dtype, false);
  Arena& A = ctx.arena;
  const i64 P = (i64)H * W;
  void* wT = A.elems((size_t)(C / 2) * 9 * C, dtype);
  void* xT = A.elems((size_t)B * P * C, dtype);
  void* oT = A.elems((size_t)B * (P / 4) * 2 * C, dtype);
  if (!ctx.fits()) return RF_ERR_WORKSPACE;
  pack_conv3(ctx, conv_w, wT, C / 2, C);
  launch_nchw_to_nhwc(ctx, x, xT, B, C, P);
  conv3x3(ctx, xT, wT, nullptr, oT, C, C / 2, ACT_NONE, OMODE_UNSHUFFLE, B, H, W, RF_K_DOWN_CONV);
  launch_nhwc_to_nchw(ctx, oT, out, B, 2 * C, P / 4);
  return finish(ctx);
}

// ---- whole model --------------------------------------------------------------------------------------

size_t rf_model_packed_bytes(int dim, int dtype, int variant) {
  if (dim <= 0 || dim % 8 || (dtype != RF_F32 && dtype != RF_BF16)) return 0;
  Layout L(nullptr);
  layout_model(L, dim, dtype, variant);
  return L.off + 256;
}

int rf_model_pack(const rf_model_weights* w, int dim, int dtype, int variant, void* packed, size_t packed_bytes,
                  void* stream) {
  if (!w || !packed) return RF_ERR_BAD_ARG;
  RF_TRY(check_model_args(dim, dtype, variant, 1, 16, 16));
  if (packed_bytes < rf_model_packed_bytes(dim, dtype, variant)) return RF_ERR_WORKSPACE;
  if ((uintptr_t)packed % 256) return RF_ERR_BAD_ARG;
  Ctx ctx = make_ctx(nullptr, 0, stream, dtype, false);
  Layout L(packed);
  PackedModel pm = layout_model(L, dim, dtype, variant);
  RF_CUDA(cudaMemcpyAsync(pm.rgb_w, w->rgb_w_host, 3 * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));
  const float* filt = variant == RF_VARIANT_ML ? w->blocks[0].flca_filt : w->blocks[0].flca_filt;
  if (!filt) return RF_ERR_BAD_ARG;
  copy_f32(ctx, filt, pm.haar, 16);
  if (!w->embedding_w || !w->embedding_b || !w->conv_out_w || !w->conv_out_b) return RF_ERR_BAD_ARG;
  launch_pack3(ctx, w->embedding_w, pm.embed_w, RF_F32, 9, 4, dim, 1, 9, 36, (i64)4 * dim, dim, 1, 0);
  copy_f32(ctx, w->embedding_b, pm.embed_b, dim);
  for (int i = 0; i < 7; ++i) pack_block(ctx, w->blocks[i], pm.blocks[i], variant);
  for (int n = 0; n < 3; ++n) {
    const int C = dim << n;
    if (!w->down_w[n] || !w->up_w[n] || !w->up_b[n] || !w->reduce_w[n] || !w->reduce_b[n]) return RF_ERR_BAD_ARG;
    pack_conv3(ctx, w->down_w[n], pm.down_w[n], C / 2, C);
    const int Co = dim << (2 - n), Ci = 2 * Co;
    // ConvTranspose2d weight [Ci,Co,2,2] -> T [(2i+j)][Co][Ci]
    launch_pack3(ctx, w->up_w[n], pm.up_w[n], dtype, 4, Co, Ci, 1, 4, (i64)Co * 4, (i64)Co * Ci, Ci, 1, 0);
    launch_pack3(ctx, w->up_b[n], pm.up_b[n], RF_F32, 4, 1, Co, 0, 0, 1, Co, 0, 1, 0);
    copy_T(ctx, w->reduce_w[n], pm.red_w[n], (i64)Co * 2 * Co);
    copy_f32(ctx, w->reduce_b[n], pm.red_b[n], Co);
  }
  launch_pack3(ctx, w->conv_out_w, pm.head_w, RF_F32, 9, dim, 12, 1, 9, (i64)dim * 9, (i64)dim * 12, 12, 1, 0);
  copy_f32(ctx, w->conv_out_b, pm.head_b, 12);
  RF_CUDA(cudaMemsetAsync(pm.head_wt, 0, (size_t)16 * 9 * dim * esize(dtype), ctx.stream));
  RF_CUDA(cudaMemsetAsync(pm.head_b16, 0, 16 * sizeof(float), ctx.stream));
  pack_conv3(ctx, w->conv_out_w, pm.head_wt, 12, dim);
  copy_f32(ctx, w->conv_out_b, pm.head_b16, 12);
  return finish(ctx);
}

size_t rf_rawformer_workspace_bytes(int dim, int dtype, int variant, int B, int H, int W) {
  if (check_model_args(dim, dtype, variant, B, H, W) != RF_OK) return 0;
  Ctx ctx = make_ctx(nullptr, 0, nullptr, dtype, true);
  Layout L(nullptr);
  PackedModel pm = layout_model(L, dim, dtype, variant);
  model_forward(ctx, pm, variant, nullptr, nullptr, B, H, W);
  return ctx.arena.peak + 4096;
}

int rf_rawformer_forward(const void* packed, int dim, int dtype, int variant, const float* raw, float* out, int B, int H,
                         int W, void* workspace, size_t workspace_bytes, void* stream) {
  if (!packed || !raw || !out || !workspace) return RF_ERR_BAD_ARG;
  RF_TRY(check_model_args(dim, dtype, variant, B, H, W));
  if ((uintptr_t)packed % 256 || (uintptr_t)raw % 16 || (uintptr_t)out % 16) return RF_ERR_BAD_ARG;
  if (workspace_bytes < rf_rawformer_workspace_bytes(dim, dtype, variant, B, H, W) - 4096) return RF_ERR_WORKSPACE;
  Ctx ctx = make_ctx(workspace, workspace_bytes, stream, dtype, false);
  Layout L(const_cast<void*>(packed));
  PackedModel pm = layout_model(L, dim, dtype, variant);
  RF_TRY(model_forward(ctx, pm, variant, raw, out, B, H, W));
  return finish(ctx);
}

// ---- row-tiled single frame ----------------------------------------------------------------------------------------

int rf_band_out_rows(const rf_band* band, int* out_rows_host, int* interior_row0_host) {
  if (!band || band->nranks < 1 || band->rank < 0 || band->rank >= band->nranks) return RF_ERR_BAD_ARG;
  const int ht = band->rank > 0 ? BAND_HALO : 0, hb = band->rank + 1 < band->nranks ? BAND_HALO : 0;
  if (out_rows_host) *out_rows_host = 2 * (ht + band->rows / 2 + hb);
  if (interior_row0_host) *interior_row0_host = 2 * ht;
  return RF_OK;
}

// dry run of rank `rank`'s plan: workspace peak and comm-region bytes
static int band_dry_run(int dim, int dtype, int variant, int H, int W, const rf_band& rb, size_t* ws, size_t* comm) {
  Ctx ctx = make_ctx(nullptr, 0, nullptr, dtype, true);
  Band bd;
  band_from_abi(rb, bd, true);
  ctx.band = &bd;
  Layout L(nullptr);
  PackedModel pm = layout_model(L, dim, dtype, variant);
  RF_TRY(model_forward_band(ctx, pm, variant, nullptr, nullptr, H, W, rb));
  if (ws) *ws = ctx.arena.peak + 4096;
  if (comm) *comm = bd.mail_off + 256;
  return RF_OK;
}

size_t rf_band_comm_bytes(int dim, int dtype, int variant, int H, int W, int nranks) {
  if (nranks < 1 || nranks > RF_BAND_MAX_RANKS || H % 16 || H / 16 < nranks * BAND_HALO) return 0;
  // mailbox sizes depend on (dim, W, nranks) only; size them with rank 0's plan of an even split
  rf_band rb;
  memset(&rb, 0, sizeof(rb));
  rb.rank = 0; rb.nranks = nranks; rb.row0 = 0;
  rb.rows = nranks == 1 ? H : (H / 16 / nranks) * 16;
  if (check_band_args(dim, dtype, variant, H, W, &rb) != RF_OK) return 0;
  size_t comm = 0;
  if (band_dry_run(dim, dtype, variant, H, W, rb, nullptr, &comm) != RF_OK) return 0;
  return comm;
}

size_t rf_rawformer_band_workspace_bytes(int dim, int dtype, int variant, int H, int W, const rf_band* band) {
  if (check_band_args(dim, dtype, variant, H, W, band) != RF_OK) return 0;
  size_t ws = 0;
  if (band_dry_run(dim, dtype, variant, H, W, *band, &ws, nullptr) != RF_OK) return 0;
  return ws;
}

int rf_rawformer_forward_band(const void* packed, int dim, int dtype, int variant, const float* raw, float* out, int H,
                              int W, const rf_band* band, void* workspace, size_t workspace_bytes, void* stream) {
  if (!packed || !raw || !out || !workspace) return RF_ERR_BAD_ARG;
  RF_TRY(check_band_args(dim, dtype, variant, H, W, band));
  if ((uintptr_t)packed % 256 || (uintptr_t)raw % 16 || (uintptr_t)out % 16) return RF_ERR_BAD_ARG;
  for (int r = 0; r < band->nranks; ++r)
    if (!band->comm[r] || (uintptr_t)band->comm[r] % 256) return RF_ERR_BAD_ARG;
  if (workspace_bytes < rf_rawformer_band_workspace_bytes(dim, dtype, variant, H, W, band) - 4096) return RF_ERR_WORKSPACE;
  if (!tcgen05_enabled()) return RF_ERR_UNSUPPORTED;
  Ctx ctx = make_ctx(workspace, workspace_bytes, stream, dtype, false);
  Band bd;
  band_from_abi(*band, bd, false);
  ctx.band = &bd;
  Layout L(const_cast<void*>(packed));
  PackedModel pm = layout_model(L, dim, dtype, variant);
  RF_TRY(model_forward_band(ctx, pm, variant, raw, out, H, W, *band));
  return finish(ctx);
}

int rf_rawformer_forward_band_profiled(const void* packed, int dim, int dtype, int variant, const float* raw, float* out,
                                       int H, int W, const rf_band* band, void* workspace, size_t workspace_bytes,
                                       void* stream, float* kernel_ms_host, int* kernel_id_host, int cap, int* n_host) {
  if (cap <= 0) return RF_ERR_BAD_ARG;
  RF_TRY(profile_begin((cudaStream_t)stream, cap));
  int st = rf_rawformer_forward_band(packed, dim, dtype, variant, raw, out, H, W, band, workspace, workspace_bytes, stream);
  int st2 = profile_end(kernel_ms_host, kernel_id_host, cap, n_host);
  return st != RF_OK ? st : st2;
}

int rf_rawformer_forward_profiled(const void* packed, int dim, int dtype, int variant, const float* raw, float* out, int B,
                                  int H, int W, void* workspace, size_t workspace_bytes, void* stream, float* kernel_ms_host,
                                  int* kernel_id_host, int cap, int* n_host) {
  if (cap <= 0) return RF_ERR_BAD_ARG;
  RF_TRY(profile_begin((cudaStream_t)stream, cap));
  int st = rf_rawformer_forward(packed, dim, dtype, variant, raw, out, B, H, W, workspace, workspace_bytes, stream);
  int st2 = profile_end(kernel_ms_host, kernel_id_host, cap, n_host);
  return st != RF_OK ? st : st2;
}

}  // extern "C"
