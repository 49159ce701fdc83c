// LayerNorm -> 1x1 conv -> depthwise 3x3 of the transformer branch as ONE dense 3x3 convolution on the tensor cores
// (bf16 mode, C = 32 and C = 64: the two full-resolution stages of RawFormer-S, the first stage of RawFormer-L).
//
//   conv_ffn   (FLCA_RF.py:204-209, :253):  out = x + pointwise2( gelu( depthwise( pointwise1( norm2(x) ) ) ) )
//   Attention  (FLCA_RF.py:223)          :  q|k|v = qkv_dwconv( qkv( norm1(x) ) )
//   Conv_out   (FLCA_RF.py:276)          :  out = lrelu( conv3x3(x) )                 (the same pipeline without the norm)
//
// A 1x1 conv followed by a depthwise 3x3 is linear, so it IS a dense 3x3 conv with the weights
//     Weff[n][tap][c] = dw[n][tap] * W[n][c] * gamma[c]                    (made once at pack time, k_pack_lnconv)
// applied to xhat = (x - mean) * rstd.  The depthwise convolution zero-pads the HIDDEN tensor (bias included), so the bias
// term of an output pixel is dw_b + sum over the taps that fall inside the image of dw[tap] * (b + W beta): a table of 9
// vectors indexed by the pixel's border state (3 row states x 3 column states).
//
// Why: on the CUDA cores the depthwise 3x3 (+ GELU) costs 17-20 cycles per pixel per 64 channels per SM whether its
// input comes from HBM or from shared memory (k_dw_tma, k_ffn_fused, k_dwqk_gram are all bound by
// that loop); as part of the contraction it costs 9x the tensor FLOPs of the 1x1 conv, which the tensor pipe has to
// spare: 18 (C = 32) / 36 (C = 64) tcgen05.mma per 128-pixel tile, ~60 cycles each whatever N <= 96 is.
//
// One persistent CTA per SM, 576 threads, NO CTA-wide barrier in the tile loop (the mbarriers count one arrival per warp,
// the warps drift apart by as much as the buffer depths allow):
//   warp 0      TMA producer: per tile ONE 4-d box = the (16+2) x (8+2) halo patch of x (pixel-major, 64 / 128-byte
//               swizzled) and ONE box of the per-pixel LayerNorm statistics (sum, sumsq); zero fill outside the image
//   warp 1      MMA issuer, one elected lane: 9 taps x C/16 k-steps out of the re-laid-out patch (no-swizzle K-major
//               operand, rows = pixels: the taps are nine start addresses into the same buffer, see rf_tc_gemm.cu), then
//               the second contraction of the tile (pointwise2, or the self-Gram of [q|k])
//   warps 2-17  compute, per tile: (A) both accumulators into registers, (B) FFN: acc3 of the tile before + bias + residual
//               -> bf16 -> global; acc1 + bias table -> FFN: GELU -> bf16 -> g tile (SWIZZLE_128B K-major operand of
//               pointwise2); QKV: q|k -> bf16 -> g tile (the same bytes are the MN-major operand of the Gram), v -> global;
//               (C) normalise + re-lay the patch two tiles ahead ([pixel][C] -> [C/8][pixel][16 B], one thread per pixel).
// q and k never reach HBM, nor does the 2C-wide hidden tensor of the FFN; the Gram and the squared norms of the CTA's
// pixels stay in tensor memory until the CTA's last tile and go to its own partial slot (k_attn_reduce sums the slots in
// order: bit-reproducible).  At C = 64 nine taps of 64 output channels are all the weights that fit: the FFN runs as two
// launches (hidden channels 0-63, then 64-127 on top of the first launch's output), q|k|v as three (q|k of heads 0-3, of
// heads 4-7 -- the Gram is block-diagonal over the heads -- and v).
//
// Measured (RawFormer-S full frame, B200): FFN 486 -> 228 us and qkv (1x1 + depthwise + Gram) 469 -> 218 us per stage-0
// block, 249 -> 206 and 251 -> 212 us per stage-1 block; Conv_out 88 -> 67 us at stage 1; with project_out in front of the
// stage-0 FFN (p.proj below) 312 us against 224 + 127 for the pair.  RAWFORMER_B200_LNCONV_DBG=1 prints the per-phase cycle
// counters of compute warp 0.  The tile loop is a dependency chain on which every mbarrier wait costs 50-200 cycles, even on a
// phase that completed long ago: a completion that causally follows another one stands for it (see the comments at the waits).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rf_kernels.cuh"
#include "rf_tma.cuh"
#include "rf_dw_math.cuh"

namespace rf {

constexpr int LC_CW = 16;                   // compute warps
constexpr int LC_CT = LC_CW * 32;           // compute threads
constexpr int LC_THREADS = LC_CT + 64;      // + producer warp + MMA warp
constexpr int LC_NPIX = 180;                // (16 + 2) x (8 + 2) patch pixels, row pitch 10
constexpr int LC_NT = 3;                    // re-laid-out patch buffers
constexpr int LC_MAXR = 6;                  // raw patch buffers (FFN: a patch lives until its tile's residual add)
constexpr uint32_t LC_ST_BYTES = 18 * 12 * 8, LC_ST_STRIDE = 1792;
// FFN: GELU + pointwise2 + residual epilogues (C = 64: one half of the 2C hidden channels per launch); QKV (C = 32): q|k -> Gram,
// v -> global; QK / V (C = 64): the same split over three launches (q|k of heads 0-3, of heads 4-7, v): 9 taps of 64 output
// channels are all the weights that fit next to the patches in shared memory
// CONV: the block's plain Conv_out 3x3 (C -> C, + bias, LeakyReLU(0.2), FLCA_RF.py:276) through the same pipeline: no
// LayerNorm in the re-layout, no statistics box, one bias row for all border states
// CAT (C = 32): Conv_out(channel_reduce(cat(xmod * s, x2))) (FLCA_RF.py:274-276) is linear before the LeakyReLU, so it is ONE
// dense 3x3 conv of the 2C-channel pair (xmod | x2) with the per-image weights sum_m Wout[n][tap][m] * Wred[m][k] * s[k]: two
// patches per tile, K = 2C per tap, N = C; the channel_reduce GEMM and its output tensor disappear
enum { LC_FFN = 0, LC_QKV = 1, LC_QK = 2, LC_V = 3, LC_CONV = 4, LC_CAT = 5 };

struct LcP {
  const float* btab;     // [9][n_tab] bias by border state (3 * row state + column state; state 0 = first, 1 = inner, 2 = last)
  const float* b2;       // FFN: [C] pointwise2 bias (NULL: none -- second half of the hidden channels)
  const bf16* resid;     // FFN, C = 64: residual [B,H,W,C] read from global memory (C = 32: the patch's own centre pixels)
  bf16* out;             // FFN: out [B,H,W,C]; QKV / V: v [B,H,W,C]
  int n_tab;             // channels of the whole conv
  int tab_stride;        // pitch of btab (0: one bias row for every border state)
  int t0, t1, tn;        // this launch computes conv channels [t0, t0 + tn) and, if tn < N, [t1, t1 + tn)
  int ch0;               // QKV / QK: first q (= k) channel of this launch
  float* gram_part;      // QKV: [slot][C][C/8] per-head diagonal blocks of q^T k of the CTA's pixels
  float* sq_part;        // QKV: [slot][2C]
  float invC, eps;
  int H, W, B;
  int ylo, yhi;          // QKV: rows that count for the statistics (row-tiled forward: the band's interior)
  int tiles_x, tiles_y, total_tiles;
  int nr;                // raw patch buffers in use
  int proj;              // FFN, C = 32: project_out fused in front (x1 = x + Mw v + proj_b never reaches HBM), see below
  const float* proj_b;   // [C]
  int wpre;              // static weights are fetched before the dependency wait (debugging aid: RAWFORMER_B200_LNCONV_WPRE=0)
  int own_stats;         // C = 32: no statistics tensor, the re-layout thread computes (sum, sumsq) of its pixel itself
  int sched;             // 0: the first contraction of tile i+1 runs under tile i's accumulator loads; 1: after them
  unsigned long long* dbg;   // debugging aid (RAWFORMER_B200_LNCONV_DBG=1): cycles per phase of compute thread 0, per CTA
};

template <int MODE, int C>
struct LcCfg {
  static constexpr int N = MODE == LC_QKV ? 3 * C : ((MODE == LC_CONV || MODE == LC_CAT) ? C : 64);   // conv output channels per launch
  static constexpr int CIN = MODE == LC_CAT ? 2 * C : C;         // conv input channels (CAT: two C-channel patches)
  static constexpr bool LN = MODE != LC_CONV && MODE != LC_CAT;                    // normalise the patch while re-laying it
  static constexpr int GU = (MODE == LC_V || MODE == LC_CONV || MODE == LC_CAT) ? 0 : (MODE == LC_QKV ? 2 * C / 8 : N / 8);   // 8-channel units that go to the g tile
  static constexpr bool SECOND = MODE != LC_V && MODE != LC_CONV && MODE != LC_CAT;                   // a second contraction follows (pointwise2 / Gram)
  static constexpr bool RES_RAW = C == 32;                       // FFN residual from the patch (else from global memory)
  static constexpr int NCH = CIN / 8;                            // 16-byte units per patch pixel
  static constexpr int LBO_PX = 184;                             // chunk pitch in pixels (a multiple of 128 bytes)
  static constexpr uint32_t LBO = LBO_PX * 16;
  static constexpr uint32_t RAW_SRC = (LC_NPIX * C * 2 + 1023) & ~1023u;   // one source's patch
  static constexpr uint32_t RAW_BYTES = LC_NPIX * CIN * 2;
  static constexpr uint32_t RAW_STRIDE = MODE == LC_CAT ? 2 * RAW_SRC : RAW_SRC;   // (the patches land 64 / 128-byte swizzled)
  static constexpr uint32_t T_STRIDE = (NCH * LBO + 1023) & ~1023u;
  static constexpr uint32_t WTAP = N * CIN * 2;                  // bytes of one tap's [N][CIN] weights
  static constexpr uint32_t W_BYTES = 9 * WTAP;
  static constexpr uint32_t W2_BYTES = MODE == LC_FFN ? C * 128 : 0;
  static constexpr uint32_t G_BYTES = 16384;                     // 128 pixels x 128 B
  static constexpr int ACC1_STRIDE = N <= 64 ? 64 : 128;         // two accumulators of the first contraction
  static constexpr int ACC3_COL = 2 * ACC1_STRIDE;               // FFN: the [128 x C] pointwise2 accumulator; QKV: the Gram
  static constexpr int TMEM_COLS = N <= 64 ? 256 : 512;
  static constexpr int UPW = N / 8 / 4;                          // 8-column units per compute warp in the first epilogue
  static constexpr uint32_t BT_BYTES = 9 * N * 4;
  static constexpr size_t smem(int nr) {
    return 1024 + W_BYTES + W2_BYTES + G_BYTES + 2048 + LC_NT * T_STRIDE + (size_t)nr * (RAW_STRIDE + LC_ST_STRIDE) +
           BT_BYTES + C * 4 + 256 + 64;
  }
};

__device__ __forceinline__ bool lc_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void lc_warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0 && !lc_test_wait(bar, parity)) mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ float4 lc_lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lc_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lc_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lc_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32-byte global accesses (sm_100: 256-bit vector ld / st): a thread's two adjacent 16-byte units of a pixel as ONE full
// 32-byte sector per lane instead of two half-sector instructions
__device__ __forceinline__ void lc_ldg256(const void* ptr, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(ptr));
}
__device__ __forceinline__ void lc_stg256(void* ptr, const uint32_t (&a)[4], const uint32_t (&b)[4]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]),
               "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3])
               : "memory");
}
// MN-major SWIZZLE_128B operand (the Gram reads the g tile with pixels as the contraction axis), see rf_qk_gram.cu
__device__ __forceinline__ uint64_t lc_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ bool lc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// tcgen05.mma with the two 64-bit shared-memory descriptors given as 32-bit halves (the high halves are constants, the low
// halves base + compile-time offset: one 32-bit add per operand instead of a 64-bit carry chain)
__device__ __forceinline__ void lc_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int MODE, int C, bool DBG>
__global__ void __launch_bounds__(LC_THREADS, 1)
k_lnconv(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapS,
         const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapW2,
         const __grid_constant__ CUtensorMap mapM, const LcP p) {
  using K = LcCfg<MODE, C>;
  constexpr int N = K::N;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;
  const uint32_t sW2 = sW + K::W_BYTES;
  const uint32_t sG = sW2 + K::W2_BYTES;
  const uint32_t sZ = sG + K::G_BYTES;                          // 2 KB of zeros: rows 64..127 of the Gram operand
  const uint32_t sT = sZ + 2048;
  const uint32_t sRaw = sT + LC_NT * K::T_STRIDE;
  const uint32_t sSt = sRaw + (uint32_t)p.nr * K::RAW_STRIDE;
  const uint32_t sBT = sSt + (uint32_t)p.nr * LC_ST_STRIDE;
  const uint32_t sB2 = sBT + K::BT_BYTES;
  const uint32_t bars = (sB2 + C * 4 + 7u) & ~7u;
  auto raw_full = [&](int i) { return bars + 8u * i; };
  auto raw_free = [&](int i) { return bars + 8u * (LC_MAXR + i); };
  auto t_full = [&](int i) { return bars + 8u * (2 * LC_MAXR + i); };
  const uint32_t mma1_done0 = bars + 8u * (2 * LC_MAXR + LC_NT), mma2_done = mma1_done0 + 16, drained0 = mma1_done0 + 24,
                 g_full = mma1_done0 + 40, w_full = mma1_done0 + 48, tmem_slot = mma1_done0 + 56;
  auto mma1_done = [&](int i) { return mma1_done0 + 8u * i; };
  auto drained = [&](int i) { return drained0 + 8u * i; };
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- project_out fused in front of the FFN (p.proj; C = 32): per tile the v patch (its own ring) times the image's folded
  //      attention / project_out weights Mw on the tensor cores -> x1acc (tensor memory, two 128-row groups over the 180 patch
  //      pixels).  Part 1 (all compute warps, three patches ahead): x1acc is read with the tile's other accumulators while the
  //      tensor pipe is empty; x1 = x + x1acc + bias, rounded to bf16, goes over x (the block input) in the raw ring -- it is
  //      the residual of the second epilogue.  Part 2 (two patches ahead) is the ordinary re-layout with norm2's statistics
  //      computed from the pixel's own bf16 values.  Measured: 363 us against 224 (FFN) + 127 (project_out GEMM) per stage-0
  //      block launch by launch, 5.65 against 5.71 ms per RawFormer-S frame in the graph (two launches and 3 C-passes less).
  constexpr bool PJ = MODE == LC_FFN && C == 32;
  constexpr int NV = 3, X1_COL = 192;
  const uint32_t pj_bars = (tmem_slot + 16u + 7u) & ~7u;            // v_full[3], v_free[3], mma0_done, x1_free, x1_ready[2]
  auto v_full = [&](int i) { return pj_bars + 8u * i; };
  auto v_free = [&](int i) { return pj_bars + 8u * (NV + i); };
  const uint32_t mma0_done = pj_bars + 8u * (2 * NV), x1_free = mma0_done + 8, x1_ready = x1_free + 8;
  const uint32_t sPB = (x1_ready + 16u + 15u) & ~15u;                // (x1_ready: two barriers) proj bias [C] floats
  const uint32_t sMw = (sPB + C * 4 + 1023u) & ~1023u;               // [C rows][C] bf16, K-major 64-byte swizzled
  const uint32_t sV = sMw + 2048u;                                   // NV v patches (+ 4 KB: the second row group reads 256 rows)

  if (tid == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapS);
    tma_prefetch_desc(&mapW);
    if (MODE == LC_FFN || MODE == LC_CAT) tma_prefetch_desc(&mapW2);
    // barriers the compute warps arrive on count one arrival per warp: no CTA-wide barrier in the tile loop, the warps
    // drift apart by as much as the buffer depths allow
    for (int i = 0; i < LC_MAXR; ++i) {
      mbar_init(raw_full(i), 1);
      mbar_init(raw_free(i), LC_CW);
    }
    for (int i = 0; i < LC_NT; ++i) mbar_init(t_full(i), LC_CW);
    mbar_init(mma1_done(0), 1);
    mbar_init(mma1_done(1), 1);
    mbar_init(mma2_done, 1);
    mbar_init(drained(0), LC_CW);
    mbar_init(drained(1), LC_CW);
    mbar_init(g_full, LC_CW);
    mbar_init(w_full, 1);
    if (PJ && p.proj) {
      for (int i = 0; i < NV; ++i) {
        mbar_init(v_full(i), 1);
        mbar_init(v_free(i), 1);
      }
      mbar_init(mma0_done, 1);
      mbar_init(x1_free, LC_CW);
      mbar_init(x1_ready, LC_CW);
      mbar_init(x1_ready + 8, LC_CW);
    }
    fence_barrier_init();
  }
  if (MODE == LC_QKV || MODE == LC_QK) {
    for (uint32_t o = (uint32_t)tid * 16u; o < 2048u; o += LC_THREADS * 16u) lc_sts128(sZ + o, 0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)K::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  // What does not depend on the kernel before -- the packed weights (not CAT's per-image ones, nor Mw) and the bias tables -- is
  // fetched BEFORE the dependency wait: under the previous kernel's tail (programmatic dependent launch)
  auto load_weights = [&]() {
    mbar_expect_tx(w_full, K::W_BYTES + K::W2_BYTES + ((PJ && p.proj) ? (uint32_t)(C * C * 2) : 0u));
    for (int tap = 0; tap < 9; ++tap) {
      tma_load_3d(sW + (uint32_t)tap * K::WTAP, &mapW, w_full, 0, tap, p.t0);
      if (p.tn < N) tma_load_3d(sW + (uint32_t)tap * K::WTAP + (uint32_t)p.tn * K::CIN * 2, &mapW, w_full, 0, tap, p.t1);
    }
    if (MODE == LC_FFN) tma_load_3d(sW2, &mapW2, w_full, p.t0, 0, 0);     // pointwise2 columns of these hidden channels
  };
  if (MODE != LC_CAT && p.wpre && tid == 0 && (int)blockIdx.x < p.total_tiles) load_weights();
  {
    float* bt = reinterpret_cast<float*>(smem_raw + (sBT - smem_u32(smem_raw)));
    for (int i = tid; i < 9 * N; i += LC_THREADS) {
      const int idx = i / N, j = i - idx * N;
      bt[i] = __ldg(p.btab + idx * p.tab_stride + (j < p.tn ? p.t0 + j : p.t1 + j - p.tn));
    }
    if (MODE == LC_FFN) {
      float* b2 = reinterpret_cast<float*>(smem_raw + (sB2 - smem_u32(smem_raw)));
      for (int i = tid; i < C; i += LC_THREADS) b2[i] = p.b2 ? __ldg(p.b2 + i) : 0.f;
      if (PJ && p.proj) {
        float* pb = reinterpret_cast<float*>(smem_raw + (sPB - smem_u32(smem_raw)));
        for (int i = tid; i < C; i += LC_THREADS) pb[i] = p.proj_b ? __ldg(p.proj_b + i) : 0.f;
      }
    }
  }
  pdl_wait();
  __syncthreads();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n = first < p.total_tiles ? (p.total_tiles - first + stride - 1) / stride : 0;
  // tile coordinates of this CTA's tiles without per-tile divisions: the walk advances by `stride` tiles
  struct TileIter {
    int tx, ty, b;
  };
  const int step_x = stride % p.tiles_x, step_y = stride / p.tiles_x;
  auto tile_first = [&]() {
    TileIter t;
    t.tx = first % p.tiles_x;
    const int r = first / p.tiles_x;
    t.ty = r % p.tiles_y;
    t.b = r / p.tiles_y;
    return t;
  };
  auto tile_next = [&](TileIter& t) {
    t.tx += step_x;
    if (t.tx >= p.tiles_x) { t.tx -= p.tiles_x; ++t.ty; }
    t.ty += step_y;
    while (t.ty >= p.tiles_y) { t.ty -= p.tiles_y; ++t.b; }
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && n > 0) {
      if (MODE == LC_CAT || !p.wpre) load_weights();
      if (PJ && p.proj) tma_load_3d(sMw, &mapM, w_full, 0, 0, 0);
      int rb = 0, vb = 0;
      uint32_t rph = 0, vph = 0;
      TileIter tl = tile_first();
      for (int i = 0; i < n; ++i) {
        if (i >= p.nr) mbar_wait(raw_free(rb), rph ^ 1u);
        const int px0 = tl.tx * 8, py0 = tl.ty * 16, b = tl.b;
        tile_next(tl);
        const uint32_t rfull = raw_full(rb);      // (p.proj: the v patch of the tile lands on the same barrier)
        mbar_expect_tx(rfull, K::RAW_BYTES + ((K::LN && !p.own_stats) ? LC_ST_BYTES : 0u) + ((PJ && p.proj) ? K::RAW_BYTES : 0u));
        tma_load_4d(sRaw + (uint32_t)rb * K::RAW_STRIDE, &mapX, raw_full(rb), 0, px0 - 1, py0 - 1, b);
        if (MODE == LC_CAT) tma_load_4d(sRaw + (uint32_t)rb * K::RAW_STRIDE + K::RAW_SRC, &mapW2, raw_full(rb), 0, px0 - 1, py0 - 1, b);
        if (K::LN && !p.own_stats) tma_load_3d(sSt + (uint32_t)rb * LC_ST_STRIDE, &mapS, raw_full(rb), 2 * (px0 - 2), py0 - 1, b);
        if (++rb == p.nr) { rb = 0; rph ^= 1u; }
        if (PJ && p.proj) {                  // (mapS is the v tensor's patch map here)
          if (i >= NV) mbar_wait(v_free(vb), vph ^ 1u);
          tma_load_4d(sV + (uint32_t)vb * K::RAW_SRC, &mapS, rfull, 0, px0 - 1, py0 - 1, b);
          if (++vb == NV) { vb = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // A tcgen05.ld queues behind the MMAs in flight (measured: ~900 cycles behind an 18-MMA batch), so the accumulators
    // are read only while the tensor pipe is empty: the first contraction of tile i+1 is issued when every compute warp
    // has drained tile i's accumulators into registers (`drained`), the second contraction of tile i when its g tile is
    // complete; the compute warps wait for both before the next tile's loads.  One accumulator of each kind suffices.
    // The whole warp walks the loop and ONE elected lane issues (elect.sync): under a plain `lane == 0` branch the
    // compiler wraps every tcgen05.mma in a loop over the active lanes (vector -> uniform register moves), ~20 instructions
    // per MMA on a thread that shares its scheduler with four busy compute warps: the issue loop alone took 1700-2400
    // cycles per 18-MMA tile.
    if (n > 0) {
      const bool leader = lc_elect_one();
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t idesc1 = make_idesc_m128(N);
      // A: no-swizzle K-major, LBO = chunk pitch, SBO = patch row pitch (10 pixels x 16 B); B: the resident weights
      const uint32_t a_hi = (160u >> 4) | (1u << 14), a_lo0 = (K::LBO >> 4) << 16;
      const uint64_t wdesc0 = make_kmajor_desc(sW, K::CIN);
      const uint32_t b_hi = (uint32_t)(wdesc0 >> 32), b_lo0 = (uint32_t)wdesc0;
      // project_out of patch j (p.proj): rows = patch pixels 0..255 of the v patch as it landed (64-byte swizzled K-major),
      // two row groups, N = C, K = C -> x1acc
      // Waits are what the tile loop pays for (150-200 cycles each, on the chain): the patches of tile j (x and v) land on ONE
      // barrier; x1acc is free once the compute warps have arrived on `drained` of the tile whose phase A read it -- which the
      // first contraction issued just before has waited for (schedule 1) -- so only the first patches wait for x1_free
      int v_b = 0, m0_rb = 0;
      uint32_t m0_rph = 0;
      auto mma0 = [&](int j) {
        if (!(PJ && p.proj)) return;
        mbar_wait(raw_full(m0_rb), m0_rph);
        if (++m0_rb == p.nr) { m0_rb = 0; m0_rph ^= 1u; }
        if (j >= 1 && j <= 3) mbar_wait(x1_free, (uint32_t)((j - 1) & 1));     // part 1 of patch j-1 (before the tile loop) has read x1acc
        tc_fence_after();
        if (leader) {
          const uint64_t ad = make_kmajor_desc(sV + (uint32_t)v_b * K::RAW_SRC, C), bd = make_kmajor_desc(sMw, C);
          const uint32_t idesc0 = make_idesc_m128(C);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
#pragma unroll
            for (int kk = 0; kk < C / 16; ++kk)
              lc_umma(tmem_base + (uint32_t)(X1_COL + g * C), (uint32_t)ad + (uint32_t)(g * ((128 * C * 2) >> 4)) + 2u * kk,
                      (uint32_t)(ad >> 32), (uint32_t)bd + 2u * kk, (uint32_t)(bd >> 32), idesc0, kk ? 1u : 0u);
          }
          umma_commit(mma0_done);
          umma_commit(v_free(v_b));
        }
        __syncwarp();
        if (++v_b == NV) v_b = 0;
      };
      int tb = 0;
      uint32_t tph = 0;
      long long t_issue = 0;
      auto mma1 = [&](int i) {
        const int s = i & 1;
        mbar_wait(t_full(tb), tph);          // (implied by the wait below in schedule 1; dropping it measured neutral: kept)
        // tile i-2's accumulator (the one this batch overwrites) is in registers; schedule 1: tile i-1's too -- which implies
        // the former (every warp arrives tile by tile), so ONE wait either way
        if (p.sched == 1 && i >= 1) mbar_wait(drained(s ^ 1), (uint32_t)(((i - 1) >> 1) & 1));
        else if (i >= 2) mbar_wait(drained(s), (uint32_t)(((i >> 1) + 1) & 1));
        tc_fence_after();
        if (DBG) t_issue = clock64();
        if (leader) {
          const uint32_t a_lo = a_lo0 | (((sT + (uint32_t)tb * K::T_STRIDE) & 0x3FFFF) >> 4);
          const uint32_t d_tmem = tmem_base + (uint32_t)(s * K::ACC1_STRIDE);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int kk = 0; kk < K::CIN / 16; ++kk)
              lc_umma(d_tmem, a_lo + (uint32_t)(((tap / 3) * 10 + (tap % 3)) + 2 * kk * K::LBO_PX), a_hi,
                      b_lo0 + (uint32_t)tap * (K::WTAP >> 4) + 2u * (uint32_t)kk, b_hi, idesc1, (tap | kk) ? 1u : 0u);
          }
          umma_commit(mma1_done(s));
        }
        __syncwarp();
        if (++tb == LC_NT) { tb = 0; tph ^= 1u; }
      };
      mma0(0);
      if (n > 1) mma0(1);
      if (n > 2) mma0(2);
      if (DBG) {
        mma1(0);
        const long long t1 = clock64();
        mbar_wait(mma1_done(0), 0);
        if (leader) {
          p.dbg[(gridDim.x + blockIdx.x) * 8 + 0] = (unsigned long long)(clock64() - t_issue);
          p.dbg[(gridDim.x + blockIdx.x) * 8 + 2] = (unsigned long long)(t1 - t_issue);
        }
      } else {
        mma1(0);
      }
      if (n > 3) mma0(3);
      if (p.sched == 0 && n > 1) mma1(1);
      const uint32_t idesc3 = make_idesc_m128(C);
      // Gram: kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128 (rows 64.. are zeros), N = 64 (32 q | 32 k)
      const uint32_t idescg = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t gdesc = make_sw128_desc(sG), w2desc = make_sw128_desc(sW2);
      for (int i = 0; i < n; ++i) {
        if (p.sched == 1 && i + 1 < n) {
          if (DBG && i == 8) {
            mma1(i + 1);
            mbar_wait(mma1_done((i + 1) & 1), (uint32_t)(((i + 1) >> 1) & 1));
            if (leader) p.dbg[(gridDim.x + blockIdx.x) * 8 + 1] = (unsigned long long)(clock64() - t_issue);
          } else {
            mma1(i + 1);
          }
        }
        if (i + 4 < n) mma0(i + 4);          // (under the compute warps' first epilogue; x1acc was read in their phase A)
        if (K::SECOND) {
          mbar_wait(g_full, (uint32_t)(i & 1));
          tc_fence_after();
        }
        if (K::SECOND && leader) {
          if (MODE == LC_FFN) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              lc_umma(tmem_base + (uint32_t)K::ACC3_COL, (uint32_t)gdesc + 2u * k, (uint32_t)(gdesc >> 32), (uint32_t)w2desc + 2u * k,
                      (uint32_t)(w2desc >> 32), idesc3, k ? 1u : 0u);
          } else {
            // MN-major SWIZZLE_128B operand, 16 pixels (16 rows of 128 B) per k-step; LBO = distance to the zero block
            const uint32_t g_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t a0 = sG + (uint32_t)k * 2048u;
              const uint32_t lo = ((a0 & 0x3FFFF) >> 4) | (((sZ - a0) >> 4) << 16);
              lc_umma(tmem_base + (uint32_t)K::ACC3_COL, lo, g_hi, lo, g_hi, idescg, (i | k) ? 1u : 0u);
            }
          }
          umma_commit(mma2_done);
        }
        __syncwarp();
        if (p.sched == 0 && i + 2 < n) mma1(i + 2);
      }
    }
  } else {
    // ================= compute warps =================
    const int ctid = tid - 64;
    const int q = warp & 3, wi = (warp - 2) >> 2;        // TMEM lane quadrant; index among the quadrant's four warps
    const int r = q * 32 + lane;                         // accumulator row = tile pixel
    const int ty = r >> 3, tx = r & 7;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    // second epilogue (FFN): C/32 adjacent 8-column units of acc3 per warp; at C = 64 a pixel's piece is one full 32-byte
    // sector (256-bit global accesses).  (C = 32 with two units on two of a quadrant's four warps: slower, 252 vs 230 us.)
    constexpr int UB = C / 32;
    constexpr bool eb_on = true;
    long long tph[DBG ? 8 : 1] = {0}, tlast = DBG ? clock64() : 0;
    auto mark = [&](int k) {
      if (DBG && ctid == 0) {
        const long long now = clock64();
        tph[DBG ? k : 0] += now - tlast;
        tlast = now;
      }
    };
    auto warp_arrive = [&](uint32_t bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };

    // ---- normalise + re-lay patch j: [pixel][C] -> [C/8][pixel][16 B] ----
    // (its buffer was last read by the MMAs of tile j-3, whose completion this warp observed before that tile's loads)
    // One thread per patch pixel (statistics, rsqrt and index arithmetic once per pixel instead of once per 16 bytes: the
    // tile loop is bound by instruction issue); the 180 pixels go to a window of the 512 compute threads that rotates from
    // tile to tile so that every warp does the same work on average.  TMA delivers the patch 64-byte swizzled (16-byte unit
    // u of patch pixel q sits at unit u ^ ((q >> 1) & 3)), which makes these 64-byte-strided loads -- and the residual
    // loads of the second epilogue -- conflict-free.
    int rl_rb = 0, rl_tb = 0, rl_rot = 0;
    uint32_t rl_rph = 0;
    // ---- project_out in front (p.proj), part 1: x1acc is loaded in phase A of the tile loop (a tcgen05.ld queues behind the
    //      MMAs in flight: 387 us with the load in phase C), the arithmetic runs in phase C.  ALL compute warps take part (on six
    //      warps per patch, a whole pixel per thread: 463 us): warp (q, wi) owns channels 8 wi .. 8 wi + 7 of the patch pixels
    //      32 q + lane (row group 0) and, q < 2, 128 + 32 q + lane (row group 1)
    TileIter trl = tile_first();
    int rl_j = 0, p1_rb = 0, p1_j = 0;
    uint32_t p1_rph = 0;
    auto proj_load = [&](uint32_t (&a)[2][8]) {
      tmem_ld8(tq + (uint32_t)(X1_COL + wi * 8), a[0]);
      if (q < 2) tmem_ld8(tq + (uint32_t)(X1_COL + C + wi * 8), a[1]);
    };
    auto proj_part1 = [&](const uint32_t (&a)[2][8]) {
      const float4 b0 = lc_lds128f(sPB + (uint32_t)wi * 32u), b1 = lc_lds128f(sPB + (uint32_t)wi * 32u + 16u);
      const float pb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int px = g * 128 + q * 32 + lane;
        if ((g == 0 || q < 2) && px < LC_NPIX) {
          const uint32_t ua = sRaw + (uint32_t)p1_rb * K::RAW_STRIDE + (uint32_t)px * (C * 2) + (uint32_t)((wi ^ ((px >> 1) & 3)) * 16);
          const uint4 xv = lc_lds128(ua);
          const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f0 = __uint_as_float(a[g][2 * e]) + pb[2 * e] + __uint_as_float(xw[e] << 16);
            const float f1 = __uint_as_float(a[g][2 * e + 1]) + pb[2 * e + 1] + __uint_as_float(xw[e] & 0xffff0000u);
            o[e] = lc_pack(f0, f1);                   // x1 as the project_out GEMM would have stored it
          }
          lc_sts128(ua, o[0], o[1], o[2], o[3]);
        }
      }
      // (two barriers by patch parity: part 1 runs a step ahead of the re-layout that waits for it; with one barrier a warp
      // still waiting for patch j could see the barrier already in the phase of patch j + 2)
      warp_arrive(x1_ready + 8u * (uint32_t)(p1_j & 1));
      ++p1_j;
      if (++p1_rb == p.nr) { p1_rb = 0; p1_rph ^= 1u; }
    };
    // (p.proj: the patch is x1, complete once every warp's part 1 has arrived; norm2's statistics from its bf16 values)
    auto relayout = [&]() {
      int py0 = 0, px0 = 0;
      if (PJ && p.proj) {
        lc_warp_wait(x1_ready + 8u * (uint32_t)(rl_j & 1), (uint32_t)((rl_j >> 1) & 1), lane);
        ++rl_j;
        py0 = trl.ty * 16 - 1; px0 = trl.tx * 8 - 1;
        tile_next(trl);
      } else {
        lc_warp_wait(raw_full(rl_rb), rl_rph, lane);
      }
      mark(5);
      const int px = (ctid - rl_rot) & (LC_CT - 1);
      if (px < LC_NPIX) {
        const uint32_t src = sRaw + (uint32_t)rl_rb * K::RAW_STRIDE + (uint32_t)px * (C * 2);
        const uint32_t dst = sT + (uint32_t)rl_tb * K::T_STRIDE + (uint32_t)px * 16u;
        const int py = px / 10, pxx = px - py * 10;
        uint4 v[K::NCH];
#pragma unroll
        for (int k = 0; k < K::NCH; ++k)     // (CAT: units C/8.. come from the second source's patch)
          v[k] = lc_lds128(src + (k >= C / 8 ? K::RAW_SRC : 0u) + (uint32_t)(((k % (C / 8)) ^ (C == 32 ? (px >> 1) & 3 : px & 7)) * 16));
        float sum = 0.f, ssq = 0.f;
        if (K::LN && C == 32 && p.own_stats) {
          // no statistics tensor (an encoder block's input: nobody's epilogue produced them): from the pixel's own values
          float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < K::NCH; ++k) {
            const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
              s2 = __fadd2_rn(s2, a);
              q2 = __ffma2_rn(a, a, q2);
            }
          }
          sum = s2.x + s2.y;
          ssq = q2.x + q2.y;
        } else if (K::LN) {
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sum), "=f"(ssq)
                       : "r"(sSt + (uint32_t)rl_rb * LC_ST_STRIDE + (uint32_t)(py * 12 + pxx + 1) * 8u));
        }
        const float mu = sum * p.invC;
        const float rs = rsqrtf(fmaxf(ssq * p.invC - mu * mu, 0.f) + p.eps);
        const float nm = -mu * rs;
        const float2 r2 = make_float2(rs, rs), n2 = make_float2(nm, nm);
#pragma unroll
        for (int k = 0; k < K::NCH; ++k) {
          const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (K::LN) {
              const float2 a = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
              const float2 f = __ffma2_rn(a, r2, n2);
              o[e] = lc_pack(f.x, f.y);
              if (PJ && p.proj) {                     // outside the image the conv sees zeros (x1 there is the bias)
                const int y = py0 + py, x = px0 + pxx;
                if (!(y >= 0 && y < p.H && x >= 0 && x < p.W)) o[e] = 0u;
              }
            } else {
              o[e] = w[e];
            }
          }
          lc_sts128(dst + (uint32_t)k * K::LBO, o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async();                                 // generic stores -> tensor-core (async proxy) reads
      warp_arrive(t_full(rl_tb));
      if (!(MODE == LC_FFN && K::RES_RAW) && lane == 0) mbar_arrive(raw_free(rl_rb));   // (else it stays for the residual add)
      if (++rl_rb == p.nr) { rl_rb = 0; rl_rph ^= 1u; }
      if (++rl_tb == LC_NT) rl_tb = 0;
      rl_rot = (rl_rot + 192) & (LC_CT - 1);
      mark(6);
    };

    // ---- FFN: second epilogue of the tile before: acc3 (in v3) + bias + residual -> bf16 -> global ----
    TileIter ta = tile_first(), tb2 = ta;
    int eb_rb = 0;
    // (C = 64) the residual of the tile tb2 points at, from global memory: requested one tile before its use (a load issued
    // in the same step would expose its full latency in every warp)
    uint4 rres[UB];
    auto res_prefetch = [&]() {
      const int y = tb2.ty * 16 + ty, x = tb2.tx * 8 + tx, b = tb2.b;
#pragma unroll
      for (int t = 0; t < UB; ++t) rres[t] = make_uint4(0u, 0u, 0u, 0u);
      if (y < p.H && x < p.W) {
        const bf16* src = p.resid + ((((i64)b * p.H + y) * p.W + x) * C + wi * UB * 8);
        if (UB == 2) lc_ldg256(src, rres[0], rres[UB - 1]);
        else rres[0] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    };
    auto epi_b = [&](uint32_t (&v3)[UB][8]) {
      const int y = tb2.ty * 16 + ty, x = tb2.tx * 8 + tx, b = tb2.b;
      tile_next(tb2);
      uint32_t ob[UB][4];
#pragma unroll
      for (int t = 0; t < UB; ++t) {
        if (!eb_on) break;
        const int u = wi * UB + t;
        const int pr = (ty + 1) * 10 + tx + 1;
        uint4 res = rres[t];
        if (K::RES_RAW) res = lc_lds128(sRaw + (uint32_t)eb_rb * K::RAW_STRIDE + (uint32_t)(pr * C * 2 + ((u ^ ((pr >> 1) & 3)) << 4)));
        const float4 b0 = lc_lds128f(sB2 + (uint32_t)u * 32u), b1 = lc_lds128f(sB2 + (uint32_t)u * 32u + 16u);
        const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float f0 = __uint_as_float(v3[t][2 * e]) + bias[2 * e] + __uint_as_float(rw[e] << 16);
          const float f1 = __uint_as_float(v3[t][2 * e + 1]) + bias[2 * e + 1] + __uint_as_float(rw[e] & 0xffff0000u);
          ob[t][e] = lc_pack(f0, f1);
        }
      }
      if (y < p.H && x < p.W) {
        bf16* dst = p.out + ((((i64)b * p.H + y) * p.W + x) * C + wi * UB * 8);
        if (UB == 2) lc_stg256(dst, ob[0], ob[UB - 1]);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(ob[0][0], ob[0][1], ob[0][2], ob[0][3]);
      }
      if (K::RES_RAW) {
        warp_arrive(raw_free(eb_rb));
        if (++eb_rb == p.nr) eb_rb = 0;
      }
    };

    // (p.proj) part 1 runs THREE patches ahead and the re-layout two: a warp's re-layout never waits for the other warps'
    // part 1 of the same step (that would take away the slack that lets the warps drift apart)
    for (int j = 0; j < 3 && j < n; ++j) {
      if (PJ && p.proj) {
        lc_warp_wait(raw_full(p1_rb), p1_rph, lane);
        lc_warp_wait(mma0_done, (uint32_t)(j & 1), lane);
        tc_fence_after();
        uint32_t a[2][8];
        proj_load(a);
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive(x1_free);
        proj_part1(a);
      }
      if (j < 2) relayout();
    }
    for (int i = 0; i < n; ++i) {
      // ---- A: both contractions that feed this step are complete -> accumulators into registers ----
      mark(0);
      // (with a second contraction: its commit for tile i-1 came after this tile's first contraction was issued, in both
      // schedules, and a commit covers everything issued before it: one wait instead of two)
      if (!(K::SECOND && i >= 1)) lc_warp_wait(mma1_done(i & 1), (uint32_t)((i >> 1) & 1), lane);
      mark(1);
      if (K::SECOND && i >= 1) lc_warp_wait(mma2_done, (uint32_t)((i - 1) & 1), lane);
      // (p.proj) x1acc of the patch two tiles ahead is read here too, with the tensor pipe empty
      // Its MMAs were issued before the second contraction of tile i-1, whose completion was just awaited (a commit covers
      // everything issued before it); the issuer had waited for the patch's barrier: no wait of its own here, except for i = 0
      const bool pj_on = PJ && p.proj && i + 3 < n;
      if (pj_on && i == 0) lc_warp_wait(mma0_done, 1u, lane);
      tc_fence_after();
      mark(2);
      uint32_t v[K::UPW][8], v3[UB][8], xa[2][8];
#pragma unroll
      for (int t = 0; t < K::UPW; ++t) tmem_ld8(tq + (uint32_t)((i & 1) * K::ACC1_STRIDE + (wi * K::UPW + t) * 8), v[t]);
      if (MODE == LC_FFN && i >= 1 && eb_on) {
#pragma unroll
        for (int t = 0; t < UB; ++t) tmem_ld8(tq + (uint32_t)(K::ACC3_COL + (wi * UB + t) * 8), v3[t]);
      }
      if (PJ) {
        if (pj_on) proj_load(xa);
      }
      tmem_ld_wait();
      tc_fence_before();
      warp_arrive(drained(i & 1));
      mark(3);
      // ---- B: epilogues on registers ----
      if (MODE == LC_FFN && i >= 1) epi_b(v3);
      if (MODE == LC_FFN && !K::RES_RAW) res_prefetch();     // tile i's residual: a whole tile ahead of its use
      mark(7);
      const int y = ta.ty * 16 + ty, x = ta.tx * 8 + tx, b = ta.b;
      tile_next(ta);
      const int rs_ = y == 0 ? 0 : (y == p.H - 1 ? 2 : 1), cs_ = x == 0 ? 0 : (x == p.W - 1 ? 2 : 1);
      const uint32_t bt = sBT + (uint32_t)((rs_ * 3 + cs_) * N) * 4u;
      const bool inside = y < p.H && x < p.W;
      const bool counts = inside && y >= p.ylo && y < p.yhi;
      const uint32_t grow = sG + (uint32_t)r * 128u;
      uint32_t ov[K::UPW][4];
#pragma unroll
      for (int t = 0; t < K::UPW; ++t) {
        const int u = wi * K::UPW + t;
        const float4 b0 = lc_lds128f(bt + (uint32_t)u * 32u), b1 = lc_lds128f(bt + (uint32_t)u * 32u + 16u);
        const float2 h[4] = {__fadd2_rn(make_float2(__uint_as_float(v[t][0]), __uint_as_float(v[t][1])), make_float2(b0.x, b0.y)),
                             __fadd2_rn(make_float2(__uint_as_float(v[t][2]), __uint_as_float(v[t][3])), make_float2(b0.z, b0.w)),
                             __fadd2_rn(make_float2(__uint_as_float(v[t][4]), __uint_as_float(v[t][5])), make_float2(b1.x, b1.y)),
                             __fadd2_rn(make_float2(__uint_as_float(v[t][6]), __uint_as_float(v[t][7])), make_float2(b1.z, b1.w))};
        uint32_t o[4];
        if (MODE == LC_FFN) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 g = gelu_tanh2(h[e]);
            o[e] = lc_pack(g.x, g.y);
          }
          lc_sts128(grow + (((uint32_t)u ^ (uint32_t)(r & 7)) << 4), o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (MODE == LC_CONV || MODE == LC_CAT) o[e] = lc_pack(lrelu_f(h[e].x), lrelu_f(h[e].y));
            else o[e] = lc_pack(h[e].x, h[e].y);
          }
          if (u < K::GU) {                                  // q|k: pixels that do not count contribute exact zeros
            if (!counts) o[0] = o[1] = o[2] = o[3] = 0u;
            lc_sts128(grow + (((uint32_t)u ^ (uint32_t)(r & 7)) << 4), o[0], o[1], o[2], o[3]);
          } else if ((MODE == LC_V || MODE == LC_CONV || MODE == LC_CAT) && K::UPW == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) ov[t][e] = o[e];
          } else if (inside) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((((i64)b * p.H + y) * p.W + x) * C + (u - K::GU) * 8));
            *dst = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      if ((MODE == LC_V || MODE == LC_CONV || MODE == LC_CAT) && K::UPW == 2 && inside)   // this warp's two adjacent units: one 32-byte store per pixel
        lc_stg256(p.out + ((((i64)b * p.H + y) * p.W + x) * C + wi * K::UPW * 8), ov[0], ov[K::UPW - 1]);
      if (K::SECOND) {
        fence_proxy_async();
        warp_arrive(g_full);
      }
      mark(4);
      // ---- C: the patch two tiles ahead (p.proj: first part 1 of the patch three tiles ahead, off the A -> B -> pointwise2
      //      chain that bounds the tile loop) ----
      if (PJ) {
        if (pj_on) proj_part1(xa);
      }
      if (i + 2 < n) relayout();
    }
    if (K::SECOND && n > 0) {
      lc_warp_wait(mma2_done, (uint32_t)((n - 1) & 1), lane);
      tc_fence_after();
    }
    if (MODE == LC_FFN && n > 0) {
      uint32_t v3[UB][8];
#pragma unroll
      if (eb_on) {
#pragma unroll
        for (int t = 0; t < UB; ++t) tmem_ld8(tq + (uint32_t)(K::ACC3_COL + (wi * UB + t) * 8), v3[t]);
      }
      tmem_ld_wait();
      epi_b(v3);
    }
    if (DBG && ctid == 0)
      for (int k = 0; k < 8; ++k) p.dbg[blockIdx.x * 8 + k] = (unsigned long long)tph[DBG ? k : 0];

    // ---- QKV / QK read-out: per-head diagonal blocks of q^T k and the squared norms -> this CTA's slot ----
    // accumulator rows / columns 0..31 = q channels ch0.., 32..63 = k channels ch0..; a head has c = C/8 channels
    if ((MODE == LC_QKV || MODE == LC_QK) && wi == 0) {
      constexpr int c = C >> 3, NQ = 32;
      const int row = r;
      float* gp = p.gram_part + (i64)blockIdx.x * C * c;
      float* sp = p.sq_part + (i64)blockIdx.x * 2 * C;
      const int h = row < NQ ? row / c : -1;           // head of this q row (local to the launch)
      for (int cc = 0; cc < 2 * NQ; cc += 16) {
        uint32_t v[16];
        if (n > 0) {
          tmem_ld16(tq + (uint32_t)(K::ACC3_COL + cc), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (row < 2 * NQ) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = cc + j;
            if (col == row) sp[row < NQ ? p.ch0 + row : C + p.ch0 + row - NQ] = __uint_as_float(v[j]);
            const int kc = col - NQ;                   // k channel (local) of this accumulator column
            if (h >= 0 && kc >= h * c && kc < (h + 1) * c) gp[(i64)(p.ch0 + row) * c + (kc - h * c)] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)K::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// pack time: Weff [N][9][K] (T) and the border-state bias table [9][N]; one warp per output channel n
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_lnconv(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ bias, const float* __restrict__ dw, const float* __restrict__ dwb,
              bf16* __restrict__ cw, float* __restrict__ btab, int N, int Kc) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float bb = 0.f;
  for (int k = lane; k < Kc; k += 32) {
    const float w = W[(i64)n * Kc + k];
    bb = fmaf(w, beta[k], bb);
    const float wg = w * gamma[k];
    for (int tap = 0; tap < 9; ++tap) cw[((i64)n * 9 + tap) * Kc + k] = __float2bfloat16_rn(wg * dw[n * 9 + tap]);
  }
  bb = warp_sum(bb) + (bias ? bias[n] : 0.f);
  if (lane < 9) {
    const int rs = lane / 3, cs = lane % 3;
    float s = dwb ? dwb[n] : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const bool valid = !(rs == 0 && ky == 0) && !(rs == 2 && ky == 2) && !(cs == 0 && kx == 0) && !(cs == 2 && kx == 2);
      if (valid) s = fmaf(dw[n * 9 + tap], bb, s);
    }
    btab[lane * N + n] = s;
  }
}

void launch_pack_lnconv(Ctx& ctx, const float* W, const float* gamma, const float* beta, const float* bias, const float* dw,
                        const float* dwb, void* cw, float* btab, int N, int K) {
  if (ctx.dry || !W || !gamma || !beta || !dw || !cw || ctx.dtype != RF_BF16) return;
  ScopedLaunch sl(RF_K_WEIGHT_PACK);
  k_pack_lnconv<<<cdiv(N, 8), 256, 0, ctx.stream>>>(W, gamma, beta, bias, dw, dwb, (bf16*)cw, btab, N, K);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool lnconv_enabled() {
  static int enabled = -1;                // debugging aid: RAWFORMER_B200_NO_LNCONV=1 keeps the separate kernels
  if (enabled < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_LNCONV");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled != 0;
}

bool lnconv_supported(const Ctx& ctx, int C, int H, int W) {
  static int c64 = -1;                  // debugging aid: RAWFORMER_B200_LNCONV_C64=0 keeps C = 64 on the separate kernels
  if (c64 < 0) {
    const char* e = getenv("RAWFORMER_B200_LNCONV_C64");
    c64 = (e && e[0] == '0') ? 0 : 1;
  }
  return lnconv_enabled() && tcgen05_enabled() && ctx.dtype == RF_BF16 && (C == 32 || (C == 64 && c64)) && (W & 1) == 0 && H >= 2 &&
         W >= 2;
}

// one launch: conv channels [t0, t0 + tn) (and [t1, t1 + tn) when tn < N) of the n_tab-channel dense conv cw / btab
struct LcSel {
  int n_tab, t0, t1, tn, ch0;
};
// FFN with project_out fused in front (C = 32, one image): x1 = x + Mw v + pb
struct LcProj {
  const void* v = nullptr;     // [H,W,C]
  const void* Mw = nullptr;    // T [C][C] (this image's softmax(attention) folded into project_out)
  const float* pb = nullptr;   // [C]
};
template <int MODE, int C>
static int lnconv_launch(Ctx& ctx, const void* x, const float* stats, const void* cw, const float* btab, const LcSel& sel,
                         const void* W2, const float* b2, const void* resid, void* out, float* gram_part, float* sq_part, int B,
                         int H, int W, int slot_cap, const LcProj* pj = nullptr) {
  using K = LcCfg<MODE, C>;
  LcP p;
  memset(&p, 0, sizeof(p));
  p.btab = btab; p.b2 = b2; p.resid = (const bf16*)resid; p.out = (bf16*)out; p.gram_part = gram_part; p.sq_part = sq_part;
  p.n_tab = sel.n_tab; p.tab_stride = MODE == LC_CONV ? 0 : sel.n_tab; p.t0 = sel.t0; p.t1 = sel.t1; p.tn = sel.tn; p.ch0 = sel.ch0;
  p.invC = 1.0f / (float)C; p.eps = 1e-5f;
  p.H = H; p.W = W; p.B = B;
  p.ylo = 0; p.yhi = H;
  if ((MODE == LC_QKV || MODE == LC_QK) && ctx.band != nullptr) { p.ylo = ctx.band->ht; p.yhi = ctx.band->ht + ctx.band->rows_in; }
  p.tiles_x = cdiv(W, 8); p.tiles_y = cdiv(H, 16);
  const i64 total = (i64)p.tiles_x * p.tiles_y * B;
  if (total <= 0 || total > 0x7fffffff) return 0;
  p.total_tiles = (int)total;
  p.nr = C == 32 ? (MODE == LC_FFN ? (pj != nullptr ? 6 : 5) : 3) : 2;     // (project_out in front: part 1 runs a patch further ahead)
  {
    // measured (RawFormer-S stage 0): FFN 236 us with schedule 1 / 262 us with 0; QKV 273 / 238 us
    static int sched = -1;
    if (sched < 0) {
      const char* e = getenv("RAWFORMER_B200_LNCONV_SCHED");
      sched = e ? atoi(e) : 2;
    }
    p.sched = sched == 2 ? (MODE == LC_FFN ? 1 : 0) : (sched & 1);
  }
  // (p.proj: + its barriers and bias, the 1 KB-aligned Mw tile, three v patches and the 4 KB the second row group reads past them)
  p.proj = (pj != nullptr && MODE == LC_FFN && C == 32) ? 1 : 0;
  if (p.proj) p.sched = 1;                 // (the x1acc hand-over relies on the order of schedule 1, see the issuer)
  p.proj_b = p.proj ? pj->pb : nullptr;
  const size_t smem = K::smem(p.nr) + (p.proj ? 1024 + 2048 + 3 * (size_t)K::RAW_SRC + 4096 + 256 : 0);
  if (smem > 232448) return 0;
  {
    static int wpre = -1;
    if (wpre < 0) {
      const char* e = getenv("RAWFORMER_B200_LNCONV_WPRE");
      wpre = (e && e[0] == '0') ? 0 : 1;
    }
    p.wpre = wpre;
  }
  p.own_stats = (K::LN && stats == nullptr) ? 1 : 0;
  if (p.own_stats && C != 32) return 0;
  if ((K::LN && ((uintptr_t)stats & 15)) || ((uintptr_t)x & 15)) return 0;
  CUtensorMap mX, mS, mW, mW2, mM;
  {
    const i64 d[4] = {C, W, H, B};
    const i64 s[4] = {1, C, (i64)C * W, (i64)C * W * H};
    const int bx[4] = {C, 10, 18, 1};
    if (!make_map_ex(&mX, x, 4, d, s, bx, 2, C * 2)) return 0;
  }
  {
    // statistics as a [B, H, 2W] fp32 tensor: the box of a tile is its patch pixels' (sum, sumsq) pairs; it starts one
    // pixel left of the patch so that its global start address is 16-byte aligned; zero fill outside the image
    const i64 d[3] = {2 * (i64)W, H, B};
    const i64 s[3] = {1, 2 * (i64)W, (i64)2 * W * H};
    const int bx[3] = {24, 18, 1};
    if (K::LN && !p.own_stats && !make_map_ex(&mS, stats, 3, d, s, bx, 4, 0)) return 0;
    if (!K::LN || p.own_stats) mS = mX;
    if (p.proj) {                          // mapS = the v tensor's patch map, mapM = Mw [C][C]
      const i64 dv[4] = {C, W, H, B};
      const i64 sv[4] = {1, C, (i64)C * W, (i64)C * W * H};
      const int bv[4] = {C, 10, 18, 1};
      if (((uintptr_t)pj->v & 15) || !make_map_ex(&mS, pj->v, 4, dv, sv, bv, 2, C * 2)) return 0;
      const i64 dm[3] = {C, C, 1};
      const i64 sm[3] = {1, C, (i64)C * C};
      const int bm[3] = {C, C, 1};
      if (((uintptr_t)pj->Mw & 15) || !make_map_ex(&mM, pj->Mw, 3, dm, sm, bm, 2, C * 2)) return 0;
    }
  }
  {
    const i64 d[3] = {K::CIN, 9, sel.n_tab};
    const i64 s[3] = {1, K::CIN, (i64)9 * K::CIN};
    const int bx[3] = {K::CIN, 1, sel.tn};
    if (!make_map_ex(&mW, cw, 3, d, s, bx, 2, K::CIN * 2)) return 0;
  }
  if (MODE == LC_FFN) {
    const i64 d[3] = {2 * C, C, 1};
    const i64 s[3] = {1, 2 * C, (i64)2 * C * C};
    const int bx[3] = {64, C, 1};
    if (!make_map_ex(&mW2, W2, 3, d, s, bx, 2, 128)) return 0;
  } else if (MODE == LC_CAT) {           // the second source's patch (W2 = x2 here)
    const i64 d[4] = {C, W, H, B};
    const i64 s[4] = {1, C, (i64)C * W, (i64)C * W * H};
    const int bx[4] = {C, 10, 18, 1};
    if (((uintptr_t)W2 & 15) || !make_map_ex(&mW2, W2, 4, d, s, bx, 2, C * 2)) return 0;
  } else {
    mW2 = mW;
  }
  if (!p.proj) mM = mX;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_lnconv<MODE, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_lnconv<MODE, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return 0;
    attr_set = true;
  }
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if ((MODE == LC_QKV || MODE == LC_QK) && grid > slot_cap) grid = slot_cap;
  static int dbg_on = -1;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on < 0) {
    const char* e = getenv("RAWFORMER_B200_LNCONV_DBG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on && cudaMalloc(&dbg_buf, 8 * 8 * 2048) != cudaSuccess) dbg_on = 0;
  }
  p.dbg = dbg_on ? dbg_buf : nullptr;
  if (dbg_on) {
    launch_pdl(k_lnconv<MODE, C, true>, dim3(grid), dim3(LC_THREADS), smem, ctx.stream, mX, mS, mW, mW2, mM, p);
    static unsigned long long h[8 * 2048];
    cudaStreamSynchronize(ctx.stream);
    cudaMemcpy(h, dbg_buf, sizeof(unsigned long long) * 8 * 2 * grid, cudaMemcpyDeviceToHost);
    {
      double m0 = 0, m1 = 0, m2 = 0;
      for (int i = 0; i < grid; ++i) { m0 += (double)h[(grid + i) * 8] / grid; m1 += (double)h[(grid + i) * 8 + 1] / grid; m2 += (double)h[(grid + i) * 8 + 2] / grid; }
      fprintf(stderr, "[lnconv] 18-MMA batch first issue->complete: first tile %.0f cycles (issue loop alone %.0f), tile 9 (steady state, schedule 1) %.0f cycles\n", m0, m2, m1);
    }
    double a[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) a[j] += (double)h[i * 8 + j] / grid;
    const double tiles = (double)cdiv(p.total_tiles, grid);
    fprintf(stderr, "[lnconv mode %d C=%d %dx%d tiles/CTA %.0f] cycles/tile (warp 2): loop %.0f wait_mma1 %.0f wait_mma2 %.0f tmem_ld %.0f "
            "epilogue 2 (+ residual prefetch) %.0f epilogue 1 %.0f | relayout: wait_raw %.0f work %.0f\n",
            MODE, C, H, W, tiles, a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles, a[7] / tiles, a[4] / tiles, a[5] / tiles,
            a[6] / tiles);
  } else {
    launch_pdl(k_lnconv<MODE, C, false>, dim3(grid), dim3(LC_THREADS), smem, ctx.stream, mX, mS, mW, mW2, mM, p);
  }
  return grid;
}

// out = x + conv_ffn(norm2(x)); stats: [rows] (sum, sumsq) of x's rows; cw / btab from launch_pack_lnconv
bool launch_lnconv_ffn(Ctx& ctx, const void* x, const float* stats, const void* cw, const float* btab, const void* W2,
                       const float* b2, void* out, int B, int H, int W, int C) {
  if (!lnconv_supported(ctx, C, H, W)) return false;
  const double rows = (double)B * H * W;
  if (C == 32) {
    ScopedLaunch sl(RF_K_FFN_FUSED, rows * C * 2.0 * 2.0 + rows * 8.0, rows * (2.0 * 9 * C * 2 * C + 2.0 * 2 * C * C));
    const LcSel sel{2 * C, 0, 0, 2 * C, 0};
    return lnconv_launch<LC_FFN, 32>(ctx, x, stats, cw, btab, sel, W2, b2, nullptr, out, nullptr, nullptr, B, H, W, 0) > 0;
  }
  // C = 64: hidden channels 0..63, then 64..127 on top of the first launch's output (its own residual, in place)
  for (int half = 0; half < 2; ++half) {
    ScopedLaunch sl(RF_K_FFN_FUSED, rows * C * 2.0 * (half ? 3.0 : 2.0) + rows * 8.0, rows * (2.0 * 9 * C * C + 2.0 * C * C));
    const LcSel sel{2 * C, 64 * half, 0, 64, 0};
    if (lnconv_launch<LC_FFN, 64>(ctx, x, stats, cw, btab, sel, W2, half ? nullptr : b2, half ? out : x, out, nullptr, nullptr, B, H,
                                  W, 0) <= 0)
      return false;
  }
  return true;
}

// ONE image: v = third part of qkv_dwconv(qkv(norm1(x))); Gram / squared norms of q, k into per-CTA partial slots.
// Returns the number of slots written (= CTAs), 0 if unsupported.
int launch_lnconv_qkv(Ctx& ctx, const void* x, const float* stats, const void* cw, const float* btab, void* v, float* gram_part,
                      float* sq_part, int H, int W, int C, int slot_cap) {
  if (!lnconv_supported(ctx, C, H, W)) return 0;
  const double rows = (double)H * W;
  if (C == 32) {
    ScopedLaunch sl(RF_K_QKV_FUSED, rows * C * 2.0 * 2.0 + rows * 8.0, rows * (2.0 * 9 * C * 3 * C + 2.0 * (2 * C) * (2 * C)));
    const LcSel sel{3 * C, 0, 0, 3 * C, 0};
    return lnconv_launch<LC_QKV, 32>(ctx, x, stats, cw, btab, sel, nullptr, nullptr, nullptr, v, gram_part, sq_part, 1, H, W,
                                     slot_cap);
  }
  // C = 64: q|k of heads 0-3, q|k of heads 4-7 (the Gram is block-diagonal over the heads), then v
  int ns = 0;
  for (int half = 0; half < 2; ++half) {
    ScopedLaunch sl(RF_K_QKV_FUSED, rows * C * 2.0 + rows * 8.0, rows * (2.0 * 9 * C * 64 + 2.0 * 64 * 64));
    const LcSel sel{3 * C, 32 * half, C + 32 * half, 32, 32 * half};
    const int g = lnconv_launch<LC_QK, 64>(ctx, x, stats, cw, btab, sel, nullptr, nullptr, nullptr, nullptr, gram_part, sq_part, 1, H,
                                           W, slot_cap);
    if (g <= 0 || (half && g != ns)) return 0;
    ns = g;
  }
  {
    ScopedLaunch sl(RF_K_QKV_FUSED, rows * C * 2.0 * 2.0 + rows * 8.0, rows * (2.0 * 9 * C * 64));
    const LcSel sel{3 * C, 2 * C, 0, 64, 0};
    if (lnconv_launch<LC_V, 64>(ctx, x, stats, cw, btab, sel, nullptr, nullptr, nullptr, v, nullptr, nullptr, 1, H, W, 0) <= 0) return 0;
  }
  return ns;
}

// Conv_out of a Conv_Transformer: out = LeakyReLU_0.2(conv3x3(x) + bias), C -> C (FLCA_RF.py:276); w = T [C][9][C]
bool launch_lnconv_conv3(Ctx& ctx, const void* x, const void* w, const float* bias, void* out, int B, int H, int W, int C) {
  if (!lnconv_supported(ctx, C, H, W)) return false;
  static int on = -1;                     // debugging aid: RAWFORMER_B200_LNCONV_CONV=0 keeps Conv_out on k_tc_gemm
  if (on < 0) {
    const char* e = getenv("RAWFORMER_B200_LNCONV_CONV");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on) return false;
  const double rows = (double)B * H * W;
  ScopedLaunch sl(RF_K_CONV3X3_LC, rows * C * 2.0 * 2.0, rows * 2.0 * 9 * C * C);
  const LcSel sel{C, 0, 0, C, 0};
  if (C == 32) return lnconv_launch<LC_CONV, 32>(ctx, x, nullptr, w, bias, sel, nullptr, nullptr, nullptr, out, nullptr, nullptr, B, H, W, 0) > 0;
  return lnconv_launch<LC_CONV, 64>(ctx, x, nullptr, w, bias, sel, nullptr, nullptr, nullptr, out, nullptr, nullptr, B, H, W, 0) > 0;
}

// ---------------------------------------------------------------------------------------------
// channel_reduce folded into Conv_out (C = 32): pack-time product P2[n][tap][k] = sum_m Wout[n][m][tap] * Wred[m][k] (fp32) and
// the border-state bias table bt[idx][n] = bout[n] + sum_{valid taps} sum_m Wout[n][m][tap] * bred[m]; per image the first C
// input channels are scaled by the squeeze-excite vector and the result is rounded to bf16 (k_cat_scale)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_cat(const float* __restrict__ wout, const float* __restrict__ bout, const float* __restrict__ wred,
           const float* __restrict__ bred, float* __restrict__ p2, float* __restrict__ bt, int C) {
  // wout [C][C][3][3] (PyTorch), wred [C][2C]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = C * 9 * 2 * C;
  if (i < total) {
    const int k = i % (2 * C), tap = (i / (2 * C)) % 9, n = i / (18 * C);
    float acc = 0.f;
    for (int m = 0; m < C; ++m) acc = fmaf(wout[((i64)n * C + m) * 9 + tap], wred[(i64)m * 2 * C + k], acc);
    p2[i] = acc;
  }
  if (i < 9 * C) {
    const int n = i % C, idx = i / C, rs = idx / 3, cs = idx % 3;
    float acc = bout ? bout[n] : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const bool valid = !(rs == 0 && ky == 0) && !(rs == 2 && ky == 2) && !(cs == 0 && kx == 0) && !(cs == 2 && kx == 2);
      if (!valid || !bred) continue;
      for (int m = 0; m < C; ++m) acc = fmaf(wout[((i64)n * C + m) * 9 + tap], bred[m], acc);
    }
    bt[i] = acc;
  }
}
void launch_pack_cat(Ctx& ctx, const float* wout, const float* bout, const float* wred, const float* bred, float* p2, float* bt,
                     int C) {
  if (ctx.dry || !wout || !wred || !p2 || !bt) return;
  ScopedLaunch sl(RF_K_WEIGHT_PACK);
  k_pack_cat<<<cdiv(C * 9 * 2 * C, 256), 256, 0, ctx.stream>>>(wout, bout, wred, bred, p2, bt, C);
}

__global__ void __launch_bounds__(256)
k_cat_scale(const float* __restrict__ p2, const float* __restrict__ scale, bf16* __restrict__ weff, int C, int n_per_image) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per_image) return;
  const int k = i % (2 * C);
  const float s = k < C ? scale[b * C + k] : 1.f;
  weff[(i64)b * n_per_image + i] = __float2bfloat16_rn(p2[i] * s);
}
// weff [B][C][9][2C] (T) = P2 with the first C input channels scaled by scale[b]
void launch_cat_scale(Ctx& ctx, const float* p2, const float* scale, void* weff, int B, int C) {
  if (ctx.dry) return;
  const int n = C * 9 * 2 * C;
  ScopedLaunch sl(RF_K_SE_FINALIZE, 6.0 * B * n);
  launch_pdl(k_cat_scale, dim3(cdiv(n, 256), B), dim3(256), 0, ctx.stream, p2, scale, (bf16*)weff, C, n);
}

bool lnconv_cat_supported(const Ctx& ctx, int C, int H, int W) {
  static int on = -1;                     // debugging aid: RAWFORMER_B200_LNCONV_CAT=0 keeps channel_reduce as its own GEMM
  if (on < 0) {
    const char* e = getenv("RAWFORMER_B200_LNCONV_CAT");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on && C == 32 && lnconv_supported(ctx, C, H, W);
}

// ONE image: out = LeakyReLU_0.2(Conv_out(channel_reduce(cat(xmod * s, x2)))); weff = this image's [C][9][2C] weights
bool launch_lnconv_cat(Ctx& ctx, const void* xmod, const void* x2, const void* weff, const float* btab, void* out, int H, int W,
                       int C) {
  if (!lnconv_cat_supported(ctx, C, H, W)) return false;
  const double rows = (double)H * W;
  ScopedLaunch sl(RF_K_CONV3X3_LC, rows * C * 2.0 * 3.0, rows * 2.0 * 9 * 2 * C * C);
  const LcSel sel{C, 0, 0, C, 0};
  return lnconv_launch<LC_CAT, 32>(ctx, xmod, nullptr, weff, btab, sel, x2, nullptr, nullptr, out, nullptr, nullptr, 1, H, W, 0) > 0;
}

// ONE image, C = 32: out = x1 + conv_ffn(norm2(x1)) with x1 = x + Mw v + proj_b computed per halo patch inside the kernel
// (project_out fused in front: x1 never reaches HBM)
bool lnconv_proj_supported(const Ctx& ctx, int C, int H, int W) {
  static int on = -1;                     // debugging aid: RAWFORMER_B200_LNCONV_PROJ=0 keeps project_out as its own GEMM
  if (on < 0) {
    const char* e = getenv("RAWFORMER_B200_LNCONV_PROJ");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on && C == 32 && lnconv_supported(ctx, C, H, W);
}
bool launch_lnconv_ffn_proj(Ctx& ctx, const void* x, const void* v, const void* Mw, const float* proj_b, const void* cw,
                            const float* btab, const void* W2, const float* b2, void* out, int H, int W, int C) {
  if (!lnconv_proj_supported(ctx, C, H, W)) return false;
  const double rows = (double)H * W;
  ScopedLaunch sl(RF_K_FFN_FUSED, rows * C * 2.0 * 3.0, rows * (2.0 * 9 * C * 2 * C + 2.0 * 2 * C * C + 2.0 * C * C * 1.4));
  const LcSel sel{2 * C, 0, 0, 2 * C, 0};
  LcProj pj;
  pj.v = v; pj.Mw = Mw; pj.pb = proj_b;
  return lnconv_launch<LC_FFN, 32>(ctx, x, nullptr, cw, btab, sel, W2, b2, nullptr, out, nullptr, nullptr, 1, H, W, 0, &pj) > 0;
}

}  // namespace rf
