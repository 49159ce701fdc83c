// q|k half of Attention.qkv_dwconv fused with the transposed attention's statistics (bf16 mode, C in {32, 64}):
//
//   q|k = depthwise3x3(qkv_pre[:, :2C]) + bias          FLCA_RF.py:223 (first two thirds of the 3C channels)
//   G   = sum_p q_p k_p^T (per-head diagonal blocks),  |q_i|^2, |k_j|^2 over the pixels        FLCA_RF.py:228-230
//
// q and k are ONLY ever used for those reductions (attn . v followed by project_out is one 1x1 conv with a per-image
// weight), so this kernel never writes them: the depthwise output tile goes to shared memory in the
// MN-major SWIZZLE_128B UMMA operand layout ([pixel][64 channels], 16-byte units XOR-swizzled by pixel & 7 -- what a TMA
// box of k_tc_gram lands as), and tcgen05 accumulates the self-Gram of [q|k] (its upper-right block is q^T k, its diagonal
// the squared norms of the bf16-rounded values) in tensor memory over all tiles of the CTA.  HBM traffic of the qkv
// depthwise + Gram step drops from 3C in + 3C out + 2C in to 3C in + C out (v is written by the plain depthwise kernel).
// Every CTA ends with plain stores of its partial statistics into its own slot; k_attn_reduce sums the slots in order.
//
// Geometry as in rf_ffn_fused.cu: tiles of TH x TW output pixels, input patch (TH+2) x PW pixels with PW = TW + 2 a
// multiple of 8 (swizzle phase = column & 7); thread = (column, 4 channels) slides a 3-row window down the patch.
#include <stdlib.h>
#include <string.h>

#include "rf_kernels.cuh"
#include "rf_tma.cuh"
#include "rf_dw_math.cuh"

namespace rf {

constexpr int QG_TH = 10;            // output rows per tile ((TH + 2) % 3 == 0)

struct QkGramP {
  const float* w;        // [9][wpitch] depthwise taps; channels 0 .. 2C-1 are q|k
  const float* bias;     // [>= 2C]
  float* gram_part;      // [nslot][C][C/8]: per-head diagonal blocks of q^T k of this CTA's pixels
  float* sq_part;        // [nslot][2C]
  int wpitch;
  int H, W, C;
  int ylo, yhi;          // rows that count (row-tiled forward: the band's interior; else 0, H)
  int TW, PW, nvec, cthreads;
  int tiles_x, total_tiles;
  int packed;            // 64-channel chunk tiles of a q|k pixel: 1 (2C = 64) or 2 (2C = 128)
  int nk;                // MMAs per tile = TH*PW/16
  uint32_t in_bytes, in_stride;   // TMA bytes of one patch; bytes per patch buffer
  uint32_t ct_bytes;     // bytes of one chunk tile of the g buffer = TH*PW*128
};

__device__ __forceinline__ uint64_t qg_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;  // LBO: next 64-element group along M/N
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: next 8-pixel group along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(512, 1)
k_dwqk_gram(const __grid_constant__ CUtensorMap mapIn, const QkGramP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t g_bytes = (uint32_t)p.packed * p.ct_bytes;
  const uint32_t sG = base;                                   // 2 g buffers (1024-aligned: ct_bytes is a multiple of 1024)
  const uint32_t sZ = sG + 2u * g_bytes;                      // 2 KB of zeros (packed == 1: the second 64-row operand group)
  const uint32_t sIn = sZ + 2048u;                            // 2 patch buffers
  const uint32_t bars = sIn + 2u * p.in_stride;
  const uint32_t in_full0 = bars, g_full0 = bars + 16, g_free0 = bars + 32, done_bar = bars + 48, tmem_slot = bars + 56;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cwarps = p.cthreads >> 5;                         // compute warps; warp `cwarps` is the control warp

  if (tid == 0) {
    tma_prefetch_desc(&mapIn);
    for (int i = 0; i < 7; ++i) mbar_init(bars + 8u * i, 1);
    fence_barrier_init();
  }
  // zero the g buffers (their two pad columns per row are never written and must not pollute the Gram) and the zero block
  for (uint32_t o = (uint32_t)tid * 16u; o < 2u * g_bytes + 2048u; o += blockDim.x * 16u)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sG + o), "r"(0u) : "memory");
  fence_proxy_async();
  if (warp == cwarps) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();

  const int first = blockIdx.x, stride = gridDim.x;
  const int ntiles = first < p.total_tiles ? (p.total_tiles - first + stride - 1) / stride : 0;

  if (warp == cwarps) {
    // ================= control thread: patch loads + Gram MMAs =================
    if (lane == 0 && ntiles > 0) {
      auto issue_in = [&](int i) {
        const int t = first + i * stride;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int buf = i & 1;
        mbar_expect_tx(in_full0 + 8u * buf, p.in_bytes);
        tma_load_3d(sIn + (uint32_t)buf * p.in_stride, &mapIn, in_full0 + 8u * buf, 0, tx * p.TW - 1, ty * QG_TH - 1);
      };
      issue_in(0);
      if (ntiles > 1) issue_in(1);
      // kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16), M = N = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      for (int i = 0; i < ntiles; ++i) {
        const int buf = i & 1;
        mbar_wait(g_full0 + 8u * buf, (i >> 1) & 1);       // g tile i written; patch buffer `buf` no longer read
        tc_fence_after();
        if (i + 2 < ntiles) issue_in(i + 2);
        const uint32_t gb = sG + (uint32_t)buf * g_bytes;
        for (int k = 0; k < p.nk; ++k) {
          const uint32_t a0 = gb + (uint32_t)k * 2048u;     // 16 pixels = 16 rows of 128 B
          const uint32_t lbo = p.packed == 2 ? p.ct_bytes : sZ - a0;
          const uint64_t d = qg_mn_desc(a0, lbo);
          umma_f16(tmem_base, d, d, idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(g_free0 + 8u * buf);
      }
      umma_commit(done_bar);
    }
  } else if (warp < cwarps) {
    // ================= compute warps: depthwise 3x3 of q|k -> g tile =================
    const int x = tid / p.nvec, cv = tid - x * p.nvec;          // column (0 .. TW-1), 4-channel vector
    const int CC = 2 * p.C;
    const uint32_t cstep = (uint32_t)CC * 2u, pitch = (uint32_t)p.PW * cstep;
    const uint32_t toff = (uint32_t)x * cstep + (uint32_t)cv * 8u;
    // g tile: chunk tile cv / 16, 16-byte unit ((cv % 16) >> 1) ^ (column & 7) of the pixel's 128-byte row
    const uint32_t goff = (uint32_t)(cv >> 4) * p.ct_bytes + (uint32_t)x * 128u +
                          ((uint32_t)(((cv & 15) >> 1) ^ (x & 7)) << 4) + (uint32_t)(cv & 1) * 8u;
    float2 wv[9][2], bs[2];
    {
      const int c0 = cv * 4;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w + (i64)k * p.wpitch + c0));
        wv[k][0] = make_float2(w4.x, w4.y);
        wv[k][1] = make_float2(w4.z, w4.w);
      }
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
      bs[0] = make_float2(b4.x, b4.y);
      bs[1] = make_float2(b4.z, b4.w);
    }
    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1;
      const int t = first + i * stride;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      mbar_wait(in_full0 + 8u * buf, (i >> 1) & 1);
      if (i >= 2) mbar_wait(g_free0 + 8u * buf, ((i >> 1) + 1) & 1);   // the MMAs of tile i-2 have read this g buffer
      uint32_t src = sIn + (uint32_t)buf * p.in_stride + toff;
      uint32_t dst = sG + (uint32_t)buf * ((uint32_t)p.packed * p.ct_bytes) + goff - 2u * (uint32_t)p.PW * 128u;
      const int xo = tx * p.TW + x;
      // output rows [lo, hi) of this tile count: inside the image (and the band's interior), column inside the image
      const int y0 = ty * QG_TH;
      const int lo = max(p.ylo - y0, 0);
      const unsigned nrows = xo < p.W ? (unsigned)max(min(QG_TH, min(p.yhi, p.H) - y0) - lo, 0) : 0u;
      float2 acc[3][2];
#pragma unroll
      for (int a = 0; a < 3; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int g = 0; g < (QG_TH + 2) / 3; ++g) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = 3 * g + j;                   // patch row: feeds outputs r (ky=0), r-1 (ky=1), r-2 (ky=2)
          float2 v[3][2];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) unpack_bf16x4(lds64(src + (uint32_t)kx * cstep), v[kx]);
          src += pitch;
          float2* aN = acc[j];
          float2* aM = acc[(j + 2) % 3];
          float2* aD = acc[(j + 1) % 3];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            aN[k] = __ffma2_rn(wv[2][k], v[2][k], __ffma2_rn(wv[1][k], v[1][k], __ffma2_rn(wv[0][k], v[0][k], bs[k])));
            aM[k] = __ffma2_rn(wv[5][k], v[2][k], __ffma2_rn(wv[4][k], v[1][k], __ffma2_rn(wv[3][k], v[0][k], aM[k])));
            aD[k] = __ffma2_rn(wv[8][k], v[2][k], __ffma2_rn(wv[7][k], v[1][k], __ffma2_rn(wv[6][k], v[0][k], aD[k])));
          }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(aD[0].x, aD[0].y), h1 = __floats2bfloat162_rn(aD[1].x, aD[1].y);
          const bool ok = (unsigned)(r - 2 - lo) < nrows;          // pixels that do not count contribute exact zeros
          const uint32_t q0 = ok ? *reinterpret_cast<uint32_t*>(&h0) : 0u, q1 = ok ? *reinterpret_cast<uint32_t*>(&h1) : 0u;
          if (r >= 2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst), "r"(q0), "r"(q1) : "memory");
          dst += (uint32_t)p.PW * 128u;
        }
      }
      fence_proxy_async();                           // g tile (generic stores) -> tensor cores (async proxy)
      asm volatile("bar.sync 1, %0;" ::"r"(p.cthreads) : "memory");
      if (tid == 0) mbar_arrive(g_full0 + 8u * buf);
    }
    // ---- read-out: per-head diagonal blocks of q^T k and the squared norms -> this CTA's slot ----
    if (warp < 4) {
      const int C = p.C, c = C >> 3;
      const int row = warp * 32 + lane;              // accumulator row = channel of [q|k]
      if (ntiles > 0) {
        mbar_wait(done_bar, 0);
        tc_fence_after();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* gp = p.gram_part + (i64)blockIdx.x * C * c;
      float* sp = p.sq_part + (i64)blockIdx.x * 2 * C;
      const int h = row < C ? row / c : -1;
      for (int cc = 0; cc < 128; cc += 16) {
        uint32_t v[16];
        if (ntiles > 0) {
          tmem_ld16(taddr + cc, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (row < 2 * C) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = cc + j;
            if (col == row) sp[row] = __uint_as_float(v[j]);
            const int kc = col - C;                  // k channel of this accumulator column
            if (h >= 0 && kc >= h * c && kc < (h + 1) * c) gp[(i64)row * c + (kc - h * c)] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == cwarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool qk_gram_enabled() {
  static int enabled = -1;                // debugging aid: RAWFORMER_B200_NO_QK_GRAM=1 keeps the depthwise + Gram kernels
  if (enabled < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_QK_GRAM");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled != 0;
}

bool qk_gram_supported(const Ctx& ctx, int C) {
  return qk_gram_enabled() && tcgen05_enabled() && ctx.dtype == RF_BF16 && (C == 32 || C == 64);
}

// ONE image: qkv_pre [H][W][3C] bf16.  Returns the number of partial slots written (= CTAs), 0 if unsupported.
int launch_dwqk_gram(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, float* gram_part, float* sq_part,
                     int H, int W, int C, int slot_cap) {
  if (!qk_gram_supported(ctx, C)) return 0;
  QkGramP p;
  memset(&p, 0, sizeof(p));
  p.w = dw_w; p.bias = dw_b; p.gram_part = gram_part; p.sq_part = sq_part; p.wpitch = 3 * C;
  p.H = H; p.W = W; p.C = C;
  p.ylo = 0; p.yhi = H;
  if (ctx.band != nullptr) { p.ylo = ctx.band->ht; p.yhi = ctx.band->ht + ctx.band->rows_in; }
  p.nvec = 2 * C / 4;                              // 16 or 32 four-channel vectors per q|k pixel
  p.PW = C == 32 ? 32 : 16;
  p.TW = p.PW - 2;
  p.cthreads = p.TW * p.nvec;                      // 480 or 448
  p.packed = C == 32 ? 1 : 2;
  p.nk = QG_TH * p.PW / 16;
  p.ct_bytes = (uint32_t)(QG_TH * p.PW * 128);
  p.in_bytes = (uint32_t)((QG_TH + 2) * p.PW * 2 * C * 2);
  p.in_stride = (p.in_bytes + 1023u) & ~1023u;
  p.tiles_x = cdiv(W, p.TW);
  const i64 total = (i64)p.tiles_x * cdiv(H, QG_TH);
  if (total <= 0 || total > 0x7fffffff) return 0;
  p.total_tiles = (int)total;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (grid > slot_cap) grid = slot_cap;
  const size_t smem = 1024 + 2 * (size_t)p.packed * p.ct_bytes + 2048 + 2 * (size_t)p.in_stride + 128;
  if (smem > 232448) return 0;
  CUtensorMap m;
  const i64 d[3] = {2 * (i64)C, W, H};
  const i64 st[3] = {1, 3 * (i64)C, 3 * (i64)C * W};
  const int bx[3] = {2 * C, p.PW, QG_TH + 2};
  if (!make_map_ex(&m, qkv_pre, 3, d, st, bx, 2, 0)) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_dwqk_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return 0;
    attr_set = true;
  }
  const double px = (double)H * W;
  ScopedLaunch sl(RF_K_QKV_FUSED, px * 2 * C * 2.0, px * (36.0 * C + 2.0 * (2 * C) * (2 * C)));
  launch_pdl(k_dwqk_gram, dim3(grid), dim3(p.cthreads + 32), smem, ctx.stream, m, p);
  return grid;
}

}  // namespace rf
