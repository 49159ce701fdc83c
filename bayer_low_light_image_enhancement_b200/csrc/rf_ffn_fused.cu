// Fused conv-FFN of the transformer branch (bf16 mode, C in {32, 64}: the two full-resolution stages):
//
//     out = x1 + pointwise2( gelu( depthwise3x3( pointwise1( norm2(x1) ) ) ) )          FLCA_RF.py:204-209, :253
//
// ONE kernel; the 2C-wide hidden tensor (the largest of the block) lives only in shared / tensor memory.  HBM traffic per
// pixel is C in (+20 % halo, mostly L2 hits) + C out = 4C bytes instead of the 22C bytes of pointwise1 -> depthwise ->
// pointwise2 as three kernels.
//
// Geometry.  A tile is TH x 30 output pixels; its halo'd input patch is (TH+2) x 32 pixels (row pitch 32 = a multiple of
// 8, so the 128-byte-swizzle phase of a pixel depends on its column only).  Hidden channels are processed in chunks of 64.
// Per (tile, chunk) "step":
//   MMA1  acc1[g] (TMEM) = Xpatch[128 px of group g x C] * W1f[64 x C]^T             tcgen05, operands by TMA
//   E1    acc1 -> folded LayerNorm (rstd, mean per pixel) + bias -> 0 outside the image (the conv's zero padding is on
//         the HIDDEN tensor) -> bf16 -> hidden tile in shared memory ([px][64 ch], 16-byte units XOR-swizzled by px & 7)
//   DW    depthwise 3x3 + bias + erf-GELU on the CUDA cores: thread = (column, 4 channels), 3-row sliding window, packed
//         FFMA2 -- the loop of rf_dw_tma.cu on the shared-memory tile -> bf16 -> g tile ([px][64 ch], same swizzle = the
//         SWIZZLE_128B K-major UMMA operand layout)
//   MMA3  acc3[g] (TMEM) (+)= G[128 px x 64] * W2[:, chunk]^T[C x 64]
// and per tile   E3: acc3 + bias + residual (x1, re-read through L2) -> bf16 -> global.
//
// Warp roles: warps 0-15 compute (E1 / DW / E3, separated by named barriers), warp 16 lane 0 issues TMA and MMAs.  The
// control thread issues MMA1 of step s+1 as soon as E1 of step s has drained acc1 (it runs under DW of step s) and MMA3
// of step s as soon as the g tile is complete (it runs under E1 of step s+1); E3 of a tile is deferred until after E1 of
// the next step, so no compute warp ever waits for the tensor pipe in steady state.  The x patch is double-buffered.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rf_kernels.cuh"
#include "rf_tma.cuh"
#include "rf_dw_math.cuh"

namespace rf {

constexpr int FF_TW = 30;            // output columns per tile
constexpr int FF_PW = 32;            // patch columns (row pitch in pixels)
constexpr int FF_CC = 64;            // hidden-channel chunk
constexpr int FF_CWARPS = 15;        // compute warps: 30 columns x 16 four-channel vectors = 480 threads
constexpr int FF_CTHREADS = FF_CWARPS * 32;
constexpr int FF_THREADS = FF_CTHREADS + 32;      // + the control warp: 512 threads, 128 registers each
constexpr int FF_MAXG = 5;           // MMA row groups of 128 patch pixels (TH + 2 <= 20)
constexpr int FF_E3T = 4;            // 16-column E3 tasks per warp (nm3 * C/16 <= 12 over >= 3 warps per quadrant)

struct FfnP {
  const float* cs;      // [2C] row sums of the LayerNorm-folded pointwise1 weights
  const float* b1;      // [2C] folded pointwise1 bias
  const float* stats;   // [B*H*W][npart] float2 (sum, sumsq) of the rows of x
  const float* dw_w;    // [9][2C]
  const float* dw_b;    // [2C]
  const float* b2;      // [C]
  const bf16* x;        // [B,H,W,C] (residual)
  bf16* out;            // [B,H,W,C]
  float invC, eps;
  int npart;
  int H, W, C, B;
  int TH, nchunk, nm1, nm3;
  int tiles_x, tiles_y, total_tiles;
  int nx;               // x patch buffers (1 or 2)
  uint32_t xrow;        // bytes of a patch pixel (C*2 = swizzle span: 64 or 128)
  uint32_t x_bytes;     // bytes one TMA box delivers = (TH+2)*32*xrow
  uint32_t x_stride;    // bytes per patch buffer = nm1*128*xrow
  uint32_t s_bytes;     // bytes of the statistics box of one tile = (TH+2)*34*8 (34 pixels per row: the box starts one pixel
                        // left of the patch so that its global start address is 16-byte aligned)
  uint32_t s_stride;    // s_bytes rounded up to 128
  uint32_t r_bytes;     // bytes of the residual box of one tile = TH*30*C*2 (0: residual read from global memory in E3)
  uint32_t w_bytes;     // bytes of all resident weights
  int acc3_col, tmem_cols;
  unsigned long long* dbg;   // debugging aid (RAWFORMER_B200_FFN_DBG=1): per-CTA cycles spent per phase by compute thread 0
};

__device__ __forceinline__ void ff_decode(const FfnP& p, int t, int& tx, int& ty, int& b) {
  tx = t % p.tiles_x;
  const int r = t / p.tiles_x;
  ty = r % p.tiles_y;
  b = r / p.tiles_y;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// one warp polls the mbarrier with ONE lane (hundreds of threads polling the same barrier word serialise: ~550 cycles for
// 512 pollers), the others wait at the warp barrier
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

template <bool DBG>
__global__ void __launch_bounds__(FF_THREADS, 1)
k_ffn_fused(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1,
            const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapS,
            const __grid_constant__ CUtensorMap mapR, const FfnP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base;                                                    // nx patch buffers (K-major, swizzled rows)
  const uint32_t sH = sX + (uint32_t)p.nx * p.x_stride;                        // hidden tile: (TH+2)*32 px x 128 B
  const uint32_t sG = sH + (uint32_t)(p.TH + 2) * FF_PW * 128u;                // g tile: nm3*128 px x 128 B
  const uint32_t sW1 = sG + (uint32_t)p.nm3 * 16384u;                          // nchunk x [64 rows x xrow]
  const uint32_t sW2 = sW1 + (uint32_t)p.nchunk * FF_CC * p.xrow;              // nchunk x [C rows x 128 B]
  const uint32_t sC = sW2 + (uint32_t)p.nchunk * p.C * 128u;                   // cs[2C], b1[2C], b2[C] floats
  const uint32_t sS = sC + (uint32_t)(5 * p.C) * 4u;                           // nx x [(TH+2)*32 px] float2 (sum, sumsq) by TMA
  const uint32_t sR = sS + (uint32_t)p.nx * p.s_stride;                        // residual tile [TH][30 px][C] by TMA
  const uint32_t bars = sR + ((p.r_bytes + 127u) & ~127u);
  const uint32_t w_full = bars, x_full0 = bars + 8, mma1_done = bars + 24, acc1_free = bars + 32, g_full = bars + 40,
                 mma3_done = bars + 48, r_full = bars + 56, tmem_slot = bars + 64;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW1);
    tma_prefetch_desc(&mapW2);
    tma_prefetch_desc(&mapS);
    if (p.r_bytes) tma_prefetch_desc(&mapR);
    for (int i = 0; i < 8; ++i) mbar_init(bars + 8u * i, 1);
    fence_barrier_init();
  }
  if (warp == FF_CWARPS) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  {
    float* sc = reinterpret_cast<float*>(smem_raw + (sC - smem_u32(smem_raw)));
    for (int i = tid; i < 2 * p.C; i += FF_THREADS) {
      sc[i] = __ldg(p.cs + i);
      sc[2 * p.C + i] = __ldg(p.b1 + i);
    }
    for (int i = tid; i < p.C; i += FF_THREADS) sc[4 * p.C + i] = __ldg(p.b2 + i);
  }
  __syncthreads();

  const int first = blockIdx.x, stride = gridDim.x;
  const int ntiles = first < p.total_tiles ? (p.total_tiles - first + stride - 1) / stride : 0;
  const int S = ntiles * p.nchunk;
  const int ksteps1 = p.C >> 4;

  if (warp == FF_CWARPS) {
    // ================= control thread: TMA + MMA issue =================
    if (lane == 0 && ntiles > 0) {
      auto issue_x = [&](int tix) {
        int tx, ty, b;
        ff_decode(p, first + tix * stride, tx, ty, b);
        const int buf = p.nx == 2 ? (tix & 1) : 0;
        mbar_expect_tx(x_full0 + 8u * buf, p.x_bytes + p.s_bytes);
        tma_load_4d(sX + (uint32_t)buf * p.x_stride, &mapX, x_full0 + 8u * buf, 0, tx * FF_TW - 1, ty * p.TH - 1, b);
        tma_load_3d(sS + (uint32_t)buf * p.s_stride, &mapS, x_full0 + 8u * buf, 2 * (tx * FF_TW - 2), ty * p.TH - 1, b);
      };
      mbar_expect_tx(w_full, p.w_bytes);
      for (int c = 0; c < p.nchunk; ++c) {
        tma_load_3d(sW1 + (uint32_t)c * FF_CC * p.xrow, &mapW1, w_full, 0, c * FF_CC, 0);
        tma_load_3d(sW2 + (uint32_t)c * p.C * 128u, &mapW2, w_full, c * FF_CC, 0, 0);
      }
      issue_x(0);
      mbar_wait(w_full, 0);
      const uint32_t idesc1 = make_idesc_m128(FF_CC), idesc3 = make_idesc_m128(p.C);
      const int bk1 = (int)(p.xrow >> 1);
      auto issue_mma1 = [&](int s) {
        const int tix = s / p.nchunk, c = s - tix * p.nchunk;
        const int buf = p.nx == 2 ? (tix & 1) : 0;
        if (c == 0) {
          mbar_wait(x_full0 + 8u * buf, (p.nx == 2 ? (tix >> 1) : tix) & 1);
          tc_fence_after();
        }
        const uint64_t bdesc = make_kmajor_desc(sW1 + (uint32_t)c * FF_CC * p.xrow, bk1);
        for (int g = 0; g < p.nm1; ++g) {
          const uint64_t adesc = make_kmajor_desc(sX + (uint32_t)buf * p.x_stride + (uint32_t)g * 128u * p.xrow, bk1);
          for (int k = 0; k < ksteps1; ++k)
            umma_f16(tmem_base + (uint32_t)(g * FF_CC), adesc + 2u * k, bdesc + 2u * k, idesc1, k ? 1u : 0u);
        }
        umma_commit(mma1_done);
        // the other patch buffer was last read by the MMA1s of tile tix-1, which completed before E1 of its last chunk
        if (c == 0 && p.nx == 2 && tix + 1 < ntiles) issue_x(tix + 1);
      };
      issue_mma1(0);
      long long cph[4] = {0, 0, 0, 0}, clast = DBG ? clock64() : 0;
      auto cmark = [&](int i) {
        if (DBG) {
          const long long now = clock64();
          cph[i] += now - clast;
          clast = now;
        }
      };
      for (int s = 0; s < S; ++s) {
        const int tix = s / p.nchunk, c = s - tix * p.nchunk;
        mbar_wait(acc1_free, s & 1);                 // E1(s) has drained acc1 (and MMA1(s) is complete)
        tc_fence_after();
        cmark(0);
        if (p.nx == 1 && c == p.nchunk - 1 && tix + 1 < ntiles) issue_x(tix + 1);
        if (s + 1 < S) issue_mma1(s + 1);
        cmark(1);
        mbar_wait(g_full, s & 1);                    // DW(s) has written the g tile (and E3 of the previous tile is done)
        tc_fence_after();
        cmark(2);
        const uint64_t bdesc = make_sw128_desc(sW2 + (uint32_t)c * p.C * 128u);
        for (int g = 0; g < p.nm3; ++g) {
          const uint64_t adesc = make_sw128_desc(sG + (uint32_t)g * 16384u);
          for (int k = 0; k < FF_CC / 16; ++k)
            umma_f16(tmem_base + (uint32_t)(p.acc3_col + g * p.C), adesc + 2u * k, bdesc + 2u * k, idesc3, (c | k) ? 1u : 0u);
        }
        umma_commit(mma3_done);
        if (p.r_bytes && c == p.nchunk - 1) {
          // residual tile of THIS tile for its (deferred) E3: the compute warps finished E3 of the previous tile before
          // DW of this step, i.e. before g_full -- the buffer is free
          int tx, ty, b;
          ff_decode(p, first + tix * stride, tx, ty, b);
          mbar_expect_tx(r_full, p.r_bytes);
          tma_load_4d(sR, &mapR, r_full, 0, tx * FF_TW, ty * p.TH, b);
        }
        cmark(3);
      }
      if (DBG) {
        for (int i = 0; i < 4; ++i) p.dbg[(gridDim.x + blockIdx.x) * 8 + i] = (unsigned long long)cph[i];
      }
    }
  } else {
    // ================= compute warps =================
    const int q = warp & 3, wi = warp >> 2;          // TMEM lane quadrant; index of this warp among the quadrant's warps
    const int nq = q < (FF_CWARPS & 3) ? (FF_CWARPS >> 2) + 1 : (FF_CWARPS >> 2);     // warps that share this quadrant
    const int dx = tid >> 4, cv = tid & 15;          // depthwise role: output column (0..29), 4-channel vector
    // shared-memory offsets of this thread's three input columns (swizzle phase = column & 7) and of its g-tile column
    uint32_t off[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
      off[kx] = (uint32_t)(dx + kx) * 128u + ((uint32_t)((cv >> 1) ^ ((dx + kx) & 7)) << 4) + (uint32_t)(cv & 1) * 8u;
    const uint32_t goff = (uint32_t)dx * 128u + ((uint32_t)((cv >> 1) ^ (dx & 7)) << 4) + (uint32_t)(cv & 1) * 8u;
    const uint32_t lane_sw = (uint32_t)(lane & 7);
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const int sl16 = p.C >> 4;                       // 16-column slices of an acc3 group
    const int e3_tasks = p.nm3 * sl16;

    long long tph[DBG ? 12 : 1] = {0}, tlast = DBG ? clock64() : 0;
    auto mark = [&](int i) {
      if (DBG && tid == 0) {
        const long long now = clock64();
        tph[i] += now - tlast;
        tlast = now;
      }
    };

    // ---- E3 of tile (tx, ty, b): acc3 + bias + residual -> bf16 -> global; the residual loads are issued early ----
    uint4 rres[FF_E3T][2];
    auto e3_coords = [&](int i, int tx, int ty, int b, int& g, int& sl, bool& ok, i64& o) {
      const int task = wi + i * nq;
      g = task / sl16; sl = task - g * sl16;
      const int pp = g * 128 + q * 32 + lane;
      const int oy = pp >> 5, ox = pp & 31;
      const int y = ty * p.TH + oy, x = tx * FF_TW + ox;
      ok = task < e3_tasks && ox < FF_TW && oy < p.TH && y < p.H && x < p.W;
      o = ((((i64)b * p.H + y) * p.W + x) * p.C) + sl * 16;
    };
    auto e3_prefetch = [&](int tx, int ty, int b) {
#pragma unroll
      for (int i = 0; i < FF_E3T; ++i) {
        int g, sl; bool ok; i64 o;
        e3_coords(i, tx, ty, b, g, sl, ok, o);
        rres[i][0] = rres[i][1] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          rres[i][0] = __ldg(reinterpret_cast<const uint4*>(p.x + o));
          rres[i][1] = __ldg(reinterpret_cast<const uint4*>(p.x + o) + 1);
        }
      }
    };
    auto epilogue3 = [&](int tx, int ty, int b, int tix_done) {
      mark(8);
      if (p.r_bytes) warp_wait(r_full, tix_done & 1, lane);
      else e3_prefetch(tx, ty, b);                  // all residual loads of this warp in flight at once (L2 hits)
      mark(9);
#pragma unroll
      for (int i = 0; i < FF_E3T; ++i) {
        if (wi + i * nq < e3_tasks) {               // warp-uniform
          int g, sl; bool ok; i64 o;
          e3_coords(i, tx, ty, b, g, sl, ok, o);
          uint32_t v[16];
          tmem_ld16(tq + (uint32_t)(p.acc3_col + g * p.C + sl * 16), v);
          tmem_ld_wait();
          mark(10);
          if (ok) {
            const uint32_t bb = sC + (uint32_t)(4 * p.C + sl * 16) * 4u;
            if (p.r_bytes) {
              const int pp = g * 128 + q * 32 + lane;
              const uint32_t ra = sR + (uint32_t)(((pp >> 5) * FF_TW + (pp & 31)) * p.C + sl * 16) * 2u;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rres[i][0].x), "=r"(rres[i][0].y), "=r"(rres[i][0].z),
                           "=r"(rres[i][0].w) : "r"(ra));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rres[i][1].x), "=r"(rres[i][1].y), "=r"(rres[i][1].z),
                           "=r"(rres[i][1].w) : "r"(ra + 16u));
            }
            uint32_t ow[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 ba = lds128f(bb + h * 32u), bc = lds128f(bb + h * 32u + 16u);
              const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bc.x, bc.y, bc.z, bc.w};
              const uint32_t rw[4] = {rres[i][h].x, rres[i][h].y, rres[i][h].z, rres[i][h].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float f0 = __uint_as_float(v[h * 8 + 2 * j]) + bias[2 * j] + __uint_as_float(rw[j] << 16);
                const float f1 = __uint_as_float(v[h * 8 + 2 * j + 1]) + bias[2 * j + 1] + __uint_as_float(rw[j] & 0xffff0000u);
                __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
                ow[h * 4 + j] = *reinterpret_cast<uint32_t*>(&hh);
              }
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + o);
            dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
          }
          mark(11);
        }
      }
    };

    int s = 0;
    int ptx = 0, pty = 0, pb = 0;                     // coordinates of the previous tile (deferred E3)
    for (int tix = 0; tix < ntiles; ++tix) {
      int tx, ty, b;
      ff_decode(p, first + tix * stride, tx, ty, b);
      float rs[FF_MAXG], nmu[FF_MAXG];
      unsigned inside_mask = 0;
      for (int c = 0; c < p.nchunk; ++c, ++s) {
        mark(0);
        // ---- E1: acc1 -> hidden tile.  Tasks = (row group g, 8-channel slice j): j = wi, wi + nq, ... ----
        warp_wait(mma1_done, s & 1, lane);
        tc_fence_after();
        mark(1);
        if (c == 0) {
          // LayerNorm statistics of this thread's patch pixels (one per MMA row group): landed with the patch (the MMA1s
          // of this tile waited for that barrier, and their completion was just observed)
          const uint32_t sbuf = sS + (uint32_t)(p.nx == 2 ? (tix & 1) : 0) * p.s_stride;
#pragma unroll
          for (int g = 0; g < FF_MAXG; ++g) {
            rs[g] = 0.f; nmu[g] = 0.f;
            if (g < p.nm1) {
              const int pp = g * 128 + q * 32 + lane;
              const int hy = pp >> 5, hx = pp & 31;
              const int y = ty * p.TH - 1 + hy, x = tx * FF_TW - 1 + hx;
              const bool in = hy < p.TH + 2 && y >= 0 && y < p.H && x >= 0 && x < p.W;
              float sum = 0.f, ssq = 0.f;
              if (hy < p.TH + 2)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sum), "=f"(ssq) : "r"(sbuf + (uint32_t)(hy * 34 + hx + 1) * 8u));
              const float mu = sum * p.invC;
              rs[g] = rsqrtf(fmaxf(ssq * p.invC - mu * mu, 0.f) + p.eps);
              nmu[g] = -rs[g] * mu;
              inside_mask |= (in ? 1u : 0u) << g;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < FF_MAXG; ++g) {
          if (g < p.nm1) {
            const int pp = g * 128 + q * 32 + lane;
            const bool rowok = pp < (p.TH + 2) * FF_PW;          // warp-uniform
            const float2 r2 = make_float2(rs[g], rs[g]), n2 = make_float2(nmu[g], nmu[g]);
            const bool in = (inside_mask >> g) & 1u;
            const uint32_t row = sH + (uint32_t)pp * 128u;
            uint32_t v[3][8];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int j = wi + t * nq;
              if (j < 8) tmem_ld8(tq + (uint32_t)(g * FF_CC + j * 8), v[t]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int j = wi + t * nq;
              if (j < 8 && rowok) {
                const uint32_t cb = sC + (uint32_t)(c * FF_CC + j * 8) * 4u;
                const float4 c0 = lds128f(cb), c1 = lds128f(cb + 16u);
                const float4 d0 = lds128f(cb + (uint32_t)(2 * p.C) * 4u), d1 = lds128f(cb + (uint32_t)(2 * p.C) * 4u + 16u);
                const float2 csv[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
                const float2 b1v[4] = {make_float2(d0.x, d0.y), make_float2(d0.z, d0.w), make_float2(d1.x, d1.y), make_float2(d1.z, d1.w)};
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 a = make_float2(__uint_as_float(v[t][2 * e]), __uint_as_float(v[t][2 * e + 1]));
                  const float2 f = __ffma2_rn(r2, a, __ffma2_rn(n2, csv[e], b1v[e]));
                  __nv_bfloat162 hh = __floats2bfloat162_rn(f.x, f.y);
                  pk[e] = in ? *reinterpret_cast<uint32_t*>(&hh) : 0u;
                }
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)j ^ lane_sw) << 4)), "r"(pk[0]),
                             "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
              }
            }
          }
        }
        mark(2);
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(FF_CTHREADS) : "memory");     // hidden tile complete, acc1 drained
        if (tid == 0) mbar_arrive(acc1_free);
        mark(3);
        // ---- the g tile is free once MMA3 of the previous step has read it; E3 of a finished tile rides here ----
        if (s > 0) {
          warp_wait(mma3_done, (s - 1) & 1, lane);
          tc_fence_after();
          mark(4);
          if (c == 0) epilogue3(ptx, pty, pb, tix - 1);
          mark(5);
        }
        // ---- DW: depthwise 3x3 + bias + GELU on the hidden tile -> g tile ----
        {
          float2 wv[9][2], bs[2];
          const int c0 = c * FF_CC + cv * 4;
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.dw_w + (i64)k * 2 * p.C + c0));
            wv[k][0] = make_float2(w4.x, w4.y);
            wv[k][1] = make_float2(w4.z, w4.w);
          }
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c0));
          bs[0] = make_float2(b4.x, b4.y);
          bs[1] = make_float2(b4.z, b4.w);
          uint32_t src = sH;
          uint32_t dst = sG + goff - 2u * FF_PW * 128u;          // row of output r - 2
          float2 acc[3][2];
#pragma unroll
          for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
          const int ngroups = (p.TH + 2) / 3;
#pragma unroll 1
          for (int g = 0; g < ngroups; ++g) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int r = 3 * g + j;                   // patch row: feeds outputs r (ky=0), r-1 (ky=1), r-2 (ky=2)
              float2 v[3][2];
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) unpack_bf16x4(lds64(src + off[kx]), v[kx]);
              src += FF_PW * 128u;
              float2* aN = acc[j];
              float2* aM = acc[(j + 2) % 3];
              float2* aD = acc[(j + 1) % 3];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                aN[k] = __ffma2_rn(wv[2][k], v[2][k], __ffma2_rn(wv[1][k], v[1][k], __ffma2_rn(wv[0][k], v[0][k], bs[k])));
                aM[k] = __ffma2_rn(wv[5][k], v[2][k], __ffma2_rn(wv[4][k], v[1][k], __ffma2_rn(wv[3][k], v[0][k], aM[k])));
                aD[k] = __ffma2_rn(wv[8][k], v[2][k], __ffma2_rn(wv[7][k], v[1][k], __ffma2_rn(wv[6][k], v[0][k], aD[k])));
              }
              const float2 o0 = gelu_erf2(aD[0]), o1 = gelu_erf2(aD[1]);
              __nv_bfloat162 h0 = __floats2bfloat162_rn(o0.x, o0.y), h1 = __floats2bfloat162_rn(o1.x, o1.y);
              if (r >= 2)
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                             "r"(*reinterpret_cast<uint32_t*>(&h1)) : "memory");
              dst += FF_PW * 128u;
            }
          }
        }
        mark(6);
        fence_proxy_async();                                      // g tile (generic stores) -> MMA3 (async proxy)
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(FF_CTHREADS) : "memory");
        if (tid == 0) mbar_arrive(g_full);
        mark(7);
      }
      ptx = tx; pty = ty; pb = b;
    }
    if (S > 0) {
      warp_wait(mma3_done, (S - 1) & 1, lane);
      tc_fence_after();
      epilogue3(ptx, pty, pb, ntiles - 1);
    }
    if (DBG && tid == 0) {
      for (int i = 0; i < 8; ++i) p.dbg[blockIdx.x * 8 + i] = (unsigned long long)tph[DBG ? i : 0];
      for (int i = 0; i < 4; ++i) p.dbg[(gridDim.x + blockIdx.x) * 8 + 4 + i] = (unsigned long long)tph[DBG ? 8 + i : 0];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FF_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static bool ffn_fused_enabled() {
  static int enabled = -1;                // debugging aid: RAWFORMER_B200_NO_FUSED_FFN=1 keeps the three-kernel FFN
  if (enabled < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_FUSED_FFN");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled != 0;
}

bool ffn_fused_supported(const Ctx& ctx, int C, int W) {
  // C = 64 (two hidden chunks per tile, smaller tiles) works but measures slower than the three kernels it replaces
  // (314 vs 250 us at stage 1 of RawFormer-S): opt-in with RAWFORMER_B200_FUSED_FFN_C64=1
  static int c64 = -1;
  if (c64 < 0) {
    const char* e = getenv("RAWFORMER_B200_FUSED_FFN_C64");
    c64 = (e && e[0] == '1') ? 1 : 0;
  }
  return ffn_fused_enabled() && tcgen05_enabled() && ctx.dtype == RF_BF16 && ctx.band == nullptr && (C == 32 || (C == 64 && c64)) &&
         (W & 1) == 0;
}

// out = x + conv_ffn(norm2(x)), norm2 folded into W1f / cs / b1 (see GemmP); stats: [rows][npart] (sum, sumsq) of x's rows
bool launch_ffn_fused(Ctx& ctx, const void* x, const void* W1f, const float* cs, const float* b1, const float* stats, int npart,
                      const float* dw_w, const float* dw_b, const void* W2, const float* b2, void* out, int B, int H, int W,
                      int C) {
  if (!ffn_fused_supported(ctx, C, W) || npart != 1) return false;
  FfnP p;
  memset(&p, 0, sizeof(p));
  p.cs = cs; p.b1 = b1; p.stats = stats; p.npart = npart; p.invC = 1.0f / (float)C; p.eps = 1e-5f;
  p.dw_w = dw_w; p.dw_b = dw_b; p.b2 = b2; p.x = (const bf16*)x; p.out = (bf16*)out;
  p.H = H; p.W = W; p.C = C; p.B = B;
  p.TH = C == 32 ? 13 : 7;              // (TH + 2) % 3 == 0; sized so that two patch buffers fit (see the smem budget below)
  p.nchunk = 2 * C / FF_CC;
  p.nm1 = cdiv((p.TH + 2) * FF_PW, 128);
  p.nm3 = cdiv(p.TH * FF_PW, 128);
  if (p.nm1 > FF_MAXG || cdiv(p.nm3 * (C / 16), 3) > FF_E3T) return false;
  p.tiles_x = cdiv(W, FF_TW); p.tiles_y = cdiv(H, p.TH);
  const i64 total = (i64)p.tiles_x * p.tiles_y * B;
  if (total > 0x7fffffff || total <= 0) return false;
  p.total_tiles = (int)total;
  p.xrow = (uint32_t)C * 2;
  p.x_bytes = (uint32_t)((p.TH + 2) * FF_PW) * p.xrow;
  p.x_stride = (uint32_t)p.nm1 * 128u * p.xrow;
  p.s_bytes = (uint32_t)((p.TH + 2) * 34) * 8u;
  p.s_stride = (p.s_bytes + 127u) & ~127u;
  p.w_bytes = (uint32_t)p.nchunk * (FF_CC * p.xrow + (uint32_t)C * 128u);
  // residual tile staged by TMA (RAWFORMER_B200_FFN_RTMA=1) instead of 32-byte loads through L2 in E3: measured slower
  // (the box takes > 7000 cycles to land behind the patch loads: 485 vs 470 us per stage-0 launch), so off by default
  static int rtma = -1;
  if (rtma < 0) {
    const char* e = getenv("RAWFORMER_B200_FFN_RTMA");
    rtma = (e && e[0] == '1') ? 1 : 0;
  }
  p.r_bytes = rtma ? (uint32_t)(p.TH * FF_TW * C * 2) : 0u;
  p.acc3_col = p.nm1 * FF_CC;
  int cols = p.acc3_col + p.nm3 * C;
  if (cols > 512) return false;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols *= 2;
  const size_t fixed = 1024 + (size_t)(p.TH + 2) * FF_PW * 128 + (size_t)p.nm3 * 16384 + p.w_bytes + (size_t)5 * C * 4 + 128;
  if (fixed + 2 * ((size_t)p.x_stride + p.s_stride) + ((p.r_bytes + 127u) & ~127u) > 232448) p.r_bytes = 0;   // no room: E3 reads x from L2
  p.nx = fixed + 2 * ((size_t)p.x_stride + p.s_stride) + ((p.r_bytes + 127u) & ~127u) <= 232448 ? 2 : 1;
  const size_t smem = fixed + (size_t)p.nx * ((size_t)p.x_stride + p.s_stride) + ((p.r_bytes + 127u) & ~127u);
  if (smem > 232448) return false;
  CUtensorMap mX, mW1, mW2, mS, mR;
  {
    // statistics as a [B, H, 2W] fp32 tensor (row pitch 8W bytes: W even): the box of a tile is its patch pixels'
    // (sum, sumsq) pairs; zero fill outside the image
    const i64 dS[3] = {2 * (i64)W, H, B};
    const i64 sS[3] = {1, 2 * (i64)W, (i64)2 * W * H};
    const int bS[3] = {2 * 34, p.TH + 2, 1};
    if (((uintptr_t)stats & 15) || !make_map_ex(&mS, stats, 3, dS, sS, bS, 4, 0)) return false;
  }
  const i64 dX[4] = {C, W, H, B};
  const i64 sX[4] = {1, C, (i64)C * W, (i64)C * W * H};
  const int bX[4] = {C, FF_PW, p.TH + 2, 1};
  if (!make_map_ex(&mX, x, 4, dX, sX, bX, 2, (int)p.xrow)) return false;
  const int bR[4] = {C, FF_TW, p.TH, 1};
  if (!make_map_ex(&mR, x, 4, dX, sX, bR, 2, 0)) return false;
  const i64 dW1[3] = {C, 2 * C, 1};
  const i64 sW1[3] = {1, C, (i64)2 * C * C};
  const int bW1[3] = {C, FF_CC, 1};
  if (!make_map_ex(&mW1, W1f, 3, dW1, sW1, bW1, 2, (int)p.xrow)) return false;
  const i64 dW2[3] = {2 * C, C, 1};
  const i64 sW2[3] = {1, 2 * C, (i64)2 * C * C};
  const int bW2[3] = {FF_CC, C, 1};
  if (!make_map_ex(&mW2, W2, 3, dW2, sW2, bW2, 2, 128)) return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_ffn_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_ffn_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return false;
    attr_set = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  const double rows = (double)B * H * W;
  static int dbg_on = -1;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on < 0) {
    const char* e = getenv("RAWFORMER_B200_FFN_DBG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on && cudaMalloc(&dbg_buf, 8 * 8 * 512) != cudaSuccess) dbg_on = 0;
  }
  p.dbg = dbg_on ? dbg_buf : nullptr;
  {
    ScopedLaunch sl(RF_K_FFN_FUSED, rows * C * 2.0 * 2.0 + rows * 8.0 * npart, rows * (8.0 * C * C + 36.0 * C));
    if (dbg_on) launch_pdl(k_ffn_fused<true>, dim3(grid), dim3(FF_THREADS), smem, ctx.stream, mX, mW1, mW2, mS, mR, p);
    else launch_pdl(k_ffn_fused<false>, dim3(grid), dim3(FF_THREADS), smem, ctx.stream, mX, mW1, mW2, mS, mR, p);
  }
  if (dbg_on) {
    static unsigned long long h[8 * 512];
    cudaStreamSynchronize(ctx.stream);
    cudaMemcpy(h, dbg_buf, sizeof(unsigned long long) * 8 * 2 * grid, cudaMemcpyDeviceToHost);
    double a[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) a[j] += (double)h[i * 8 + j] / grid;
    const double steps = (double)cdiv(p.total_tiles, grid) * p.nchunk;
    fprintf(stderr, "[ffn_fused C=%d %dx%d tiles %d TH %d nx %d] cycles/step: prep %.0f wait_mma1 %.0f E1 %.0f bar1 %.0f wait_mma3 %.0f E3 %.0f DW %.0f bar2 %.0f\n",
            C, H, W, p.total_tiles, p.TH, p.nx, a[0] / steps, a[1] / steps, a[2] / steps, a[3] / steps, a[4] / steps, a[5] / steps,
            a[6] / steps, a[7] / steps);
    double cc[4] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 4; ++j) cc[j] += (double)h[(grid + i) * 8 + j] / grid;
    double e3[4] = {0};
    for (int i = 0; i < grid; ++i)
      for (int j = 0; j < 4; ++j) e3[j] += (double)h[(grid + i) * 8 + 4 + j] / grid;
    fprintf(stderr, "   E3 detail: before %.0f r_full %.0f tmem_ld %.0f math+store %.0f\n", e3[0] / steps, e3[1] / steps, e3[2] / steps, e3[3] / steps);
    fprintf(stderr, "   control: wait_acc1_free %.0f issue_mma1(+x wait) %.0f wait_g_full %.0f issue_mma3 %.0f\n", cc[0] / steps, cc[1] / steps,
            cc[2] / steps, cc[3] / steps);
  }
  return true;
}

}  // namespace rf
