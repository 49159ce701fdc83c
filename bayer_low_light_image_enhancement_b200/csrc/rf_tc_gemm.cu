// tcgen05 / TMEM / TMA contraction kernel for the bf16 mode (sm_100a only).
//
//   D[128 pixels x BN] (fp32, tensor memory) += A[128 x 64] (smem, TMA, 128B swizzle) * W[BN x 64]^T (smem, TMA)
//
// One CTA = one 128-row pixel tile x one BN-column slice.  Warp roles (192 threads):
//   warp 0   : TMA producer  (one elected lane; full/empty mbarrier ring of `stages` k-blocks)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma.cta_group::1.kind::f16)
//   warps 2-9: epilogue  (tcgen05.ld 32x32b -> bias / activation / residual -> bf16 -> global); two warps share
//              each TMEM lane quadrant and split the 16-column chunks between them
// A operand forms:
//   rows  : 3-d tensor map [K, M, B]; k-blocks run over A1 then A2 (K-concatenation = the cat fusions)
//   conv3 : 4-d tensor map [C, W, H, B]; the M tile is a th x tw pixel patch and every 3x3 tap is the same box
//           shifted by (dx,dy) -- out-of-bounds rows/columns are zero-filled by TMA = the conv's zero padding.
// K tails (K % 64 != 0) are zero-filled by TMA on both operands; the MMA loop skips all-zero 16-wide slices.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "rf_kernels.cuh"
#include "rf_tma.cuh"

namespace rf {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 320;      // 2 control warps + 8 epilogue warps
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_MAX_STAGES = 8;

struct TcParams {
  const float* bias;
  const bf16* R;
  bf16* Y;
  i64 ldr, ldy;
  int M, N, B;         // rows per image, total output columns, images
  int BN;              // columns per CTA (multiple of 16, <= 256)
  int act, amode, omode;
  int H, W;            // image size of the rows (conv / convT / unshuffle addressing)
  int tw, th;          // conv patch (tw*th == 128)
  int tw_log2;
  int tiles_x;         // conv: patches per image row
  int kb1, kb2;        // k-blocks from A1 (per tap for conv) and A2
  int K1, K2;          // K of A1 (per tap for conv: Cin) and A2
  int taps;            // 1 or 9
  int stages;
  int w_per_image;     // 1: weight tensor map coordinate 2 = image index
  int ksplit;          // split-K factor: blockIdx.z = b*ksplit + split (OMODE_ATOMIC_F32)
  int tmem_cols;       // allocated TMEM columns (two accumulator buffers)
  int acc_cols;        // column offset of the second accumulator buffer
  int tiles_m, tiles_n, total_tiles;
  // halo conv (3x3, Cin <= 64): ONE TMA box per tile fetches the (th+2) x (tw+2) halo patch (pixel-major); a warp
  // re-lays it channel-chunk-major ([Cin/8][pixel][16 B]), which is the canonical NO-SWIZZLE K-major UMMA layout with
  // pixels as rows -- so the nine taps are nine descriptors into the same buffer (start shifted by (dy*pitch+dx)*16 B)
  int halo, halo_bytes, nh, nt, npix, pitch;
  int w_resident;      // all k-blocks of W (one N tile) stay in shared memory for the whole kernel
  int bk;              // K elements per k-block: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
  int tma_store;       // OMODE_ROWS: stage the bf16 output tile in shared memory and write it with TMA
  int nslab;           // 64-column slabs of the staged tile
  int nbuf;            // staging buffers (1 or 2); tile ti uses buffer ti % nbuf
  int r_tma;           // the residual tile is TMA-loaded into the staging buffer and updated in place
  // folded LayerNorm (see GemmP) and row statistics of the output
  const float* ln_stats;
  const float* ln_cs;
  int ln_npart;
  float ln_invC, ln_eps;
  float* stats_out;    // [rows][tiles_n] float2 (sum, sumsq) of the rounded outputs (tma_store mode only)
};

__device__ __forceinline__ uint32_t make_idesc(int n) { return make_idesc_m128(n); }
// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct TileCoord {
  int b, split, n0, m0, px0, py0, kb_begin, nkb;
};

// tile index -> coordinates; N slices of the same pixel tile are adjacent so that they share the A tile through L2
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
  TileCoord tc;
  int ny = 0, r = t;
  if (p.tiles_n > 1) { ny = t % p.tiles_n; r = t / p.tiles_n; }
  int mx = r, z = 0;
  if (p.B * p.ksplit > 1) { mx = r % p.tiles_m; z = r / p.tiles_m; }
  tc.b = z; tc.split = 0;
  if (p.ksplit > 1) { tc.b = z / p.ksplit; tc.split = z - tc.b * p.ksplit; }
  tc.n0 = ny * p.BN;
  tc.m0 = 0; tc.px0 = 0; tc.py0 = 0;
  if (p.amode == AMODE_CONV3) {
    const int ty = mx / p.tiles_x;
    tc.px0 = (mx - ty * p.tiles_x) * p.tw;
    tc.py0 = ty * p.th;
  } else {
    tc.m0 = mx * TC_BM;
  }
  const int nkb_all = p.taps * p.kb1 + p.kb2;
  tc.kb_begin = 0; tc.nkb = nkb_all;
  if (p.ksplit > 1) {
    const int kb_per = (nkb_all + p.ksplit - 1) / p.ksplit;
    tc.kb_begin = tc.split * kb_per;
    tc.nkb = min(nkb_all, tc.kb_begin + kb_per) - tc.kb_begin;
  }
  return tc;
}

// Persistent, warp-specialised: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA producer
// runs ahead across tile boundaries through the smem ring; the accumulator is double-buffered in tensor memory so the
// epilogue of tile i overlaps the MMAs of tile i+1.
template <int OM, bool LN, bool HR, bool ST, int ACT, bool HALO>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_tc_gemm(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
          const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapY,
          const __grid_constant__ CUtensorMap mapR, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A tile 16 KB][stages x W tile BN*128 B][nbuf x nslab x 16 KB residual/output staging][barriers]
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = TC_BM * 2u * p.bk, w_bytes = (uint32_t)p.BN * 2u * p.bk;
  const uint32_t halo_stride = ((uint32_t)p.halo_bytes + 127u) & ~127u;
  const uint32_t sA = base;
  const uint32_t sW = base + (HALO ? (((uint32_t)(p.nh + p.nt) * halo_stride + 1023u) & ~1023u) : p.stages * a_bytes);
  const uint32_t sT = sA + (uint32_t)p.nh * halo_stride;   // halo conv: re-laid-out patches
  const uint32_t sY = sW + (p.w_resident ? 0u : p.stages * w_bytes);            // 1024-aligned (a_bytes, w_bytes are multiples of 1024)
  const uint32_t y_bytes = (uint32_t)p.nslab * 16384u;
  const uint32_t sWres = sY + (uint32_t)p.nbuf * y_bytes; // resident W: (taps*kb1 + kb2) k-blocks of w_bytes (1024-aligned)
  const uint32_t sStat = sWres + (p.w_resident ? (uint32_t)(p.taps * p.kb1 + p.kb2) * w_bytes : 0u);  // [128] float2 (stats_out)
  const uint32_t bars = sStat + (p.stats_out ? 1024u : 0u);  // 8-byte aligned
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  auto yfull_bar = [&](int q) { return bars + 8u * (2 * TC_MAX_STAGES + 4 + q); };
  auto yempty_bar = [&](int q) { return bars + 8u * (2 * TC_MAX_STAGES + 8 + q); };
  const uint32_t wres_bar = bars + 8u * (2 * TC_MAX_STAGES + 12);
  auto afull_bar = [&](int q) { return bars + 8u * (2 * TC_MAX_STAGES + 13 + q); };
  auto aempty_bar = [&](int q) { return bars + 8u * (2 * TC_MAX_STAGES + 17 + q); };
  const uint32_t tmem_slot = bars + 8u * (2 * TC_MAX_STAGES + 21);
  const uint32_t sDesc = bars + 8u * (2 * TC_MAX_STAGES + 22);   // halo conv: precomputed UMMA descriptor pairs (<= 1 KB)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA1);
    if (p.kb2) tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapW);
    if (p.tma_store) tma_prefetch_desc(&mapY);
    if (p.r_tma) tma_prefetch_desc(&mapR);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS);      // one arrival per epilogue warp
    }
    for (int q = 0; q < 4; ++q) {
      mbar_init(yfull_bar(q), 1);
      mbar_init(yempty_bar(q), 1);
    }
    mbar_init(wres_bar, 1);
    for (int q = 0; q < 4; ++q) {
      mbar_init(afull_bar(q), 1);
      mbar_init(aempty_bar(q), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();        // (after the TMEM allocation, see launch_pdl)
  pdl_wait();           // everything above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0 && HALO) {
    // ================= halo conv: TMA producer =================
    // keeps nh patch loads in flight; the patches are re-laid-out by the epilogue warps (256 threads, a few 16-byte units
    // each), which release the landing slot through empty_bar
    if (lane == 0) {
      if (p.w_resident) {
        const int nkb_all = p.taps * p.kb1;
        mbar_expect_tx(wres_bar, (uint32_t)nkb_all * w_bytes);
        for (int i = 0; i < nkb_all; ++i) {
          const int tap = i / p.kb1, cb = i - tap * p.kb1;
          tma_load_3d(sWres + i * w_bytes, &mapW, wres_bar, cb * p.bk, tap, 0);
        }
      }
      int s = 0, wrap = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tn = decode_tile(p, t);
        if (wrap > 0) mbar_wait(empty_bar(s), (wrap - 1) & 1);
        mbar_expect_tx(full_bar(s), (uint32_t)p.halo_bytes);
        tma_load_4d(sA + s * halo_stride, &mapA1, full_bar(s), 0, tn.px0 - 1, tn.py0 - 1, tn.b);
        if (++s == p.nh) { s = 0; ++wrap; }
      }
    }
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t tx = a_bytes + (p.w_resident ? 0u : w_bytes);
      int s = 0, wrap = 0;  // ring slot and how often the ring has wrapped
      int ti = 0;
      if (p.w_resident) {
        // the whole weight matrix of this CTA's (only) N tile, once: W rows cost as much TMA time as A rows, and the
        // HBM-bound shapes have as many W rows as A rows per tile
        const int nkb_all = p.taps * p.kb1 + p.kb2;
        mbar_expect_tx(wres_bar, (uint32_t)nkb_all * w_bytes);
        for (int i = 0; i < nkb_all; ++i) {
          const uint32_t dstW = sWres + i * w_bytes;
          if (p.amode == AMODE_CONV3) {
            const int tap = i / p.kb1, cb = i - tap * p.kb1;
            tma_load_3d(dstW, &mapW, wres_bar, cb * p.bk, tap, 0);
          } else if (i < p.kb1) {
            tma_load_3d(dstW, &mapW, wres_bar, i * p.bk, 0, 0);
          } else {
            tma_load_3d(dstW, &mapW, wres_bar, p.K1 + (i - p.kb1) * p.bk, 0, 0);
          }
        }
      }
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        if (tc.nkb <= 0) continue;
        if (p.r_tma) {
          // residual tile -> staging buffer ti % nbuf (the epilogue adds the accumulator in place); the buffer is free
          // once the TMA store of its previous tile has finished reading it (yempty, arrived by the storing thread)
          const int q = ti % p.nbuf, use = ti / p.nbuf;
          if (use > 0) mbar_wait(yempty_bar(q), (use - 1) & 1);
          int ns = 0;
          for (int sl = 0; sl < p.nslab; ++sl) ns += (tc.n0 + sl * 64 < p.N) ? 1 : 0;
          mbar_expect_tx(yfull_bar(q), (uint32_t)ns * 16384u);
          for (int sl = 0; sl < ns; ++sl) {
            const uint32_t dst = sY + q * y_bytes + sl * 16384u;
            if (p.amode == AMODE_CONV3) tma_load_4d(dst, &mapR, yfull_bar(q), tc.n0 + sl * 64, tc.px0, tc.py0, tc.b);
            else tma_load_3d(dst, &mapR, yfull_bar(q), tc.n0 + sl * 64, tc.m0, tc.b);
          }
        }
        ++ti;
        // k-block walk with counters only (this single thread's dependent-instruction chain paces the whole CTA when a
        // tile has many small k-blocks -- no divisions in here)
        int i = tc.kb_begin;
        int tap = 0, cb = i, tdy = 0, tdx = 0;
        if (p.amode == AMODE_CONV3 && i > 0) { tap = i / p.kb1; cb = i - tap * p.kb1; tdy = tap / 3; tdx = tap - 3 * tdy; }
        for (int il = 0; il < tc.nkb; ++il, ++i) {
          if (wrap > 0) mbar_wait(empty_bar(s), (wrap - 1) & 1);
          mbar_expect_tx(full_bar(s), tx);
          const uint32_t dstA = sA + s * a_bytes, dstW = sW + s * w_bytes;
          if (p.amode == AMODE_CONV3) {
            tma_load_4d(dstA, &mapA1, full_bar(s), cb * p.bk, tc.px0 + tdx - 1, tc.py0 + tdy - 1, tc.b);
            if (!p.w_resident) tma_load_3d(dstW, &mapW, full_bar(s), cb * p.bk, tap, tc.n0);
            if (++cb == p.kb1) {
              cb = 0; ++tap;
              if (++tdx == 3) { tdx = 0; ++tdy; }
            }
          } else if (i < p.kb1) {
            tma_load_3d(dstA, &mapA1, full_bar(s), i * p.bk, tc.m0, tc.b);
            if (!p.w_resident) tma_load_3d(dstW, &mapW, full_bar(s), i * p.bk, tc.n0, p.w_per_image ? tc.b : 0);
          } else {
            const int j = i - p.kb1;
            tma_load_3d(dstA, &mapA2, full_bar(s), j * p.bk, tc.m0, tc.b);
            if (!p.w_resident) tma_load_3d(dstW, &mapW, full_bar(s), p.K1 + j * p.bk, tc.n0, p.w_per_image ? tc.b : 0);
          }
          if (++s == p.stages) { s = 0; ++wrap; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the tile / k-block loops (warp-uniform control flow and descriptor arithmetic, which the compiler
    // keeps on the uniform datapath) and ONE elected lane issues the tcgen05 instructions.  Under a plain `lane == 0` branch
    // every tcgen05.mma was wrapped in a loop over the active lanes with vector -> uniform register moves: ~20 instructions
    // per MMA on a thread that shares its scheduler with busy epilogue warps (measured in rf_lnconv.cu: 95-135 cycles per MMA
    // instead of the tensor pipe's 59).
    {
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
      const bool leader = elected != 0;
      auto umma_f16 = [&](uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc_, uint32_t accumulate) {
        if (leader) rf::umma_f16(d_tmem, adesc, bdesc, idesc_, accumulate);
      };
      auto umma_commit = [&](uint32_t bar) {
        if (leader) rf::umma_commit(bar);
      };
      const uint32_t idesc = make_idesc(p.BN);
      int s = 0, wrap = 0, ti = 0;
      if (HALO) {
        // table of (A descriptor relative to the patch buffer, B descriptor) for the 9 taps x Cin/16 k-steps
        const uint64_t dbase0 = ((uint64_t)((uint32_t)(p.npix * 16) >> 4) << 16) | ((uint64_t)((uint32_t)(p.pitch * 16) >> 4) << 32) |
                                ((uint64_t)1 << 46);
        const int ksteps = p.K1 >> 4, kpb = p.bk >> 4;
        const uint64_t wdesc0 = make_kmajor_desc(sWres, p.bk);
        const uint32_t wstep = w_bytes >> 4;
        int i = 0;
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - 3 * dy;
          for (int kk = 0; kk < ksteps; ++kk, ++i) {
            const uint64_t ad = dbase0 | (uint64_t)(uint32_t)((dy * p.pitch + dx) + 2 * kk * p.npix);
            const uint64_t bd = wdesc0 + (uint64_t)((tap * p.kb1 + kk / kpb) * wstep + 2u * (kk % kpb));
            if (leader)
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sDesc + (uint32_t)i * 16u), "r"((uint32_t)ad),
                           "r"((uint32_t)(ad >> 32)), "r"((uint32_t)bd), "r"((uint32_t)(bd >> 32))
                           : "memory");
          }
        }
      }
      __syncwarp();
      if (p.w_resident) {
        mbar_wait(wres_bar, 0);
        tc_fence_after();
      }
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        if (tc.nkb <= 0) continue;
        const int acc = ti & 1, u = ti >> 1;
        if (u > 0) {                                  // wait until the epilogue drained this accumulator buffer
          mbar_wait(tempty_bar(acc), (u - 1) & 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
        if (HALO) {
          // nine taps x Cin/16 k-steps straight out of the re-laid-out patch: rows = pixels (8-row groups = patch rows,
          // SBO = pitch*16 B), the two 8-channel chunks of a k-step are LBO = npix*16 B apart, no swizzle
          mbar_wait(afull_bar(s), wrap & 1);
          tc_fence_after();
          const uint32_t At = sT + s * halo_stride;
          const uint64_t dbase = ((uint64_t)((uint32_t)(p.npix * 16) >> 4) << 16) | ((uint64_t)((uint32_t)(p.pitch * 16) >> 4) << 32) |
                                 ((uint64_t)1 << 46);
          // the single issuing thread's instruction chain per MMA is the critical path of the CTA (a tcgen05.mma costs
          // ~59 issue cycles by itself): the tap / k-step loops are fully unrolled for the four channel counts of the
          // halo form, so every descriptor is the base plus a compile-time constant (patch pitch 10, 180 patch pixels)
          const uint32_t aoff = (At & 0x3FFFF) >> 4;
          const uint64_t abase = dbase + (uint64_t)aoff;
          const uint64_t wdesc0 = make_kmajor_desc(sWres, p.bk);
          const uint32_t wstep = w_bytes >> 4;
          auto issue_all = [&](auto KS, auto KPB) {          // one k-block per tap in the halo form (Cin <= 64)
            constexpr int ksteps = decltype(KS)::value, kpb = decltype(KPB)::value;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t adesc = abase + (uint64_t)(uint32_t)(((tap / 3) * 10 + (tap % 3)) + 2 * kk * 180);
                const uint64_t bdesc = wdesc0 + (uint64_t)((uint32_t)(tap + kk / kpb) * wstep + 2u * (uint32_t)(kk % kpb));
                umma_f16(d_tmem, adesc, bdesc, idesc, (tap | kk) ? 1u : 0u);
              }
            }
          };
          const bool std_patch = p.pitch == 10 && p.npix == 180 && p.kb1 == 1;
          if (std_patch && p.K1 == 32 && p.bk == 32) issue_all(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{});
          else if (std_patch && p.K1 == 64 && p.bk == 64) issue_all(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{});
          else if (std_patch && p.K1 == 48 && p.bk == 64) issue_all(std::integral_constant<int, 3>{}, std::integral_constant<int, 4>{});
          else if (std_patch && p.K1 == 16 && p.bk == 32) issue_all(std::integral_constant<int, 1>{}, std::integral_constant<int, 2>{});
          else {
            const int nmma = 9 * (p.K1 >> 4);
            for (int i = 0; i < nmma; ++i) {
              uint32_t alo, ahi, blo, bhi;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(alo), "=r"(ahi), "=r"(blo), "=r"(bhi)
                           : "r"(sDesc + (uint32_t)i * 16u));
              const uint64_t adesc = ((uint64_t)ahi << 32) | (uint64_t)(alo + aoff);
              const uint64_t bdesc = ((uint64_t)bhi << 32) | (uint64_t)blo;
              umma_f16(d_tmem, adesc, bdesc, idesc, i ? 1u : 0u);
            }
          }
          umma_commit(aempty_bar(s));         // the patch buffer is free when these MMAs retire
          if (++s == p.nt) { s = 0; ++wrap; }
          umma_commit(tfull_bar(acc));
          ++ti;
          continue;
        }
        int i = tc.kb_begin;
        int cb = p.amode == AMODE_CONV3 ? i % p.kb1 : 0;
        for (int il = 0; il < tc.nkb; ++il, ++i) {
          mbar_wait(full_bar(s), wrap & 1);
          tc_fence_after();
          // valid K in this block (zero-filled beyond): skip all-zero 16-wide slices
          int kvalid;
          if (p.amode == AMODE_CONV3) {
            kvalid = min(p.bk, p.K1 - cb * p.bk);
            if (++cb == p.kb1) cb = 0;
          } else if (i < p.kb1) {
            kvalid = min(p.bk, p.K1 - i * p.bk);
          } else {
            kvalid = min(p.bk, p.K2 - (i - p.kb1) * p.bk);
          }
          const int ksteps = (kvalid + 15) >> 4;
          const uint64_t adesc = make_kmajor_desc(sA + s * a_bytes, p.bk);
          const uint64_t bdesc = make_kmajor_desc(p.w_resident ? sWres + i * w_bytes : sW + s * w_bytes, p.bk);
          for (int k = 0; k < ksteps; ++k) {
            // advance 32 B (16 bf16) inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (il | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));          // frees the smem stage when these MMAs retire
          if (++s == p.stages) { s = 0; ++wrap; }
        }
        umma_commit(tfull_bar(acc));          // accumulator complete
        ++ti;
      }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    // Specialised at compile time (OM, LN, HR, ST, ACT): the tiles of the HBM-bound shapes are small (128 x 32..96), so
    // the epilogue's instruction count per tile bounds the kernel -- no runtime mode switches inside the chunk code.
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;       // which half of the 16-column chunks this warp handles
    const int r = quad * 32 + lane;         // row of the tile
    const uint32_t sw = (uint32_t)(r & 7);  // 128-byte-swizzle phase of this row in the staging tile
    const int rty = r >> p.tw_log2, rtx = r & (p.tw - 1);   // conv patch coordinates of this row
    int ti = 0;
    float4 ln_next = make_float4(0.f, 0.f, 0.f, 0.f);
    auto row_of = [&](const TileCoord& tn, bool& ok, int& oy, int& ox) -> i64 {
      if (p.amode == AMODE_CONV3) {
        oy = tn.py0 + rty;
        ox = tn.px0 + rtx;
        ok = oy < p.H && ox < p.W;
        return (i64)tn.b * p.M + (i64)oy * p.W + ox;
      }
      const int m = tn.m0 + r;
      ok = m < p.M;
      if (OM == OMODE_CONVT || OM == OMODE_UNSHUFFLE) { oy = m / p.W; ox = m - oy * p.W; }
      return (i64)tn.b * p.M + m;
    };
    // (sum, sumsq) of this thread's row in tile t (folded LayerNorm); rows outside the tensor give zeros
    auto ln_fetch = [&](int t) {
      const TileCoord tn = decode_tile(p, t);
      bool ok;
      int yy, xx;
      const i64 row = row_of(tn, ok, yy, xx);
      // NO arithmetic on the loaded values here: they are consumed one tile later, so the load latency is hidden
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + row * p.ln_npart;
        const float2 a = __ldg(st);
        v.x = a.x; v.y = a.y;
        if (p.ln_npart > 1) {
          const float2 c = __ldg(st + 1);
          v.z = c.x; v.w = c.y;
          for (int q = 2; q < p.ln_npart; ++q) {   // (more than two N tiles of partials: not a hot shape)
            const float2 t2 = __ldg(st + q);
            v.z += t2.x; v.w += t2.y;
          }
        }
      }
      return v;
    };
    // halo conv: the 8 epilogue warps re-lay patch j (pixel-major, as TMA delivers it) into the channel-chunk-major
    // [Cin/8][pixel][16 B] form (the no-swizzle K-major UMMA layout) one tile AHEAD of their epilogue, so the MMAs of tile
    // j run under the epilogue of tile j-1
    const int etid = threadIdx.x - 64;             // 0..255 over the epilogue warps
    int rl_j = 0, rl_sh = 0, rl_wh = 0, rl_st = 0, rl_wt = 0;
    const int rl_total = HALO ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto relayout_next = [&]() {
      if (rl_j >= rl_total) return;                // (uniform over the epilogue warps)
      const int nch = p.K1 >> 3, items = p.npix * nch;     // 16-byte units per pixel (2, 4, 6 or 8)
      const bool pow2 = (nch & (nch - 1)) == 0;
      const int nch_log2 = nch == 8 ? 3 : (nch == 4 ? 2 : 1);
      mbar_wait_sleep(full_bar(rl_sh), rl_wh & 1);                        // the patch has landed
      if (rl_wt > 0) mbar_wait_sleep(aempty_bar(rl_st), (rl_wt - 1) & 1); // its destination has been consumed by the MMAs
      const uint32_t src = sA + rl_sh * halo_stride, dst = sT + rl_st * halo_stride;
      for (int it = etid; it < items; it += 256) {
        int q, c8;
        if (pow2) { q = it >> nch_log2; c8 = it & (nch - 1); }
        else { q = it / nch; c8 = it - q * nch; }
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(src + (uint32_t)it * 16u));
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(c8 * p.npix + q) * 16u), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      }
      fence_proxy_async();                                   // generic stores -> tensor-core (async proxy) reads
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (etid == 0) {
        mbar_arrive(afull_bar(rl_st));
        mbar_arrive(empty_bar(rl_sh));
      }
      ++rl_j;
      if (++rl_sh == p.nh) { rl_sh = 0; ++rl_wh; }
      if (++rl_st == p.nt) { rl_st = 0; ++rl_wt; }
    };
    if (HALO) relayout_next();                   // tile 0
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      if (tc.nkb <= 0) continue;
      const int acc = ti & 1, u = ti >> 1;
      const int b = tc.b, n0 = tc.n0;
      if (HALO) relayout_next();                 // tile ti + 1
      // the storing thread hands the staging buffer of the previous tile back to the producer as soon as its TMA store
      // has finished READING it -- before waiting for this tile's accumulator, so that the producer (which fetches the
      // residual of a later tile into that buffer before issuing that tile's operands) can never wait on this tile
      if (HR && warp == 2 && lane == 0 && ti > 0) {
        tma_store_wait_read<0>();
        mbar_arrive(yempty_bar((ti - 1) % p.nbuf));
      }
      // folded LayerNorm: this tile's row statistics were fetched one tile ago (the global-load latency hides under the
      // previous tile's epilogue); fetch the next tile's now
      float ln_rs = 1.f, ln_nm = 0.f;
      if (LN) {
        if (ti == 0) ln_next = ln_fetch(t);
        const float4 cur = ln_next;
        if (t + (int)gridDim.x < p.total_tiles) ln_next = ln_fetch(t + gridDim.x);
        const float mu = (cur.x + cur.z) * p.ln_invC;
        ln_rs = rsqrtf(fmaxf((cur.y + cur.w) * p.ln_invC - mu * mu, 0.f) + p.ln_eps);
        ln_nm = -ln_rs * mu;
      }
      bool row_ok;
      int oy = 0, ox = 0;
      const i64 orow = row_of(tc, row_ok, oy, ox);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      constexpr bool STAGED = OM == OMODE_ROWS || OM == OMODE_UNSHUFFLE;   // output tile staged in smem, TMA store
      const int yq = ti % p.nbuf;
      const uint32_t srow = sY + yq * y_bytes + (uint32_t)r * 128u;   // this row in slab 0 of the staging buffer
      mbar_wait_sleep(tfull_bar(acc), u & 1);
      tc_fence_after();
      if (STAGED) {
        if (HR) {
          mbar_wait_sleep(yfull_bar(yq), (ti / p.nbuf) & 1); // this tile's residual has landed in the staging buffer
        } else {
          // the staging buffer is free once the TMA store of tile ti - nbuf has finished READING it
          if (warp == 2 && lane == 0) {
            if (p.nbuf >= 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      float st_sum = 0.f, st_sq = 0.f;
      // one 16-column chunk of this thread's row: v = its accumulator values
      auto chunk = [&](const int c, const uint32_t (&v)[16]) {
        const int n = n0 + c;
        const int nvalid = p.N - n;          // < 16 only in the last chunk when N % 16 == 8
        if (nvalid <= 0) return;
        if (OM != OMODE_ROWS && OM != OMODE_UNSHUFFLE && !row_ok) return;   // (staged modes: TMA clips outside the tensor)
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (LN) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (j < nvalid) {
              const float4 cq = __ldg(reinterpret_cast<const float4*>(p.ln_cs + n + j));
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              f[j] = fmaf(ln_rs, f[j], fmaf(ln_nm, cq.x, bq.x));
              f[j + 1] = fmaf(ln_rs, f[j + 1], fmaf(ln_nm, cq.y, bq.y));
              f[j + 2] = fmaf(ln_rs, f[j + 2], fmaf(ln_nm, cq.z, bq.z));
              f[j + 3] = fmaf(ln_rs, f[j + 3], fmaf(ln_nm, cq.w, bq.w));
            }
          }
        } else if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (j < nvalid) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
              f[j] += bq.x; f[j + 1] += bq.y; f[j + 2] += bq.z; f[j + 3] += bq.w;
            }
          }
        }
        if (ACT == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.2f * f[j]);
        } else if (ACT == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (ACT == ACT_TANH_RES) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = 0.2f * tanh_fast(f[j]);
        }
        if (OM == OMODE_ROWS) {
          // staging tile: 64-column slabs of 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk
          // position j ^ (r & 7); a quarter warp (8 consecutive rows) hits all 32 banks once
          const uint32_t rowp = srow + (uint32_t)(c >> 6) * 16384u;
          const uint32_t j0 = (uint32_t)(c & 63) >> 3;
          const uint32_t a0 = rowp + ((j0 ^ sw) << 4), a1 = rowp + (((j0 + 1) ^ sw) << 4);
          if (HR) {                                          // residual, TMA-loaded into the same place
            uint4 rr0, rr1;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr0.x), "=r"(rr0.y), "=r"(rr0.z), "=r"(rr0.w) : "r"(a0));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr1.x), "=r"(rr1.y), "=r"(rr1.z), "=r"(rr1.w) : "r"(a1));
            const uint32_t w0[4] = {rr0.x, rr0.y, rr0.z, rr0.w}, w1[4] = {rr1.x, rr1.y, rr1.z, rr1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[2 * j] += __uint_as_float(w0[j] << 16);
              f[2 * j + 1] += __uint_as_float(w0[j] & 0xffff0000u);
              f[8 + 2 * j] += __uint_as_float(w1[j] << 16);
              f[8 + 2 * j + 1] += __uint_as_float(w1[j] & 0xffff0000u);
            }
          }
          uint32_t q0[4], q1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
            q0[j] = *reinterpret_cast<uint32_t*>(&h0);
            q1[j] = *reinterpret_cast<uint32_t*>(&h1);
          }
          if (ST) {                                          // statistics of the ROUNDED values (what the consumer reads)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x0 = __uint_as_float(q0[j] << 16), x1 = __uint_as_float(q0[j] & 0xffff0000u);
              st_sum += x0 + x1;
              st_sq = fmaf(x0, x0, fmaf(x1, x1, st_sq));
            }
            if (nvalid > 8) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x0 = __uint_as_float(q1[j] << 16), x1 = __uint_as_float(q1[j] & 0xffff0000u);
                st_sum += x0 + x1;
                st_sq = fmaf(x0, x0, fmaf(x1, x1, st_sq));
              }
            }
          }
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(q0[0]), "r"(q0[1]), "r"(q0[2]), "r"(q0[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(q1[0]), "r"(q1[1]), "r"(q1[2]), "r"(q1[3]) : "memory");
        } else if (OM == OMODE_ATOMIC_F32) {
          float* yf = reinterpret_cast<float*>(p.Y) + orow * p.ldy + n;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n + j < p.N) atomicAdd(yf + j, f[j]);
        } else if (OM == OMODE_HEAD) {
          // conv_out: channel n = 4*ch + 2*i + j of packed pixel (oy, ox) -> out[b, ch, 2*oy + i, 2*ox + j], fp32 NCHW
          float* outp = reinterpret_cast<float*>(p.Y);
          const i64 Wo = 2 * (i64)p.W, Ho = 2 * (i64)p.H;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float a = f[4 * ch + 2 * i], bq = f[4 * ch + 2 * i + 1];
              *reinterpret_cast<float2*>(outp + (((i64)b * 3 + ch) * Ho + 2 * oy + i) * Wo + 2 * ox) =
                  make_float2(fmaxf(a, 0.2f * a), fmaxf(bq, 0.2f * bq));
            }
        } else if (OM == OMODE_CONVT) {
          const int Co = p.N >> 2;
          const int ij = n / Co, co = n - ij * Co;   // 16-wide chunk never straddles ij (Co % 16 == 0)
          const i64 dst = (((i64)b * 2 * p.H + 2 * oy + (ij >> 1)) * (2 * p.W) + 2 * ox + (ij & 1)) * p.ldy + co;
          float o8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = f[j];
          store8(p.Y + dst, o8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = f[8 + j];
          store8(p.Y + dst + 8, o8);
        } else {
          // pixel-unshuffle: source pixel (rty, rtx) of the patch feeds output pixel (rty/2, rtx/2) of the (th/2 x tw/2)
          // output tile, channel n*4 + 2*(rty&1) + (rtx&1).  The output tile is staged as 64-channel slabs of 128-byte
          // rows (SWIZZLE_128B over the output-pixel index) and written with TMA instead of 2-byte scattered stores.
          const int opix = (rty >> 1) * (p.tw >> 1) + (rtx >> 1);
          const int sub = 2 * (rty & 1) + (rtx & 1);
          const uint32_t obase = sY + yq * y_bytes + (uint32_t)opix * 128u;
          const uint32_t osw = (uint32_t)(opix & 7);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nvalid) {
              const int ch = (c + j) * 4 + sub;                     // channel inside this N tile's 4*BN output channels
              const uint32_t a = obase + (uint32_t)(ch >> 6) * 16384u + ((((uint32_t)(ch & 63) >> 3) ^ osw) << 4) +
                                 (uint32_t)(ch & 7) * 2u;
              const __nv_bfloat16 hv = __float2bfloat16_rn(f[j]);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
            }
          }
        }
      };
      // two chunks per TMEM round trip (32 accumulator registers in flight)
      for (int c = half * 16; c < p.BN; c += 64) {
        uint32_t va[16], vb[16];
        const bool two = c + 32 < p.BN;      // warp-uniform
        tmem_ld16(taddr + c, va);
        if (two) tmem_ld16(taddr + c + 32, vb);
        tmem_ld_wait();
        chunk(c, va);
        if (two) chunk(c + 32, vb);
      }
      // this warp is done reading the accumulator buffer: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (STAGED) {
        if (ST && half == 1)
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sStat + 8u * r), "f"(st_sum), "f"(st_sq) : "memory");
        fence_proxy_async();                               // staged tile (generic stores) -> TMA store (async proxy)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ST && half == 0 && row_ok) {
          float ox2, oy2;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ox2), "=f"(oy2) : "r"(sStat + 8u * r));
          reinterpret_cast<float2*>(p.stats_out)[orow * p.tiles_n + n0 / p.BN] = make_float2(st_sum + ox2, st_sq + oy2);
        }
        if (warp == 2 && lane == 0) {
          const uint32_t sYq = sY + yq * y_bytes;
          if (OM == OMODE_UNSHUFFLE) {
            for (int sl = 0; sl < p.nslab; ++sl) {
              if (4 * n0 + sl * 64 >= 4 * p.N) break;
              tma_store_4d(&mapY, sYq + sl * 16384u, 4 * n0 + sl * 64, tc.px0 >> 1, tc.py0 >> 1, b);
            }
          } else {
            for (int sl = 0; sl < p.nslab; ++sl) {
              if (n0 + sl * 64 >= p.N) break;
              if (p.amode == AMODE_CONV3) tma_store_4d(&mapY, sYq + sl * 16384u, n0 + sl * 64, tc.px0, tc.py0, b);
              else tma_store_3d(&mapY, sYq + sl * 16384u, n0 + sl * 64, tc.m0, b);
            }
          }
          tma_store_commit();
        }
      }
      ++ti;
    }
  }
  if ((OM == OMODE_ROWS || OM == OMODE_UNSHUFFLE) && warp == 2 && lane == 0) tma_store_wait_all<0>();   // smem must outlive the last store's reads
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int pick_bn(int N) {
  if (N % 8) return 0;
  if (N <= 256) return (N + 15) & ~15;   // N % 16 == 8: the last 8 columns are zero weights (TMA OOB) and masked stores
  if (N % 16) return 0;
  // several N tiles: multiples of 64 (the TMA-store slabs of neighbouring tiles must not overlap); the last tile may
  // be ragged (W rows beyond N are zero-filled by TMA, stores are clipped).  Least padded work, then the widest tile.
  int best = 0;
  i64 best_pad = 0;
  for (int bn = 256; bn >= 128; bn -= 64) {
    const i64 pad = (i64)((N + bn - 1) / bn) * bn;
    if (best == 0 || pad < best_pad) { best = bn; best_pad = pad; }
  }
  return best;
}


// ---------------------------------------------------------------------------------------------
// Gram kernel of the transposed attention:  G[i][j] += sum_p q[p][i] * k[p][j]   (FLCA_RF.py:230, before softmax)
// q,k live interleaved in the NHWC depthwise output [P][2C] (q = channels [0,C), k = [C,2C)).  Pixels are the
// contraction axis, so both operands are "MN-major" for UMMA: a TMA box of 64 channels x 128 pixels lands as 128 rows
// (pixels) of 128 B, which is exactly the canonical MN-major SWIZZLE_128B atom stack (SBO = 1024 B between 8-pixel
// groups, LBO = one 16 KB chunk tile between 64-channel groups).  One CTA owns a 128x128 tile of G and a slice of the
// pixel range (split-K); the accumulator stays in TMEM over the whole slice and only the per-head diagonal blocks
// are added to global memory.
// ---------------------------------------------------------------------------------------------
constexpr int GR_STAGES = 8;                   // maximum ring depth (barrier slots); the launcher picks nst <= 8
constexpr int GR_THREADS = 192;               // TMA warp, MMA warp, 4 epilogue warps
constexpr int GR_PIX = 128;                    // pixels per k-block
constexpr uint32_t GR_CHUNK = GR_PIX * 128;    // bytes of one 64-channel x 128-pixel chunk tile

__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes = GR_CHUNK) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;  // LBO: next 64-element group along M/N
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: next 8-row group along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(GR_THREADS)
k_tc_gram(const __grid_constant__ CUtensorMap mapQK, float* __restrict__ G, int C, i64 P, int ksplit, int packed, int nst) {
  // packed = 0: q rows m0.. and k rows n0.. are different channel ranges: 2 + 2 chunks per k-block.
  // packed = 1|2 (2C <= 64 | 128): the whole q|k pixel row fits `packed` 64-channel chunks, A = B = those chunks and the
  //   accumulator is the Gram of [q|k] with itself, whose upper-right block is q^T k -- a quarter / half of the TMA
  //   traffic and shared-memory fill of the general form.  packed = 1 points the second 64-row group of both operands at
  //   a zeroed chunk behind the ring.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (packed ? (uint32_t)packed : 4u) * GR_CHUNK;
  const uint32_t zero_chunk = base + (uint32_t)nst * stage_bytes;          // packed == 1 only
  const uint32_t bars = zero_chunk + (packed == 1 ? GR_CHUNK : 0u);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (GR_STAGES + s); };
  const uint32_t tmem_full = bars + 8u * (2 * GR_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * GR_STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int c = C >> 3;  // head width
  // heads touched by the rows / columns of this tile: skip tiles without a common head
  const int hr0 = m0 / c, hr1 = (min(m0 + 128, C) - 1) / c, hc0 = n0 / c, hc1 = (min(n0 + 128, C) - 1) / c;
  if (!packed && (hr1 < hc0 || hc1 < hr0)) return;
  const int nkb_all = (int)((P + GR_PIX - 1) / GR_PIX);
  const int kb_per = (nkb_all + ksplit - 1) / ksplit;
  const int kb_begin = blockIdx.z * kb_per;
  const int nkb = min(nkb_all, kb_begin + kb_per) - kb_begin;
  if (nkb <= 0) return;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQK);
    for (int s = 0; s < GR_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (packed == 1) {
    for (uint32_t o = threadIdx.x * 16u; o < GR_CHUNK; o += GR_THREADS * 16u)
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(zero_chunk + o), "r"(0u) : "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, wrap = 0;
      for (int il = 0; il < nkb; ++il) {
        if (wrap > 0) mbar_wait(empty_bar(s), (wrap - 1) & 1);
        mbar_expect_tx(full_bar(s), stage_bytes);
        const uint32_t dst = base + s * stage_bytes;
        const int p0 = (kb_begin + il) * GR_PIX;
        if (packed) {
          tma_load_3d(dst, &mapQK, full_bar(s), 0, p0, 0);
          if (packed == 2) tma_load_3d(dst + GR_CHUNK, &mapQK, full_bar(s), 64, p0, 0);
        } else {
          tma_load_3d(dst, &mapQK, full_bar(s), m0, p0, 0);
          tma_load_3d(dst + GR_CHUNK, &mapQK, full_bar(s), m0 + 64, p0, 0);
          tma_load_3d(dst + 2 * GR_CHUNK, &mapQK, full_bar(s), C + n0, p0, 0);
          tma_load_3d(dst + 3 * GR_CHUNK, &mapQK, full_bar(s), C + n0 + 64, p0, 0);
        }
        if (++s == nst) { s = 0; ++wrap; }
      }
    }
  } else if (warp == 1) {
    {
      // the whole warp walks the loop, ONE elected lane issues (see the MMA issuer of k_tc_gemm)
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
      const bool leader = elected != 0;
      // kind::f16, D=f32, A=B=bf16, both MN-major (bits 15, 16), M = N = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      int s = 0, wrap = 0;
      for (int il = 0; il < nkb; ++il) {
        mbar_wait(full_bar(s), wrap & 1);
        tc_fence_after();
        const uint32_t a0 = base + s * stage_bytes, b0 = packed ? a0 : a0 + 2 * GR_CHUNK;
        const uint32_t lbo = packed == 1 ? zero_chunk - a0 : GR_CHUNK;
        const uint64_t adesc = make_sw128_mn_desc(a0, lbo), bdesc = make_sw128_mn_desc(b0, lbo);
        if (leader) {
#pragma unroll
          for (int k = 0; k < GR_PIX / 16; ++k)   // 16 pixels = 16 rows of 128 B = 2048 B per MMA
            umma_f16(tmem_base, adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (2048 >> 4)), idesc, (il | k) ? 1u : 0u);
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == nst) { s = 0; ++wrap; }
      }
      if (leader) umma_commit(tmem_full);
    }
  } else {
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;   // q channel
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int h = row < C ? row / c : -1;
    for (int cc = 0; cc < 128; cc += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + cc, v);
      tmem_ld_wait();
      if (h < 0) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = packed ? cc + j - C : n0 + cc + j;   // k channel (packed: accumulator column = C + k channel)
        if (col >= 0 && col < C && col / c == h)          // this split's slot: part[z][row][col - h*c] (plain store)
          G[((i64)blockIdx.z * C + row) * c + (col - h * c)] = __uint_as_float(v[j]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool g_tc_enabled = true;
static bool g_tc_checked = false;
void set_tcgen05_enabled(bool on) { g_tc_enabled = on; g_tc_checked = true; }
bool tcgen05_enabled() {
  if (!g_tc_checked) {
    g_tc_checked = true;
    const char* e = getenv("RAWFORMER_B200_NO_TCGEN05");
    if (e && e[0] == '1') g_tc_enabled = false;
  }
  return g_tc_enabled;
}

int launch_gemm_tcgen05(Ctx& ctx, const GemmP& g) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16) return -1;
  const int BN = pick_bn(g.N);
  if (BN == 0) return -1;
  if (g.omode != OMODE_ATOMIC_F32 && (g.K1 % 8 || g.K2 % 8)) return -1;  // (Gram: pitch is padded, TMA zero-fills the K tail)
  if (g.lda1 % 8 || (g.A2 && g.lda2 % 8) || (g.ldw % 8)) return -1;
  if (g.omode == OMODE_CONVT && ((g.N / 4) % 16)) return -1;
  if (g.amode == AMODE_CONV3 && g.A2) return -1;
  if (g.amode != AMODE_CONV3 && (g.lda1 < g.K1 || (g.A2 && g.lda2 < g.K2))) return -1;
  if (g.omode == OMODE_ATOMIC_F32 && (g.amode != AMODE_ROWS || g.A2 || g.bias || g.R || g.act != ACT_NONE)) return -1;

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.bias = g.bias; p.R = (const bf16*)g.R; p.Y = (bf16*)g.Y; p.ldr = g.ldr; p.ldy = g.ldy;
  p.M = g.M; p.N = g.N; p.B = g.B; p.BN = BN; p.act = g.act; p.amode = g.amode; p.omode = g.omode; p.H = g.H; p.W = g.W;
  p.w_per_image = g.w_img != 0;
  p.ln_stats = g.ln_stats; p.ln_cs = g.ln_cs; p.ln_npart = g.ln_npart; p.ln_eps = g.ln_eps;
  p.ln_invC = g.ln_C > 0 ? 1.0f / (float)g.ln_C : 0.f;
  CUtensorMap mA1, mA2, mW, mY, mR;
  // k-block width: narrow (32) blocks when every operand source has K <= 32 -- half the smem per stage, so the ring is
  // twice as deep for the same bytes (the 9-tap implicit conv and the C = 32 row GEMMs are TMA-latency bound)
  const int kmax = g.amode == AMODE_CONV3 ? g.K1 / 9 : (g.K1 > g.K2 ? g.K1 : g.K2);
  static int halo_ok = -1;                // debugging aid: RAWFORMER_B200_NO_HALO=1 keeps the per-tap TMA conv
  if (halo_ok < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_HALO");
    halo_ok = (e && e[0] == '1') ? 0 : 1;
  }
  const bool halo_shape = (g.omode == OMODE_ROWS && g.act == ACT_LRELU && !g.R && !g.ln_stats) || g.omode == OMODE_UNSHUFFLE ||
                          g.omode == OMODE_HEAD;
  p.halo = (halo_ok && halo_shape && g.amode == AMODE_CONV3 && (kmax == 16 || kmax == 32 || kmax == 48 || kmax == 64) &&
            g.N <= 256 &&
            (size_t)9 * BN * kmax * 2 <= 81920) ? 1 : 0;
  const int BK = (kmax <= 32 && g.omode != OMODE_ATOMIC_F32) ? 32 : 64;
  p.bk = BK;
  const int kswz = BK * 2;
  const bool unshuf = g.omode == OMODE_UNSHUFFLE && g.amode == AMODE_CONV3 && !(g.H & 1) && !(g.W & 1) && g.ldy % 8 == 0;
  p.tma_store = ((g.omode == OMODE_ROWS && g.ldy % 8 == 0 && (!g.R || g.ldr % 8 == 0)) || unshuf) ? 1 : 0;
  p.nslab = p.tma_store ? cdiv(unshuf ? 4 * BN : BN, 64) : 0;
  p.r_tma = (p.tma_store && g.R != nullptr) ? 1 : 0;
  int grid_x;
  if (g.amode == AMODE_CONV3) {
    const int Cin = g.K1 / 9;
    p.taps = 9; p.K1 = Cin; p.K2 = 0; p.kb1 = cdiv(Cin, BK); p.kb2 = 0;
    // patch shape: minimise padded area
    int best_tw = 16;
    i64 best = -1;
    for (int tw = 8; tw <= 128; tw *= 2) {
      const int th = 128 / tw;
      const i64 area = (i64)cdiv(g.W, tw) * tw * cdiv(g.H, th) * th;
      if (best < 0 || area < best) { best = area; best_tw = tw; }
    }
    if (p.halo) best_tw = 8;              // 8-pixel patch rows = the 8-row core-matrix groups of the UMMA layout
    p.tw = best_tw; p.th = 128 / best_tw;
    p.tw_log2 = 0;
    while ((1 << p.tw_log2) < p.tw) ++p.tw_log2;
    p.tiles_x = cdiv(g.W, p.tw);
    grid_x = p.tiles_x * cdiv(g.H, p.th);
    const i64 dA[4] = {Cin, g.W, g.H, g.B};
    const i64 sA[4] = {1, g.lda1, g.lda1 * g.W, g.lda1 * g.W * g.H};
    const int bA[4] = {BK, p.tw, p.th, 1};
    if (p.halo) {
      p.pitch = p.tw + 2; p.npix = (p.tw + 2) * (p.th + 2);
      p.halo_bytes = p.npix * Cin * 2;
      const int bH[4] = {Cin, p.tw + 2, p.th + 2, 1};
      if (!make_map_ex(&mA1, g.A1, 4, dA, sA, bH, 2, 0)) return -1;
    } else if (!make_map_ex(&mA1, g.A1, 4, dA, sA, bA, 2, kswz)) return -1;
    mA2 = mA1;
    const i64 dW[3] = {Cin, 9, g.N};
    const i64 sW[3] = {1, Cin, (i64)9 * Cin};
    const int bW[3] = {BK, 1, BN};
    if (!make_map_ex(&mW, g.Wt, 3, dW, sW, bW, 2, kswz)) return -1;
    if (unshuf) {
      // output [B, H/2, W/2, 4N]: one (th/2 x tw/2)-pixel box of 64 channels per slab
      const i64 dY[4] = {4 * (i64)g.N, g.W / 2, g.H / 2, g.B};
      const i64 sY[4] = {1, g.ldy, g.ldy * (g.W / 2), g.ldy * (g.W / 2) * (g.H / 2)};
      const int bY[4] = {64, p.tw / 2, p.th / 2, 1};
      if (!make_map(&mY, g.Y, 4, dY, sY, bY)) return -1;
    } else if (p.tma_store) {
      const i64 dY[4] = {g.N, g.W, g.H, g.B};
      const i64 sY[4] = {1, g.ldy, g.ldy * g.W, g.ldy * g.W * g.H};
      const int bY[4] = {64, p.tw, p.th, 1};
      if (!make_map(&mY, g.Y, 4, dY, sY, bY)) return -1;
      if (p.r_tma) {
        const i64 sR[4] = {1, g.ldr, g.ldr * g.W, g.ldr * g.W * g.H};
        if (!make_map(&mR, g.R, 4, dY, sR, bY)) return -1;
      }
    }
  } else {
    p.taps = 1; p.K1 = g.K1; p.K2 = g.A2 ? g.K2 : 0;
    p.kb1 = cdiv(g.K1, BK); p.kb2 = g.A2 ? cdiv(g.K2, BK) : 0;
    p.tw = 128; p.th = 1; p.tw_log2 = 7; p.tiles_x = 1;
    grid_x = cdiv(g.M, TC_BM);
    const i64 dA[3] = {g.K1, g.M, g.B};
    const i64 sA[3] = {1, g.lda1, g.a1_img ? g.a1_img : g.lda1 * g.M};
    const int bA[3] = {BK, TC_BM, 1};
    if (!make_map_ex(&mA1, g.A1, 3, dA, sA, bA, 2, kswz)) return -1;
    if (g.A2) {
      const i64 dA2[3] = {g.K2, g.M, g.B};
      const i64 sA2[3] = {1, g.lda2, g.lda2 * g.M};
      if (!make_map_ex(&mA2, g.A2, 3, dA2, sA2, bA, 2, kswz)) return -1;
    } else {
      mA2 = mA1;
    }
    const i64 K = g.K1 + (g.A2 ? g.K2 : 0);
    const i64 ldw = g.ldw ? g.ldw : K;
    const i64 dW[3] = {K, g.N, g.w_img ? g.B : 1};
    const i64 sW[3] = {1, ldw, g.w_img ? g.w_img : ldw * g.N};
    const int bW[3] = {BK, BN, 1};
    if (!make_map_ex(&mW, g.Wt, 3, dW, sW, bW, 2, kswz)) return -1;
    if (p.tma_store) {
      const i64 dY[3] = {g.N, g.M, g.B};
      const i64 sY[3] = {1, g.ldy, g.ldy * g.M};
      const int bY[3] = {64, TC_BM, 1};
      if (!make_map(&mY, g.Y, 3, dY, sY, bY)) return -1;
      if (p.r_tma) {
        const i64 sR[3] = {1, g.ldr, g.ldr * g.M};
        if (!make_map(&mR, g.R, 3, dY, sR, bY)) return -1;
      }
    }
  }
  if (!p.tma_store) mY = mW;
  if (!p.r_tma) mR = mW;
  p.stats_out = p.tma_store ? g.stats_out : nullptr;
  const int nkb_all = p.taps * p.kb1 + p.kb2;
  p.ksplit = (g.omode == OMODE_ATOMIC_F32 && g.ksplit > 1) ? (g.ksplit < nkb_all ? g.ksplit : nkb_all) : 1;
  // smem: operand ring + residual/output staging.  Two CTAs per SM (two producer / MMA / epilogue sets) when the
  // accumulators fit twice in tensor memory AND a >= 2-deep ring plus the staging fits in half the shared memory.
  // resident W: one N tile, a weight matrix shared by all tiles of the CTA, at most 40 KB
  const int nkb_tile0 = p.taps * p.kb1 + p.kb2;
  const size_t wres_bytes = (size_t)nkb_tile0 * BN * 2 * BK;
  p.w_resident = (cdiv(g.N, BN) == 1 && (!g.w_img || g.B == 1) && wres_bytes <= (p.halo ? 81920u : 40960u) &&
                  g.omode != OMODE_ATOMIC_F32) ? 1 : 0;
  if (p.halo && !p.w_resident) return -1;
  const size_t stage_bytes = ((size_t)TC_BM + (p.w_resident ? 0 : (size_t)BN)) * 2 * BK;
  int cols = 32;
  while (cols < BN) cols *= 2;
  const size_t staging1 = (size_t)p.nslab * 16384;
  const size_t fixed = 1024 + 1024 + 8 * (2 * TC_MAX_STAGES + 22) + 1024 + (p.w_resident ? wres_bytes : 0);
  const int nkb_tile = p.taps * p.kb1 + p.kb2;
  const int want = nkb_tile > 1 ? 3 : 2;
  static int nbuf_r = -1;                 // debugging aid: RAWFORMER_B200_RBUF=2|3 (residual staging depth)
  if (nbuf_r < 0) {
    const char* e = getenv("RAWFORMER_B200_RBUF");
    nbuf_r = (e && (e[0] == '2' || e[0] == '3')) ? e[0] - '0' : 3;
  }
  int ctas_per_sm = 1, stages = 0;
  p.nbuf = 1;
  const size_t halo_stride = ((size_t)p.halo_bytes + 127) & ~(size_t)127;
  if (p.halo) {
    // patch rings: nh TMA landing buffers + nt re-laid-out buffers; two CTAs per SM when everything fits twice
    p.nbuf = p.tma_store ? 2 : 1;
    p.nh = 3; p.nt = 2;
    size_t need = fixed + p.nbuf * staging1 + (size_t)(p.nh + p.nt) * halo_stride + 1024;
    if (2 * cols <= 256 && need <= 115712) {
      ctas_per_sm = 2;
    } else {
      if (need > 232448) { p.nh = 2; need = fixed + p.nbuf * staging1 + (size_t)(p.nh + p.nt) * halo_stride + 1024; }
      if (need > 232448) return -1;
    }
    stages = 3;   // (>= nh: sizes the full/empty barrier initialisation; the operand ring itself is unused)
  }
  if (!p.halo && 2 * cols <= 256) {
    for (int nb = p.tma_store ? (p.r_tma ? nbuf_r : 2) : 1; nb >= 1 && stages == 0; --nb) {
      if (nb == 1 && p.r_tma) break;                       // in-place residual wants >= two buffers: one CTA per SM
      const size_t used = fixed + nb * staging1;
      if (used + want * stage_bytes <= 115712) {
        ctas_per_sm = 2; p.nbuf = nb;
        stages = (int)((115712 - used) / stage_bytes);
      }
    }
  }
  if (!p.halo && stages == 0) {
    ctas_per_sm = 1;
    p.nbuf = p.tma_store ? (p.r_tma ? nbuf_r : 2) : 1;
    // keep two staging buffers if at all possible (a 2-deep operand ring is enough for that)
    while (p.nbuf > 2 && fixed + p.nbuf * staging1 + want * stage_bytes > 232448) --p.nbuf;
    if (p.nbuf == 2 && fixed + 2 * staging1 + 2 * stage_bytes > 232448) p.nbuf = 1;
    stages = (int)((232448 - fixed - p.nbuf * staging1) / stage_bytes);
  }
  const size_t staging = p.nbuf * staging1;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.acc_cols = cols;
  p.tmem_cols = 2 * cols;                     // BN <= 256 -> <= 512 columns
  p.tiles_m = grid_x;
  p.tiles_n = cdiv(g.N, BN);
  p.total_tiles = p.tiles_m * p.tiles_n * g.B * p.ksplit;
  const size_t smem = fixed + staging + (p.halo ? (size_t)(p.nh + p.nt) * halo_stride + 1024 : (size_t)p.stages * stage_bytes);
  const int K = g.K1 + g.K2;
  const double es = 2.0, rows = (double)g.B * g.M;
  const double abytes = rows * (g.amode == AMODE_CONV3 ? g.K1 / 9 : K) * es;
  const double bytes = abytes + rows * g.N * es * (g.R ? 2.0 : 1.0) + (double)g.N * K * es * (g.w_img ? g.B : 1);
  const int max_ctas = num_sms() * ctas_per_sm;
  const int grid = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
  const bool ln = g.ln_stats != nullptr, hr = p.r_tma != 0;
  bool st = p.stats_out != nullptr;
  if (ln && (!g.ln_cs || !g.bias)) return -1;
#define RF_TC_LAUNCH_H(OM, LN, HR, ST, ACT, HALO)                                                                         \
  do {                                                                                                                      \
    static bool attr = false;                                                                                               \
    auto kern = k_tc_gemm<OM, LN, HR, ST, ACT, HALO>;                                                                       \
    if (!attr) {                                                                                                            \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;   \
      attr = true;                                                                                                          \
    }                                                                                                                       \
    ScopedLaunch sl(g.kernel_id, bytes, 2.0 * rows * g.N * K);                                                              \
    launch_pdl(kern, dim3(grid), dim3(TC_THREADS), smem, ctx.stream, mA1, mA2, mW, mY, mR, p);                              \
    return p.stats_out ? p.tiles_n : 0;                                                                                     \
  } while (0)
  // the halo-conv code (patch re-layout in the epilogue warps, descriptor-table MMA loop) is compiled only into the
  // three shapes that use it, so the row-GEMM instantiations stay lean
#define RF_TC_LAUNCH(OM, LN, HR, ST, ACT) RF_TC_LAUNCH_H(OM, LN, HR, ST, ACT, false)
#define RF_TC_LAUNCH_CONV(OM, ACT)                          \
  do {                                                      \
    if (p.halo) RF_TC_LAUNCH_H(OM, false, false, false, ACT, true); \
    RF_TC_LAUNCH_H(OM, false, false, false, ACT, false);    \
  } while (0)
  if (g.omode == OMODE_ROWS) {
    if (!p.tma_store) return -1;
    if (st && (ln || g.act != ACT_NONE)) { st = false; p.stats_out = nullptr; }   // (the caller runs a stats pass instead)
    if (ln) {
      if (hr || g.act != ACT_NONE) return -1;
      RF_TC_LAUNCH(OMODE_ROWS, true, false, false, ACT_NONE);
    }
    if (g.act == ACT_NONE) {
      if (hr && st) RF_TC_LAUNCH(OMODE_ROWS, false, true, true, ACT_NONE);
      if (hr) RF_TC_LAUNCH(OMODE_ROWS, false, true, false, ACT_NONE);
      if (st) RF_TC_LAUNCH(OMODE_ROWS, false, false, true, ACT_NONE);
      RF_TC_LAUNCH(OMODE_ROWS, false, false, false, ACT_NONE);
    }
    if (g.act == ACT_LRELU && !hr) RF_TC_LAUNCH_CONV(OMODE_ROWS, ACT_LRELU);
    if (g.act == ACT_RELU && !hr) RF_TC_LAUNCH(OMODE_ROWS, false, false, false, ACT_RELU);
    if (g.act == ACT_TANH_RES && hr) RF_TC_LAUNCH(OMODE_ROWS, false, true, false, ACT_TANH_RES);
    return -1;
  }
  if (ln || g.act != ACT_NONE || g.R) return -1;
  if (g.omode == OMODE_CONVT) RF_TC_LAUNCH(OMODE_CONVT, false, false, false, ACT_NONE);
  if (g.omode == OMODE_UNSHUFFLE) {
    if (!p.tma_store) return -1;
    RF_TC_LAUNCH_CONV(OMODE_UNSHUFFLE, ACT_NONE);
  }
  if (g.omode == OMODE_ATOMIC_F32) RF_TC_LAUNCH(OMODE_ATOMIC_F32, false, false, false, ACT_NONE);
  if (g.omode == OMODE_HEAD && g.N == 16 && g.amode == AMODE_CONV3) RF_TC_LAUNCH_CONV(OMODE_HEAD, ACT_NONE);
#undef RF_TC_LAUNCH
#undef RF_TC_LAUNCH_H
#undef RF_TC_LAUNCH_CONV
  return -1;
}

int gram_max_splits() { return 4 * num_sms(); }

// qk: bf16 NHWC [P][2C] (q | k) of ONE image; G: per-split partials [ksplit][C][C/8] (see rf_kernels.cuh)
int launch_gram_tcgen05(Ctx& ctx, const void* qkv, float* G, int C, i64 P) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16 || C % 8) return 0;
  CUtensorMap m;
  const i64 d[3] = {2 * (i64)C, P, 1};
  const i64 st[3] = {1, 2 * (i64)C, 2 * (i64)C * P};
  const int bx[3] = {64, GR_PIX, 1};
  if (!make_map(&m, qkv, 3, d, st, bx)) return 0;
  const int packed = 2 * C <= 64 ? 1 : (2 * C <= 128 ? 2 : 0);
  const int mt = packed ? 1 : cdiv(C, 128);
  const int nkb = (int)((P + GR_PIX - 1) / GR_PIX);
  // packed: two CTAs per SM (<= 113 KB each), ring as deep as that allows; general: one CTA per SM, 3 stages of 64 KB
  const int nst = packed == 1 ? 5 : (packed == 2 ? 3 : 3);
  // split-K: enough CTAs to fill the machine, but at least ~24 k-blocks each (TMEM allocation, barrier set-up and the
  // atomic read-out of the accumulator are per-CTA costs that dominated the small stages)
  int ksplit = (packed ? 4 : 2) * num_sms() / (mt * mt);
  if (ksplit > nkb / 24) ksplit = nkb / 24;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > nkb) ksplit = nkb;
  // no EMPTY split: the kernel gives every split ceil(nkb / ksplit) k-blocks, so the last split(s) can end up without any
  // (nkb = 1024, ksplit = 42: 41 x 25 >= 1024) -- such a CTA returns without writing its partial slot, and the ordered
  // reduction that sums `ksplit` slots would read whatever the workspace held there
  while (ksplit > 1 && (ksplit - 1) * cdiv(nkb, ksplit) >= nkb) --ksplit;
  const size_t smem = 1024 + (size_t)nst * (packed ? packed : 4) * GR_CHUNK + (packed == 1 ? GR_CHUNK : 0) + 8 * (2 * GR_STAGES + 2);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_tc_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return 0;
    attr_set = true;
  }
  ScopedLaunch sl(RF_K_GEMM_GRAM, 4.0 * C * P, 2.0 * P * C * (C / 8.0));
  launch_pdl(k_tc_gram, dim3(mt, mt, ksplit), dim3(GR_THREADS), smem, ctx.stream, m, G, C, P, ksplit, packed, nst);
  return ksplit;
}

}  // namespace rf
