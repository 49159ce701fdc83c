// Small-K 3x3 convolutions of the 4-channel fp32 maps on the tensor cores (bf16 mode, sm_100a):
//   MODE 0 (FLCA, FLCA_RF.py:150-158): the three guidance convolutions low(LL), high(|high|), chroma(cr,cb) are ONE
//          GEMM  pre[128 pixels x 3*Cc] = im2col(G)[128 x 36] * W^T, K padded to 48 (3 k-steps of tcgen05.mma); the
//          epilogue reads the three pre-activations of every feature element from tensor memory, applies
//          x * (1 + a*sigmoid(low) + b*tanh(high) + g*sigmoid(chr)) with MUFU.TANH only, accumulates the squeeze-excite
//          channel sums and writes the tile back through shared memory with a TMA store.
//   MODE 1 (embedding, FLCA_RF.py:303,338): out[128 x d] = im2col(x_ds)[128 x 36] * W_embed^T + bias.
// The CUDA-core kernels these replace spent 36 FMA + 3 transcendentals per element (instruction-bound at 9 % of HBM);
// here the FMAs are 3 MMAs per 128-pixel tile and the epilogue is ~10 instructions per element.
//
// One persistent CTA per SM (512 threads), fixed channel chunk Cc per CTA, software pipeline over 128-pixel tiles:
//   build A(i+1) (registers prefetched one tile earlier -> bf16 -> swizzled smem) | TMA load feat(i+1)
//   tcgen05.mma(i+1) -> TMEM buffer (i+1)&1             (async, one thread)
//   epilogue(i): tcgen05.ld TMEM buffer i&1, feat(i) from smem, result in place, TMA store
// so the MMA, the TMA traffic and the guidance gathers of the next tiles are in flight under the epilogue math.
#include "rf_kernels.cuh"
#include "rf_tma.cuh"

namespace rf {

constexpr int IT_THREADS = 512;
constexpr int IT_SLOTS = 32;          // == FLCA_SLOTS (rf_flca.cu): partial-sum slots per image
constexpr int IT_NF = 5;              // feature/output ring depth: loads run 3 tiles ahead, stores drain 1 tile behind

struct Im2colTcParams {
  const float* G;        // [B][H][W][4] fp32 (FLCA guidance / packed frame x_ds)
  const float* w;        // [9][4][C] fp32 taps (FLCA: maps LL, |high|, cr, cb; embed: the 4 input channels)
  const float* coef;     // FLCA: alpha, beta, gamma; embed: bias[C]
  float* partial;        // FLCA: [B][IT_SLOTS][C] channel sums (atomicAdd)
  int H, W, C, Cc, nchunks, B;
  int tiles_per_img, total_tiles, lanes;
  int parts;             // epilogue column split: warps 4*part .. 4*part+3 own 8*UPT*part .. channels of every row
  int swz;               // swizzle of the feature/output tile: 128, 64 or 0
  uint32_t row_bytes;    // Cc * 2
  int tmem_cols;         // columns of one accumulator buffer
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int MODE, int UPT>
__global__ void __launch_bounds__(IT_THREADS, 1)
k_im2col_tc(const __grid_constant__ CUtensorMap mapIn, const __grid_constant__ CUtensorMap mapOut, const Im2colTcParams p) {
  constexpr int SEG = MODE == 0 ? 3 : 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                         // 2 x 16 KB im2col tiles (K-major, SWIZZLE_128B)
  const uint32_t sW = base + 2 * 16384;             // SEG*Cc rows x 128 B (<= 24 KB)
  const uint32_t sF = sW + 24576;                   // IT_NF x 16 KB feature / output tiles (ring, updated in place)
  const uint32_t bars = sF + IT_NF * 16384;         // acc_full[2], feat_full[IT_NF], tmem slot
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (bars + 64 - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x % p.nchunks, lane_id = blockIdx.x / p.nchunks;
  const int N = SEG * p.Cc;
  const i64 P = (i64)p.H * p.W;

  if (tid == 0) {
    tma_prefetch_desc(&mapOut);
    if (MODE == 0) tma_prefetch_desc(&mapIn);
    for (int i = 0; i < 2 + IT_NF; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(bars + 64, (uint32_t)(2 * p.tmem_cols));

  // ---- W tile (once): row n = seg*Cc + cl, K index = tap*4 + map; sigmoid inputs are pre-halved (sigmoid(a) =
  //      0.5*tanh(a/2) + 0.5), so the epilogue needs MUFU.TANH only -------------------------------------------------
  for (int item = tid; item < N * 6; item += IT_THREADS) {
    const int n = item / 6, j = item - n * 6;       // 16-byte chunk j = taps 2j, 2j+1
    const int seg = n / p.Cc, cl = n - seg * p.Cc;
    const int c = chunk * p.Cc + cl;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int tap = 2 * j + (e >> 2), m = e & 3;
      float wv = 0.f;
      if (tap < 9) {
        bool use;
        float sc = 1.f;
        if (MODE == 0) {
          use = seg == 0 ? m == 0 : (seg == 1 ? m == 1 : m >= 2);
          sc = seg == 1 ? 1.f : 0.5f;
        } else {
          use = true;
        }
        if (use) wv = p.w[(i64)(tap * 4 + m) * p.C + c] * sc;
      }
      v[e] = wv;
    }
    uint4 q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    sts128(sW + (uint32_t)n * 128 + (uint32_t)((j ^ (n & 7)) << 4), q);
  }

  // ---- im2col builder: 4 threads per tile row; thread q gathers taps 2q, 2q+1 (chunk q), q == 0 also tap 8 ---------
  const int arow = tid >> 2, aq = tid & 3;
  float4 g0, g1, g2;                                 // prefetched guidance of the next tile to build
  auto load_G = [&](int gt) {                        // gt = global tile index
    const int b = gt / p.tiles_per_img, tl = gt - b * p.tiles_per_img;
    const i64 pix = (i64)tl * 128 + arow;
    g0 = g1 = g2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < P) {
      const int y = (int)(pix / p.W), x = (int)(pix - (i64)y * p.W);
      const float4* Gb = reinterpret_cast<const float4*>(p.G) + (i64)b * P;
      auto tap_at = [&](int tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) return __ldg(Gb + (i64)yy * p.W + xx);
        return make_float4(0.f, 0.f, 0.f, 0.f);
      };
      g0 = tap_at(2 * aq);
      g1 = tap_at(2 * aq + 1);
      if (aq == 0) g2 = tap_at(8);
    }
  };
  auto store_A = [&](int stage) {
    const uint32_t rowp = sA + stage * 16384 + (uint32_t)arow * 128;
    const uint32_t sw = (uint32_t)(arow & 7);
    sts128(rowp + ((aq ^ sw) << 4),
           make_uint4(pack_bf16x2(g0.x, g0.y), pack_bf16x2(g0.z, g0.w), pack_bf16x2(g1.x, g1.y), pack_bf16x2(g1.z, g1.w)));
    if (aq == 0) sts128(rowp + ((4u ^ sw) << 4), make_uint4(pack_bf16x2(g2.x, g2.y), pack_bf16x2(g2.z, g2.w), 0u, 0u));
    if (aq == 1) sts128(rowp + ((5u ^ sw) << 4), make_uint4(0u, 0u, 0u, 0u));
  };
  auto issue_feat = [&](int gt, int stage) {         // one thread
    const int b = gt / p.tiles_per_img, tl = gt - b * p.tiles_per_img;
    mbar_expect_tx(bars + 16 + 8 * stage, 128u * p.row_bytes);
    tma_load_3d(sF + stage * 16384, &mapIn, bars + 16 + 8 * stage, chunk * p.Cc, tl * 128, b);
  };
  const uint32_t idesc = make_idesc_m128(N);
  auto issue_mma = [&](int stage, uint32_t tmem_base) {   // one thread
    const uint64_t adesc = make_sw128_desc(sA + stage * 16384), bdesc = make_sw128_desc(sW);
    const uint32_t d = tmem_base + (uint32_t)(stage * p.tmem_cols);
#pragma unroll
    for (int k = 0; k < 3; ++k) umma_f16(d, adesc + 2u * k, bdesc + 2u * k, idesc, k ? 1u : 0u);
    umma_commit(bars + 8 * stage);
  };

  // ---- epilogue role of this thread -------------------------------------------------------------------------------
  const int quad = warp & 3, part = warp >> 2;
  const bool epi = part < p.parts;
  const int erow = quad * 32 + lane;
  const uint32_t esw = p.swz == 128 ? (uint32_t)(erow & 7) : (p.swz == 64 ? (uint32_t)((erow >> 1) & 3) : 0u);
  float cf[MODE == 0 ? 4 : 8 * UPT];                 // FLCA: base, h0, h1, h2;  embed: bias of this thread's channels
  if (MODE == 0) {
    const float k0 = p.coef[0], k1 = p.coef[1], k2 = p.coef[2];
    cf[0] = 1.f + 0.5f * k0 + 0.5f * k2; cf[1] = 0.5f * k0; cf[2] = k1; cf[3] = 0.5f * k2;
  } else if (epi) {
#pragma unroll
    for (int e = 0; e < 8 * UPT; ++e) cf[e] = p.coef[chunk * p.Cc + part * UPT * 8 + e];
  }
  float csum[8 * UPT];
#pragma unroll
  for (int e = 0; e < 8 * UPT; ++e) csum[e] = 0.f;
  int cur_b = -1;
  auto flush = [&](int b) {                          // warp-reduce the 32 rows of this thread's channels
    if (epi) {
#pragma unroll
      for (int e = 0; e < 8 * UPT; ++e) {
        float v = warp_sum(csum[e]);
        if (lane == 0 && v != 0.f)
          atomicAdd(p.partial + ((i64)b * IT_SLOTS + (lane_id % IT_SLOTS)) * p.C + chunk * p.Cc + part * UPT * 8 + e, v);
        csum[e] = 0.f;
      }
    }
  };

  fence_proxy_async();                               // W tile (generic stores) -> tensor-core (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int t0 = lane_id, tstep = p.lanes;
  const int n_my = t0 < p.total_tiles ? (p.total_tiles - 1 - t0) / tstep + 1 : 0;
  if (n_my > 0) {
    load_G(t0);
    store_A(0);
    if (n_my > 1) load_G(t0 + tstep);
    if (MODE == 0 && tid == 0) {
      for (int j = 0; j < 3 && j < n_my; ++j) issue_feat(t0 + j * tstep, j);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_mma(0, tmem_base);
    }
  }
  for (int i = 0; i < n_my; ++i) {
    const int a = i & 1;                             // A stage / TMEM buffer of tile i
    const int fs = i % IT_NF;                        // feature stage of tile i
    const int gt = t0 + i * tstep;
    if (tid == 32) {
      // (bulk-async groups are per thread: this thread also issues the stores)  store(i-2) has finished reading
      // feature stage (i+3) % IT_NF, so the load that runs 3 tiles ahead may overwrite it
      tma_store_wait_read<1>();
      if (MODE == 0 && i + 3 < n_my) issue_feat(gt + 3 * tstep, (i + 3) % IT_NF);
    }
    if (i + 1 < n_my) {
      store_A(a ^ 1);                                // A stage a^1: MMA(i-1) completed before epilogue(i-1) started
      if (i + 2 < n_my) load_G(gt + 2 * tstep);      // stays in flight under the epilogue
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_mma(a ^ 1, tmem_base);                 // TMEM buffer a^1 was drained by epilogue(i-1)
      }
    }
    const int b = gt / p.tiles_per_img, tl = gt - b * p.tiles_per_img;
    if (MODE == 0 && b != cur_b) {
      if (cur_b >= 0) flush(cur_b);
      cur_b = b;
    }
    mbar_wait(bars + 8 * a, (i >> 1) & 1);
    tc_fence_after();
    if (MODE == 0) mbar_wait(bars + 16 + 8 * fs, (i / IT_NF) & 1);
    if (epi) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.tmem_cols);
      const uint32_t frow = sF + fs * 16384 + (uint32_t)erow * p.row_bytes;
#pragma unroll
      for (int u = 0; u < UPT; ++u) {
        const int ui = part * UPT + u;               // 8-channel unit inside the chunk
        const uint32_t faddr = frow + ((((uint32_t)ui) ^ esw) << 4);
        uint32_t v0[8], v1[8], v2[8];
        tmem_ld8(taddr + ui * 8, v0);
        if (MODE == 0) {
          tmem_ld8(taddr + p.Cc + ui * 8, v1);
          tmem_ld8(taddr + 2 * p.Cc + ui * 8, v2);
        }
        uint4 fq = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 0) fq = lds128(faddr);
        tmem_ld_wait();
        float o[8];
        if (MODE == 0) {
          const uint32_t fw[4] = {fq.x, fq.y, fq.z, fq.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float f = (e & 1) ? __uint_as_float(fw[e >> 1] & 0xffff0000u) : __uint_as_float(fw[e >> 1] << 16);
            float m = fmaf(cf[1], tanh_fast(__uint_as_float(v0[e])), cf[0]);
            m = fmaf(cf[2], tanh_fast(__uint_as_float(v1[e])), m);
            m = fmaf(cf[3], tanh_fast(__uint_as_float(v2[e])), m);
            o[e] = f * m;
            csum[u * 8 + e] += o[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v0[e]) + cf[u * 8 + e];
        }
        sts128(faddr, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                 pack_bf16x2(o[6], o[7])));
      }
    }
    tc_fence_before();
    fence_proxy_async();                             // output tile (generic stores) -> TMA store (async proxy)
    __syncthreads();
    if (tid == 32) {
      tma_store_3d(&mapOut, sF + fs * 16384, chunk * p.Cc, tl * 128, b);
      tma_store_commit();
    }
  }
  if (tid == 32) tma_store_wait_all<0>();
  if (MODE == 0 && cur_b >= 0) flush(cur_b);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(2 * p.tmem_cols));
  }
}

bool im2col_tc_supported(const Ctx& ctx, int C) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16) return false;
  if (C % 32 && C % 48) return false;
  const int Cc = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 48);
  return C / Cc <= num_sms();
}

static bool run_im2col_tc(Ctx& ctx, int mode, const void* feat, const float* G, const float* w, const float* coef, void* out,
                          float* partial, int B, int H, int W, int C) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16) return false;
  int Cc, parts, upt;
  if (C % 64 == 0) { Cc = 64; parts = 4; upt = 2; }
  else if (C % 32 == 0) { Cc = 32; parts = 4; upt = 1; }
  else if (C % 48 == 0) { Cc = 48; parts = 3; upt = 2; }
  else return false;
  const i64 P = (i64)H * W;
  Im2colTcParams p;
  p.G = G; p.w = w; p.coef = coef; p.partial = partial;
  p.H = H; p.W = W; p.C = C; p.Cc = Cc; p.nchunks = C / Cc; p.B = B;
  p.tiles_per_img = (int)cdivl(P, 128);
  const i64 total = (i64)p.tiles_per_img * B;
  if (total > 0x7fffffff || p.nchunks > num_sms()) return false;
  p.total_tiles = (int)total;
  p.lanes = num_sms() / p.nchunks;
  if (p.lanes > p.total_tiles) p.lanes = p.total_tiles;
  p.parts = parts;
  p.row_bytes = (uint32_t)Cc * 2;
  p.swz = p.row_bytes == 128 ? 128 : (p.row_bytes == 64 ? 64 : 0);
  const int N = (mode == 0 ? 3 : 1) * Cc;
  int cols = 32;
  while (cols < N) cols *= 2;
  p.tmem_cols = cols;
  CUtensorMap mIn, mOut;
  const i64 d[3] = {C, P, B};
  const i64 st[3] = {1, C, (i64)C * P};
  const int bx[3] = {Cc, 128, 1};
  if (!make_map_ex(&mOut, out, 3, d, st, bx, 2, p.swz)) return false;
  if (mode == 0) {
    if (!make_map_ex(&mIn, feat, 3, d, st, bx, 2, p.swz)) return false;
  } else {
    mIn = mOut;
  }
  const size_t smem = 1024 + 2 * 16384 + 24576 + IT_NF * 16384 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_im2col_tc<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_im2col_tc<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_im2col_tc<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_im2col_tc<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return false;
    attr_set = true;
  }
  const int grid = p.lanes * p.nchunks;
  if (mode == 0) {
    if (upt == 1) k_im2col_tc<0, 1><<<grid, IT_THREADS, smem, ctx.stream>>>(mIn, mOut, p);
    else k_im2col_tc<0, 2><<<grid, IT_THREADS, smem, ctx.stream>>>(mIn, mOut, p);
  } else {
    if (upt == 1) k_im2col_tc<1, 1><<<grid, IT_THREADS, smem, ctx.stream>>>(mIn, mOut, p);
    else k_im2col_tc<1, 2><<<grid, IT_THREADS, smem, ctx.stream>>>(mIn, mOut, p);
  }
  return true;
}

// xmod = feat * (1 + a*sig(conv(LL)) + b*tanh(conv(yh)) + g*sig(conv(cr,cb))); partial [B][32][C] += channel sums
bool launch_flca_mod_tc(Ctx& ctx, const void* feat, const float* G, const float* w36, const float* abg, void* xmod,
                        float* partial, int B, int Hf, int Wf, int C) {
  return run_im2col_tc(ctx, 0, feat, G, w36, abg, xmod, partial, B, Hf, Wf, C);
}

// out = conv3x3(x_ds; 4 -> d) + bias, bf16 NHWC
bool launch_embed_tc(Ctx& ctx, const float* x_ds, const float* w, const float* b, void* out, int B, int h, int w_, int d) {
  return run_im2col_tc(ctx, 1, nullptr, x_ds, w, b, out, nullptr, B, h, w_, d);
}

}  // namespace rf
