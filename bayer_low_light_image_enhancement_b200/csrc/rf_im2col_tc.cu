// Small-K 3x3 convolutions of the 4-channel fp32 maps on the tensor cores (bf16 mode, sm_100a):
//   MODE 0 (FLCA, FLCA_RF.py:150-158): the three guidance convolutions low(LL), high(|high|), chroma(cr,cb) are ONE
//          GEMM  pre[128 pixels x 3*Cc] = patch(G) * W^T; the epilogue reads the three pre-activations of every feature
//          element from tensor memory, applies x * (1 + a*sigmoid(low) + b*tanh(high) + g*sigmoid(chr)) with MUFU.TANH
//          only, accumulates the squeeze-excite channel sums and writes the tile back through shared memory by TMA.
//   MODE 1 (embedding, FLCA_RF.py:303,338): out[128 x d] = patch(x_ds) * W_embed^T + bias.
//   MODE 2 (FLCA_Pyramid level, ML_RF.py:157-166): xs = x * (ga*sigmoid(conv(LL_l)) + gb*tanh(conv(hi_l))), gates per image.
//   MODE 3 (FLCA_Pyramid chroma, ML_RF.py:171-176): xs = x * gc*sigmoid(conv(cr, cb)).
//
// No im2col is ever built.  The 4 fp32 maps of a pixel are stored once per stage as 8 bf16 = [hi(4) | lo(4)] (hi = bf16(v),
// lo = bf16(v - hi): 16 significant bits) = ONE 16-byte unit per pixel.  A TMA box of the (16+2) x (8+2) halo patch then
// lands as [pixel][16 B], which is the canonical NO-SWIZZLE K-major UMMA operand with pixels as rows and one tap as one
// 8-wide K chunk: core matrix = 8 consecutive pixels, SBO = patch pitch * 16 B, and the second K chunk of a k-step is
// simply ANOTHER TAP, LBO = (tap offset difference) * 16 B.  Nine taps = 5 tcgen05.mma (K = 16) per 128-pixel tile whose
// A descriptors differ only in start address and LBO; out-of-image pixels are TMA zero fill = conv padding.
// The weights are laid out the same way by the CTA once: [K chunk][n][16 B].
//
// One persistent CTA per SM (512 threads), fixed channel chunk Cc per CTA, 8 x 16 pixel tiles, ONE CTA barrier per tile:
//   thread 32: TMA loads (guidance patch + feature tile) 3 tiles ahead, TMA store of the finished tile
//   thread 0 : tcgen05.mma of tile i+1 into the other TMEM buffer while
//   all warps: epilogue of tile i (tcgen05.ld, feature tile from smem, result in place)
#include <stdlib.h>

#include "rf_kernels.cuh"
#include "rf_tma.cuh"

namespace rf {

constexpr int IT_THREADS = 512;
constexpr int IT_SLOTS = 320;         // == FLCA_SLOTS (rf_flca.cu): partial-sum slots per image; slot = CTA index within its
                                      // channel chunk (< num_sms <= 160): every CTA owns its slot -> plain stores, and the
                                      // ordered sum over the slots (k_se_fold) makes the channel sums bit-reproducible
constexpr int IT_NGS = 4;             // guidance patch ring depth
constexpr int IT_TW = 8, IT_TH = 16;  // tile = 8 x 16 pixels (patch rows = the 8-row core-matrix groups)
constexpr int IT_PITCH = IT_TW + 2, IT_NPIX = (IT_TW + 2) * (IT_TH + 2);
constexpr uint32_t IT_PATCH_BYTES = IT_NPIX * 16;                 // 2880
constexpr uint32_t IT_PATCH_STRIDE = (IT_PATCH_BYTES + 127u) & ~127u;

struct Im2colTcParams {
  const float* w;        // [9][wmaps][C] fp32 taps (FLCA: maps LL, |high|, cr, cb; embed: the 4 input channels; ML: 6 maps)
  const float* coef;     // FLCA: alpha, beta, gamma; embed: bias[C]; ML: gates [B][6]
  int wmaps;             // weight maps per tap (4, or 6 for the pyramid variant)
  int seg_of[4];         // pre-activation segment fed by guidance slot m of the 16-byte pixel (-1: unused)
  int wmap_of[4];        // weight map of slot m
  float scale_of[4];     // 0.5 for sigmoid inputs (sigmoid(a) = 0.5*tanh(a/2) + 0.5), else 1
  int coef_off;          // ML: index of the first gate used inside gates[b][6]
  float* partial;        // FLCA: [B][lanes][C] channel sums, slot = CTA index within the chunk (pre-zeroed; plain stores)
  int ylo, yhi;          // FLCA: rows that contribute to the channel sums (row-tiled forward: the band's interior)
  int H, W, C, Cc, nchunks, B;
  int tiles_x, tiles_y, tiles_per_img, total_tiles, lanes;
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

// fp32 x4 per pixel -> [hi x4 | lo x4] bf16 (16 bytes per pixel)
__global__ void __launch_bounds__(256)
k_split_bf16x8(const float4* __restrict__ g, uint4* __restrict__ out, i64 n, int stride4) {
  pdl_trigger();
  pdl_wait();
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const float4 v = g[i * stride4];
    const float h0 = __bfloat162float(__float2bfloat16_rn(v.x)), h1 = __bfloat162float(__float2bfloat16_rn(v.y));
    const float h2 = __bfloat162float(__float2bfloat16_rn(v.z)), h3 = __bfloat162float(__float2bfloat16_rn(v.w));
    out[i] = make_uint4(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3), pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
  }
}
void launch_split_bf16x8(Ctx& ctx, const float* g4, void* out16, i64 npix, int stride_floats) {
  if (ctx.dry || npix <= 0) return;
  ScopedLaunch sl(RF_K_GUIDANCE, 32.0 * npix);
  const unsigned gx = (unsigned)(cdivl(npix, 256) < 8 * num_sms() ? cdivl(npix, 256) : 8 * num_sms());
  launch_pdl(k_split_bf16x8, dim3(gx), dim3(256), 0, ctx.stream, (const float4*)g4, (uint4*)out16, npix, stride_floats / 4);
}

// TWO: two CTAs per SM (ring depth 4 instead of 5, <= 64 registers): the tile loop is a latency chain (accumulator ready ->
// tcgen05.ld -> MUFU -> store -> CTA barrier -> TMA store) with 8 elements per thread per tile; a second resident CTA fills it
template <int MODE, int UPT, bool TWO>
__global__ void __launch_bounds__(IT_THREADS, TWO ? 2 : 1)
k_im2col_tc(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapIn,
            const __grid_constant__ CUtensorMap mapOut, const Im2colTcParams p) {
  constexpr int SEG = MODE == 0 ? 3 : (MODE == 2 ? 2 : 1);
  constexpr bool MODULATE = MODE != 1;               // multiply the feature tile (vs. bias epilogue)
  constexpr int IT_NF = TWO ? 4 : 5;                 // feature/output ring depth: loads run IT_NF - 2 tiles ahead, stores drain 1 behind
  constexpr int AHEAD = IT_NF - 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sF = base;                                   // IT_NF x 16 KB feature / output tiles (ring, in place)
  const uint32_t sW = sF + IT_NF * 16384;                     // [10 K chunks][N rows][16 B] (<= 30 KB)
  const uint32_t sG = sW + 30720;                             // IT_NGS guidance patches
  const uint32_t bars = sG + IT_NGS * IT_PATCH_STRIDE;        // acc_full[2], g_full[IT_NGS], f_full[IT_NF], tmem slot
  auto acc_bar = [&](int a) { return bars + 8u * a; };
  auto g_bar = [&](int q) { return bars + 8u * (2 + q); };
  auto f_bar = [&](int q) { return bars + 8u * (2 + IT_NGS + q); };
  const uint32_t tmem_slot = bars + 8u * (2 + IT_NGS + IT_NF);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* s_red = reinterpret_cast<float*>(smem_raw + (tmem_slot + 64 - smem_u32(smem_raw)));   // [16 warps][16] flush scratch
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x % p.nchunks, lane_id = blockIdx.x / p.nchunks;
  const int N = SEG * p.Cc;

  if (tid == 0) {
    tma_prefetch_desc(&mapG);
    tma_prefetch_desc(&mapOut);
    if (MODULATE) tma_prefetch_desc(&mapIn);
    for (int i = 0; i < 2 + IT_NGS + IT_NF; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)(2 * p.tmem_cols));
  pdl_wait();           // barrier set-up and the TMEM allocation overlapped the previous kernel's tail

  // ---- W (once): K chunk kc <-> tap kc for kc < 8, chunk 8 = zeros (it pairs tap 7's pixels a second time), chunk 9 =
  //      tap 8; inside a chunk [hi weights of the 4 maps | the same weights again for the lo parts].  Row n = seg*Cc + cl.
  //      Sigmoid inputs are pre-halved (sigmoid(a) = 0.5*tanh(a/2) + 0.5), so the epilogue needs MUFU.TANH only -----------
  for (int item = tid; item < N * 10; item += IT_THREADS) {
    const int kc = item / N, n = item - kc * N;
    const int seg = n / p.Cc, cl = n - seg * p.Cc;
    const int c = chunk * p.Cc + cl;
    const int tap = kc < 8 ? kc : (kc == 9 ? 8 : -1);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (tap >= 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (p.seg_of[m] == seg) v[m] = p.w[(i64)(tap * p.wmaps + p.wmap_of[m]) * p.C + c] * p.scale_of[m];
    }
    const uint32_t q0 = pack_bf16x2(v[0], v[1]), q1 = pack_bf16x2(v[2], v[3]);
    sts128(sW + (uint32_t)(kc * N + n) * 16u, make_uint4(q0, q1, q0, q1));
  }

  // ---- tile bookkeeping -------------------------------------------------------------------------------------------
  const int t0 = lane_id, tstep = p.lanes;
  const int n_my = t0 < p.total_tiles ? (p.total_tiles - 1 - t0) / tstep + 1 : 0;
  auto decode = [&](int gt, int& b, int& ty, int& tx) {
    b = gt / p.tiles_per_img;
    const int r = gt - b * p.tiles_per_img;
    ty = r / p.tiles_x;
    tx = r - ty * p.tiles_x;
  };
  auto issue_loads = [&](int j) {                    // thread 32: guidance patch + feature tile of my j-th tile
    int b, ty, tx;
    decode(t0 + j * tstep, b, ty, tx);
    const int gs = j % IT_NGS, fs = j % IT_NF;
    mbar_expect_tx(g_bar(gs), IT_PATCH_BYTES);
    // (x, 8 bf16) is one flattened dimension of the map: a patch row is ONE 160-byte line instead of ten 16-byte ones
    // (TMA time goes with the number of lines); element-wise out-of-bounds fill keeps the zero padding exact
    tma_load_3d(sG + gs * IT_PATCH_STRIDE, &mapG, g_bar(gs), (tx * IT_TW - 1) * 8, ty * IT_TH - 1, b);
    if (MODULATE) {
      mbar_expect_tx(f_bar(fs), 128u * p.row_bytes);
      tma_load_4d(sF + fs * 16384, &mapIn, f_bar(fs), chunk * p.Cc, tx * IT_TW, ty * IT_TH, b);
    }
  };
  const uint32_t idesc = make_idesc_m128(N);
  // no-swizzle K-major descriptors: start | LBO (K-chunk stride) | SBO (8-row-group stride) | version
  auto nosw_desc = [&](uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46);
  };
  // warp 0, all lanes (warp-uniform descriptor arithmetic stays on the uniform datapath), ONE elected lane issues: under a
  // `tid == 0` branch every tcgen05.mma cost ~20 instructions, and warp 0 -- an epilogue warp like the others -- reached the
  // tile's barrier that much later than everybody else
  bool mma_leader = false;
  if (warp == 0) {
    uint32_t e;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(e));
    mma_leader = e != 0;
  }
  auto issue_mma = [&](int j, uint32_t tmem_base) {          // warp 0: the 5 k-steps of my j-th tile
    mbar_wait(g_bar(j % IT_NGS), (j / IT_NGS) & 1);
    tc_fence_after();
    const uint32_t Gp = sG + (j % IT_NGS) * IT_PATCH_STRIDE;
    const uint32_t d = tmem_base + (uint32_t)((j & 1) * p.tmem_cols);
    // k-step s: taps (2s, 2s+1); the last one: (tap 7 again with zero weights, tap 8)
    const int ta[5] = {0, 2, 4, 6, 7};                          // first tap of each k-step
    const int tb[5] = {1, 3, 5, 7, 8};                          // second tap
#pragma unroll
    for (int s5 = 0; s5 < 5; ++s5) {
      const uint32_t oa = (uint32_t)((ta[s5] / 3) * IT_PITCH + ta[s5] % 3), ob = (uint32_t)((tb[s5] / 3) * IT_PITCH + tb[s5] % 3);
      const uint64_t adesc = nosw_desc(Gp + oa * 16u, (ob - oa) * 16u, IT_PITCH * 16u);
      const uint64_t bdesc = nosw_desc(sW + (uint32_t)(2 * s5 * N) * 16u, (uint32_t)N * 16u, 128u);
      if (mma_leader) umma_f16(d, adesc, bdesc, idesc, s5 ? 1u : 0u);
    }
    if (mma_leader) umma_commit(acc_bar(j & 1));
    __syncwarp();
  };

  // ---- epilogue role of this thread -------------------------------------------------------------------------------
  const int quad = warp & 3, part = warp >> 2;
  const bool epi = part < p.parts;
  const int erow = quad * 32 + lane;                 // tile row = patch pixel (py, px) = (erow >> 3, erow & 7)
  const uint32_t esw = p.swz == 128 ? (uint32_t)(erow & 7) : (p.swz == 64 ? (uint32_t)((erow >> 1) & 3) : 0u);
  float cf[MODE == 1 ? 8 * UPT : 4];                 // modulate: base, h0, h1, h2;  embed: bias of this thread's channels
  auto set_coef = [&](int b) {                       // modulation = cf0 + cf1*tanh(a0) + cf2*tanh(a1) + cf3*tanh(a2)
    if (MODE == 0) {
      const float k0 = p.coef[0], k1 = p.coef[1], k2 = p.coef[2];
      cf[0] = 1.f + 0.5f * k0 + 0.5f * k2; cf[1] = 0.5f * k0; cf[2] = k1; cf[3] = 0.5f * k2;
    } else if (MODE == 2) {
      const float k0 = p.coef[b * 6 + p.coef_off], k1 = p.coef[b * 6 + p.coef_off + 1];
      cf[0] = 0.5f * k0; cf[1] = 0.5f * k0; cf[2] = k1; cf[3] = 0.f;
    } else if (MODE == 3) {
      const float k0 = p.coef[b * 6 + p.coef_off];
      cf[0] = 0.5f * k0; cf[1] = 0.5f * k0; cf[2] = 0.f; cf[3] = 0.f;
    }
  };
  if (MODE == 0) {
    set_coef(0);
  } else if (MODE == 1 && epi) {
#pragma unroll
    for (int e = 0; e < 8 * UPT; ++e) cf[e] = p.coef[chunk * p.Cc + part * UPT * 8 + e];
  }
  float csum[8 * UPT];
#pragma unroll
  for (int e = 0; e < 8 * UPT; ++e) csum[e] = 0.f;
  int cur_b = -1;
  auto flush = [&](int b) {                          // block-uniform call: channel sums of image b -> this CTA's slot
    // warp-reduce the 32 rows of this thread's channels (fixed shuffle tree), park the warp's values, then ONE thread per
    // channel adds the four lane-quadrant warps in quadrant order and stores to the CTA's own slot: no atomics
    if (epi) {
#pragma unroll
      for (int e = 0; e < 8 * UPT; ++e) {
        const float v = warp_sum(csum[e]);
        if (lane == 0) s_red[warp * 16 + e] = v;
        csum[e] = 0.f;
      }
    }
    __syncthreads();
    if (tid < p.parts * 8 * UPT) {
      const int pt = tid / (8 * UPT), e = tid - pt * (8 * UPT);
      float v = 0.f;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) v += s_red[(pt * 4 + qd) * 16 + e];
      p.partial[((i64)b * p.lanes + lane_id) * p.C + chunk * p.Cc + pt * UPT * 8 + e] = v;
    }
    __syncthreads();
  };

  fence_proxy_async();                               // W (generic stores) -> tensor-core (async proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();                                     // (after the TMEM allocation, see launch_pdl)

  if (tid == 32) {
    for (int j = 0; j < AHEAD && j < n_my; ++j) issue_loads(j);
  }
  if (warp == 0 && n_my > 0) issue_mma(0, tmem_base);

  for (int i = 0; i < n_my; ++i) {
    const int a = i & 1;                             // TMEM buffer of tile i
    const int fs = i % IT_NF;                        // feature stage of tile i
    if (tid == 32) {
      // (bulk-async groups are per thread: this thread also issues the stores)  store(i-2) has finished reading feature
      // stage (i+3) % IT_NF, so the load that runs 3 tiles ahead may overwrite it; the guidance slot (i+3) % IT_NGS was
      // consumed by the MMAs of tile i-1, complete before the epilogue of tile i-1 started
      tma_store_wait_read<1>();
      if (i + AHEAD < n_my) issue_loads(i + AHEAD);
    }
    if (warp == 0 && i + 1 < n_my) issue_mma(i + 1, tmem_base);   // TMEM buffer a^1 was drained by epilogue(i-1)
    if (MODULATE) {
      const int b = (t0 + i * tstep) / p.tiles_per_img;
      if (b != cur_b) {
        if (MODE == 0 && cur_b >= 0) flush(cur_b);
        if (MODE != 0) set_coef(b);
        cur_b = b;
      }
    }
    mbar_wait(acc_bar(a), (i >> 1) & 1);
    tc_fence_after();
    if (MODULATE) mbar_wait(f_bar(fs), (i / IT_NF) & 1);
    float smask = 1.f;                               // MODE 0: does this thread's pixel count towards the channel sums
    if (MODE == 0 && p.yhi - p.ylo < p.H) {
      const int gt = t0 + i * tstep;
      const int ty = (gt % p.tiles_per_img) / p.tiles_x;
      smask = (unsigned)(ty * IT_TH + (erow >> 3) - p.ylo) < (unsigned)(p.yhi - p.ylo) ? 1.f : 0.f;
    }
    if (epi) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.tmem_cols);
      const uint32_t frow = sF + fs * 16384 + (uint32_t)erow * p.row_bytes;
#pragma unroll
      for (int u = 0; u < UPT; ++u) {
        const int ui = part * UPT + u;               // 8-channel unit inside the chunk
        const uint32_t faddr = frow + ((((uint32_t)ui) ^ esw) << 4);
        uint32_t v0[8], v1[8], v2[8];
        tmem_ld8(taddr + ui * 8, v0);
        if (SEG >= 2) tmem_ld8(taddr + p.Cc + ui * 8, v1);
        if (SEG >= 3) tmem_ld8(taddr + 2 * p.Cc + ui * 8, v2);
        uint4 fq = make_uint4(0u, 0u, 0u, 0u);
        if (MODULATE) fq = lds128(faddr);
        tmem_ld_wait();
        float o[8];
        if (MODULATE) {
          const uint32_t fw[4] = {fq.x, fq.y, fq.z, fq.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float f = (e & 1) ? __uint_as_float(fw[e >> 1] & 0xffff0000u) : __uint_as_float(fw[e >> 1] << 16);
            float m = fmaf(cf[1], tanh_fast(__uint_as_float(v0[e])), cf[0]);
            if (SEG >= 2) m = fmaf(cf[2], tanh_fast(__uint_as_float(v1[e])), m);
            if (SEG >= 3) m = fmaf(cf[3], tanh_fast(__uint_as_float(v2[e])), m);
            o[e] = f * m;
            if (MODE == 0) csum[u * 8 + e] = fmaf(o[e], smask, csum[u * 8 + e]);
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
    __syncthreads();                                 // (also: TMEM buffer a is drained before MMA(i+2) is issued)
    if (tid == 32) {
      int bb, ty, tx;
      decode(t0 + i * tstep, bb, ty, tx);
      tma_store_4d(&mapOut, sF + fs * 16384, chunk * p.Cc, tx * IT_TW, ty * IT_TH, bb);
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

static bool im2col_two_enabled() {
  static int on = -1;                     // debugging aid: RAWFORMER_B200_IM2COL_TWO=0 keeps one CTA per SM
  if (on < 0) {
    const char* e = getenv("RAWFORMER_B200_IM2COL_TWO");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

bool im2col_tc_supported(const Ctx& ctx, int C) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16) return false;
  if (C % 32 && C % 48) return false;
  const int Cc = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 48);
  return C / Cc <= num_sms();
}

// G16: [B][H][W] x 16 bytes ([hi x4 | lo x4] bf16 of the 4 fp32 maps, launch_split_bf16x8)
static bool run_im2col_tc(Ctx& ctx, int mode, const void* feat, const void* G16, const float* w, const float* coef, void* out,
                          float* partial, int B, int H, int W, int C, int level = 0) {
  if (!tcgen05_enabled() || ctx.dtype != RF_BF16 || !G16) return false;
  int Cc, parts, upt;
  if (C % 64 == 0) { Cc = 64; parts = 4; upt = 2; }
  else if (C % 32 == 0) { Cc = 32; parts = 4; upt = 1; }
  else if (C % 48 == 0) { Cc = 48; parts = 3; upt = 2; }
  else return false;
  Im2colTcParams p;
  p.w = w; p.coef = coef; p.partial = partial;
  p.wmaps = 4; p.coef_off = 0;
  for (int m = 0; m < 4; ++m) { p.seg_of[m] = -1; p.wmap_of[m] = m; p.scale_of[m] = 1.f; }
  if (mode == 0) {          // FLCA: LL -> sigmoid, |high| -> tanh, (cr, cb) -> sigmoid
    p.seg_of[0] = 0; p.seg_of[1] = 1; p.seg_of[2] = 2; p.seg_of[3] = 2;
    p.scale_of[0] = 0.5f; p.scale_of[2] = 0.5f; p.scale_of[3] = 0.5f;
  } else if (mode == 1) {   // embedding: all four input channels -> one pre-activation
    for (int m = 0; m < 4; ++m) p.seg_of[m] = 0;
  } else if (mode == 2) {   // pyramid level l on the (LL1, hi1, LL2, hi2) pixels: weight maps 2l, 2l+1 of 6
    p.wmaps = 6; p.coef_off = 2 * level;
    p.seg_of[2 * level] = 0; p.wmap_of[2 * level] = 2 * level; p.scale_of[2 * level] = 0.5f;
    p.seg_of[2 * level + 1] = 1; p.wmap_of[2 * level + 1] = 2 * level + 1;
  } else {                  // pyramid chroma on the (cr, cb, mag, 0) pixels: weight maps 4, 5 of 6
    p.wmaps = 6; p.coef_off = 4;
    p.seg_of[0] = 0; p.wmap_of[0] = 4; p.scale_of[0] = 0.5f;
    p.seg_of[1] = 0; p.wmap_of[1] = 5; p.scale_of[1] = 0.5f;
  }
  p.H = H; p.W = W; p.C = C; p.Cc = Cc; p.nchunks = C / Cc; p.B = B;
  p.ylo = 0; p.yhi = H;
  if (ctx.band != nullptr) { p.ylo = ctx.band->ht; p.yhi = ctx.band->ht + ctx.band->rows_in; }
  p.tiles_x = cdiv(W, IT_TW); p.tiles_y = cdiv(H, IT_TH);
  p.tiles_per_img = p.tiles_x * p.tiles_y;
  const i64 total = (i64)p.tiles_per_img * B;
  if (total > 0x7fffffff || p.nchunks > num_sms()) return false;
  p.total_tiles = (int)total;
  const int N = (mode == 0 ? 3 : (mode == 2 ? 2 : 1)) * Cc;
  int cols = 32;
  while (cols < N) cols *= 2;
  p.tmem_cols = cols;
  // two CTAs per SM where two double-buffered accumulators fit in tensor memory (FLCA modulation at C = 32)
  const bool two = im2col_two_enabled() && mode == 0 && upt == 1 && 4 * cols <= 512;
  p.lanes = (two ? 2 : 1) * num_sms() / p.nchunks;
  if (p.lanes > IT_SLOTS) p.lanes = IT_SLOTS;
  if (p.lanes > p.total_tiles) p.lanes = p.total_tiles;
  p.parts = parts;
  p.row_bytes = (uint32_t)Cc * 2;
  p.swz = p.row_bytes == 128 ? 128 : (p.row_bytes == 64 ? 64 : 0);
  CUtensorMap mG, mIn, mOut;
  const i64 dg[3] = {(i64)8 * W, H, B};
  const i64 sg[3] = {1, (i64)8 * W, (i64)8 * W * H};
  const int bg[3] = {8 * IT_PITCH, IT_TH + 2, 1};
  if (!make_map_ex(&mG, G16, 3, dg, sg, bg, 2, 0)) return false;
  const i64 d[4] = {C, W, H, B};
  const i64 st[4] = {1, C, (i64)C * W, (i64)C * W * H};
  const int bx[4] = {Cc, IT_TW, IT_TH, 1};
  if (!make_map_ex(&mOut, out, 4, d, st, bx, 2, p.swz)) return false;
  if (mode != 1) {
    if (!make_map_ex(&mIn, feat, 4, d, st, bx, 2, p.swz)) return false;
  } else {
    mIn = mOut;
  }
  const int nf = two ? 4 : 5;
  const size_t smem = 1024 + nf * 16384 + 30720 + IT_NGS * IT_PATCH_STRIDE + 8 * (2 + IT_NGS + nf) + 64 + 16 * 16 * 4 + 64;
#define RF_IT_LAUNCH(M, U, T)                                                                                               \
  do {                                                                                                                       \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      if (cudaFuncSetAttribute(k_im2col_tc<M, U, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
        return false;                                                                                                        \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    launch_pdl(k_im2col_tc<M, U, T>, dim3(grid), dim3(IT_THREADS), smem, ctx.stream, mG, mIn, mOut, p);                      \
  } while (0)
  const int grid = p.lanes * p.nchunks;
  if (mode == 0) { if (two) RF_IT_LAUNCH(0, 1, true); else if (upt == 1) RF_IT_LAUNCH(0, 1, false); else RF_IT_LAUNCH(0, 2, false); }
  else if (mode == 1) { if (upt == 1) RF_IT_LAUNCH(1, 1, false); else RF_IT_LAUNCH(1, 2, false); }
  else if (mode == 2) { if (upt == 1) RF_IT_LAUNCH(2, 1, false); else RF_IT_LAUNCH(2, 2, false); }
  else { if (upt == 1) RF_IT_LAUNCH(3, 1, false); else RF_IT_LAUNCH(3, 2, false); }
#undef RF_IT_LAUNCH
  return true;
}

// xmod = feat * (1 + a*sig(conv(LL)) + b*tanh(conv(yh)) + g*sig(conv(cr,cb))); partial [B][IT_SLOTS][C] = channel sums per CTA
bool launch_flca_mod_tc(Ctx& ctx, const void* feat, const void* G16, const float* w36, const float* abg, void* xmod,
                        float* partial, int B, int Hf, int Wf, int C, int* used_slots) {
  if (used_slots) {              // == p.lanes of run_im2col_tc: the slots (per image) the kernel writes
    const int Cc = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 48);
    const i64 total = (i64)cdiv(Wf, IT_TW) * cdiv(Hf, IT_TH) * B;
    const bool two = im2col_two_enabled() && Cc == 32;     // (== run_im2col_tc: mode 0, one unit per thread, N = 96)
    int lanes = (two ? 2 : 1) * num_sms() / (C / Cc);
    if (lanes > IT_SLOTS) lanes = IT_SLOTS;
    if (lanes > total) lanes = (int)total;
    *used_slots = lanes;
  }
  return run_im2col_tc(ctx, 0, feat, G16, w36, abg, xmod, partial, B, Hf, Wf, C);
}

// out = conv3x3(x_ds; 4 -> d) + bias, bf16 NHWC; x16 = launch_split_bf16x8(x_ds)
bool launch_embed_tc(Ctx& ctx, const void* x16, const float* w, const float* b, void* out, int B, int h, int w_, int d) {
  return run_im2col_tc(ctx, 1, nullptr, x16, w, b, out, nullptr, B, h, w_, d);
}

// FLCA_Pyramid spatial steps: mode 0 = level `level` (GA16 = [hi|lo] of (LL1, hi1, LL2, hi2)), mode 1 = chroma
// (GB16 = [hi|lo] of (cr, cb, mag, 0)); gates [B][6]
bool launch_pyr_spatial_tc(Ctx& ctx, const void* x, const void* G16, const float* w54, const float* gates, void* xs, int mode,
                           int level, int B, int Hf, int Wf, int C) {
  return run_im2col_tc(ctx, mode == 0 ? 2 : 3, x, G16, w54, gates, xs, nullptr, B, Hf, Wf, C, level);
}

}  // namespace rf
