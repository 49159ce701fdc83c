// Fused  LayerNorm -> 1x1 conv -> depthwise 3x3  (bf16 mode, C in {32, 64}: the two full-resolution stages).
//   MODE 1: qkv = qkv_dwconv(qkv(norm1(x)))  -> q|k [.,2C], v [.,C] and the squared norms of q,k   (FLCA_RF.py:221-229)
//   MODE 2: h = gelu(depthwise(pointwise1(norm2(x))))                                              (FLCA_RF.py:204-207)
// The 1x1-conv output (3C / 2C channels wide -- the largest tensor of the block) never goes to HBM: per tile the CTA
// fetches the halo'd patch of the C-channel input with ONE TMA box (zero fill outside the image), runs the 1x1 conv of
// the patch for its channel chunk on the tensor cores (patch pixels = UMMA rows, K = C, weights resident in shared
// memory), applies the folded LayerNorm (per-pixel mean / rstd from the statistics buffer) + bias in the TMEM read-out,
// writes the bf16 result as the depthwise input tile in shared memory (rows padded by 16 B: conflict-free for the
// row-per-thread writes) and then runs exactly the sliding-window depthwise loop of rf_dw_tma.cu on it.
// HBM traffic per pixel: 1.2*C (input with halo) + Cn (output) instead of C + Cn (GEMM) + 1.2*Cn + Cn (depthwise).
// Pixels outside the image must be ZERO in the depthwise input (conv padding), not conv1x1(0) = bias: masked on read-out.
#include "rf_kernels.cuh"
#include "rf_tma.cuh"
#include "rf_dw_math.cuh"

namespace rf {

constexpr int PD_TH = 16;            // output rows per tile; TH + 2 multiple of 3 (window rotation)
constexpr int PD_THREADS = 512;

struct PwDwParams {
  const float* cs;       // [Cn] row sums of the folded 1x1 weights
  const float* pbias;    // [Cn] folded 1x1 bias
  const float* stats;    // [B*H*W][npart] float2 (sum, sumsq) of the input rows
  int npart;
  float invC, eps;
  const float* w;        // [9][Cn] depthwise taps
  const float* bias;     // [Cn] depthwise bias
  bf16* out;             // [B,H,W,Cn]; MODE 1: q|k [B,H,W,C2]
  bf16* vout;            // MODE 1: v [B,H,W,Cn-C2]
  float* sumsq;          // MODE 1: [B][C2]
  int H, W, C, Cn, C2, B;
  int CC, nvec, TW;      // output-channel chunk, 4-channel vectors per pixel of a chunk, tile width
  int npix, nmt;         // halo patch pixels (TH+2)*(TW+2), 128-row MMA tiles covering them
  int tiles_x, tiles_y, nchunks, sp_tiles, lanes;
  uint32_t xrow;         // bytes per input pixel (C*2 = swizzle span: 64 or 128)
  uint32_t x_bytes;      // TMA bytes of one patch = npix * xrow
  uint32_t dpx;          // depthwise-tile pixel pitch in bytes (CC*2 + 16)
};

template <int MODE>
__global__ void __launch_bounds__(PD_THREADS, 1)
k_pwdw(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const PwDwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base;                                          // input patch, nmt*128 rows of xrow bytes (K-major, swizzled)
  const uint32_t sWt = sX + (uint32_t)p.nmt * 128u * p.xrow;         // resident 1x1 weights of this chunk: CC rows of xrow bytes
  const uint32_t sD = sWt + (((uint32_t)p.CC * p.xrow + 1023u) & ~1023u);   // depthwise input tile: npix pixels of dpx bytes
  const uint32_t sC = sD + (((uint32_t)p.npix * p.dpx + 127u) & ~127u);     // cs[CC], pbias[CC] floats
  const uint32_t bars = sC + 2u * p.CC * 4u;                         // x_full, acc_full, w_full; tile coords; tmem slot
  uint8_t* tail = smem_raw + (bars - smem_u32(smem_raw));
  int4* s_tile = reinterpret_cast<int4*>(tail + 32);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(tail + 48);
  float* s_sum = reinterpret_cast<float*>(tail + 64);                // [CC] flush scratch (MODE 1)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x % p.nchunks, lane_id = blockIdx.x / p.nchunks;
  const int stride = p.lanes;

  if (tid == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    mbar_init(bars + 16, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(bars + 48, 512);
  for (int i = tid; i < p.CC; i += PD_THREADS) {
    float* sc = reinterpret_cast<float*>(smem_raw + (sC - smem_u32(smem_raw)));
    sc[i] = p.cs[chunk * p.CC + i];
    sc[p.CC + i] = p.pbias[chunk * p.CC + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto issue_x = [&](int t) {                                        // thread 0: patch of spatial tile t
    int r = t;
    const int tx = r % p.tiles_x; r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    const int b = r / p.tiles_y;
    s_tile[0] = make_int4(tx, ty, b, 0);
    mbar_expect_tx(bars, p.x_bytes);
    tma_load_4d(sX, &mapX, bars, 0, tx * p.TW - 1, ty * PD_TH - 1, b);
  };
  if (tid == 0) {
    mbar_expect_tx(bars + 16, (uint32_t)p.CC * p.xrow);
    tma_load_3d(sWt, &mapW, bars + 16, 0, chunk * p.CC, 0);
    if (lane_id < p.sp_tiles) issue_x(lane_id);
  }

  // ---- depthwise role of this thread (as in rf_dw_tma.cu) -----------------------------------------------------------
  const int dx = tid / p.nvec, cv = tid - dx * p.nvec;
  const bool active = dx < p.TW;
  float2 wv[9][2], bs[2];
  float sq[4] = {0.f, 0.f, 0.f, 0.f};
  const int c0 = chunk * p.CC + cv * 4;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(p.w + (i64)k * p.Cn + c0);
      wv[k][0] = make_float2(w4.x, w4.y);
      wv[k][1] = make_float2(w4.z, w4.w);
    }
    const float4 b4 = *reinterpret_cast<const float4*>(p.bias + c0);
    bs[0] = make_float2(b4.x, b4.y);
    bs[1] = make_float2(b4.z, b4.w);
  }
  const bool is_qk = MODE == 1 && c0 < p.C2;
  const int ocn = MODE == 1 ? (is_qk ? p.C2 : p.Cn - p.C2) : p.Cn;
  bf16* const obase = (MODE == 1 && !is_qk ? p.vout : p.out) + (MODE == 1 && !is_qk ? c0 - p.C2 : c0);
  const i64 opitch = (i64)p.W * ocn;
  int cur_b = -1;
  const uint32_t pitch = (uint32_t)(p.TW + 2) * p.dpx;           // bytes per halo row of the depthwise tile
  const uint32_t toff = (uint32_t)dx * p.dpx + (uint32_t)cv * 8u;
  auto flush = [&](int b) {
    for (int i = tid; i < p.CC; i += PD_THREADS) s_sum[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&s_sum[cv * 4 + k], sq[k]);
    }
    __syncthreads();
    for (int i = tid; i < p.CC; i += PD_THREADS) {
      const int c = chunk * p.CC + i;
      if (c < p.C2 && s_sum[i] != 0.f) atomicAdd(p.sumsq + (i64)b * p.C2 + c, s_sum[i]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sq[k] = 0.f;
  };

  // ---- 1x1-conv read-out role: warp w reads TMEM lane quadrant w & 3 of MMA tiles (w >> 2), (w >> 2) + 4, ... -------
  const int quad = warp & 3;
  const uint32_t idesc = make_idesc_m128(p.CC);
  const int ksteps = p.C >> 4;
  const int bkk = (int)(p.xrow >> 1);                                // K elements per row = C (one k-block)
  const int hw = p.TW + 2;

  int it = 0;
  for (int t = lane_id; t < p.sp_tiles; t += stride, ++it) {
    // (1) tensor cores: conv1x1 of the whole patch for this chunk
    if (tid == 0) {
      if (it == 0) mbar_wait(bars + 16, 0);
      mbar_wait(bars, it & 1);
      tc_fence_after();
      const uint64_t bdesc = make_kmajor_desc(sWt, bkk);
      for (int mt = 0; mt < p.nmt; ++mt) {
        const uint64_t adesc = make_kmajor_desc(sX + (uint32_t)mt * 128u * p.xrow, bkk);
        for (int k = 0; k < ksteps; ++k)
          umma_f16(tmem_base + (uint32_t)(mt * p.CC), adesc + 2u * k, bdesc + 2u * k, idesc, k ? 1u : 0u);
      }
      umma_commit(bars + 8);
    }
    mbar_wait(bars + 8, it & 1);                                     // accumulators complete (and the patch buffer is free)
    tc_fence_after();
    const int4 tc = s_tile[0];
    const int tx = tc.x, ty = tc.y, b = tc.z;
    __syncthreads();                                                 // everybody has read the tile coordinates
    if (tid == 0 && t + stride < p.sp_tiles) issue_x(t + stride);    // next patch lands under the read-out + depthwise
    if (MODE == 1 && b != cur_b) {
      if (cur_b >= 0 && is_qk) flush(cur_b);
      cur_b = b;
    }
    // (2) read-out: folded LayerNorm + bias, zero outside the image, bf16 -> depthwise input tile
    {
      const float* sc = reinterpret_cast<const float*>(smem_raw + (sC - smem_u32(smem_raw)));
      for (int mt = warp >> 2; mt < p.nmt; mt += 4) {
        const int q = mt * 128 + quad * 32 + lane;                   // patch pixel of this thread
        const int hy = q / hw, hx = q - hy * hw;
        const int y = ty * PD_TH - 1 + hy, x = tx * p.TW - 1 + hx;
        const bool inside = q < p.npix && y >= 0 && y < p.H && x >= 0 && x < p.W;
        float rs = 0.f, nm = 0.f;
        if (inside) {
          const float2* st = reinterpret_cast<const float2*>(p.stats) + (((i64)b * p.H + y) * p.W + x) * p.npart;
          float su = 0.f, sq2 = 0.f;
          for (int i = 0; i < p.npart; ++i) {
            const float2 v = __ldg(st + i);
            su += v.x; sq2 += v.y;
          }
          const float mu = su * p.invC;
          rs = rsqrtf(fmaxf(sq2 * p.invC - mu * mu, 0.f) + p.eps);
          nm = -rs * mu;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * p.CC);
        const uint32_t drow = sD + (uint32_t)q * p.dpx;
        for (int c = 0; c < p.CC; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          if (q < p.npix) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float a0 = 0.f, a1 = 0.f;
              if (inside) {
                a0 = fmaf(rs, __uint_as_float(v[j]), fmaf(nm, sc[c + j], sc[p.CC + c + j]));
                a1 = fmaf(rs, __uint_as_float(v[j + 1]), fmaf(nm, sc[c + j + 1], sc[p.CC + c + j + 1]));
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(drow + (uint32_t)c * 2u), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3])
                         : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(drow + (uint32_t)c * 2u + 16u), "r"(pk[4]),
                         "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                         : "memory");
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                                 // depthwise tile complete; TMEM drained
    // (3) depthwise 3x3 on the tile (sliding 3-row window, packed FFMA2)
    if (active) {
      uint32_t src = sD + toff;
      const int xo = tx * p.TW + dx;
      bf16* optr = obase + (((i64)b * p.H + (i64)ty * PD_TH) * p.W + xo) * ocn - 2 * opitch;
      const unsigned rows_ok = xo < p.W ? (unsigned)min(PD_TH, p.H - ty * PD_TH) : 0u;
      float2 acc[3][2];
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int g = 0; g < (PD_TH + 2) / 3; ++g) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = 3 * g + j;
          float2 v[3][2];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) unpack_bf16x4(lds64(src + (uint32_t)kx * p.dpx), v[kx]);
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
          const bool ok = (unsigned)(r - 2) < rows_ok;
          float2 o0 = aD[0], o1 = aD[1];
          if (MODE == 2) {
            o0 = gelu_erf2(o0);
            o1 = gelu_erf2(o1);
          }
          uint2 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
          h[0] = __floats2bfloat162_rn(o0.x, o0.y);
          h[1] = __floats2bfloat162_rn(o1.x, o1.y);
          if (MODE == 1) {
            if (!ok) pk = make_uint2(0u, 0u);
            if (is_qk) {
              float2 rv[2];
              unpack_bf16x4(pk, rv);
              sq[0] = fmaf(rv[0].x, rv[0].x, sq[0]);
              sq[1] = fmaf(rv[0].y, rv[0].y, sq[1]);
              sq[2] = fmaf(rv[1].x, rv[1].x, sq[2]);
              sq[3] = fmaf(rv[1].y, rv[1].y, sq[3]);
            }
          }
          if (ok) *reinterpret_cast<uint2*>(optr) = pk;
          optr += opitch;
        }
      }
    }
    __syncthreads();                                                 // the depthwise tile may be overwritten
  }
  if (MODE == 1 && cur_b >= 0 && is_qk) flush(cur_b);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// mode 1: qkv path (Cn = 3C, C2 = 2C); mode 2: FFN path (Cn = 2C).  x: bf16 [B,H,W,C] (the raw block input / x1);
// Wf: folded 1x1 weights T [Cn][C]; cs / pbias: [Cn]; stats: [rows][npart] float2.  false = shape not supported.
static bool run_pwdw(Ctx& ctx, int mode, const void* x, const void* Wf, const float* cs, const float* pbias, const float* stats,
                     int npart, const float* dw_w, const float* dw_b, void* out, void* vout, float* sumsq, int B, int H, int W,
                     int C, int Cn, int C2) {
  static int enabled = -1;                // debugging aid: RAWFORMER_B200_NO_PWDW=1 keeps the separate GEMM + depthwise
  if (enabled < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_PWDW");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  if (!enabled || !tcgen05_enabled() || ctx.dtype != RF_BF16) return false;
  if (C != 32 && C != 64) return false;
  const int CC = (mode == 1) ? C : 64;    // qkv: a chunk must not straddle the q|k / v boundary; FFN: 2C is a multiple of 64
  if (Cn % CC || (mode == 1 && C2 % CC)) return false;
  PwDwParams p;
  p.cs = cs; p.pbias = pbias; p.stats = stats; p.npart = npart; p.invC = 1.0f / (float)C; p.eps = 1e-5f;
  p.w = dw_w; p.bias = dw_b; p.out = (bf16*)out; p.vout = (bf16*)vout; p.sumsq = sumsq;
  p.H = H; p.W = W; p.C = C; p.Cn = Cn; p.C2 = C2; p.B = B;
  p.CC = CC; p.nvec = CC / 4; p.TW = PD_THREADS / p.nvec;
  p.npix = (PD_TH + 2) * (p.TW + 2);
  p.nmt = cdiv(p.npix, 128);
  if (p.nmt * CC > 512) return false;
  p.tiles_x = cdiv(W, p.TW); p.tiles_y = cdiv(H, PD_TH); p.nchunks = Cn / CC;
  const i64 sp = (i64)p.tiles_x * p.tiles_y * B;
  if (sp > 0x7fffffff || p.nchunks > num_sms()) return false;
  p.sp_tiles = (int)sp;
  p.lanes = num_sms() / p.nchunks;
  if (p.lanes > p.sp_tiles) p.lanes = p.sp_tiles;
  p.xrow = (uint32_t)C * 2;
  p.x_bytes = (uint32_t)p.npix * p.xrow;
  p.dpx = (uint32_t)CC * 2 + 16;
  const size_t smem = 1024 + (size_t)p.nmt * 128 * p.xrow + (((size_t)CC * p.xrow + 1023) & ~(size_t)1023) +
                      (((size_t)p.npix * p.dpx + 127) & ~(size_t)127) + 2 * CC * 4 + 64 + CC * 4 + 64;
  if (smem > 227 * 1024) return false;
  CUtensorMap mX, mW;
  const i64 dX[4] = {C, W, H, B};
  const i64 sX[4] = {1, C, (i64)C * W, (i64)C * W * H};
  const int bX[4] = {C, p.TW + 2, PD_TH + 2, 1};
  if (!make_map_ex(&mX, x, 4, dX, sX, bX, 2, (int)p.xrow)) return false;
  const i64 dW[3] = {C, Cn, 1};
  const i64 sW[3] = {1, C, (i64)C * Cn};
  const int bW[3] = {C, CC, 1};
  if (!make_map_ex(&mW, Wf, 3, dW, sW, bW, 2, (int)p.xrow)) return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_pwdw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_pwdw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return false;
    attr_set = true;
  }
  const int grid = p.lanes * p.nchunks;
  if (mode == 1) k_pwdw<1><<<grid, PD_THREADS, smem, ctx.stream>>>(mX, mW, p);
  else k_pwdw<2><<<grid, PD_THREADS, smem, ctx.stream>>>(mX, mW, p);
  return true;
}

bool pwdw_supported(const Ctx& ctx, int C) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RAWFORMER_B200_NO_PWDW");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled && tcgen05_enabled() && ctx.dtype == RF_BF16 && (C == 32 || C == 64);
}

bool launch_qkv_dw_fused(Ctx& ctx, const void* x, const void* Wf, const float* cs, const float* pbias, const float* stats,
                         int npart, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq, int B, int H, int W,
                         int C) {
  return run_pwdw(ctx, 1, x, Wf, cs, pbias, stats, npart, dw_w, dw_b, qk, v, sumsq, B, H, W, C, 3 * C, 2 * C);
}

bool launch_pw1_dw_fused(Ctx& ctx, const void* x, const void* Wf, const float* cs, const float* pbias, const float* stats,
                         int npart, const float* dw_w, const float* dw_b, void* h, int B, int H, int W, int C) {
  return run_pwdw(ctx, 2, x, Wf, cs, pbias, stats, npart, dw_w, dw_b, h, nullptr, nullptr, B, H, W, C, 2 * C, 0);
}

}  // namespace rf
