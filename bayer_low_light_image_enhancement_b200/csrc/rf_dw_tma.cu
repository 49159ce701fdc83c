// TMA-staged depthwise 3x3 on bf16 NHWC activations (sm_100a).
//
// A persistent CTA (512 threads, one per SM) walks tiles of TH x TW pixels x CC channels.  The halo'd input tile
// ((TH+2) x (TW+2) x CC, dense) is fetched by ONE 4-d TMA box per tile into a double-buffered shared-memory stage --
// TMA's out-of-bounds zero fill is the convolution's zero padding, so the inner loop has no bounds checks -- and the
// next tile's box is in flight while this one is computed.  Each thread owns 4 channels of one tile column and slides
// a 3-row window down the tile: every input row costs 3 conflict-free LDS.64 and feeds the three output rows it
// touches; the 36 tap weights stay in registers.  L2->SM traffic is (1+2/TH)(1+2/TW) x the input instead of the 3.75x
// of the register-strip kernel this replaces (rf_dw.cu keeps the fp32 parity path).
//   MODE 0: out = dw(in) + bias
//   MODE 1: same on qkv_pre [.,3C], written as q|k [.,2C] and v [.,C]; additionally sumsq[b][2C] += squared norms of the
//           (bf16-rounded) q,k channels
//           (Attention.qkv_dwconv + F.normalize statistics, FLCA_RF.py:223-229)
//   MODE 2: out = gelu_erf(dw(in) + bias)            (conv_ffn.depthwise + nn.GELU, FLCA_RF.py:206-207); see gelu_erf2
#include "rf_kernels.cuh"
#include "rf_tma.cuh"
#include "rf_dw_math.cuh"

namespace rf {

constexpr int DT_TH = 16;          // output rows per tile; TH + 2 must be a multiple of 3 (window rotation)
constexpr int DT_THREADS = 512;

struct DwTmaParams {
  const float* w;      // [9][Cn]
  const float* bias;   // [Cn]
  bf16* out;           // [B,H,W,Cn]; MODE 1: q|k [B,H,W,C2]
  bf16* vout;          // MODE 1: v [B,H,W,Cn-C2]
  float* sumsq;        // MODE 1: [B][C2] (atomic accumulation; used only when sq_part is null)
  float* sq_part;      // MODE 1: [B][lanes][C2] per-CTA partial squared norms (plain stores: every CTA owns a slot; the
                       // caller sums the slots in a fixed order -> bit-reproducible), pre-zeroed
  int ylo, yhi;        // MODE 1: rows that contribute to sumsq (row-tiled forward: the band's interior; else 0, H)
  int H, W, Cn, C2;
  int wpitch;          // channels per tap row of w (== Cn unless the kernel runs on a channel sub-range of a wider tensor)
  int CC, nvec, TW;    // channel chunk, 4-channel vectors per pixel of a chunk, tile width
  int tiles_x, tiles_y, nchunks, B;
  int sp_tiles;        // spatial tiles per chunk = B * tiles_y * tiles_x
  int lanes;           // CTAs per chunk: CTA = (chunk = blockIdx.x % nchunks, lane = blockIdx.x / nchunks)
  uint32_t tile_bytes; // (TH+2)*(TW+2)*CC*2
  uint32_t buf_stride; // tile_bytes rounded up to 128
};

template <int MODE>
__global__ void __launch_bounds__(DT_THREADS, 1)
k_dw_tma(const __grid_constant__ CUtensorMap mapIn, const DwTmaParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bars = base + 2 * p.buf_stride;                 // two mbarriers
  uint8_t* tail = smem_raw + (bars + 16 - smem_u32(smem_raw));
  int4* s_tile = reinterpret_cast<int4*>(tail);                  // [2] decoded tile (tx, ty, chunk, b) of each stage
  float* s_sum = reinterpret_cast<float*>(tail + 32);            // [CC] flush scratch (MODE 1)
  float4* s_sq = reinterpret_cast<float4*>(tail + 32 + ((sizeof(float) * p.CC + 15) & ~15));   // [DT_THREADS] (MODE 1)
  const int tid = threadIdx.x;
  const int x = tid / p.nvec, cv = tid - x * p.nvec;
  const bool active = x < p.TW;

  if (tid == 0) {
    tma_prefetch_desc(&mapIn);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();

  // Every CTA owns ONE channel chunk and walks the spatial tiles lane, lane + lanes, ...  CTAs of the same lane run
  // the other chunks of the same pixels at the same time, so a 128-byte line that holds two chunks is fetched from
  // DRAM once.  The tap weights stay in registers for the whole kernel.
  const int chunk = blockIdx.x % p.nchunks, lane_id = blockIdx.x / p.nchunks;
  const int stride = p.lanes;
  // the issuing thread decodes the tile once and publishes the coordinates with the stage (the mbarrier's
  // release/acquire orders the plain store before the consumers' loads)
  auto issue = [&](int t, int s) {
    int r = t;
    const int tx = r % p.tiles_x; r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    const int b = r / p.tiles_y;
    s_tile[s] = make_int4(tx, ty, b, 0);
    mbar_expect_tx(bars + 8 * s, p.tile_bytes);
    tma_load_4d(base + s * p.buf_stride, &mapIn, bars + 8 * s, chunk * p.CC, tx * p.TW - 1, ty * DT_TH - 1, b);
  };
  if (tid == 0) {
    if (lane_id < p.sp_tiles) issue(lane_id, 0);
    if (lane_id + stride < p.sp_tiles) issue(lane_id + stride, 1);
  }

  float2 wv[9][2], bs[2];
  float sq[4] = {0.f, 0.f, 0.f, 0.f};
  const int c0 = chunk * p.CC + cv * 4;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(p.w + (i64)k * p.wpitch + c0);
      wv[k][0] = make_float2(w4.x, w4.y);
      wv[k][1] = make_float2(w4.z, w4.w);
    }
    const float4 b4 = *reinterpret_cast<const float4*>(p.bias + c0);
    bs[0] = make_float2(b4.x, b4.y);
    bs[1] = make_float2(b4.z, b4.w);
  }
  const bool is_qk = MODE == 1 && c0 < p.C2;
  // MODE 1 splits the channels into two dense tensors (chunks never straddle C2: CC divides C)
  const int ocn = MODE == 1 ? (is_qk ? p.C2 : p.Cn - p.C2) : p.Cn;
  bf16* const obase = (MODE == 1 && !is_qk ? p.vout : p.out) + (MODE == 1 && !is_qk ? c0 - p.C2 : c0);
  const i64 opitch = (i64)p.W * ocn;
  int cur_b = -1;
  const uint32_t pitch = (uint32_t)(p.TW + 2) * p.CC * 2;       // bytes per halo row
  const uint32_t toff = (uint32_t)(x * p.CC + cv * 4) * 2;      // this thread's column (kx = 0) inside a halo row
  const uint32_t cstep = (uint32_t)p.CC * 2;                    // bytes between horizontally adjacent pixels

  // MODE 1: add this thread's squared-norm partials of image b to sumsq (block-uniform call)
  auto flush = [&](int b) {
    if (p.sq_part != nullptr) {
      // fixed-order combine: thread (x, cv) parks its 4 partials, channel c then adds the TW columns in column order
      s_sq[tid] = make_float4(sq[0], sq[1], sq[2], sq[3]);
      __syncthreads();
      for (int i = tid; i < p.CC; i += DT_THREADS) {
        const int c = chunk * p.CC + i;
        if (c < p.C2) {
          const float* col = reinterpret_cast<const float*>(s_sq) + (i >> 2) * 4 + (i & 3);
          float a0 = 0.f, a1 = 0.f;
          int x2 = 0;
          for (; x2 + 2 <= p.TW; x2 += 2) {
            a0 += col[(x2 * p.nvec) * 4];
            a1 += col[((x2 + 1) * p.nvec) * 4];
          }
          if (x2 < p.TW) a0 += col[(x2 * p.nvec) * 4];
          p.sq_part[((i64)b * p.lanes + lane_id) * p.C2 + c] = a0 + a1;
        }
      }
      __syncthreads();
    } else {
      for (int i = tid; i < p.CC; i += DT_THREADS) s_sum[i] = 0.f;
      __syncthreads();
      if (active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(&s_sum[cv * 4 + k], sq[k]);
      }
      __syncthreads();
      for (int i = tid; i < p.CC; i += DT_THREADS) {
        const int c = chunk * p.CC + i;
        if (c < p.C2 && s_sum[i] != 0.f) atomicAdd(p.sumsq + (i64)b * p.C2 + c, s_sum[i]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sq[k] = 0.f;
  };

  int it = 0;
  for (int t = lane_id; t < p.sp_tiles; t += stride, ++it) {
    const int s = it & 1;
    mbar_wait(bars + 8 * s, (it >> 1) & 1);
    const int4 tc = s_tile[s];
    const int tx = tc.x, ty = tc.y, b = tc.z;
    if (MODE == 1 && b != cur_b) {
      if (cur_b >= 0 && is_qk) flush(cur_b);     // is_qk is uniform over the CTA (one chunk per CTA)
      cur_b = b;
    }
    if (active) {
      uint32_t src = base + s * p.buf_stride + toff;
      const int xo = tx * p.TW + x;
      bf16* optr = obase + (((i64)b * p.H + (i64)ty * DT_TH) * p.W + xo) * ocn - 2 * opitch;   // row of output r - 2
      const unsigned rows_ok = xo < p.W ? (unsigned)min(DT_TH, p.H - ty * DT_TH) : 0u;
      // MODE 1: tile rows [nlo, nlo + nrows) count towards the squared norms
      const int nlo = max(p.ylo - ty * DT_TH, 0);
      const unsigned nrows = xo < p.W ? (unsigned)max(min(DT_TH, p.yhi - ty * DT_TH) - nlo, 0) : 0u;
      float2 acc[3][2];
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int g = 0; g < (DT_TH + 2) / 3; ++g) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = 3 * g + j;                    // halo row: feeds outputs r (ky=0), r-1 (ky=1), r-2 (ky=2)
          float2 v[3][2];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) unpack_bf16x4(lds64(src + (uint32_t)kx * cstep), v[kx]);
          src += pitch;
          float2* aN = acc[j];                 // output r      (r % 3 == j)
          float2* aM = acc[(j + 2) % 3];       // output r - 1
          float2* aD = acc[(j + 1) % 3];       // output r - 2
          // rows beyond the tile (r >= TH for ky=0, r > TH for ky=1) and before it (r < 1, r < 2) only produce
          // accumulators that are never emitted, so the FMAs run unconditionally (no divergence, no branches)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            aN[k] = __ffma2_rn(wv[2][k], v[2][k], __ffma2_rn(wv[1][k], v[1][k], __ffma2_rn(wv[0][k], v[0][k], bs[k])));
            aM[k] = __ffma2_rn(wv[5][k], v[2][k], __ffma2_rn(wv[4][k], v[1][k], __ffma2_rn(wv[3][k], v[0][k], aM[k])));
            aD[k] = __ffma2_rn(wv[8][k], v[2][k], __ffma2_rn(wv[7][k], v[1][k], __ffma2_rn(wv[6][k], v[0][k], aD[k])));
          }
          // output o = r - 2; o < 0 wraps to a huge unsigned, so one compare masks both ends.  Everything but the
          // store is unconditional (straight-line code); masked rows are zeroed so the norms stay exact.
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
            if (is_qk) {
              // rows outside the image / the band's interior are zeroed so the norms stay exact
              const uint2 pn = (unsigned)(r - 2 - nlo) < nrows ? pk : make_uint2(0u, 0u);
              float2 rv[2];
              unpack_bf16x4(pn, rv);
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
    __syncthreads();                                  // every read of stage s (tile and coordinates) is done
    if (tid == 0 && t + 2 * stride < p.sp_tiles) issue(t + 2 * stride, s);
  }
  if (MODE == 1 && cur_b >= 0 && is_qk) flush(cur_b);
}

// false when the shape is not supported by the TMA path (caller falls back to the register-strip kernel)
static bool run_dw_tma(Ctx& ctx, int mode, const void* in, const float* w, const float* bias, void* out, void* vout,
                       float* sumsq, int B, int H, int W, int Cn, int C2, float* sq_part = nullptr, int* nslots = nullptr,
                       int in_pitch = 0, int wpitch = 0) {
  if (in_pitch == 0) in_pitch = Cn;
  if (wpitch == 0) wpitch = Cn;
  if (Cn % 8 || !tcgen05_enabled()) return false;
  static const int kCC[] = {64, 96, 48, 128, 32, 80, 112, 72, 56, 40, 24, 16, 8};
  const int div = mode == 1 ? C2 / 2 : Cn;   // MODE 1: a chunk must not straddle the q|k / v boundary
  int CC = 0;
  for (int c : kCC)
    if (div % c == 0) { CC = c; break; }
  if (CC == 0) return false;
  DwTmaParams p;
  p.w = w; p.bias = bias; p.out = (bf16*)out; p.vout = (bf16*)vout; p.sumsq = sumsq; p.sq_part = sq_part;
  p.H = H; p.W = W; p.Cn = Cn; p.C2 = C2; p.B = B; p.wpitch = wpitch;
  p.ylo = 0; p.yhi = H;
  if (ctx.band != nullptr) { p.ylo = ctx.band->ht; p.yhi = ctx.band->ht + ctx.band->rows_in; }
  p.CC = CC; p.nvec = CC / 4;
  p.TW = DT_THREADS / p.nvec;
  if (p.TW > 254) p.TW = 254;
  p.tiles_x = cdiv(W, p.TW); p.tiles_y = cdiv(H, DT_TH); p.nchunks = Cn / CC;
  const i64 sp = (i64)p.tiles_x * p.tiles_y * B;
  if (sp > 0x7fffffff || p.nchunks > num_sms()) return false;
  p.sp_tiles = (int)sp;
  p.lanes = num_sms() / p.nchunks;
  if (p.lanes > p.sp_tiles) p.lanes = p.sp_tiles;
  p.tile_bytes = (uint32_t)((DT_TH + 2) * (p.TW + 2) * CC * 2);
  p.buf_stride = (p.tile_bytes + 127u) & ~127u;
  const size_t smem = 128 + 2 * (size_t)p.buf_stride + 16 + 32 + sizeof(float) * CC + 16 + (mode == 1 ? sizeof(float4) * DT_THREADS : 0);
  if (nslots) *nslots = p.lanes;
  if (smem > 227 * 1024) return false;
  CUtensorMap m;
  const i64 d[4] = {Cn, W, H, B};
  const i64 st[4] = {1, in_pitch, (i64)in_pitch * W, (i64)in_pitch * W * H};
  const int bx[4] = {CC, p.TW + 2, DT_TH + 2, 1};
  if (!make_map_ex(&m, in, 4, d, st, bx, 2, 0)) return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_dw_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_dw_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_dw_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return false;
    attr_set = true;
  }
  const int grid = p.lanes * p.nchunks;
  if (mode == 0) launch_pdl(k_dw_tma<0>, dim3(grid), dim3(DT_THREADS), smem, ctx.stream, m, p);
  else if (mode == 1) launch_pdl(k_dw_tma<1>, dim3(grid), dim3(DT_THREADS), smem, ctx.stream, m, p);
  else launch_pdl(k_dw_tma<2>, dim3(grid), dim3(DT_THREADS), smem, ctx.stream, m, p);
  return true;
}

bool launch_dwconv_tma(Ctx& ctx, const void* in, const float* dw_w, const float* dw_b, void* out, int gelu, int B, int H, int W,
                       int Cn) {
  return run_dw_tma(ctx, gelu ? 2 : 0, in, dw_w, dw_b, out, nullptr, nullptr, B, H, W, Cn, 0);
}

// depthwise 3x3 (+bias) of Cn channels that are a sub-range of a wider NHWC tensor (row pitch in_pitch elements, tap rows of
// wpitch channels): `in`, `dw_w`, `dw_b` point at the first channel of the range; out is dense [B,H,W,Cn]
bool launch_dwconv_tma_sub(Ctx& ctx, const void* in, int in_pitch, const float* dw_w, int wpitch, const float* dw_b, void* out,
                           int B, int H, int W, int Cn) {
  return run_dw_tma(ctx, 0, in, dw_w, dw_b, out, nullptr, nullptr, B, H, W, Cn, 0, nullptr, nullptr, in_pitch, wpitch);
}

bool launch_dwqkv_tma(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq,
                      int B, int H, int W, int C, float* sq_part, int* nslots) {
  return run_dw_tma(ctx, 1, qkv_pre, dw_w, dw_b, qk, v, sumsq, B, H, W, 3 * C, 2 * C, sq_part, nslots);
}

}  // namespace rf
