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
//   MODE 2: out = gelu_erf(dw(in) + bias)            (conv_ffn.depthwise + nn.GELU, FLCA_RF.py:206-207)
#include "rf_kernels.cuh"
#include "rf_tma.cuh"

namespace rf {

constexpr int DT_TH = 16;          // output rows per tile; TH + 2 must be a multiple of 3 (window rotation)
constexpr int DT_THREADS = 512;

struct DwTmaParams {
  const float* w;      // [9][Cn]
  const float* bias;   // [Cn]
  bf16* out;           // [B,H,W,Cn]; MODE 1: q|k [B,H,W,C2]
  bf16* vout;          // MODE 1: v [B,H,W,Cn-C2]
  float* sumsq;        // MODE 1: [B][C2]
  int H, W, Cn, C2;
  int CC, nvec, TW;    // channel chunk, 4-channel vectors per pixel of a chunk, tile width
  int tiles_x, tiles_y, nchunks, B;
  int total_tiles;
  uint32_t tile_bytes; // (TH+2)*(TW+2)*CC*2
  uint32_t buf_stride; // tile_bytes rounded up to 128
};

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void unpack_bf16x4(const uint2& t, float (&v)[4]) {
  v[0] = __uint_as_float(t.x << 16);
  v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16);
  v[3] = __uint_as_float(t.y & 0xffff0000u);
}

template <int MODE>
__global__ void __launch_bounds__(DT_THREADS, 1)
k_dw_tma(const __grid_constant__ CUtensorMap mapIn, const DwTmaParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bars = base + 2 * p.buf_stride;                 // two mbarriers
  float* s_sum = reinterpret_cast<float*>(smem_raw + (bars + 16 - smem_u32(smem_raw)));   // [CC] flush scratch (MODE 1)
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

  const int stride = gridDim.x;
  auto issue = [&](int t, int s) {
    int r = t;
    const int tx = r % p.tiles_x; r /= p.tiles_x;
    const int ty = r % p.tiles_y; r /= p.tiles_y;
    const int chunk = r % p.nchunks;
    const int b = r / p.nchunks;
    mbar_expect_tx(bars + 8 * s, p.tile_bytes);
    tma_load_4d(base + s * p.buf_stride, &mapIn, bars + 8 * s, chunk * p.CC, tx * p.TW - 1, ty * DT_TH - 1, b);
  };
  if (tid == 0) {
    if ((int)blockIdx.x < p.total_tiles) issue(blockIdx.x, 0);
    if ((int)blockIdx.x + stride < p.total_tiles) issue(blockIdx.x + stride, 1);
  }

  float wv[9][4], bs[4];
  float sq[4] = {0.f, 0.f, 0.f, 0.f};
  int cur_key = -1;
  const uint32_t pitch = (uint32_t)(p.TW + 2) * p.CC * 2;       // bytes per halo row
  const uint32_t toff = (uint32_t)(x * p.CC + cv * 4) * 2;      // this thread's column (kx = 0) inside a halo row

  // MODE 1: add this thread's squared-norm partials of image/chunk `key` to sumsq (block-uniform call)
  auto flush = [&](int key) {
    const int chunk = key % p.nchunks, b = key / p.nchunks;
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
#pragma unroll
    for (int k = 0; k < 4; ++k) sq[k] = 0.f;
  };

  int it = 0;
  for (int t = blockIdx.x; t < p.total_tiles; t += stride, ++it) {
    int r0 = t;
    const int tx = r0 % p.tiles_x; r0 /= p.tiles_x;
    const int ty = r0 % p.tiles_y; r0 /= p.tiles_y;
    const int chunk = r0 % p.nchunks;
    const int b = r0 / p.nchunks;
    const int key = b * p.nchunks + chunk;
    const int c0 = chunk * p.CC + cv * 4;
    if (key != cur_key) {
      if (MODE == 1 && cur_key >= 0) flush(cur_key);
      cur_key = key;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) load4(p.w + (i64)k * p.Cn + c0, wv[k]);
        load4(p.bias + c0, bs);
      }
    }
    const int s = it & 1;
    mbar_wait(bars + 8 * s, (it >> 1) & 1);
    if (active) {
      const uint32_t src = base + s * p.buf_stride + toff;
      const int xo = tx * p.TW + x;
      const bool x_ok = xo < p.W;
      const bool is_qk = MODE == 1 && c0 < p.C2;
      // MODE 1 splits the channels into two dense tensors (chunks never straddle C2: CC divides C)
      const int ocn = MODE == 1 ? (is_qk ? p.C2 : p.Cn - p.C2) : p.Cn;
      const int oc0 = MODE == 1 && !is_qk ? c0 - p.C2 : c0;
      bf16* orow = (MODE == 1 && !is_qk ? p.vout : p.out) + (((i64)b * p.H + (i64)ty * DT_TH) * p.W + xo) * ocn + oc0;
      const i64 opitch = (i64)p.W * ocn;
      const int rows_ok = min(DT_TH, p.H - ty * DT_TH);
      float acc[3][4];
#pragma unroll 1
      for (int g = 0; g < (DT_TH + 2) / 3; ++g) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = 3 * g + j;                    // halo row: feeds outputs r (ky=0), r-1 (ky=1), r-2 (ky=2)
          float v[3][4];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) unpack_bf16x4(lds64(src + (uint32_t)r * pitch + (uint32_t)kx * p.CC * 2), v[kx]);
          float* aN = acc[j];                 // output r      (r % 3 == j)
          float* aM = acc[(j + 2) % 3];       // output r - 1
          float* aD = acc[(j + 1) % 3];       // output r - 2
          if (r < DT_TH) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              aN[k] = fmaf(wv[2][k], v[2][k], fmaf(wv[1][k], v[1][k], fmaf(wv[0][k], v[0][k], bs[k])));
          }
          if (r >= 1 && r <= DT_TH) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              aM[k] = fmaf(wv[5][k], v[2][k], fmaf(wv[4][k], v[1][k], fmaf(wv[3][k], v[0][k], aM[k])));
          }
          if (r >= 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              aD[k] = fmaf(wv[8][k], v[2][k], fmaf(wv[7][k], v[1][k], fmaf(wv[6][k], v[0][k], aD[k])));
            const int o = r - 2;
            if (x_ok && o < rows_ok) {
              float ov[4] = {aD[0], aD[1], aD[2], aD[3]};
              if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ov[k] = gelu_erf_fast(ov[k]);
              }
              uint2 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
              h[0] = __floats2bfloat162_rn(ov[0], ov[1]);
              h[1] = __floats2bfloat162_rn(ov[2], ov[3]);
              if (is_qk) {
                float rv[4];
                unpack_bf16x4(pk, rv);
#pragma unroll
                for (int k = 0; k < 4; ++k) sq[k] = fmaf(rv[k], rv[k], sq[k]);
              }
              *reinterpret_cast<uint2*>(orow + (i64)o * opitch) = pk;
            }
          }
        }
      }
    }
    __syncthreads();                                  // every read of stage s is done
    if (tid == 0 && t + 2 * stride < p.total_tiles) issue(t + 2 * stride, s);
  }
  if (MODE == 1 && cur_key >= 0) flush(cur_key);
}

// false when the shape is not supported by the TMA path (caller falls back to the register-strip kernel)
static bool run_dw_tma(Ctx& ctx, int mode, const void* in, const float* w, const float* bias, void* out, void* vout,
                       float* sumsq, int B, int H, int W, int Cn, int C2) {
  if (Cn % 8 || !tcgen05_enabled()) return false;
  static const int kCC[] = {64, 96, 48, 128, 32, 80, 112, 72, 56, 40, 24, 16, 8};
  const int div = mode == 1 ? C2 / 2 : Cn;   // MODE 1: a chunk must not straddle the q|k / v boundary
  int CC = 0;
  for (int c : kCC)
    if (div % c == 0) { CC = c; break; }
  if (CC == 0) return false;
  DwTmaParams p;
  p.w = w; p.bias = bias; p.out = (bf16*)out; p.vout = (bf16*)vout; p.sumsq = sumsq;
  p.H = H; p.W = W; p.Cn = Cn; p.C2 = C2; p.B = B;
  p.CC = CC; p.nvec = CC / 4;
  p.TW = DT_THREADS / p.nvec;
  if (p.TW > 254) p.TW = 254;
  p.tiles_x = cdiv(W, p.TW); p.tiles_y = cdiv(H, DT_TH); p.nchunks = Cn / CC;
  const i64 total = (i64)p.tiles_x * p.tiles_y * p.nchunks * B;
  if (total > 0x7fffffff) return false;
  p.total_tiles = (int)total;
  p.tile_bytes = (uint32_t)((DT_TH + 2) * (p.TW + 2) * CC * 2);
  p.buf_stride = (p.tile_bytes + 127u) & ~127u;
  const size_t smem = 128 + 2 * (size_t)p.buf_stride + 16 + sizeof(float) * CC;
  if (smem > 227 * 1024) return false;
  CUtensorMap m;
  const i64 d[4] = {Cn, W, H, B};
  const i64 st[4] = {1, Cn, (i64)Cn * W, (i64)Cn * W * H};
  const int bx[4] = {CC, p.TW + 2, DT_TH + 2, 1};
  if (!make_map_ex(&m, in, 4, d, st, bx, 2, false)) return false;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_dw_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_dw_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_dw_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return false;
    attr_set = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (mode == 0) k_dw_tma<0><<<grid, DT_THREADS, smem, ctx.stream>>>(m, p);
  else if (mode == 1) k_dw_tma<1><<<grid, DT_THREADS, smem, ctx.stream>>>(m, p);
  else k_dw_tma<2><<<grid, DT_THREADS, smem, ctx.stream>>>(m, p);
  return true;
}

bool launch_dwconv_tma(Ctx& ctx, const void* in, const float* dw_w, const float* dw_b, void* out, int gelu, int B, int H, int W,
                       int Cn) {
  return run_dw_tma(ctx, gelu ? 2 : 0, in, dw_w, dw_b, out, nullptr, nullptr, B, H, W, Cn, 0);
}

bool launch_dwqkv_tma(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq,
                      int B, int H, int W, int C) {
  return run_dw_tma(ctx, 1, qkv_pre, dw_w, dw_b, qk, v, sumsq, B, H, W, 3 * C, 2 * C);
}

}  // namespace rf
