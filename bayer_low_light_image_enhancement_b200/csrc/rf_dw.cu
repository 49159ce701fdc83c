// Strip-mined depthwise 3x3 on NHWC activations.
//
// One thread owns 4 channels of a strip of L horizontally adjacent pixels and slides a 3-column window along the
// strip: every input column (3 rows x 4 channels) is loaded once and feeds the three outputs it touches, the 36 tap
// weights stay in registers.  That is 3(L+2)/L vector loads per output instead of 9 + weights, which is what moves
// these kernels from L1-bound to HBM-bound.
//   MODE 0: out = [gelu](dw(in) + bias), NHWC                       (conv_ffn.depthwise + GELU, FLCA_RF.py:206-207)
//   MODE 1: in = qkv_pre [.,3C] -> dw(in)+bias split into out = q|k [.,2C] and vout = v [.,C] (both dense NHWC) and the
//           squared norms of the q,k channels -> sumsq[b][2C]  (Attention.qkv_dwconv + F.normalize statistics,
//           FLCA_RF.py:223-229); the Gram q k^T is then a tensor-core kernel with MN-major operands (rf_tc_gemm.cu)
#include "rf_kernels.cuh"

namespace rf {

template <typename T> struct RawVec;
template <> struct RawVec<bf16> {
  typedef uint2 type;
  __device__ static __forceinline__ uint2 zero() { return make_uint2(0u, 0u); }
  __device__ static __forceinline__ void unpack(const uint2& t, float (&v)[4]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};
template <> struct RawVec<float> {
  typedef float4 type;
  __device__ static __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static __forceinline__ void unpack(const float4& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};

template <typename T, int MODE, int L>
__global__ void __launch_bounds__(384, 2)
k_dw_strip(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
           T* __restrict__ vout, float* __restrict__ sumsq, int gelu, int H, int W, int Cn, int C, i64 Ppad, i64 total) {
  // block = (Cn/4 channel groups, SY strips); total = strips per image
  __shared__ float s_sq[MODE == 1 ? 1024 : 1];  // [SY][2C] partial squared norms (SY*2C <= 1024)
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const i64 b = blockIdx.y;
  const i64 strip = (i64)blockIdx.x * blockDim.y + threadIdx.y;
  const int c0 = threadIdx.x * 4;
  if (MODE == 1) {
    for (int i = tid; i < (int)blockDim.y * 2 * C; i += nthr) s_sq[i] = 0.f;
    __syncthreads();
  }
  if (strip < total) {
    const int SW = (W + L - 1) / L;
    const int x0 = (int)(strip % SW) * L, y = (int)(strip / SW);
    const i64 P = (i64)H * W;
    const T* img = in + b * P * Cn;
    float wv[9][4], bs[4];
#pragma unroll
    for (int t = 0; t < 9; ++t) load4(w + t * Cn + c0, wv[t]);
    load4(bias + c0, bs);
    typedef typename RawVec<T>::type raw_t;
    float a[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int k = 0; k < 4; ++k) a[s][k] = 0.f;
    float sq[4] = {0.f, 0.f, 0.f, 0.f};
    const bool is_qk = MODE == 1 && c0 < 2 * C;
#pragma unroll
    for (int j = 0; j < L + 2; ++j) {
      const int xc = x0 - 1 + j;
      float v[3][4];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yy = y + r - 1;
        raw_t rv = RawVec<T>::zero();
        if (yy >= 0 && yy < H && xc >= 0 && xc < W) rv = *reinterpret_cast<const raw_t*>(img + ((i64)yy * W + xc) * Cn + c0);
        RawVec<T>::unpack(rv, v[r]);
      }
      float* aN = a[(j + 2) % 3];  // output q = j      (this column is its left neighbour, kx = 0)
      float* aC = a[(j + 1) % 3];  // output q = j - 1  (kx = 1)
      float* aD = a[j % 3];        // output q = j - 2  (kx = 2) -> complete after this column
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        aN[k] = fmaf(wv[6][k], v[2][k], fmaf(wv[3][k], v[1][k], fmaf(wv[0][k], v[0][k], bs[k])));
        aC[k] = fmaf(wv[7][k], v[2][k], fmaf(wv[4][k], v[1][k], fmaf(wv[1][k], v[0][k], aC[k])));
        aD[k] = fmaf(wv[8][k], v[2][k], fmaf(wv[5][k], v[1][k], fmaf(wv[2][k], v[0][k], aD[k])));
      }
      const int q = j - 2;
      if (q >= 0 && x0 + q < W) {
        float o[4] = {aD[0], aD[1], aD[2], aD[3]};
        if (MODE == 0) {
          if (gelu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = FastMath<T>::value ? gelu_erf_fast(o[k]) : gelu_erf_f(o[k]);
          }
          store4(out + (b * P + (i64)y * W + x0 + q) * Cn + c0, o);
        } else {
          // MODE 1: q,k,v all stay NHWC; squared norms of the ROUNDED q,k values (what the Gram kernel reads)
          T ob[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) from_f(ob[k], o[k]);
          if (is_qk) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float r = to_f(ob[k]);
              sq[k] = fmaf(r, r, sq[k]);
            }
          }
          const i64 pix = b * P + (i64)y * W + x0 + q;
          if (is_qk) store4(out + pix * 2 * C + c0, o);
          else store4(vout + pix * C + (c0 - 2 * C), o);
        }
      }
    }
    if (is_qk) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s_sq[threadIdx.y * 2 * C + c0 + k] = sq[k];
    }
  }
  if (MODE == 1) {
    __syncthreads();
    for (int i = tid; i < 2 * C; i += nthr) {
      float v = 0.f;
      for (int sy = 0; sy < (int)blockDim.y; ++sy) v += s_sq[sy * 2 * C + i];
      if (v != 0.f) atomicAdd(sumsq + b * 2 * C + i, v);
    }
  }
}

template <typename T, int MODE>
static void run_dw_strip(Ctx& ctx, const void* in, const float* w, const float* bias, void* out, void* vout, float* sumsq,
                         int gelu, int B, int H, int W, int Cn, int C, i64 Ppad) {
  constexpr int L = 8;
  const i64 total = (i64)H * cdiv(W, L);   // strips per image
  const int V4 = Cn / 4;
  const int SY = V4 >= 256 ? 1 : 256 / V4;
  dim3 grid((unsigned)cdivl(total, SY), B), block(V4, SY);
  k_dw_strip<T, MODE, L><<<grid, block, 0, ctx.stream>>>((const T*)in, w, bias, (T*)out, (T*)vout, sumsq, gelu, H, W, Cn, C, Ppad,
                                                      total);
}

void launch_dwconv(Ctx& ctx, const void* in, const float* dw_w, const float* dw_b, void* out, int gelu, int B, int H, int W,
                   int Cn, int kernel_id) {
  if (ctx.dry) return;
  double px = (double)B * H * W;
  ScopedLaunch sl(kernel_id, 2.0 * px * Cn * esize(ctx.dtype), 18.0 * px * Cn);
  if (ctx.dtype == RF_BF16 && launch_dwconv_tma(ctx, in, dw_w, dw_b, out, gelu, B, H, W, Cn)) return;
  if (ctx.dtype == RF_BF16) run_dw_strip<bf16, 0>(ctx, in, dw_w, dw_b, out, nullptr, nullptr, gelu, B, H, W, Cn, 0, 0);
  else run_dw_strip<float, 0>(ctx, in, dw_w, dw_b, out, nullptr, nullptr, gelu, B, H, W, Cn, 0, 0);
}

int launch_dwqkv_nhwc(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* qk, void* v, float* sumsq,
                      int B, int H, int W, int C, float* sq_part) {
  if (ctx.dry) return 0;
  double px = (double)B * H * W;
  ScopedLaunch sl(RF_K_DW_QKV_GRAM, 6.0 * px * C * esize(ctx.dtype), 54.0 * px * C);
  int nslots = 0;
  if (ctx.dtype == RF_BF16 && launch_dwqkv_tma(ctx, qkv_pre, dw_w, dw_b, qk, v, sumsq, B, H, W, C, sq_part, &nslots))
    return sq_part ? nslots : 0;
  if (ctx.band != nullptr) {   // only the TMA kernel restricts the norms to the band's interior rows
    recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return 0;
  }
  if (ctx.dtype == RF_BF16) run_dw_strip<bf16, 1>(ctx, qkv_pre, dw_w, dw_b, qk, v, sumsq, 0, B, H, W, 3 * C, C, 0);
  else run_dw_strip<float, 1>(ctx, qkv_pre, dw_w, dw_b, qk, v, sumsq, 0, B, H, W, 3 * C, C, 0);
  return 0;
}

}  // namespace rf
