// TrueColor head / tail: the learned colour front end and tone-mapping tail that
// TrueColorRawFormer.py ("variant 0") and BayerTORGBColorMultiLvl.py ("variant 1") put around the same U-Net body.
//
//   EnhancedBayerProcessor.forward      TrueColorRawFormer.py:109-142, BayerTORGBColorMultiLvl.py:103-136
//   CameraAwareColorCorrection.forward  TrueColorRawFormer.py:170-185, BayerTORGBColorMultiLvl.py:164-181
//
// fp32 NCHW throughout (these modules work on the 4-plane packed frame and on the 3-channel output image: a handful of
// channels, so they are CUDA-core kernels): a small-channel dense 3x3 convolution with fused activation / residual, the
// per-pixel white balance + colour matrix + luminance step with its per-image max, and the per-pixel gamma / 1x1 MLP /
// tone-curve tail.
#include <math.h>
#include <string.h>

#include "rf_kernels.cuh"

namespace rf {

enum { TC_ACT_NONE = 0, TC_ACT_RELU = 1, TC_ACT_SOFTPLUS = 2, TC_ACT_TANH = 3, TC_ACT_GELU = 4 };

__device__ __forceinline__ float tc_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus defaults

template <int ACT>
__device__ __forceinline__ float tc_act(float v) {
  if (ACT == TC_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == TC_ACT_SOFTPLUS) return tc_softplus(v);
  if (ACT == TC_ACT_TANH) return tanhf(v);
  if (ACT == TC_ACT_GELU) return gelu_erf_f(v);
  return v;
}

// Dense 3x3, stride 1, zero padding 1, few channels: one thread per output pixel keeps all COUT accumulators; the weights
// sit in shared memory as [Cin*9][COUT].  in_scale (optional, [Cin]): the input is multiplied per channel on load (the
// white-balance gains, applied exactly where the reference applies them: before the convolution's products).
template <int COUT, int ACT>
__global__ void __launch_bounds__(256)
k_conv3x3_small(const float* __restrict__ in, const float* __restrict__ in_scale, const float* __restrict__ w,
                const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ out, int Cin, int H, int W) {
  extern __shared__ float sw[];                      // [Cin*9][COUT], then [Cin] scales
  float* ss = sw + Cin * 9 * COUT;
  for (int i = threadIdx.x; i < Cin * 9 * COUT; i += blockDim.x) {
    const int co = i % COUT, k = i / COUT;           // k = ci*9 + tap
    sw[i] = w[(i64)co * Cin * 9 + k];
  }
  for (int i = threadIdx.x; i < Cin; i += blockDim.x) ss[i] = in_scale ? in_scale[i] : 1.f;
  __syncthreads();
  const i64 b = blockIdx.y;
  const i64 P = (i64)H * W;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int y = (int)(p / W), x = (int)(p % W);
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = bias ? bias[co] : 0.f;
  const float* ib = in + b * Cin * P;
  for (int ci = 0; ci < Cin; ++ci) {
    const float sc = ss[ci];
    const bool scaled = in_scale != nullptr;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= W) continue;
        float v = __ldg(ib + ci * P + (i64)yy * W + xx);
        if (scaled) v = __fmul_rn(v, sc);
        const float* wr = sw + (ci * 9 + dy * 3 + dx) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(wr[co], v, acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    float v = tc_act<ACT>(acc[co]);
    if (resid) v += resid[(b * COUT + co) * P + p];
    out[(b * COUT + co) * P + p] = v;
  }
}

struct TcMixP {
  float gains[4];      // applied to the four planes (variant 1: softplus(wb_gains) + 1e-6, computed by the caller)
  float cm[12];        // colour matrix [3][4]: 3x3 + bias column
  float yw[3];         // luminance weights
  int apply_gains;     // variant 0: the planes are the refined ones (gains were applied before demosaic_refine)
};

// planes [B,4,H,W] (R, G1, G2, B) -> rgb_linear [B,3,H,W], chroma_in [B,4,H,W] = (r, g, b, y_raw), ymax[b] = max y_raw
__global__ void __launch_bounds__(256)
k_tc_mix(const float* __restrict__ planes, float* __restrict__ rgb_linear, float* __restrict__ chroma_in, float* __restrict__ ymax,
         const TcMixP q, i64 P) {
  const i64 b = blockIdx.y;
  float m = -INFINITY;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (i64)gridDim.x * blockDim.x) {
    const float* pl = planes + b * 4 * P + p;
    float v0 = pl[0], v1 = pl[P], v2 = pl[2 * P], v3 = pl[3 * P];
    if (q.apply_gains) {
      v0 = __fmul_rn(v0, q.gains[0]); v1 = __fmul_rn(v1, q.gains[1]);
      v2 = __fmul_rn(v2, q.gains[2]); v3 = __fmul_rn(v3, q.gains[3]);
    }
    const float r = v0, g = __fmul_rn(0.5f, __fadd_rn(v1, v2)), bl = v3;
    float lin[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // sum over j of M[i][j] * rgb[j] in index order (einsum / matmul accumulate left to right), then the bias column
      float s = __fmul_rn(q.cm[i * 4 + 0], r);
      s = __fadd_rn(s, __fmul_rn(q.cm[i * 4 + 1], g));
      s = __fadd_rn(s, __fmul_rn(q.cm[i * 4 + 2], bl));
      lin[i] = __fadd_rn(s, q.cm[i * 4 + 3]);
      rgb_linear[(b * 3 + i) * P + p] = lin[i];
    }
    float yr = __fmul_rn(lin[0], q.yw[0]);
    yr = __fadd_rn(yr, __fmul_rn(lin[1], q.yw[1]));
    yr = __fadd_rn(yr, __fmul_rn(lin[2], q.yw[2]));
    float* ci = chroma_in + b * 4 * P + p;
    ci[0] = r; ci[P] = g; ci[2 * P] = bl; ci[3 * P] = yr;
    m = fmaxf(m, yr);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_float(ymax + b, m);
}

// y = y_raw / max(ymax, eps) (in place in plane 3 of chroma_in, and to y_out)
__global__ void __launch_bounds__(256)
k_tc_ynorm(float* __restrict__ chroma_in, const float* __restrict__ ymax, float eps, float* __restrict__ y_out, i64 P) {
  const i64 b = blockIdx.y;
  const float d = fmaxf(ymax[b], eps);
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (i64)gridDim.x * blockDim.x) {
    const float v = __fdiv_rn(chroma_in[(b * 4 + 3) * P + p], d);
    chroma_in[(b * 4 + 3) * P + p] = v;
    y_out[b * P + p] = v;
  }
}

// CameraAwareColorCorrection: per pixel x^(1/gamma) on clamp(x,0,1), 3 -> 64 -> 3 MLP (ReLU), then the shared 1 -> 32 -> 1
// tone curve per channel.  variant 0: out = clamp(sigmoid(tone(t)), 0, 1);  variant 1: out = clamp(t * (0.8 + 0.4 *
// sigmoid(tone(t))), 0, 1).   sm: w1[64][3], b1[64], w2[3][64], b2[3], w3[32], b3[32], w4[32], b4
__global__ void __launch_bounds__(256)
k_tc_color_correction(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                      const float* __restrict__ b3, const float* __restrict__ w4, const float* __restrict__ b4,
                      float* __restrict__ out, float inv_gamma, int variant, i64 P) {
  __shared__ float s[64 * 3 + 64 + 3 * 64 + 3 + 32 + 32 + 32 + 1];
  float* sw1 = s; float* sb1 = sw1 + 192; float* sw2 = sb1 + 64; float* sb2 = sw2 + 192;
  float* sw3 = sb2 + 3; float* sb3 = sw3 + 32; float* sw4 = sb3 + 32; float* sb4 = sw4 + 32;
  for (int i = threadIdx.x; i < 192; i += blockDim.x) { sw1[i] = w1[i]; sw2[i] = w2[i]; }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) { sw3[i] = w3[i]; sb3[i] = b3[i]; sw4[i] = w4[i]; }
  if (threadIdx.x < 3) sb2[threadIdx.x] = b2[threadIdx.x];
  if (threadIdx.x == 0) sb4[0] = b4[0];
  __syncthreads();
  const i64 b = blockIdx.y;
  for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (i64)gridDim.x * blockDim.x) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = powf(fminf(fmaxf(x[(b * 3 + c) * P + p], 0.f), 1.f), inv_gamma);
    float t[3] = {sb2[0], sb2[1], sb2[2]};
    for (int j = 0; j < 64; ++j) {
      float h = sb1[j];
      h = fmaf(sw1[j * 3 + 0], v[0], h);
      h = fmaf(sw1[j * 3 + 1], v[1], h);
      h = fmaf(sw1[j * 3 + 2], v[2], h);
      h = fmaxf(h, 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) t[c] = fmaf(sw2[c * 64 + j], h, t[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a = sb4[0];
      for (int j = 0; j < 32; ++j) a = fmaf(sw4[j], fmaxf(fmaf(sw3[j], t[c], sb3[j]), 0.f), a);
      const float sg = sigmoid_f(a);
      const float o = variant == 0 ? sg : t[c] * (0.8f + 0.4f * sg);
      out[(b * 3 + c) * P + p] = fminf(fmaxf(o, 0.f), 1.f);
    }
  }
}

template <int COUT>
static int launch_conv_small(const float* in, const float* in_scale, const float* w, const float* bias, const float* resid,
                             float* out, int Cin, int act, int B, int H, int W, cudaStream_t st) {
  const i64 P = (i64)H * W;
  const dim3 grid((unsigned)cdivl(P, 256), B);
  const size_t smem = sizeof(float) * ((size_t)Cin * 9 * COUT + Cin);
  ScopedLaunch sl(RF_K_MISC, 4.0 * B * P * (Cin + COUT), 18.0 * B * P * Cin * COUT);
#define RF_TC_CONV(A) k_conv3x3_small<COUT, A><<<grid, 256, smem, st>>>(in, in_scale, w, bias, resid, out, Cin, H, W)
  switch (act) {
    case TC_ACT_NONE: RF_TC_CONV(TC_ACT_NONE); break;
    case TC_ACT_RELU: RF_TC_CONV(TC_ACT_RELU); break;
    case TC_ACT_SOFTPLUS: RF_TC_CONV(TC_ACT_SOFTPLUS); break;
    case TC_ACT_TANH: RF_TC_CONV(TC_ACT_TANH); break;
    case TC_ACT_GELU: RF_TC_CONV(TC_ACT_GELU); break;
    default: return RF_ERR_BAD_ARG;
  }
#undef RF_TC_CONV
  return check_cuda(cudaGetLastError());
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_conv3x3_small(const float* in, const float* in_scale, const float* weight, const float* bias, const float* resid,
                     float* out, int Cin, int Cout, int act, int B, int H, int W, void* stream) {
  if (!in || !weight || !out) return RF_ERR_BAD_ARG;
  if (Cin <= 0 || Cin > 64 || B < 0 || H <= 0 || W <= 0) return RF_ERR_BAD_SHAPE;
  if (B == 0) return RF_OK;
  if (B > 65535) return RF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  switch (Cout) {
    case 2: return launch_conv_small<2>(in, in_scale, weight, bias, resid, out, Cin, act, B, H, W, st);
    case 3: return launch_conv_small<3>(in, in_scale, weight, bias, resid, out, Cin, act, B, H, W, st);
    case 4: return launch_conv_small<4>(in, in_scale, weight, bias, resid, out, Cin, act, B, H, W, st);
    case 16: return launch_conv_small<16>(in, in_scale, weight, bias, resid, out, Cin, act, B, H, W, st);
    case 32: return launch_conv_small<32>(in, in_scale, weight, bias, resid, out, Cin, act, B, H, W, st);
    default: return RF_ERR_UNSUPPORTED;
  }
}

int rf_truecolor_mix(const float* planes, const float* gains_host, int apply_gains, const float* color_matrix_host,
                     const float* y_weights_host, float eps, float* rgb_linear, float* chroma_in, float* y, float* ymax_ws,
                     int B, int H, int W, void* stream) {
  if (!planes || !gains_host || !color_matrix_host || !y_weights_host || !rgb_linear || !chroma_in || !y || !ymax_ws)
    return RF_ERR_BAD_ARG;
  if (B < 0 || H <= 0 || W <= 0) return RF_ERR_BAD_SHAPE;
  if (B == 0) return RF_OK;
  if (B > 65535) return RF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  TcMixP q;
  memcpy(q.gains, gains_host, sizeof(q.gains));
  memcpy(q.cm, color_matrix_host, sizeof(q.cm));
  memcpy(q.yw, y_weights_host, sizeof(q.yw));
  q.apply_gains = apply_gains ? 1 : 0;
  const i64 P = (i64)H * W;
  Ctx ctx;
  ctx.stream = st;
  launch_fill_f32(ctx, ymax_ws, -INFINITY, B);
  unsigned gx = (unsigned)(cdivl(P, 256) < 8 * num_sms() ? cdivl(P, 256) : 8 * num_sms());
  {
    ScopedLaunch sl(RF_K_MISC, 4.0 * B * P * 11);
    k_tc_mix<<<dim3(gx, B), 256, 0, st>>>(planes, rgb_linear, chroma_in, ymax_ws, q, P);
  }
  {
    ScopedLaunch sl(RF_K_MISC, 4.0 * B * P * 3);
    k_tc_ynorm<<<dim3(gx, B), 256, 0, st>>>(chroma_in, ymax_ws, eps, y, P);
  }
  return check_cuda(cudaGetLastError());
}

int rf_color_correction(const float* x, float gamma, int variant, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* w3, const float* b3, const float* w4, const float* b4, float* out, int B,
                        int H, int W, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !out) return RF_ERR_BAD_ARG;
  if (variant != 0 && variant != 1) return RF_ERR_BAD_ARG;
  if (B < 0 || H <= 0 || W <= 0) return RF_ERR_BAD_SHAPE;
  if (B == 0) return RF_OK;
  if (B > 65535) return RF_ERR_UNSUPPORTED;
  const i64 P = (i64)H * W;
  unsigned gx = (unsigned)(cdivl(P, 256) < 16 * num_sms() ? cdivl(P, 256) : 16 * num_sms());
  ScopedLaunch sl(RF_K_MISC, 4.0 * B * P * 6, 2.0 * B * P * (192 + 192 + 3 * 64));
  k_tc_color_correction<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(x, w1, b1, w2, b2, w3, b3, w4, b4, out, 1.0f / gamma, variant, P);
  return check_cuda(cudaGetLastError());
}

}  // extern "C"
