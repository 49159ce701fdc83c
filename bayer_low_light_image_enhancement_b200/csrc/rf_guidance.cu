// Frame prologue: Bayer pack + luma (fused), per-image max normalisation, Haar high-band magnitude, and the
// bilinear-resized guidance maps of every U-Net stage.  The reference recomputes the DWT of `y` in all seven
// blocks (FLCA_RF.py:140) and resizes per block (FLCA_RF.py:145-148); the results only depend on the stage, so
// they are computed once per frame and stage here.  All maps are fp32.
#include "rf_kernels.cuh"

namespace rf {

// raw [B,1,H,W] -> x_ds [B,h,w,4] (float4 per packed pixel: R,G1,G2,B = (0,0),(0,1),(1,0),(1,1)), y_raw, ymax.
// downshuffle: FLCA_RF.py:18-33;  luma: FLCA_RF.py:89-92 (no FMA contraction, same op order as the reference).
__global__ void k_pack_luma(const float* __restrict__ raw, float4* __restrict__ x_ds, float* __restrict__ y_raw,
                            float* ymax, const float* __restrict__ rgb_w, int H, int W) {
  const int h = H >> 1, w = W >> 1, w2 = w >> 1;  // each thread: 2 packed pixels
  i64 b = blockIdx.y;
  const float rw = rgb_w[0], gw = rgb_w[1], bw = rgb_w[2];
  float m = -INFINITY;
  i64 total = (i64)h * w2;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    int xv = (int)(idx % w2);
    int y = (int)(idx / w2);
    const float* p0 = raw + (b * H + 2 * y) * (i64)W + xv * 4;
    float4 t = *reinterpret_cast<const float4*>(p0);
    float4 u = *reinterpret_cast<const float4*>(p0 + W);
    float4 q0 = make_float4(t.x, t.y, u.x, u.y), q1 = make_float4(t.z, t.w, u.z, u.w);
    i64 o = (b * h + y) * (i64)w + xv * 2;
    x_ds[o] = q0;
    x_ds[o + 1] = q1;
    float g0 = __fmul_rn(0.5f, __fadd_rn(q0.y, q0.z)), g1 = __fmul_rn(0.5f, __fadd_rn(q1.y, q1.z));
    float y0 = __fadd_rn(__fadd_rn(__fmul_rn(rw, q0.x), __fmul_rn(gw, g0)), __fmul_rn(bw, q0.w));
    float y1 = __fadd_rn(__fadd_rn(__fmul_rn(rw, q1.x), __fmul_rn(gw, g1)), __fmul_rn(bw, q1.w));
    *reinterpret_cast<float2*>(y_raw + o) = make_float2(y0, y1);
    m = fmaxf(m, fmaxf(y0, y1));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_float(ymax + b, m);
}

void launch_pack_luma(Ctx& ctx, const float* raw, float* x_ds, float* y_raw, float* ymax, const float* rgb_w, int B,
                      int H, int W) {
  if (ctx.dry) return;
  i64 total = (i64)(H / 2) * (W / 4);
  unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_PACK_LUMA, (4.0 + 4.0 + 1.0) * B * H * W);
  k_pack_luma<<<dim3(gx, B), 256, 0, ctx.stream>>>(raw, (float4*)x_ds, y_raw, ymax, rgb_w, H, W);
}

// y = y_raw / max(ymax, eps); cr = r - y; cb = b - y   (FLCA_RF.py:94-96; r, b are NOT normalised)
__global__ void k_luma_finalize(const float4* __restrict__ x_ds, const float* __restrict__ y_raw,
                                const float* __restrict__ ymax, float eps, float* __restrict__ y, float* __restrict__ cr,
                                float* __restrict__ cb, i64 hw) {
  i64 b = blockIdx.y;
  float d = fmaxf(ymax[b], eps);
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    float4 q = x_ds[b * hw + i];
    float v = __fdiv_rn(y_raw[b * hw + i], d);
    y[b * hw + i] = v;
    cr[b * hw + i] = __fsub_rn(q.x, v);
    cb[b * hw + i] = __fsub_rn(q.w, v);
  }
}

void launch_luma_finalize(Ctx& ctx, const float* x_ds, const float* y_raw, const float* ymax, float eps, float* y,
                          float* cr, float* cb, int B, int h, int w) {
  if (ctx.dry) return;
  i64 hw = (i64)h * w;
  unsigned gx = (unsigned)(cdivl(hw, 256) < 8 * num_sms() ? cdivl(hw, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_LUMA_NORM, 32.0 * B * hw);
  k_luma_finalize<<<dim3(gx, B), 256, 0, ctx.stream>>>((const float4*)x_ds, y_raw, ymax, eps, y, cr, cb, hw);
}

// HaarDWT of a 1-channel map + |high| = sqrt(LH^2+HL^2+HH^2+1e-8)  (FLCA_RF.py:56-73,140-141).
// filt16 = the registered [4,1,2,2] buffer (device).  Reflect pad right/bottom for odd sizes.
__global__ void k_dwt_high(const float* __restrict__ y, const float* __restrict__ filt16, float* __restrict__ LL,
                           float* __restrict__ yh, int Hy, int Wy) {
  const int H2 = (Hy + 1) >> 1, W2 = (Wy + 1) >> 1;
  i64 b = blockIdx.y;
  __shared__ float f[16];
  if (threadIdx.x < 16) f[threadIdx.x] = filt16[threadIdx.x];
  __syncthreads();
  i64 total = (i64)H2 * W2;
  const float* p = y + b * (i64)Hy * Wy;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    int x = (int)(idx % W2), yy = (int)(idx / W2);
    int y0 = 2 * yy, y1 = 2 * yy + 1, x0 = 2 * x, x1 = 2 * x + 1;
    if (y1 >= Hy) y1 = Hy - 2;
    if (x1 >= Wy) x1 = Wy - 2;
    float a = p[(i64)y0 * Wy + x0], bb = p[(i64)y0 * Wy + x1], c = p[(i64)y1 * Wy + x0], d = p[(i64)y1 * Wy + x1];
    float s[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) s[n] = f[n * 4] * a + f[n * 4 + 1] * bb + f[n * 4 + 2] * c + f[n * 4 + 3] * d;
    LL[b * total + idx] = s[0];
    yh[b * total + idx] = sqrtf(s[1] * s[1] + s[2] * s[2] + s[3] * s[3] + 1e-8f);
  }
}

void launch_dwt_high(Ctx& ctx, const float* y, const float* filt16, float* LL, float* yh, int B, int Hy, int Wy) {
  if (ctx.dry) return;
  i64 total = (i64)((Hy + 1) / 2) * ((Wy + 1) / 2);
  unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_DWT_HIGH, 4.0 * B * Hy * Wy + 8.0 * B * total);
  k_dwt_high<<<dim3(gx, B), 256, 0, ctx.stream>>>(y, filt16, LL, yh, Hy, Wy);
}

// F.interpolate(mode='bilinear', align_corners=False) of one map at one output pixel
__device__ __forceinline__ float bilerp(const float* __restrict__ src, int Hs, int Ws, int y0, int y1, float ly, int x0,
                                        int x1, float lx) {
  float v00 = src[(i64)y0 * Ws + x0], v01 = src[(i64)y0 * Ws + x1];
  float v10 = src[(i64)y1 * Ws + x0], v11 = src[(i64)y1 * Ws + x1];
  float top = v00 * (1.f - lx) + v01 * lx;
  float bot = v10 * (1.f - lx) + v11 * lx;
  return top * (1.f - ly) + bot * ly;
}

template <int NG>
__global__ void k_guidance_stage(const float* __restrict__ LL1, const float* __restrict__ yh1, int H1, int W1,
                                 const float* __restrict__ LL2, const float* __restrict__ yh2, int H2, int W2,
                                 const float* __restrict__ cr, const float* __restrict__ cb, int Hy, int Wy,
                                 float* __restrict__ G, float* sums, int Hf, int Wf, uint4* __restrict__ G16a,
                                 uint4* __restrict__ G16b, int y_begin, int y_rows) {
  i64 b = blockIdx.y;
  const i64 total = (i64)Hf * Wf;                     // pixels of one image of G
  const i64 first = (i64)y_begin * Wf, last = first + (i64)y_rows * Wf;   // rows [y_begin, y_begin + y_rows) are produced
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // [hi x4 | lo x4] bf16 of four fp32 maps = the 16-byte pixel of the tensor-core FLCA kernels (rf_im2col_tc.cu)
  auto split4 = [](float a, float bq, float c, float d) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(bq), h2 = __float2bfloat16_rn(c),
                        h3 = __float2bfloat16_rn(d);
    __nv_bfloat162 p0 = __halves2bfloat162(h0, h1), p1 = __halves2bfloat162(h2, h3);
    __nv_bfloat162 q0 = __floats2bfloat162_rn(a - __bfloat162float(h0), bq - __bfloat162float(h1));
    __nv_bfloat162 q1 = __floats2bfloat162_rn(c - __bfloat162float(h2), d - __bfloat162float(h3));
    return make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&q0),
                      *reinterpret_cast<uint32_t*>(&q1));
  };
  for (i64 idx = first + (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < last; idx += (i64)gridDim.x * blockDim.x) {
    int x = (int)(idx % Wf), y = (int)(idx / Wf);
    int ya, yb, xa, xb;
    float ly, lx;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bilinear_taps(y, H1, Hf, ya, yb, ly);
    bilinear_taps(x, W1, Wf, xa, xb, lx);
    v[0] = bilerp(LL1 + b * (i64)H1 * W1, H1, W1, ya, yb, ly, xa, xb, lx);
    v[1] = bilerp(yh1 + b * (i64)H1 * W1, H1, W1, ya, yb, ly, xa, xb, lx);
    int o = 2;
    if (NG == 8) {
      bilinear_taps(y, H2, Hf, ya, yb, ly);
      bilinear_taps(x, W2, Wf, xa, xb, lx);
      v[2] = bilerp(LL2 + b * (i64)H2 * W2, H2, W2, ya, yb, ly, xa, xb, lx);
      v[3] = bilerp(yh2 + b * (i64)H2 * W2, H2, W2, ya, yb, ly, xa, xb, lx);
      o = 4;
    }
    bilinear_taps(y, Hy, Hf, ya, yb, ly);
    bilinear_taps(x, Wy, Wf, xa, xb, lx);
    v[o] = bilerp(cr + b * (i64)Hy * Wy, Hy, Wy, ya, yb, ly, xa, xb, lx);
    v[o + 1] = bilerp(cb + b * (i64)Hy * Wy, Hy, Wy, ya, yb, ly, xa, xb, lx);
    float* g = G + (b * total + idx) * NG;                // (G == nullptr: only the bf16 [hi | lo] pixels are wanted)
    if (G16a != nullptr) G16a[b * total + idx] = split4(v[0], v[1], v[2], v[3]);
    if (NG == 4) {
      if (G != nullptr) *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      v[6] = sqrtf(v[4] * v[4] + v[5] * v[5] + 1e-8f);  // chr_mag, ML_RF.py:172
      if (G != nullptr) {
        *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(g + 4) = make_float4(v[4], v[5], v[6], 0.f);
      }
      if (G16b != nullptr) G16b[b * total + idx] = split4(v[4], v[5], v[6], 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  if (NG == 8 && sums != nullptr) {
    // (`sums` = per-CTA partial slots [B][gridDim.x][8]; launch_sum_slots adds them in order: no float atomics)
    __shared__ float red[8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      sums[((i64)b * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = t;
    }
  }
}

void launch_guidance_stage(Ctx& ctx, const float* LL1, const float* yh1, int H1, int W1, const float* LL2,
                           const float* yh2, int H2, int W2, const float* cr, const float* cb, int Hy, int Wy, float* G,
                           int NG, float* sums, int B, int Hf, int Wf, void* G16a, void* G16b, int y_begin, int y_rows) {
  if (y_rows < 0) { y_begin = 0; y_rows = Hf; }
  i64 total = (i64)y_rows * Wf;
  unsigned gx = (unsigned)(cdivl(total, 256) < 4 * num_sms() ? cdivl(total, 256) : 4 * num_sms());
  // multi-level variant: the maps' sums as per-CTA partial slots + an ordered second stage (bit-reproducible)
  const size_t mk = ctx.arena.mark();
  float* part = (NG == 8 && sums != nullptr) ? ctx.arena.get<float>((size_t)B * gx * 8) : nullptr;
  struct Rel { Arena& a; size_t m; ~Rel() { a.release(m); } } rel{ctx.arena, mk};
  if (ctx.dry) return;
  {
    ScopedLaunch sl(RF_K_GUIDANCE, (G ? 4.0 * NG : 0.0) * B * total + (G16a ? 16.0 : 0.0) * B * total * (G16b ? 2 : 1) +
                                       4.0 * B * (2.0 * H1 * W1 + 2.0 * Hy * Wy));
    if (NG == 4)
      k_guidance_stage<4><<<dim3(gx, B), 256, 0, ctx.stream>>>(LL1, yh1, H1, W1, LL2, yh2, H2, W2, cr, cb, Hy, Wy, G, sums,
                                                            Hf, Wf, (uint4*)G16a, (uint4*)G16b, y_begin, y_rows);
    else
      k_guidance_stage<8><<<dim3(gx, B), 256, 0, ctx.stream>>>(LL1, yh1, H1, W1, LL2, yh2, H2, W2, cr, cb, Hy, Wy, G, part,
                                                            Hf, Wf, (uint4*)G16a, (uint4*)G16b, y_begin, y_rows);
  }
  if (part != nullptr) launch_sum_slots(ctx, part, (int)gx, 8, sums, 8, B);
}

}  // namespace rf
