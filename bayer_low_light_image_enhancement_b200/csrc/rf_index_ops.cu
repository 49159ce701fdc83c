// Index / wavelet operators at the C-ABI boundary (NCHW fp32 in and out) and layout helpers.
// These are pure HBM gather/scatter kernels: every global access is a 16-byte vector on the fast paths,
// the 2x2 analysis/synthesis matrices travel as kernel arguments, nothing is staged.
#include "rf_kernels.cuh"

namespace rf {

struct Mat4 { float k[16]; };  // k[n*4 + t]: sub-band n, tap t (a,b,c,d = TL,TR,BL,BR)
struct Ptr4 { float* p[4]; };

// ---------------------------------------------------------------------------------------------
// 2x2 stride-2 analysis: sub_n = sum_t K[n][t] * tap_t.  in [B,C,H,W]; sub-band n of (b,c) is the plane at
// out.p[n] + (b*ob + c)*H2*W2.  Reflect-pads right/bottom when H or W is odd (HaarDWT, FLCA_RF.py:63-66).
// ---------------------------------------------------------------------------------------------
__global__ void k_dwt2x2_vec(const float* __restrict__ in, Ptr4 out, Mat4 m, i64 planes, int C, i64 ob, int H, int W) {
  // fast path: H, W even, W % 8 == 0.  One thread = 4 output pixels of one output row.
  const int W2 = W >> 1, H2 = H >> 1, W8 = W >> 3;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes * H2 * W8;
  if (idx >= total) return;
  int xv = (int)(idx % W8);
  i64 t = idx / W8;
  int y = (int)(t % H2);
  i64 pl = t / H2;
  const float* r0 = in + (pl * H + 2 * y) * (i64)W + xv * 8;
  const float* r1 = r0 + W;
  float4 a0 = *reinterpret_cast<const float4*>(r0), a1 = *reinterpret_cast<const float4*>(r0 + 4);
  float4 b0 = *reinterpret_cast<const float4*>(r1), b1 = *reinterpret_cast<const float4*>(r1 + 4);
  float ta[4] = {a0.x, a0.z, a1.x, a1.z}, tb[4] = {a0.y, a0.w, a1.y, a1.w};
  float tc[4] = {b0.x, b0.z, b1.x, b1.z}, td[4] = {b0.y, b0.w, b1.y, b1.w};
  i64 b = pl / C;
  int c = (int)(pl % C);
  i64 o = ((b * ob + c) * H2 + y) * (i64)W2 + xv * 4;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = m.k[n * 4 + 0] * ta[i] + m.k[n * 4 + 1] * tb[i] + m.k[n * 4 + 2] * tc[i] + m.k[n * 4 + 3] * td[i];
    *reinterpret_cast<float4*>(out.p[n] + o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void k_dwt2x2_gen(const float* __restrict__ in, Ptr4 out, Mat4 m, i64 planes, int C, i64 ob, int H, int W) {
  const int W2 = (W + 1) >> 1, H2 = (H + 1) >> 1;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes * H2 * W2;
  if (idx >= total) return;
  int x = (int)(idx % W2);
  i64 t = idx / W2;
  int y = (int)(t % H2);
  i64 pl = t / H2;
  int y0 = 2 * y, y1 = 2 * y + 1, x0 = 2 * x, x1 = 2 * x + 1;
  if (y1 >= H) y1 = H - 2;  // reflect (no edge repeat): index H -> H-2
  if (x1 >= W) x1 = W - 2;
  const float* p = in + pl * (i64)H * W;
  float a = p[(i64)y0 * W + x0], b_ = p[(i64)y0 * W + x1], c_ = p[(i64)y1 * W + x0], d = p[(i64)y1 * W + x1];
  i64 b = pl / C;
  int c = (int)(pl % C);
  i64 o = ((b * ob + c) * H2 + y) * (i64)W2 + x;
#pragma unroll
  for (int n = 0; n < 4; ++n)
    out.p[n][o] = m.k[n * 4 + 0] * a + m.k[n * 4 + 1] * b_ + m.k[n * 4 + 2] * c_ + m.k[n * 4 + 3] * d;
}

static int run_dwt2x2(const float* in, Ptr4 out, const Mat4& m, int B, int C, i64 ob, int H, int W, cudaStream_t st) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return RF_OK;
  if (((H & 1) && H < 2) || ((W & 1) && W < 2)) return RF_ERR_BAD_SHAPE;
  i64 planes = (i64)B * C;
  bool vec = !(H & 1) && (W % 8 == 0) && ((uintptr_t)in % 16 == 0);
  for (int n = 0; n < 4; ++n) vec = vec && ((uintptr_t)out.p[n] % 16 == 0);
  ScopedLaunch sl(RF_K_INDEX_OP, 8.0 * planes * H * W);
  if (vec) {
    i64 total = planes * (H / 2) * (W / 8);
    k_dwt2x2_vec<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, m, planes, C, ob, H, W);
  } else {
    i64 total = planes * ((H + 1) / 2) * ((W + 1) / 2);
    k_dwt2x2_gen<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, m, planes, C, ob, H, W);
  }
  return check_cuda(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// 2x2 stride-2 synthesis: out[2y+i, 2x+j] = sum_n K[n][2i+j] * sub_n[y,x].  sub-band n of (b,c) is the plane at
// in.p[n] + (b*ib + c)*H*W;  out [B,C,2H,2W].
// ---------------------------------------------------------------------------------------------
struct CPtr4 { const float* p[4]; };

template <bool VEC>
__global__ void k_idwt2x2(CPtr4 in, float* __restrict__ out, Mat4 m, i64 planes, int C, i64 ib, int H, int W) {
  constexpr int PX = VEC ? 4 : 1;
  const int Wv = W / PX;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes * H * Wv;
  if (idx >= total) return;
  int xv = (int)(idx % Wv);
  i64 t = idx / Wv;
  int y = (int)(t % H);
  i64 pl = t / H;
  i64 b = pl / C;
  int c = (int)(pl % C);
  i64 si = ((b * ib + c) * H + y) * (i64)W + xv * PX;
  float s[4][PX];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    if (VEC) {
      float4 v = *reinterpret_cast<const float4*>(in.p[n] + si);
      s[n][0] = v.x; s[n][1 % PX] = v.y; s[n][2 % PX] = v.z; s[n][3 % PX] = v.w;
    } else {
      s[n][0] = in.p[n][si];
    }
  }
  float* o0 = out + (pl * 2 * H + 2 * y) * (i64)(2 * W) + xv * PX * 2;
  float* o1 = o0 + 2 * W;
  float r0[2 * PX], r1[2 * PX];
#pragma unroll
  for (int i = 0; i < PX; ++i) {
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      float v = m.k[0 * 4 + tp] * s[0][i] + m.k[1 * 4 + tp] * s[1][i] + m.k[2 * 4 + tp] * s[2][i] + m.k[3 * 4 + tp] * s[3][i];
      if (tp < 2) r0[2 * i + tp] = v; else r1[2 * i + tp - 2] = v;
    }
  }
  if (VEC) {
    *reinterpret_cast<float4*>(o0) = make_float4(r0[0], r0[1], r0[2 % (2 * PX)], r0[3 % (2 * PX)]);
    *reinterpret_cast<float4*>(o0 + 4) = make_float4(r0[4 % (2 * PX)], r0[5 % (2 * PX)], r0[6 % (2 * PX)], r0[7 % (2 * PX)]);
    *reinterpret_cast<float4*>(o1) = make_float4(r1[0], r1[1], r1[2 % (2 * PX)], r1[3 % (2 * PX)]);
    *reinterpret_cast<float4*>(o1 + 4) = make_float4(r1[4 % (2 * PX)], r1[5 % (2 * PX)], r1[6 % (2 * PX)], r1[7 % (2 * PX)]);
  } else {
    o0[0] = r0[0]; o0[1] = r0[1]; o1[0] = r1[0]; o1[1] = r1[1];
  }
}

static int run_idwt2x2(CPtr4 in, float* out, const Mat4& m, int B, int C, i64 ib, int H, int W, cudaStream_t st) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return RF_OK;
  i64 planes = (i64)B * C;
  bool vec = (W % 4 == 0) && ((uintptr_t)out % 16 == 0);
  for (int n = 0; n < 4; ++n) vec = vec && ((uintptr_t)in.p[n] % 16 == 0);
  ScopedLaunch sl(RF_K_INDEX_OP, 32.0 * planes * H * W);
  if (vec) {
    i64 total = planes * H * (W / 4);
    k_idwt2x2<true><<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, m, planes, C, ib, H, W);
  } else {
    i64 total = planes * H * W;
    k_idwt2x2<false><<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, m, planes, C, ib, H, W);
  }
  return check_cuda(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// pixel (un)shuffle, NCHW
// ---------------------------------------------------------------------------------------------
__global__ void k_unshuffle_gen(const float* __restrict__ in, float* __restrict__ out, i64 planes, int H, int W, int r) {
  const int h = H / r, w = W / r;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes * r * r * h * w;
  if (idx >= total) return;
  int x = (int)(idx % w);
  i64 t = idx / w;
  int y = (int)(t % h);
  t /= h;
  int ij = (int)(t % (r * r));
  i64 pl = t / (r * r);
  int i = ij / r, j = ij % r;
  out[idx] = in[(pl * H + (i64)y * r + i) * W + (i64)x * r + j];
}

// r = 2, W % 8 == 0: one thread reads 8 consecutive inputs of one input row and writes two float4 (j = 0, 1)
__global__ void k_unshuffle2_vec(const float* __restrict__ in, float* __restrict__ out, i64 planes, int H, int W) {
  const int h = H >> 1, w = W >> 1, W8 = W >> 3;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes * (i64)(2 * h) * W8;
  if (idx >= total) return;
  int xv = (int)(idx % W8);
  i64 t = idx / W8;
  int yi = (int)(t % (2 * h));
  i64 pl = t / (2 * h);
  const float* p = in + (pl * H + yi) * (i64)W + xv * 8;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  int y = yi >> 1, i = yi & 1;
  float* o = out + ((pl * 4 + 2 * i) * h + y) * (i64)w + xv * 4;
  *reinterpret_cast<float4*>(o) = make_float4(a.x, a.z, b.x, b.z);
  *reinterpret_cast<float4*>(o + (i64)h * w) = make_float4(a.y, a.w, b.y, b.w);
}

__global__ void k_shuffle_gen(const float* __restrict__ in, float* __restrict__ out, i64 planes_out, int H, int W, int r) {
  // in [planes_out*r*r, H, W] -> out [planes_out, H*r, W*r]
  const int Ho = H * r, Wo = W * r;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes_out * Ho * Wo;
  if (idx >= total) return;
  int xo = (int)(idx % Wo);
  i64 t = idx / Wo;
  int yo = (int)(t % Ho);
  i64 pl = t / Ho;
  int y = yo / r, i = yo % r, x = xo / r, j = xo % r;
  out[idx] = in[((pl * r * r + i * r + j) * H + y) * (i64)W + x];
}

__global__ void k_shuffle2_vec(const float* __restrict__ in, float* __restrict__ out, i64 planes_out, int H, int W) {
  const int W4 = W >> 2;
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = planes_out * (i64)(2 * H) * W4;
  if (idx >= total) return;
  int xv = (int)(idx % W4);
  i64 t = idx / W4;
  int yo = (int)(t % (2 * H));
  i64 pl = t / (2 * H);
  int y = yo >> 1, i = yo & 1;
  const float* p = in + ((pl * 4 + 2 * i) * H + y) * (i64)W + xv * 4;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + (i64)H * W);
  float* o = out + (pl * 2 * H + yo) * (i64)(2 * W) + xv * 8;
  *reinterpret_cast<float4*>(o) = make_float4(a.x, b.x, a.y, b.y);
  *reinterpret_cast<float4*>(o + 4) = make_float4(a.z, b.z, a.w, b.w);
}

// ---------------------------------------------------------------------------------------------
// layout conversion fp32 NCHW <-> T NHWC (32x32 smem tile transpose)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ in, T* __restrict__ out, int C, i64 HW) {
  __shared__ float tile[32][33];
  i64 b = blockIdx.z;
  i64 p0 = (i64)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    int c = c0 + k;
    i64 p = p0 + tx;
    tile[k][tx] = (c < C && p < HW) ? in[(b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    i64 p = p0 + k;
    int c = c0 + tx;
    if (c < C && p < HW) from_f(out[(b * HW + p) * C + c], tile[tx][k]);
  }
}

template <typename T>
__global__ void k_nhwc_to_nchw(const T* __restrict__ in, float* __restrict__ out, int C, i64 HW) {
  __shared__ float tile[32][33];
  i64 b = blockIdx.z;
  i64 p0 = (i64)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;
  for (int k = ty; k < 32; k += 8) {
    i64 p = p0 + k;
    int c = c0 + tx;
    tile[k][tx] = (c < C && p < HW) ? to_f(in[(b * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int c = c0 + k;
    i64 p = p0 + tx;
    if (c < C && p < HW) out[(b * C + c) * HW + p] = tile[tx][k];
  }
}

void launch_nchw_to_nhwc(Ctx& ctx, const float* in, void* out, int B, int C, i64 HW) {
  if (ctx.dry || B <= 0 || HW <= 0) return;
  ScopedLaunch sl(RF_K_LAYOUT, (4.0 + esize(ctx.dtype)) * B * C * HW);
  dim3 grid((unsigned)cdivl(HW, 32), cdiv(C, 32), B), block(32, 8);
  if (ctx.dtype == RF_BF16)
    k_nchw_to_nhwc<bf16><<<grid, block, 0, ctx.stream>>>(in, (bf16*)out, C, HW);
  else
    k_nchw_to_nhwc<float><<<grid, block, 0, ctx.stream>>>(in, (float*)out, C, HW);
}

void launch_nhwc_to_nchw(Ctx& ctx, const void* in, float* out, int B, int C, i64 HW) {
  if (ctx.dry || B <= 0 || HW <= 0) return;
  ScopedLaunch sl(RF_K_LAYOUT, (4.0 + esize(ctx.dtype)) * B * C * HW);
  dim3 grid((unsigned)cdivl(HW, 32), cdiv(C, 32), B), block(32, 8);
  if (ctx.dtype == RF_BF16)
    k_nhwc_to_nchw<bf16><<<grid, block, 0, ctx.stream>>>((const bf16*)in, out, C, HW);
  else
    k_nhwc_to_nchw<float><<<grid, block, 0, ctx.stream>>>((const float*)in, out, C, HW);
}

__global__ void k_fill_f32(float* p, float v, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_f32(Ctx& ctx, float* p, float v, i64 n) {
  if (ctx.dry || n <= 0) return;
  ScopedLaunch sl(RF_K_MISC, 4.0 * n);
  k_fill_f32<<<(unsigned)cdivl(n, 256), 256, 0, ctx.stream>>>(p, v, n);
}

template <typename T>
__global__ void k_pack3(const float* __restrict__ src, T* __restrict__ dst, int A, int Bn, int Cn, i64 sa, i64 sb, i64 sc,
                        i64 da, i64 db, i64 dc, i64 doff) {
  i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 total = (i64)A * Bn * Cn;
  if (idx >= total) return;
  int c = (int)(idx % Cn);
  i64 t = idx / Cn;
  int b = (int)(t % Bn);
  int a = (int)(t / Bn);
  from_f(dst[doff + a * da + b * db + c * dc], src[a * sa + b * sb + c * sc]);
}
void launch_pack3(Ctx& ctx, const float* src, void* dst, int dst_dtype, int A, int Bn, int Cn, i64 sa, i64 sb, i64 sc,
                  i64 da, i64 db, i64 dc, i64 doff) {
  if (ctx.dry) return;
  i64 total = (i64)A * Bn * Cn;
  if (total <= 0) return;
  ScopedLaunch sl(RF_K_WEIGHT_PACK, 8.0 * total);
  unsigned g = (unsigned)cdivl(total, 256);
  if (dst_dtype == RF_BF16)
    k_pack3<bf16><<<g, 256, 0, ctx.stream>>>(src, (bf16*)dst, A, Bn, Cn, sa, sb, sc, da, db, dc, doff);
  else
    k_pack3<float><<<g, 256, 0, ctx.stream>>>(src, (float*)dst, A, Bn, Cn, sa, sb, sc, da, db, dc, doff);
}

// ---------------------------------------------------------------------------------------------
// BayerLumaChroma on planar NCHW (the ABI-level operator; the model path uses the fused pack kernel)
// ---------------------------------------------------------------------------------------------
__global__ void k_luma_planar_pass1(const float* __restrict__ x, float* __restrict__ y, float* ymax, float rw, float gw,
                                    float bw, i64 hw) {
  i64 b = blockIdx.y;
  const float* px = x + b * 4 * hw;
  float m = -INFINITY;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    float r = px[i], g = __fmul_rn(0.5f, __fadd_rn(px[hw + i], px[2 * hw + i])), bl = px[3 * hw + i];
    float v = __fadd_rn(__fadd_rn(__fmul_rn(rw, r), __fmul_rn(gw, g)), __fmul_rn(bw, bl));
    y[b * hw + i] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_float(ymax + b, m);
}
__global__ void k_luma_planar_pass2(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ cr,
                                    float* __restrict__ cb, const float* ymax, float eps, i64 hw) {
  i64 b = blockIdx.y;
  const float* px = x + b * 4 * hw;
  float d = fmaxf(ymax[b], eps);
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    float v = __fdiv_rn(y[b * hw + i], d);
    y[b * hw + i] = v;
    cr[b * hw + i] = __fsub_rn(px[i], v);
    cb[b * hw + i] = __fsub_rn(px[3 * hw + i], v);
  }
}

// per-pixel LayerNorm over the channel axis of an NCHW tensor (ABI-level operator)
__global__ void k_layernorm_nchw(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                                 float* __restrict__ out, float eps, int mode, int C, i64 hw) {
  i64 b = blockIdx.y;
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const float* p = x + b * C * hw + i;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += p[c * hw];
  float mu = s / C;
  float v = 0.f;
  for (int c = 0; c < C; ++c) {
    float d = p[c * hw] - mu;
    v += d * d;
  }
  float rstd = 1.0f / sqrtf(v / C + eps);
  float* o = out + b * C * hw + i;
  for (int c = 0; c < C; ++c) {
    float xv = p[c * hw];
    o[c * hw] = mode == 0 ? (xv - mu) * rstd * g[c] + (bta ? bta[c] : 0.f) : xv * rstd * g[c];
  }
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_downshuffle(const float* in, float* out, int B, int C, int H, int W, int r, void* stream) {
  if (!in || !out) return RF_ERR_BAD_ARG;
  if (r < 1 || B < 0 || C < 0 || H < 0 || W < 0) return RF_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  i64 planes = (i64)B * C;
  int h = H / r, w = W / r;
  if (planes == 0 || h == 0 || w == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP, 8.0 * planes * H * W);
  if (r == 2 && H % 2 == 0 && W % 8 == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) {
    i64 total = planes * H * (W / 8);
    k_unshuffle2_vec<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, planes, H, W);
  } else {
    i64 total = planes * r * r * h * w;
    k_unshuffle_gen<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, planes, H, W, r);
  }
  return check_cuda(cudaGetLastError());
}

int rf_pixelshuffle(const float* in, float* out, int B, int C_out, int H, int W, int r, void* stream) {
  if (!in || !out) return RF_ERR_BAD_ARG;
  if (r < 1 || B < 0 || C_out < 0 || H < 0 || W < 0) return RF_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  i64 planes = (i64)B * C_out;
  if (planes == 0 || H == 0 || W == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP, 8.0 * planes * r * r * H * W);
  if (r == 2 && W % 4 == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) {
    i64 total = planes * 2 * H * (W / 4);
    k_shuffle2_vec<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, planes, H, W);
  } else {
    i64 total = planes * (i64)H * r * W * r;
    k_shuffle_gen<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(in, out, planes, H, W, r);
  }
  return check_cuda(cudaGetLastError());
}

int rf_custom_dwt(const float* in, float* out, const float* k16_host, int B, int C, int H, int W, void* stream) {
  if (!in || !out || !k16_host) return RF_ERR_BAD_ARG;
  if ((H & 1) || (W & 1) || B < 0 || C < 0) return RF_ERR_BAD_SHAPE;
  Mat4 m;
  for (int i = 0; i < 16; ++i) m.k[i] = k16_host[i];
  i64 plane = (i64)(H / 2) * (W / 2);
  Ptr4 o;
  for (int n = 0; n < 4; ++n) o.p[n] = out + (i64)n * C * plane;  // channel = n*C + c
  return run_dwt2x2(in, o, m, B, C, 4 * (i64)C, H, W, (cudaStream_t)stream);
}

int rf_custom_idwt(const float* in, float* out, const float* k16_host, int B, int C, int H, int W, void* stream) {
  if (!in || !out || !k16_host) return RF_ERR_BAD_ARG;
  if (B < 0 || C < 0 || H < 0 || W < 0) return RF_ERR_BAD_SHAPE;
  Mat4 m;
  for (int i = 0; i < 16; ++i) m.k[i] = k16_host[i];
  i64 plane = (i64)H * W;
  CPtr4 s;
  for (int n = 0; n < 4; ++n) s.p[n] = in + (i64)n * C * plane;
  return run_idwt2x2(s, out, m, B, C, 4 * (i64)C, H, W, (cudaStream_t)stream);
}

int rf_haar_dwt(const float* in, const float* filt, float* LL, float* LH, float* HL, float* HH, int B, int C, int H, int W,
                void* stream) {
  if (!in || !filt || !LL || !LH || !HL || !HH) return RF_ERR_BAD_ARG;
  if (B < 0 || C < 0 || H < 0 || W < 0) return RF_ERR_BAD_SHAPE;
  Mat4 m;
  cudaStream_t st = (cudaStream_t)stream;
  // 16 floats of the registered buffer; a tiny synchronous read keeps the kernels argument-only
  RF_CUDA(cudaMemcpyAsync(m.k, filt, sizeof(m.k), cudaMemcpyDeviceToHost, st));
  RF_CUDA(cudaStreamSynchronize(st));
  Ptr4 o = {{LL, LH, HL, HH}};
  return run_dwt2x2(in, o, m, B, C, (i64)C, H, W, st);
}

int rf_dwt_init(const float* in, float* out, int B, int C, int H, int W, void* stream) {
  if (!in || !out) return RF_ERR_BAD_ARG;
  if ((H & 1) || (W & 1) || B < 0 || C < 0) return RF_ERR_BAD_SHAPE;
  // x1=a/2, x2=c/2, x3=b/2, x4=d/2; LL=+x1+x2+x3+x4, HL=-x1-x2+x3+x4, LH=-x1+x2-x3+x4, HH=+x1-x2-x3+x4
  Mat4 m = {{0.5f, 0.5f, 0.5f, 0.5f, -0.5f, 0.5f, -0.5f, 0.5f, -0.5f, -0.5f, 0.5f, 0.5f, 0.5f, -0.5f, -0.5f, 0.5f}};
  i64 plane = (i64)(H / 2) * (W / 2);
  Ptr4 o;
  for (int n = 0; n < 4; ++n) o.p[n] = out + (i64)n * B * C * plane;  // batch = n*B + b
  return run_dwt2x2(in, o, m, B, C, (i64)C, H, W, (cudaStream_t)stream);
}

int rf_iwt_init(const float* in, float* out, int B, int C, int H, int W, void* stream) {
  if (!in || !out) return RF_ERR_BAD_ARG;
  if (B < 0 || C < 0 || H < 0 || W < 0) return RF_ERR_BAD_SHAPE;
  // rows = sub-band (x1..x4), columns = tap (TL,TR,BL,BR)
  Mat4 m = {{0.5f, 0.5f, 0.5f, 0.5f, -0.5f, 0.5f, -0.5f, 0.5f, -0.5f, -0.5f, 0.5f, 0.5f, 0.5f, -0.5f, -0.5f, 0.5f}};
  i64 plane = (i64)H * W;
  CPtr4 s;
  for (int n = 0; n < 4; ++n) s.p[n] = in + (i64)n * B * C * plane;
  return run_idwt2x2(s, out, m, B, C, (i64)C, H, W, (cudaStream_t)stream);
}

int rf_luma_chroma(const float* x_ds, float* y, float* cr, float* cb, const float* rgb_w_host, float eps, int B, int h,
                   int w, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x_ds || !y || !cr || !cb || !rgb_w_host || !workspace) return RF_ERR_BAD_ARG;
  if (workspace_bytes < sizeof(float) * (size_t)(B > 0 ? B : 1)) return RF_ERR_WORKSPACE;
  if (B <= 0 || h <= 0 || w <= 0) return B < 0 || h < 0 || w < 0 ? RF_ERR_BAD_SHAPE : RF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* ymax = (float*)workspace;
  i64 hw = (i64)h * w;
  {
    ScopedLaunch sl(RF_K_MISC);
    k_fill_f32<<<cdiv(B, 256), 256, 0, st>>>(ymax, -INFINITY, B);
  }
  unsigned gx = (unsigned)(cdivl(hw, 256) < 1184 ? cdivl(hw, 256) : 1184);
  {
    ScopedLaunch sl(RF_K_LUMA_NORM, 20.0 * B * hw);
    k_luma_planar_pass1<<<dim3(gx, B), 256, 0, st>>>(x_ds, y, ymax, rgb_w_host[0], rgb_w_host[1], rgb_w_host[2], hw);
  }
  {
    ScopedLaunch sl(RF_K_LUMA_NORM, 24.0 * B * hw);
    k_luma_planar_pass2<<<dim3(gx, B), 256, 0, st>>>(x_ds, y, cr, cb, ymax, eps, hw);
  }
  return check_cuda(cudaGetLastError());
}

int rf_layernorm(const float* in, const float* weight, const float* bias, float* out, float eps, int mode, int B, int C,
                 int H, int W, void* stream) {
  if (!in || !weight || !out) return RF_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return RF_ERR_BAD_ARG;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return (B < 0 || C <= 0 || H < 0 || W < 0) ? RF_ERR_BAD_SHAPE : RF_OK;
  i64 hw = (i64)H * W;
  ScopedLaunch sl(RF_K_LAYERNORM, 8.0 * B * C * hw);
  k_layernorm_nchw<<<dim3((unsigned)cdivl(hw, 128), B), 128, 0, (cudaStream_t)stream>>>(in, weight, bias, out, eps, mode, C,
                                                                                      hw);
  return check_cuda(cudaGetLastError());
}

}  // extern "C"
