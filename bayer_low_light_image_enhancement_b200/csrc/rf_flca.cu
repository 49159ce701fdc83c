// FLCA branch (FLCA_RF.py:136-162) and FLCA_Pyramid pieces (ML_RF.py:132-183) on NHWC activations.
//
// flca_mod: one pass over the block input.  Each thread owns 8 channels of PIX horizontally adjacent pixels; the
// 3x3 neighbourhood of the four stage-resolution guidance maps (one float4 per pixel) is gathered through L1, the
// 36 x C tap weights sit in shared memory.  The three 1->C / 2->C convolutions, sigmoid/tanh, the modulation and
// the squeeze-excite channel sums are fused; the SE scale itself is folded into channel_reduce's weights
// (fold_reduce) so the FLCA output is never re-read for a per-channel multiply.
#include "rf_kernels.cuh"

namespace rf {

constexpr int FLCA_PIX = 2;

int flca_num_partials(int C, int B, i64 P) {
  int cv = C / 8;
  int ppb = 256 / cv;
  if (ppb < 1) ppb = 1;
  i64 per_iter = (i64)ppb * FLCA_PIX;
  i64 want = cdivl(P, per_iter * 4);  // >= 4 iterations per block
  i64 cap = (i64)num_sms() * 8 / (B > 0 ? B : 1);
  if (cap < 1) cap = 1;
  i64 n = want < cap ? want : cap;
  return (int)(n < 1 ? 1 : n);
}

template <typename T>
__device__ __forceinline__ float act_sig(float x) { return sigmoid_f(x); }

// G: [B,Hf,Wf,NG] guidance.  MODE 0: FLCA (maps 0..3 of NG=4): xmod = feat*(1 + a*sig(low) + b*tanh(high) + g*sig(chr)).
// MODE 1: pyramid level (maps 2l, 2l+1 of NG=8): xs = x*(ga*sig(low_l) + gb*tanh(high_l)).
// MODE 2: pyramid chroma (maps 4,5 of NG=8): xs = x*(gc*sig(chr)).
// w: [9][NW][C] tap weights, NW = 4 (FLCA) or 6 (ML).  coef: abg[3] (MODE 0) or gates[b][6] (MODE 1/2).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_flca_mod(const T* __restrict__ feat, const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ coef,
           T* __restrict__ xmod, float* __restrict__ partial, int Hf, int Wf, int C, int level) {
  constexpr int NG = MODE == 0 ? 4 : 8;
  constexpr int NW = MODE == 0 ? 4 : 6;
  constexpr int NA = MODE == 0 ? 3 : (MODE == 1 ? 2 : 1);  // pre-activations per channel
  constexpr int PIX = FLCA_PIX;
  extern __shared__ float smem[];
  float* sw = smem;  // [9][NW][C]
  const int cv = blockDim.x, ppb = blockDim.y;
  const int tid = threadIdx.y * cv + threadIdx.x, nthr = cv * ppb;
  for (int i = tid; i < 9 * NW * C; i += nthr) sw[i] = w[i];
  __syncthreads();
  const i64 b = blockIdx.y;
  const int c0 = threadIdx.x * 8;
  float k0, k1, k2;
  if constexpr (MODE == 0) { k0 = coef[0]; k1 = coef[1]; k2 = coef[2]; }
  else if constexpr (MODE == 1) { k0 = coef[b * 6 + 2 * level]; k1 = coef[b * 6 + 2 * level + 1]; k2 = 0.f; }
  else { k0 = coef[b * 6 + 4]; k1 = 0.f; k2 = 0.f; }
  // map indices inside a guidance pixel and inside the weight row
  const int gA = MODE == 0 ? 0 : (MODE == 1 ? 2 * level : 4);
  const int gB = gA + 1;
  const int Wp = (Wf + PIX - 1) / PIX;           // pixel groups per row
  const i64 groups = (i64)Hf * Wp;
  float csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float* Gb = G + b * (i64)Hf * Wf * NG;
  for (i64 g = (i64)blockIdx.x * ppb + threadIdx.y; g < groups; g += (i64)gridDim.x * ppb) {
    const int y = (int)(g / Wp), x0 = (int)(g % Wp) * PIX;
    float acc[PIX][NA][8];
#pragma unroll
    for (int p = 0; p < PIX; ++p)
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][a][j] = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= Hf) continue;
#pragma unroll
      for (int dxx = -1; dxx <= PIX; ++dxx) {  // guidance column x0 + dxx feeds pixels p with tap dx = dxx - p
        const int xx = x0 + dxx;
        if (xx < 0 || xx >= Wf) continue;
        const float* gp = Gb + ((i64)yy * Wf + xx) * NG;
        float ga, gb, gc = 0.f, gd = 0.f;
        if constexpr (MODE == 0) {
          float4 q = *reinterpret_cast<const float4*>(gp);
          ga = q.x; gb = q.y; gc = q.z; gd = q.w;
        } else {
          float2 q = *reinterpret_cast<const float2*>(gp + gA);
          ga = q.x; gb = q.y;
        }
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
          const int dx = dxx - p;
          if (dx < -1 || dx > 1) continue;
          const int t = (dy + 1) * 3 + (dx + 1);
          const float* wr = sw + (t * NW) * C + c0;
          if constexpr (MODE == 0) {
            float wl[8], wh[8], w2[8], w3[8];
            load8(wr, wl); load8(wr + C, wh); load8(wr + 2 * C, w2); load8(wr + 3 * C, w3);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[p][0][j] = fmaf(wl[j], ga, acc[p][0][j]);
              acc[p][1][j] = fmaf(wh[j], gb, acc[p][1][j]);
              acc[p][2][j] = fmaf(w2[j], gc, fmaf(w3[j], gd, acc[p][2][j]));
            }
          } else if constexpr (MODE == 1) {
            float wl[8], wh[8];
            load8(wr + gA * C, wl); load8(wr + gB * C, wh);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[p][0][j] = fmaf(wl[j], ga, acc[p][0][j]);
              acc[p][1 % NA][j] = fmaf(wh[j], gb, acc[p][1 % NA][j]);
            }
          } else {
            float w2[8], w3[8];
            load8(wr + 4 * C, w2); load8(wr + 5 * C, w3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][0][j] = fmaf(w2[j], ga, fmaf(w3[j], gb, acc[p][0][j]));
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
      const int x = x0 + p;
      if (x >= Wf) continue;
      const i64 off = ((b * Hf + y) * (i64)Wf + x) * C + c0;
      float f[8], o[8];
      load8(feat + off, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sp;
        if constexpr (MODE == 0)
          sp = 1.f + k0 * sigmoid_f(acc[p][0][j]) + k1 * tanhf(acc[p][1][j]) + k2 * sigmoid_f(acc[p][2][j]);
        else if constexpr (MODE == 1)
          sp = k0 * sigmoid_f(acc[p][0][j]) + k1 * tanhf(acc[p][1 % NA][j]);
        else
          sp = k0 * sigmoid_f(acc[p][0][j]);
        o[j] = f[j] * sp;
        csum[j] += o[j];
      }
      store8(xmod + off, o);
    }
  }
  if (partial != nullptr) {
    // reduce csum over the ppb pixel lanes of the block (reuse the weight smem after a barrier)
    __syncthreads();
    float* red = smem;  // [ppb][C]
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y * C + c0 + j] = csum[j];
    __syncthreads();
    for (int c = tid; c < C; c += nthr) {
      float s = 0.f;
      for (int r = 0; r < ppb; ++r) s += red[r * C + c];
      partial[(b * gridDim.x + blockIdx.x) * C + c] = s;
    }
  }
}

template <typename T, int MODE>
static void run_flca_mod(Ctx& ctx, const void* feat, const float* G, const float* w, const float* coef, void* xmod,
                         float* partial, int nblk, int B, int Hf, int Wf, int C, int level, int kid) {
  const int cv = C / 8;
  int ppb = 256 / cv;
  if (ppb < 1) ppb = 1;
  constexpr int NW = MODE == 0 ? 4 : 6;
  size_t smem = sizeof(float) * (size_t)C * (9 * NW > ppb ? 9 * NW : ppb);
  auto kern = k_flca_mod<T, MODE>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double px = (double)B * Hf * Wf;
  ScopedLaunch sl(kid, px * C * 2.0 * sizeof(T) + px * 16.0, px * C * 2.0 * (MODE == 0 ? 36 : 18));
  kern<<<dim3(nblk, B), dim3(cv, ppb), smem, ctx.stream>>>((const T*)feat, G, w, coef, (T*)xmod, partial, Hf, Wf, C, level);
}

void launch_flca_mod(Ctx& ctx, const void* feat, const float* G, const float* w36, const float* abg, void* xmod,
                     float* partial, int nblk, int B, int Hf, int Wf, int C) {
  if (ctx.dry) return;
  if (ctx.dtype == RF_BF16)
    run_flca_mod<bf16, 0>(ctx, feat, G, w36, abg, xmod, partial, nblk, B, Hf, Wf, C, 0, RF_K_FLCA_MOD);
  else
    run_flca_mod<float, 0>(ctx, feat, G, w36, abg, xmod, partial, nblk, B, Hf, Wf, C, 0, RF_K_FLCA_MOD);
}

void launch_pyr_spatial(Ctx& ctx, const void* x, const float* G8, const float* w54, const float* gates, void* xs, int mode,
                        int level, int B, int Hf, int Wf, int C) {
  if (ctx.dry) return;
  int nblk = flca_num_partials(C, B, (i64)Hf * Wf);
  if (ctx.dtype == RF_BF16) {
    if (mode == 0) run_flca_mod<bf16, 1>(ctx, x, G8, w54, gates, xs, nullptr, nblk, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
    else run_flca_mod<bf16, 2>(ctx, x, G8, w54, gates, xs, nullptr, nblk, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
  } else {
    if (mode == 0) run_flca_mod<float, 1>(ctx, x, G8, w54, gates, xs, nullptr, nblk, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
    else run_flca_mod<float, 2>(ctx, x, G8, w54, gates, xs, nullptr, nblk, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
  }
}

// gates (ML_RF.py:151-156,173): per level sigmoid(W [lowmean, highmean] + b); gamma = sigmoid(w*mean(chr_mag)+b)
__global__ void k_pyr_gates(const float* __restrict__ sums, float invP, const float* __restrict__ gw,
                            const float* __restrict__ gb, const float* __restrict__ cg, float* __restrict__ gates, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* s = sums + b * 8;
  for (int l = 0; l < 2; ++l) {
    float lo = s[2 * l] * invP, hi = s[2 * l + 1] * invP;
    for (int o = 0; o < 2; ++o)
      gates[b * 6 + 2 * l + o] = sigmoid_f(gw[l * 4 + o * 2 + 0] * lo + gw[l * 4 + o * 2 + 1] * hi + gb[l * 2 + o]);
  }
  gates[b * 6 + 4] = sigmoid_f(cg[0] * (s[6] * invP) + cg[1]);
  gates[b * 6 + 5] = 0.f;
}
void launch_pyr_gates(Ctx& ctx, const float* sums, i64 P, const float* gate_w, const float* gate_b, const float* cgate,
                      float* gates, int B) {
  if (ctx.dry) return;
  ScopedLaunch sl(RF_K_MISC);
  k_pyr_gates<<<cdiv(B, 32), 32, 0, ctx.stream>>>(sums, 1.0f / (float)P, gate_w, gate_b, cgate, gates, B);
}

// per-channel sums of an NHWC tensor -> partial [B][nblk][C]   (blockDim = (C/8, ppb))
template <typename T>
__global__ void k_channel_sums(const T* __restrict__ x, float* __restrict__ partial, i64 P, int C) {
  extern __shared__ float smem[];
  const int cv = blockDim.x, ppb = blockDim.y;
  const int tid = threadIdx.y * cv + threadIdx.x, nthr = cv * ppb;
  const i64 b = blockIdx.y;
  const int c0 = threadIdx.x * 8;
  float csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (i64 p = (i64)blockIdx.x * ppb + threadIdx.y; p < P; p += (i64)gridDim.x * ppb) {
    float f[8];
    load8(x + (b * P + p) * C + c0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) smem[threadIdx.y * C + c0 + j] = csum[j];
  __syncthreads();
  for (int c = tid; c < C; c += nthr) {
    float s = 0.f;
    for (int r = 0; r < ppb; ++r) s += smem[r * C + c];
    partial[(b * gridDim.x + blockIdx.x) * C + c] = s;
  }
}
void launch_channel_sums(Ctx& ctx, const void* x, float* partial, int nblk, int B, i64 P, int C) {
  if (ctx.dry) return;
  const int cv = C / 8;
  int ppb = 256 / cv;
  if (ppb < 1) ppb = 1;
  size_t smem = sizeof(float) * (size_t)C * ppb;
  ScopedLaunch sl(RF_K_CHANNEL_SUMS, (double)B * P * C * esize(ctx.dtype));
  if (ctx.dtype == RF_BF16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_channel_sums<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_channel_sums<bf16><<<dim3(nblk, B), dim3(cv, ppb), smem, ctx.stream>>>((const bf16*)x, partial, P, C);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_channel_sums<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_channel_sums<float><<<dim3(nblk, B), dim3(cv, ppb), smem, ctx.stream>>>((const float*)x, partial, P, C);
  }
}

// squeeze-excite MLP (FLCA_RF.py:124-130,160): one block per image
__global__ void __launch_bounds__(256)
k_se_finalize(const float* __restrict__ partial, int nblk, float invP, const float* __restrict__ w1,
              const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ scale, int C, int hid) {
  extern __shared__ float smem[];
  float* mean = smem;       // [C]
  float* hbuf = smem + C;   // [hid]
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nblk; ++k) s += partial[((i64)b * nblk + k) * C + c];
    mean[c] = s * invP;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < hid; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w1[(i64)j * C + c] * mean[c];
    s = warp_sum(s);
    if (lane == 0) hbuf[j] = fmaxf(s + b1[j], 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int j = 0; j < hid; ++j) s += w2[(i64)c * hid + j] * hbuf[j];
    scale[(i64)b * C + c] = sigmoid_f(s);
  }
}
void launch_se_finalize(Ctx& ctx, const float* partial, int nblk, i64 P, const float* w1, const float* b1,
                        const float* w2, const float* b2, float* scale, int B, int C, int hid) {
  if (ctx.dry) return;
  ScopedLaunch sl(RF_K_SE_FINALIZE, 4.0 * B * nblk * C);
  k_se_finalize<<<B, 256, sizeof(float) * (C + hid), ctx.stream>>>(partial, nblk, 1.0f / (float)P, w1, b1, w2, b2, scale, C,
                                                                  hid);
}

template <typename T>
__global__ void k_fold_reduce(const float* __restrict__ red_w, const float* __restrict__ scale, T* __restrict__ wred, int C) {
  const i64 b = blockIdx.y;
  const i64 n2 = (i64)C * 2 * C;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (i64)gridDim.x * blockDim.x) {
    int k = (int)(i % (2 * C));
    float v = red_w[i];
    if (k < C) v *= scale[b * C + k];
    from_f(wred[b * n2 + i], v);
  }
}
void launch_fold_reduce(Ctx& ctx, const float* red_w, const float* scale, void* wred, int B, int C) {
  if (ctx.dry) return;
  i64 n2 = (i64)C * 2 * C;
  unsigned gx = (unsigned)(cdivl(n2, 256) < 296 ? cdivl(n2, 256) : 296);
  ScopedLaunch sl(RF_K_FOLD_REDUCE, (4.0 + esize(ctx.dtype)) * B * n2);
  if (ctx.dtype == RF_BF16)
    k_fold_reduce<bf16><<<dim3(gx, B), 256, 0, ctx.stream>>>(red_w, scale, (bf16*)wred, C);
  else
    k_fold_reduce<float><<<dim3(gx, B), 256, 0, ctx.stream>>>(red_w, scale, (float*)wred, C);
}

template <typename T>
__global__ void k_scale_channels(const T* __restrict__ x, const float* __restrict__ scale, T* __restrict__ out, i64 P, int C) {
  const i64 b = blockIdx.y;
  const int cv = C / 8;
  const i64 total = P * cv;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    int c0 = (int)(i % cv) * 8;
    i64 off = (b * P + i / cv) * C + c0;
    float f[8];
    load8(x + off, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= scale[b * C + c0 + j];
    store8(out + off, f);
  }
}
void launch_scale_channels(Ctx& ctx, const void* x, const float* scale, void* out, int B, i64 P, int C) {
  if (ctx.dry) return;
  i64 total = P * (C / 8);
  unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_MISC, 2.0 * B * P * C * esize(ctx.dtype));
  if (ctx.dtype == RF_BF16)
    k_scale_channels<bf16><<<dim3(gx, B), 256, 0, ctx.stream>>>((const bf16*)x, scale, (bf16*)out, P, C);
  else
    k_scale_channels<float><<<dim3(gx, B), 256, 0, ctx.stream>>>((const float*)x, scale, (float*)out, P, C);
}

}  // namespace rf
