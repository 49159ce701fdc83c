// FLCA branch (FLCA_RF.py:136-162) and FLCA_Pyramid pieces (ML_RF.py:132-183) on NHWC activations.
//
// flca_mod: one pass over the block input.  Each thread owns 8 channels of PIX horizontally adjacent pixels; the
// 3x3 neighbourhood of the four stage-resolution guidance maps (one float4 per pixel) is gathered through L1, the
// 36 x C tap weights sit in shared memory.  The three 1->C / 2->C convolutions, sigmoid/tanh, the modulation and
// the squeeze-excite channel sums are fused; the SE scale itself is folded into channel_reduce's weights
// (k_se_fold) so the FLCA output is never re-read for a per-channel multiply.
#include "rf_kernels.cuh"

namespace rf {

constexpr int FLCA_SLOTS = 320;  // partial-sum slots per image (== IT_SLOTS of the tensor-core kernel, which owns one slot per
                                 // CTA; this CUDA-core kernel accumulates atomically), summed in order by k_se_fold / se_finalize
constexpr int FLCA_LS = 32;      // pixels per warp strip

int flca_num_partials(int C, int B, i64 P) {
  (void)C; (void)B; (void)P;
  return FLCA_SLOTS;
}

// One warp = 32 channels (one per lane) x a strip of FLCA_LS pixels of one row.  Each lane keeps the 9 x NM tap weights
// of ITS channel in registers; the guidance column (3 rows x NM maps) is a warp-uniform load that feeds the three
// outputs it touches (sliding window), so the inner loop is register-only FFMA + 3 MUFU per output.
// G: [B,Hf,Wf,NG] guidance.  MODE 0: FLCA (maps 0..3 of NG=4): xmod = feat*(1 + a*sig(low) + b*tanh(high) + g*sig(chr)).
// MODE 1: pyramid level (maps 2l, 2l+1 of NG=8): xs = x*(ga*sig(low_l) + gb*tanh(high_l)).
// MODE 2: pyramid chroma (maps 4,5 of NG=8): xs = x*(gc*sig(chr)).
// w: [9][NW][C] tap weights, NW = 4 (FLCA) or 6 (ML).  coef: abg[3] (MODE 0) or gates[b][6] (MODE 1/2).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_flca_mod(const T* __restrict__ feat, const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ coef,
           T* __restrict__ xmod, float* __restrict__ partial, int Hf, int Wf, int C, int level, i64 tasks) {
  constexpr int NG = MODE == 0 ? 4 : 8;
  constexpr int NW = MODE == 0 ? 4 : 6;
  constexpr int NM = MODE == 0 ? 4 : 2;                      // maps used
  constexpr int NA = MODE == 0 ? 3 : (MODE == 1 ? 2 : 1);    // pre-activations
  constexpr int LS = FLCA_LS;
  const int lane = threadIdx.x & 31;
  const i64 b = blockIdx.y;
  const i64 task = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= tasks) return;
  const int CG = (C + 31) >> 5, SW = (Wf + LS - 1) / LS;
  const int cg = (int)(task % CG);
  const i64 t2 = task / CG;
  const int x0 = (int)(t2 % SW) * LS, y = (int)(t2 / SW);
  const int c = cg * 32 + lane;
  const bool c_ok = c < C;
  const int cc = c_ok ? c : C - 1;
  const int m0 = MODE == 0 ? 0 : (MODE == 1 ? 2 * level : 4);   // first map used (guidance pixel and weight row)
  float wv[9][NM];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int g = 0; g < NM; ++g) wv[t][g] = w[(t * NW + m0 + g) * C + cc];
  float k0, k1 = 0.f, k2 = 0.f;
  if constexpr (MODE == 0) { k0 = coef[0]; k1 = coef[1]; k2 = coef[2]; }
  else if constexpr (MODE == 1) { k0 = coef[b * 6 + 2 * level]; k1 = coef[b * 6 + 2 * level + 1]; }
  else { k0 = coef[b * 6 + 4]; }
  const float* Gb = G + b * (i64)Hf * Wf * NG + m0;
  const T* fb = feat + b * (i64)Hf * Wf * C;
  T* ob = xmod + b * (i64)Hf * Wf * C;
  // stage the warp's guidance window (3 rows x LS+2 columns x NM maps) in shared memory with coalesced loads, and
  // prefetch the strip's feature values, so that every global request of the warp is in flight before the FMA loop
  __shared__ float sG[8][3][LS + 2][NM];
  const int wslot = threadIdx.x >> 5;
  for (int idx = lane; idx < 3 * (LS + 2); idx += 32) {
    const int r = idx / (LS + 2), j = idx - r * (LS + 2);
    const int yy = y + r - 1, xc = x0 - 1 + j;
    float gv[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) gv[m] = 0.f;
    if (yy >= 0 && yy < Hf && xc >= 0 && xc < Wf) {
      const float* gp = Gb + ((i64)yy * Wf + xc) * NG;
      if constexpr (NM == 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(gp);
        gv[0] = q4.x; gv[1] = q4.y; gv[2] = q4.z; gv[3] = q4.w;
      } else {
        const float2 q2 = *reinterpret_cast<const float2*>(gp);
        gv[0] = q2.x; gv[1] = q2.y;
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) sG[wslot][r][j][m] = gv[m];
  }
  T fv[LS];
#pragma unroll
  for (int q = 0; q < LS; ++q) {
    if (c_ok && x0 + q < Wf) fv[q] = fb[((i64)y * Wf + x0 + q) * C + c];
    else from_f(fv[q], 0.f);
  }
  __syncwarp();
  float a[3][NA];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int q = 0; q < NA; ++q) a[s][q] = 0.f;
  float csum = 0.f;
  const int ncol = min(LS, Wf - x0) + 2;
#pragma unroll
  for (int j0 = 0; j0 < LS + 2; j0 += 3) {
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {       // the accumulator rotation is static
      const int j = j0 + jj;
      if (j >= LS + 2) break;
      if (j >= ncol) break;
      float g[3][NM];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int m = 0; m < NM; ++m) g[r][m] = sG[wslot][r][j][m];
      float* aN = a[(jj + 2) % 3];  // output q = j      (kx = 0)
      float* aC = a[(jj + 1) % 3];  // output q = j - 1  (kx = 1)
      float* aD = a[jj % 3];        // output q = j - 2  (kx = 2), complete after this column
      // pre-activation index per map: MODE 0: maps (LL, hi, cr, cb) -> (0, 1, 2, 2); MODE 1: (0, 1); MODE 2: (0, 0)
#pragma unroll
      for (int q = 0; q < NA; ++q) aN[q] = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          const int q = MODE == 0 ? (m < 2 ? m : 2) : (MODE == 1 ? m : 0);
          aN[q] = fmaf(wv[r * 3 + 0][m], g[r][m], aN[q]);
          aC[q] = fmaf(wv[r * 3 + 1][m], g[r][m], aC[q]);
          aD[q] = fmaf(wv[r * 3 + 2][m], g[r][m], aD[q]);
        }
      const int qo = j - 2;
      if (qo >= 0) {
        const int x = x0 + qo;
        float sp;
        if constexpr (MODE == 0) {
          if (FastMath<T>::value)
            sp = 1.f + k0 * sigmoid_fast(aD[0]) + k1 * tanh_fast(aD[1]) + k2 * sigmoid_fast(aD[2]);
          else
            sp = 1.f + k0 * sigmoid_f(aD[0]) + k1 * tanhf(aD[1]) + k2 * sigmoid_f(aD[2]);
        } else if constexpr (MODE == 1) {
          sp = FastMath<T>::value ? k0 * sigmoid_fast(aD[0]) + k1 * tanh_fast(aD[1 % NA])
                                  : k0 * sigmoid_f(aD[0]) + k1 * tanhf(aD[1 % NA]);
        } else {
          sp = FastMath<T>::value ? k0 * sigmoid_fast(aD[0]) : k0 * sigmoid_f(aD[0]);
        }
        if (c_ok) {
          const i64 off = ((i64)y * Wf + x) * C + c;
          const float o = to_f(fv[qo]) * sp;
          from_f(ob[off], o);
          csum += o;
        }
      }
    }
  }
  if (partial != nullptr && c_ok) atomicAdd(partial + (b * FLCA_SLOTS + (task % FLCA_SLOTS)) * C + c, csum);
}

template <typename T, int MODE>
static void run_flca_mod(Ctx& ctx, const void* feat, const float* G, const float* w, const float* coef, void* xmod,
                         float* partial, int nblk, int B, int Hf, int Wf, int C, int level, int kid) {
  (void)nblk;
  const i64 tasks = (i64)Hf * cdiv(Wf, FLCA_LS) * cdiv(C, 32);
  double px = (double)B * Hf * Wf;
  ScopedLaunch sl(kid, px * C * 2.0 * sizeof(T) + px * 16.0, px * C * 2.0 * (MODE == 0 ? 36 : 18));
  k_flca_mod<T, MODE><<<dim3((unsigned)cdivl(tasks, 8), B), 256, 0, ctx.stream>>>((const T*)feat, G, w, coef, (T*)xmod, partial,
                                                                                 Hf, Wf, C, level, tasks);
}

int launch_flca_mod(Ctx& ctx, const void* feat, const float* G, const void* G16, const float* w36, const float* abg,
                    void* xmod, float* partial, int nblk, int B, int Hf, int Wf, int C) {
  if (ctx.dry) return nblk;
  // (partial comes zero-initialised from zeroed_f32)
  if (G16 != nullptr && im2col_tc_supported(ctx, C)) {
    const double px = (double)B * Hf * Wf;
    ScopedLaunch sl(RF_K_FLCA_MOD, px * C * 4.0 + px * 16.0, px * C * 72.0);
    int used = nblk;
    if (!launch_flca_mod_tc(ctx, feat, G16, w36, abg, xmod, partial, B, Hf, Wf, C, &used) || used > nblk)
      recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return used;
  }
  if (ctx.band != nullptr) {   // only the tensor-core kernel restricts the channel sums to the band's interior rows
    recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return nblk;
  }
  if (ctx.dtype == RF_BF16)
    run_flca_mod<bf16, 0>(ctx, feat, G, w36, abg, xmod, partial, nblk, B, Hf, Wf, C, 0, RF_K_FLCA_MOD);
  else
    run_flca_mod<float, 0>(ctx, feat, G, w36, abg, xmod, partial, nblk, B, Hf, Wf, C, 0, RF_K_FLCA_MOD);
  return nblk;
}

void launch_pyr_spatial(Ctx& ctx, const void* x, const float* G8, const float* w54, const float* gates, void* xs, int mode,
                        int level, int B, int Hf, int Wf, int C) {
  if (ctx.dry) return;
  if (ctx.dtype == RF_BF16) {
    if (mode == 0) run_flca_mod<bf16, 1>(ctx, x, G8, w54, gates, xs, nullptr, 0, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
    else run_flca_mod<bf16, 2>(ctx, x, G8, w54, gates, xs, nullptr, 0, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
  } else {
    if (mode == 0) run_flca_mod<float, 1>(ctx, x, G8, w54, gates, xs, nullptr, 0, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
    else run_flca_mod<float, 2>(ctx, x, G8, w54, gates, xs, nullptr, 0, B, Hf, Wf, C, level, RF_K_PYR_SPATIAL);
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
__global__ void k_channel_sums(const T* __restrict__ x, float* __restrict__ partial, i64 P, int C, int nslots) {
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
    // (the launcher keeps the grid within the slots: every CTA owns its slot -> a plain store, bit-reproducible)
    float* dst = partial + (b * nslots + blockIdx.x % nslots) * C + c;
    if ((int)gridDim.x <= nslots) *dst = s;
    else atomicAdd(dst, s);                                               // partial is zero-initialised by the caller
  }
}
void launch_channel_sums(Ctx& ctx, const void* x, float* partial, int nblk, int B, i64 P, int C) {
  if (ctx.dry) return;
  const int cv = C / 8;
  int ppb = 256 / cv;
  if (ppb < 1) ppb = 1;
  size_t smem = sizeof(float) * (size_t)C * ppb;
  // enough CTAs to stream at HBM speed; they add into the nblk partial slots (zero-initialised by the caller)
  i64 want = cdivl(P, (i64)ppb * 64);
  unsigned gx = (unsigned)(want < 1 ? 1 : (want > 8 * num_sms() ? 8 * num_sms() : want));
  if (nblk > 0 && gx > (unsigned)nblk) gx = (unsigned)nblk;      // one CTA per partial slot (320: still > 2 CTAs per SM)
  ScopedLaunch sl(RF_K_CHANNEL_SUMS, (double)B * P * C * esize(ctx.dtype));
  if (ctx.dtype == RF_BF16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_channel_sums<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_channel_sums<bf16><<<dim3(gx, B), dim3(cv, ppb), smem, ctx.stream>>>((const bf16*)x, partial, P, C, nblk);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_channel_sums<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_channel_sums<float><<<dim3(gx, B), dim3(cv, ppb), smem, ctx.stream>>>((const float*)x, partial, P, C, nblk);
  }
}

// squeeze-excite MLP (FLCA_RF.py:124-130,160): one block per image
__global__ void __launch_bounds__(256)
k_se_finalize(const float* __restrict__ partial, int nblk, float invP, const float* __restrict__ w1,
              const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ scale, int C, int hid) {
  extern __shared__ float smem[];
  pdl_trigger();
  pdl_wait();
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
  launch_pdl(k_se_finalize, dim3(B), dim3(256), sizeof(float) * (C + hid), ctx.stream, partial, nblk, 1.0f / (float)P, w1, b1, w2,
             b2, scale, C, hid);
}

// squeeze-excite MLP + fold of its scale into channel_reduce in ONE launch (FLCA_RF.py:160-161,275-276): every CTA
// recomputes the tiny MLP (C x hid) in shared memory and then scales its slice of W_red [C][2C]
template <typename T>
__global__ void __launch_bounds__(256)
k_se_fold(const float* __restrict__ partial, int nblk, float invP, const float* __restrict__ w1, const float* __restrict__ b1,
          const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale_out,
          const float* __restrict__ red_w, T* __restrict__ wred, int C, int hid) {
  extern __shared__ float smem[];
  pdl_trigger();
  pdl_wait();
  float* mean = smem;            // [C]
  float* hbuf = smem + C;        // [hid]
  float* sc = hbuf + hid;        // [C]
  const int b = blockIdx.y;
  // channel sums over the nblk partial slots in a fixed order: 8 interleaved slot groups per channel in parallel (group
  // g adds slots g, g + 8, ...), then the 8 group sums in group order
  float* grp = sc + C;            // [8][C] scratch behind sc
  for (int idx = threadIdx.x; idx < 8 * C; idx += blockDim.x) {
    const int g = idx / C, c = idx - g * C;
    const float* pp = partial + (i64)b * nblk * C + c;
    float s0 = 0.f, s1 = 0.f;
    int k = g;
    for (; k + 8 < nblk; k += 16) {
      s0 += pp[(i64)k * C];
      s1 += pp[(i64)(k + 8) * C];
    }
    if (k < nblk) s0 += pp[(i64)k * C];
    grp[idx] = s0 + s1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s0 += grp[g * C + c];
    mean[c] = s0 * invP;
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
    // the row of W2 as float4s with four independent chains when hid % 4 == 0 (this loop was the kernel's latency)
    float s0 = b2[c], s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if ((hid & 3) == 0) {
      const float4* wr = reinterpret_cast<const float4*>(w2 + (i64)c * hid);
#pragma unroll 4
      for (int j = 0; j < hid / 4; ++j) {
        const float4 q = wr[j];
        s0 = fmaf(q.x, hbuf[4 * j], s0); s1 = fmaf(q.y, hbuf[4 * j + 1], s1);
        s2 = fmaf(q.z, hbuf[4 * j + 2], s2); s3 = fmaf(q.w, hbuf[4 * j + 3], s3);
      }
    } else {
      for (int j = 0; j < hid; ++j) s0 = fmaf(w2[(i64)c * hid + j], hbuf[j], s0);
    }
    const float s = sigmoid_f((s0 + s1) + (s2 + s3));
    sc[c] = s;
    if (blockIdx.x == 0) scale_out[(i64)b * C + c] = s;
  }
  __syncthreads();
  const i64 n2 = (i64)C * 2 * C;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (i64)gridDim.x * blockDim.x) {
    const int k = (int)(i % (2 * C));
    float v = red_w[i];
    if (k < C) v *= sc[k];
    from_f(wred[b * n2 + i], v);
  }
}
void launch_se_fold(Ctx& ctx, const float* partial, int nblk, i64 P, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* scale, const float* red_w, void* wred, int B, int C, int hid) {
  if (ctx.dry) return;
  const i64 n2 = (i64)C * 2 * C;
  unsigned gx = (unsigned)cdivl(n2, 2048);
  if (gx > (unsigned)num_sms()) gx = num_sms();
  const size_t smem = sizeof(float) * (2 * C + hid + 8 * C);
  ScopedLaunch sl(RF_K_SE_FINALIZE, 4.0 * B * nblk * C + (4.0 + esize(ctx.dtype)) * B * n2);
  if (ctx.dtype == RF_BF16)
    launch_pdl(k_se_fold<bf16>, dim3(gx, B), dim3(256), smem, ctx.stream, partial, nblk, 1.0f / (float)P, w1, b1, w2, b2, scale,
               red_w, (bf16*)wred, C, hid);
  else
    launch_pdl(k_se_fold<float>, dim3(gx, B), dim3(256), smem, ctx.stream, partial, nblk, 1.0f / (float)P, w1, b1, w2, b2, scale,
               red_w, (float*)wred, C, hid);
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
