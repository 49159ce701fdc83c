// Transformer-branch kernels on NHWC activations: LayerNorm, depthwise 3x3 (+GELU), the fused
// "depthwise-qkv + Gram + row norms" pass of the transposed attention, the per-image softmax/fold micro-kernel,
// and the 4->d embedding / d->12 head convolutions.
//
// Attention (FLCA_RF.py:221-235) is restructured so that q and k never reach HBM:
//   pass A  qkv_pre = W_qkv LN(x)            (GEMM)          -> [P,3C]
//   pass B  dw3x3(qkv_pre): v is stored; q,k only feed  G = sum_p q_p k_p^T (per head) and |q|^2, |k|^2
//   micro   attn = softmax(G / (|q||k|) * temperature);  M = W_proj * blockdiag(attn)      (C x C per image)
//   pass C  x1 = x + M v + b_proj           (GEMM with a per-image weight)
#include "rf_kernels.cuh"

namespace rf {

// ---------------------------------------------------------------------------------------------
// LayerNorm over C per pixel row.  GS lanes cooperate on one row (GS = pow2 >= C/8, <= 32).
// mode 0: nn.LayerNorm (FLCA_RF.py:183); mode 1: BiasFree (WFB/model.py:100-103, no mean subtraction in the numerator)
// ---------------------------------------------------------------------------------------------
template <typename T, int GS, int VPL>  // VPL = 8-wide vectors per lane
__global__ void __launch_bounds__(256)
k_layernorm(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta, T* __restrict__ out,
            float eps, int mode, i64 rows, int C) {
  const int lane = threadIdx.x & (GS - 1);
  const i64 row = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / GS;
  const bool row_ok = row < rows;
  const int cv = C >> 3;
  float v[VPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * GS;
    if (row_ok && vec < cv) {
      load8(x + row * C + vec * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
#pragma unroll
  for (int o = GS >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * GS;
    if (vec < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[i][j] - mu;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = GS >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)C + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * GS;
    if (row_ok && vec < cv) {
      float gg[8], bb[8], o[8];
      load8(g + vec * 8, gg);
      if (mode == 0 && bta != nullptr) load8(bta + vec * 8, bb);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (mode == 0 ? (v[i][j] - mu) : v[i][j]) * rstd * gg[j] + bb[j];
      store8(out + row * C + vec * 8, o);
    }
  }
}

template <typename T>
static void run_layernorm(Ctx& ctx, const void* x, const float* g, const float* b, void* out, float eps, int mode, i64 rows,
                          int C) {
  const int cv = C / 8;
#define RF_LN(GS, VPL)                                                                                            \
  k_layernorm<T, GS, VPL><<<(unsigned)cdivl(rows * GS, 256), 256, 0, ctx.stream>>>((const T*)x, g, b, (T*)out, eps, mode, \
                                                                                 rows, C)
  if (cv <= 4) RF_LN(4, 1);
  else if (cv <= 8) RF_LN(8, 1);
  else if (cv <= 16) RF_LN(16, 1);
  else if (cv <= 32) RF_LN(32, 1);
  else if (cv <= 64) RF_LN(32, 2);
  else if (cv <= 128) RF_LN(32, 4);
  else RF_LN(32, 8);
#undef RF_LN
}

void launch_layernorm(Ctx& ctx, const void* x, const float* g, const float* b, void* out, float eps, int mode, i64 rows,
                      int C) {
  if (ctx.dry || rows <= 0) return;
  ScopedLaunch sl(RF_K_LAYERNORM, 2.0 * rows * C * esize(ctx.dtype));
  if (ctx.dtype == RF_BF16) run_layernorm<bf16>(ctx, x, g, b, out, eps, mode, rows, C);
  else run_layernorm<float>(ctx, x, g, b, out, eps, mode, rows, C);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm folded into the following 1x1 conv (bf16 mode): per-row (sum, sumsq) and the weight fold
// ---------------------------------------------------------------------------------------------
template <typename T, int GS, int VPL>
__global__ void __launch_bounds__(256)
k_row_stats(const T* __restrict__ x, float2* __restrict__ stats, i64 rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & (GS - 1);
  const i64 row = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / GS;
  const bool row_ok = row < rows;
  const int cv = C >> 3;
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * GS;
    if (row_ok && vec < cv) {
      float v[8];
      load8(x + row * C + vec * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[j];
        q = fmaf(v[j], v[j], q);
      }
    }
  }
#pragma unroll
  for (int o = GS >> 1; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (row_ok && lane == 0) stats[row] = make_float2(s, q);
}

void launch_row_stats(Ctx& ctx, const void* x, float* stats, i64 rows, int C) {
  if (ctx.dry || rows <= 0) return;
  ScopedLaunch sl(RF_K_LAYERNORM, rows * (C * (double)esize(ctx.dtype) + 8.0));
  const int cv = C / 8;
#define RF_RS(T, GS, VPL) \
  launch_pdl(k_row_stats<T, GS, VPL>, dim3((unsigned)cdivl(rows * GS, 256)), dim3(256), 0, ctx.stream, (const T*)x, (float2*)stats, rows, C)
#define RF_RS_ALL(T)                   \
  do {                                 \
    if (cv <= 4) RF_RS(T, 4, 1);       \
    else if (cv <= 8) RF_RS(T, 8, 1);  \
    else if (cv <= 16) RF_RS(T, 16, 1); \
    else if (cv <= 32) RF_RS(T, 32, 1); \
    else if (cv <= 64) RF_RS(T, 32, 2); \
    else if (cv <= 128) RF_RS(T, 32, 4); \
    else RF_RS(T, 32, 8);              \
  } while (0)
  if (ctx.dtype == RF_BF16) RF_RS_ALL(bf16);
  else RF_RS_ALL(float);
#undef RF_RS_ALL
#undef RF_RS
}

// one warp per output row n
template <typename T>
__global__ void __launch_bounds__(256)
k_fold_ln(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
          const float* __restrict__ bias, T* __restrict__ Wf, float* __restrict__ cs, float* __restrict__ bf, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float c = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(i64)n * K + k];
    T r;
    from_f(r, w * gamma[k]);
    Wf[(i64)n * K + k] = r;
    c += to_f(r);
    bb = fmaf(w, beta[k], bb);
  }
  c = warp_sum(c);
  bb = warp_sum(bb);
  if (lane == 0) {
    cs[n] = c;
    bf[n] = (bias ? bias[n] : 0.f) + bb;
  }
}

void launch_fold_ln(Ctx& ctx, const float* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* cs,
                    float* bf, int N, int K) {
  if (ctx.dry || !W || !gamma || !beta) return;
  ScopedLaunch sl(RF_K_WEIGHT_PACK);
  if (ctx.dtype == RF_BF16) k_fold_ln<bf16><<<cdiv(N, 8), 256, 0, ctx.stream>>>(W, gamma, beta, bias, (bf16*)Wf, cs, bf, N, K);
  else k_fold_ln<float><<<cdiv(N, 8), 256, 0, ctx.stream>>>(W, gamma, beta, bias, (float*)Wf, cs, bf, N, K);
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 (+bias, optional exact-erf GELU) on NHWC; one thread = 8 channels of one pixel
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void dw3x3_at(const T* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                         int H, int W, int Cn, int y, int x, int c0, float (&acc)[8]) {
  load8(bias + c0, acc);
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      float v[8], k[8];
      load8(img + ((i64)yy * W + xx) * Cn + c0, v);
      load8(w + ((dy + 1) * 3 + dx + 1) * Cn + c0, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], k[j], acc[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass B of the attention: depthwise 3x3 on qkv_pre, store v, accumulate Gram + squared norms.
// Persistent blocks loop over tiles of TP pixels; q,k of a tile are staged in smem (fp32) and the per-head c x c
// Gram patches are accumulated in registers across all tiles of the block; one atomicAdd per entry at the end.
// stats[b] = { G[C][C] (row = q channel, col = k channel; only the per-head diagonal blocks are written), qn2[C], kn2[C] }.
// ---------------------------------------------------------------------------------------------
template <typename T, int PT, int NP>  // PT x PT register patch, NP patches per thread
__global__ void __launch_bounds__(256)
k_dwqkv_gram(const T* __restrict__ qkv, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ vout,
             float* __restrict__ stats, int H, int W, int C, int TP, int tiles_per_img) {
  extern __shared__ float smem[];
  const int ldq = C + 4;          // padded row
  float* sq = smem;               // [TP][ldq]
  float* sk = smem + TP * ldq;    // [TP][ldq]
  const i64 b = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int c = C >> 3;           // head width
  const int C3 = 3 * C, v3 = C3 >> 3;
  const i64 P = (i64)H * W;
  const T* img = qkv + b * P * C3;
  const int ppr = c / PT;                  // patches per head row
  const int npatch = 8 * ppr * ppr;

  float acc[NP][PT][PT];
#pragma unroll
  for (int a = 0; a < NP; ++a)
#pragma unroll
    for (int i = 0; i < PT; ++i)
#pragma unroll
      for (int j = 0; j < PT; ++j) acc[a][i][j] = 0.f;
  float n2[4] = {0.f, 0.f, 0.f, 0.f};  // squared norms of channels tid, tid+nthr, ... (< 2C <= 4*nthr)

  for (int tile = blockIdx.x; tile < tiles_per_img; tile += gridDim.x) {
    const i64 p0 = (i64)tile * TP;
    // phase 1: depthwise conv of the tile
    for (int it = tid; it < TP * v3; it += nthr) {
      const int pl = it / v3, c0 = (it - pl * v3) * 8;
      const i64 p = p0 + pl;
      float a8[8];
      if (p < P) {
        const int y = (int)(p / W), x = (int)(p - (i64)y * W);
        dw3x3_at(img, w, bias, H, W, C3, y, x, c0, a8);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = 0.f;
      }
      if (c0 < C) {
        store8(sq + pl * ldq + c0, a8);
      } else if (c0 < 2 * C) {
        store8(sk + pl * ldq + (c0 - C), a8);
      } else if (p < P) {
        store8(vout + (b * P + p) * C + (c0 - 2 * C), a8);
      }
    }
    __syncthreads();
    // phase 2a: squared norms
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ch = tid + r * nthr;
      if (ch < 2 * C) {
        const float* s = ch < C ? sq + ch : sk + (ch - C);
        float t = 0.f;
        for (int pl = 0; pl < TP; ++pl) {
          float v = s[pl * ldq];
          t = fmaf(v, v, t);
        }
        n2[r] += t;
      }
    }
    // phase 2b: Gram patches
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      const int pe = tid + a * nthr;
      if (pe < npatch) {
        const int h = pe / (ppr * ppr), rem = pe - h * ppr * ppr;
        const int i0 = h * c + (rem / ppr) * PT, j0 = h * c + (rem % ppr) * PT;
        for (int pl = 0; pl < TP; ++pl) {
          float qv[PT], kv[PT];
          if constexpr (PT == 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(sq + pl * ldq + i0);
            const float4 k4 = *reinterpret_cast<const float4*>(sk + pl * ldq + j0);
            qv[0] = q4.x; qv[1] = q4.y; qv[2] = q4.z; qv[3] = q4.w;
            kv[0] = k4.x; kv[1] = k4.y; kv[2] = k4.z; kv[3] = k4.w;
          } else if constexpr (PT == 2) {
            const float2 q2 = *reinterpret_cast<const float2*>(sq + pl * ldq + i0);
            const float2 k2 = *reinterpret_cast<const float2*>(sk + pl * ldq + j0);
            qv[0] = q2.x; qv[1] = q2.y; kv[0] = k2.x; kv[1] = k2.y;
          } else {
            qv[0] = sq[pl * ldq + i0];
            kv[0] = sk[pl * ldq + j0];
          }
#pragma unroll
          for (int i = 0; i < PT; ++i)
#pragma unroll
            for (int j = 0; j < PT; ++j) acc[a][i][j] = fmaf(qv[i], kv[j], acc[a][i][j]);
        }
      }
    }
    __syncthreads();
  }
  float* st = stats + b * ((i64)C * C + 2 * C);
#pragma unroll
  for (int a = 0; a < NP; ++a) {
    const int pe = tid + a * nthr;
    if (pe < npatch) {
      const int h = pe / (ppr * ppr), rem = pe - h * ppr * ppr;
      const int i0 = (rem / ppr) * PT, j0 = (rem % ppr) * PT;
#pragma unroll
      for (int i = 0; i < PT; ++i)
#pragma unroll
        for (int j = 0; j < PT; ++j) atomicAdd(st + ((i64)h * c + i0 + i) * C + h * c + j0 + j, acc[a][i][j]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ch = tid + r * nthr;
    if (ch < 2 * C) atomicAdd(st + (i64)C * C + ch, n2[r]);
  }
}

template <typename T>
static int run_dwqkv_gram(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* v, float* stats, int B,
                          int H, int W, int C) {
  const int c = C / 8;
  const int PT = (c % 4 == 0) ? 4 : ((c % 2 == 0) ? 2 : 1);
  const int npatch = 8 * (c / PT) * (c / PT);
  const int NP = cdiv(npatch, 256);
  // tile size: q,k staging <= ~96 KB
  int TP = (96 * 1024) / (2 * (C + 4) * 4);
  if (TP > 128) TP = 128;
  if (TP < 8) TP = 8;
  TP &= ~7;
  const i64 P = (i64)H * W;
  const int tiles = (int)cdivl(P, TP);
  int gx = 2 * num_sms() / (B > 0 ? B : 1);
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  const size_t smem = sizeof(float) * 2 * (size_t)TP * (C + 4);
  double px = (double)B * P;
  ScopedLaunch sl(RF_K_DW_QKV_GRAM, px * 4.0 * C * sizeof(T), px * (54.0 * C + 2.0 * C * c));
#define RF_GRAM(PT_, NP_)                                                                                              \
  do {                                                                                                                 \
    auto kern = k_dwqkv_gram<T, PT_, NP_>;                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    kern<<<dim3(gx, B), 256, smem, ctx.stream>>>((const T*)qkv_pre, dw_w, dw_b, (T*)v, stats, H, W, C, TP, tiles);      \
    return RF_OK;                                                                                                      \
  } while (0)
  if (PT == 4) {
    if (NP <= 1) RF_GRAM(4, 1);
    if (NP <= 2) RF_GRAM(4, 2);
    if (NP <= 4) RF_GRAM(4, 4);
    if (NP <= 8) RF_GRAM(4, 8);
  } else if (PT == 2) {
    if (NP <= 1) RF_GRAM(2, 1);
    if (NP <= 2) RF_GRAM(2, 2);
    if (NP <= 4) RF_GRAM(2, 4);
    if (NP <= 8) RF_GRAM(2, 8);
    if (NP <= 16) RF_GRAM(2, 16);
  } else {
    if (NP <= 1) RF_GRAM(1, 1);
    if (NP <= 4) RF_GRAM(1, 4);
    if (NP <= 16) RF_GRAM(1, 16);
  }
#undef RF_GRAM
  return RF_ERR_UNSUPPORTED;
}

void launch_dwqkv_gram(Ctx& ctx, const void* qkv_pre, const float* dw_w, const float* dw_b, void* v, float* stats, int B,
                       int H, int W, int C) {
  if (ctx.dry) return;
  int st = ctx.dtype == RF_BF16 ? run_dwqkv_gram<bf16>(ctx, qkv_pre, dw_w, dw_b, v, stats, B, H, W, C)
                                : run_dwqkv_gram<float>(ctx, qkv_pre, dw_w, dw_b, v, stats, B, H, W, C);
  if (st != RF_OK) recorder().last_cuda_error = (int)cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// softmax over the c x c Gram of one (image, head) and fold into project_out:
//   Mw[b][n][h*c + j] = sum_i proj_w[n][h*c + i] * attn[h][i][j]
// F.normalize semantics: divide by max(||.||, 1e-12) (FLCA_RF.py:228-229); temperature per head (:230).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_attn_finalize(const float* __restrict__ stats, const float* __restrict__ temperature, const float* __restrict__ proj_w,
                T* __restrict__ Mw, int C, const float* __restrict__ norms) {
  extern __shared__ float smem[];
  pdl_trigger();
  pdl_wait();
  const int c = C >> 3;
  float* attn = smem;           // [c][c+1]
  float* nq = smem + c * (c + 1);
  float* nk = nq + c;
  const int h = blockIdx.x;
  const i64 b = blockIdx.y;
  const float* st = stats + b * ((i64)C * C + 2 * C);
  const float* gram = st + ((i64)h * c) * C + h * c;   // diagonal block of head h, row pitch C
  const float* nrm = norms != nullptr ? norms + b * 2 * C : st + (i64)C * C;
  const float* qn2 = nrm + h * c;
  const float* kn2 = nrm + C + h * c;
  const int tid = threadIdx.x;
  for (int i = tid; i < c; i += blockDim.x) {
    nq[i] = fmaxf(sqrtf(qn2[i]), 1e-12f);
    nk[i] = fmaxf(sqrtf(kn2[i]), 1e-12f);
  }
  __syncthreads();
  const float temp = temperature[h];
  for (int e = tid; e < c * c; e += blockDim.x) {
    const int i = e / c, j = e - i * c;
    attn[i * (c + 1) + j] = gram[(i64)i * C + j] / (nq[i] * nk[j]) * temp;
  }
  __syncthreads();
  for (int i = tid; i < c; i += blockDim.x) {  // row softmax (c <= 64: one thread per row is fine)
    float* r = attn + i * (c + 1);
    float m = -INFINITY;
    for (int j = 0; j < c; ++j) m = fmaxf(m, r[j]);
    float s = 0.f;
    for (int j = 0; j < c; ++j) {
      float e = expf(r[j] - m);
      r[j] = e;
      s += e;
    }
    const float inv = 1.0f / s;
    for (int j = 0; j < c; ++j) r[j] *= inv;
  }
  __syncthreads();
  // the fold itself is split over gridDim.z CTAs (slices of the output rows n); the c x c softmax above is recomputed
  // by each of them (it is tiny), so the 8-CTA latency chain of the large stages becomes 8 * C/32 CTAs
  const int nper = (C + gridDim.z - 1) / gridDim.z;
  const int nb = blockIdx.z * nper, ne = min(C, nb + nper);
  for (int e = tid + nb * c; e < ne * c; e += blockDim.x) {
    const int n = e / c, j = e - n * c;
    const float* pw = proj_w + (i64)n * C + h * c;
    float s = 0.f;
    for (int i = 0; i < c; ++i) s = fmaf(pw[i], attn[i * (c + 1) + j], s);
    from_f(Mw[(b * C + n) * C + h * c + j], s);
  }
}

// Ordered second stage of the attention statistics (bit-reproducible: fixed summation order and association):
//   stats[row*C + h*c + j] = sum_z gram_part[(z*C + row)*c + j]     (per-head diagonal blocks of q^T k, h = row / c)
//   norms[ch]              = sum_slot sq_part[slot*2C + ch]          (squared norms of q | k)
__global__ void __launch_bounds__(256)
k_attn_reduce(const float* __restrict__ gram_part, int nsplit, const float* __restrict__ sq_part, int nslots,
              float* __restrict__ stats, float* __restrict__ norms, int C) {
  pdl_trigger();
  pdl_wait();
  // one WARP per output: lane l adds slots l, l + 32, ... (in that order), then a fixed xor-shuffle tree -- the same
  // association every run, whatever the slot count
  const int c = C >> 3, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int ng = gram_part != nullptr ? C * c : 0;
  const float* src;
  i64 step;
  int n;
  float* dst;
  if (i < ng) {
    const int row = i / c, j = i - row * c;
    src = gram_part + i; step = (i64)C * c; n = nsplit;
    dst = stats + (i64)row * C + (row / c) * c + j;
  } else if (sq_part != nullptr && i - ng < 2 * C) {
    src = sq_part + (i - ng); step = 2 * C; n = nslots;
    dst = norms + (i - ng);
  } else {
    return;
  }
  float a0 = 0.f, a1 = 0.f;
  int k = lane;
  for (; k + 32 < n; k += 64) {
    a0 += src[(i64)k * step];
    a1 += src[(i64)(k + 32) * step];
  }
  if (k < n) a0 += src[(i64)k * step];
  const float v = warp_sum(a0 + a1);
  if (lane == 0) *dst = v;
}

void launch_attn_reduce(Ctx& ctx, const float* gram_part, int nsplit, const float* sq_part, int nslots, float* stats, float* norms,
                        int C) {
  if (ctx.dry) return;
  if (nsplit <= 0) gram_part = nullptr;
  if (nslots <= 0) sq_part = nullptr;
  if (!gram_part && !sq_part) return;
  const int n = (gram_part ? C * (C / 8) : 0) + (sq_part ? 2 * C : 0);
  ScopedLaunch sl(RF_K_ATTN_FINALIZE, 4.0 * ((double)nsplit * C * (C / 8) + (double)nslots * 2 * C));
  launch_pdl(k_attn_reduce, dim3(cdiv(n, 8)), dim3(256), 0, ctx.stream, gram_part, nsplit, sq_part, nslots, stats, norms, C);
}

void launch_attn_finalize(Ctx& ctx, const float* stats, const float* temperature, const float* proj_w, void* Mw, int B,
                          int C, const float* norms) {
  if (ctx.dry) return;
  const int c = C / 8;
  size_t smem = sizeof(float) * (c * (c + 1) + 2 * c);
  const int nz = C >= 64 ? C / 32 : 1;
  ScopedLaunch sl(RF_K_ATTN_FINALIZE, 4.0 * B * C * c + (4.0 + esize(ctx.dtype)) * B * C * C, 2.0 * B * C * C * c);
  if (ctx.dtype == RF_BF16)
    launch_pdl(k_attn_finalize<bf16>, dim3(8, B, nz), dim3(256), smem, ctx.stream, stats, temperature, proj_w, (bf16*)Mw, C, norms);
  else
    launch_pdl(k_attn_finalize<float>, dim3(8, B, nz), dim3(256), smem, ctx.stream, stats, temperature, proj_w, (float*)Mw, C, norms);
}

// ---------------------------------------------------------------------------------------------
// embedding: dense 3x3, 4 -> d, from the fp32 packed frame (FLCA_RF.py:303,338); w [9][4][d]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_embed(const float4* __restrict__ x_ds, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
        int h, int wd, int d) {
  extern __shared__ float sw[];  // [36][d]
  for (int i = threadIdx.x; i < 36 * d; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const i64 b = blockIdx.y;
  const int cv = d >> 3;
  const i64 total = (i64)h * wd * cv;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    const i64 p = i / cv;
    const int y = (int)(p / wd), x = (int)(p % wd);
    float acc[8];
    load8(bias + c0, acc);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= wd) continue;
        const float4 q = x_ds[(b * h + yy) * (i64)wd + xx];
        const float in4[4] = {q.x, q.y, q.z, q.w};
        const float* wr = sw + ((dy + 1) * 3 + dx + 1) * 4 * d + c0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float k[8];
          load8(wr + ch * d, k);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(in4[ch], k[j], acc[j]);
        }
      }
    }
    store8(out + (b * (i64)h * wd + p) * d + c0, acc);
  }
}

void launch_embed(Ctx& ctx, const float* x_ds, const void* x16, const float* w, const float* b, void* out, int B, int h,
                  int w_, int d) {
  if (ctx.dry) return;
  i64 total = (i64)h * w_ * (d / 8);
  unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  double px = (double)B * h * w_;
  ScopedLaunch sl(RF_K_EMBED, px * (16.0 + d * esize(ctx.dtype)), 72.0 * px * d);
  if (x16 != nullptr && im2col_tc_supported(ctx, d)) {
    if (!launch_embed_tc(ctx, x16, w, b, out, B, h, w_, d)) recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return;
  }
  size_t smem = sizeof(float) * 36 * d;
  if (ctx.dtype == RF_BF16)
    k_embed<bf16><<<dim3(gx, B), 256, smem, ctx.stream>>>((const float4*)x_ds, w, b, (bf16*)out, h, w_, d);
  else
    k_embed<float><<<dim3(gx, B), 256, smem, ctx.stream>>>((const float4*)x_ds, w, b, (float*)out, h, w_, d);
}

// ---------------------------------------------------------------------------------------------
// head: dense 3x3 d -> 12, LeakyReLU(0.2), PixelShuffle(2) into fp32 NCHW [B,3,2h,2w] (FLCA_RF.py:368-369)
// w [9][d][12]; one thread = one packed pixel (12 outputs)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
k_head(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int h,
       int wd, int d) {
  extern __shared__ float sw[];  // [9*d][12]
  for (int i = threadIdx.x; i < 9 * d * 12; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const i64 b = blockIdx.y;
  const i64 P = (i64)h * wd;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int y = (int)(p / wd), x = (int)(p % wd);
  float acc[12];
#pragma unroll
  for (int n = 0; n < 12; ++n) acc[n] = bias[n];
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= wd) continue;
      const T* src = in + ((b * h + yy) * (i64)wd + xx) * d;
      const float* wt = sw + ((dy + 1) * 3 + dx + 1) * d * 12;
      for (int c0 = 0; c0 < d; c0 += 8) {
        float v[8];
        load8(src + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4* w4 = reinterpret_cast<const float4*>(wt + (c0 + j) * 12);
          float4 a = w4[0], bq = w4[1], cq = w4[2];
          acc[0] = fmaf(v[j], a.x, acc[0]); acc[1] = fmaf(v[j], a.y, acc[1]);
          acc[2] = fmaf(v[j], a.z, acc[2]); acc[3] = fmaf(v[j], a.w, acc[3]);
          acc[4] = fmaf(v[j], bq.x, acc[4]); acc[5] = fmaf(v[j], bq.y, acc[5]);
          acc[6] = fmaf(v[j], bq.z, acc[6]); acc[7] = fmaf(v[j], bq.w, acc[7]);
          acc[8] = fmaf(v[j], cq.x, acc[8]); acc[9] = fmaf(v[j], cq.y, acc[9]);
          acc[10] = fmaf(v[j], cq.z, acc[10]); acc[11] = fmaf(v[j], cq.w, acc[11]);
        }
      }
    }
  }
  // pixel shuffle: out[b, ch, 2y+i, 2x+j] = lrelu(acc[4*ch + 2*i + j])
  const i64 Wo = 2 * (i64)wd, Ho = 2 * (i64)h;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float2 v = make_float2(lrelu_f(acc[4 * ch + 2 * i]), lrelu_f(acc[4 * ch + 2 * i + 1]));
      *reinterpret_cast<float2*>(out + ((b * 3 + ch) * Ho + 2 * y + i) * Wo + 2 * x) = v;
    }
}

void launch_head(Ctx& ctx, const void* in, const float* w, const float* b, float* out, int B, int h, int w_, int d) {
  if (ctx.dry) return;
  i64 P = (i64)h * w_;
  double px = (double)B * P;
  ScopedLaunch sl(RF_K_HEAD, px * (d * esize(ctx.dtype) + 48.0), 216.0 * px * d);
  size_t smem = sizeof(float) * 9 * d * 12;
  if (ctx.dtype == RF_BF16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_head<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_head<bf16><<<dim3((unsigned)cdivl(P, 128), B), 128, smem, ctx.stream>>>((const bf16*)in, w, b, out, h, w_, d);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_head<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_head<float><<<dim3((unsigned)cdivl(P, 128), B), 128, smem, ctx.stream>>>((const float*)in, w, b, out, h, w_, d);
  }
}

// ---------------------------------------------------------------------------------------------
// ML tail (ML_RF.py:270-288,403-414).  sums[b][0..2] = sum of bilinear-upsampled (R, (G1+G2)/2, B),
// sums[b][3..5] = sum of out channels.  apply: out += 0.12*(in_mean - out_mean); then
// out += 0.03*(up(LL2) - (.299 R + .587 G + .114 B)) on the corrected values.
// ---------------------------------------------------------------------------------------------
// per-CTA partial sums -> this CTA's slot (plain stores: bit-reproducible, unlike atomics); 256 threads
template <int NK>
__device__ __forceinline__ void cta_store_sums(const float (&acc)[NK], float* slot, int nk) {
  __shared__ float red[8][NK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if ((int)threadIdx.x < nk) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    slot[threadIdx.x] = t;
  }
}
__global__ void k_sum_slots(const float* __restrict__ part, int nslots, int width, float* __restrict__ out, int ostride) {
  const int b = blockIdx.x, i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (i >= width) return;
  float a = 0.f;
  for (int s = lane; s < nslots; s += 32) a += part[((i64)b * nslots + s) * width + i];
  a = warp_sum(a);
  if (lane == 0) out[b * ostride + i] = a;
}
void launch_sum_slots(Ctx& ctx, const float* part, int nslots, int width, float* out, int ostride, int B) {
  if (ctx.dry) return;
  ScopedLaunch sl(RF_K_MISC);
  k_sum_slots<<<B, 32 * width, 0, ctx.stream>>>(part, nslots, width, out, ostride);
}

__global__ void __launch_bounds__(256)
k_tail_stats(const float* __restrict__ out, const float4* __restrict__ x_ds, float* part, int h, int wd) {
  const i64 b = blockIdx.y;
  const int Ho = 2 * h, Wo = 2 * wd;
  const i64 total = (i64)Ho * Wo;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo), y = (int)(i / Wo);
    int ya, yb, xa, xb;
    float ly, lx;
    bilinear_taps(y, h, Ho, ya, yb, ly);
    bilinear_taps(x, wd, Wo, xa, xb, lx);
    const float4 q00 = x_ds[(b * h + ya) * (i64)wd + xa], q01 = x_ds[(b * h + ya) * (i64)wd + xb];
    const float4 q10 = x_ds[(b * h + yb) * (i64)wd + xa], q11 = x_ds[(b * h + yb) * (i64)wd + xb];
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    acc[0] += w00 * q00.x + w01 * q01.x + w10 * q10.x + w11 * q11.x;
    acc[1] += w00 * (0.5f * (q00.y + q00.z)) + w01 * (0.5f * (q01.y + q01.z)) + w10 * (0.5f * (q10.y + q10.z)) +
              w11 * (0.5f * (q11.y + q11.z));
    acc[2] += w00 * q00.w + w01 * q01.w + w10 * q10.w + w11 * q11.w;
    if (out != nullptr) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) acc[3 + ch] += out[(b * 3 + ch) * total + i];
    }
  }
  const int nk = out != nullptr ? 6 : 3;
  cta_store_sums<6>(acc, part + ((i64)b * gridDim.x + blockIdx.x) * nk, nk);
}
__global__ void __launch_bounds__(256)
k_out_sums(const float* __restrict__ out, i64 ch_stride, i64 first, i64 count, float* part) {
  float acc[3] = {0, 0, 0};
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) acc[ch] += out[ch * ch_stride + first + i];
  }
  cta_store_sums<3>(acc, part + (i64)blockIdx.x * 3, 3);
}
void launch_out_sums(Ctx& ctx, const float* out, int rows_total, int row0, int rows, int Wo, float* sums3) {
  const i64 count = (i64)rows * Wo;
  const unsigned gx = (unsigned)(cdivl(count, 256) < 8 * num_sms() ? cdivl(count, 256) : 8 * num_sms());
  const size_t mk = ctx.arena.mark();
  float* part = ctx.arena.get<float>((size_t)gx * 3);
  if (!ctx.dry) {
    {
      ScopedLaunch sl(RF_K_TAIL_STATS, 12.0 * count);
      k_out_sums<<<gx, 256, 0, ctx.stream>>>(out, (i64)rows_total * Wo, (i64)row0 * Wo, count, part);
    }
    launch_sum_slots(ctx, part, (int)gx, 3, sums3, 3, 1);
  }
  ctx.arena.release(mk);
}
void launch_tail_stats(Ctx& ctx, const float* out, const float* x_ds, float* sums, int B, int h, int w_) {
  const i64 total = 4 * (i64)h * w_;
  const unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  const int nk = (out != nullptr || ctx.dry) ? 6 : 3;
  const size_t mk = ctx.arena.mark();
  float* part = ctx.arena.get<float>((size_t)B * gx * 6);
  if (!ctx.dry) {
    {
      ScopedLaunch sl(RF_K_TAIL_STATS, (out ? 12.0 : 0.0) * B * total + 16.0 * B * h * w_);
      k_tail_stats<<<dim3(gx, B), 256, 0, ctx.stream>>>(out, (const float4*)x_ds, part, h, w_);
    }
    launch_sum_slots(ctx, part, (int)gx, nk, sums, 8, B);
  }
  ctx.arena.release(mk);
}

__global__ void __launch_bounds__(256)
k_tail_apply(float* __restrict__ out, const float* __restrict__ sums, const float* __restrict__ LL2, int H2, int W2, int h,
             int wd, int y_off, int rows) {
  const i64 b = blockIdx.y;
  const int Ho = 2 * h, Wo = 2 * wd;
  const i64 total = (i64)rows * Wo;                      // elements per channel of `out` (rows [y_off, y_off + rows) of the frame)
  const float inv = 1.0f / (float)((i64)Ho * Wo);        // the means are the whole frame's
  float corr[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) corr[ch] = 0.12f * (sums[b * 8 + ch] * inv - sums[b * 8 + 3 + ch] * inv);
  const float* ll = LL2 + b * (i64)H2 * W2;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo), y = (int)(i / Wo) + y_off;
    int ya, yb, xa, xb;
    float ly, lx;
    bilinear_taps(y, H2, Ho, ya, yb, ly);
    bilinear_taps(x, W2, Wo, xa, xb, lx);
    const float top = ll[(i64)ya * W2 + xa] * (1.f - lx) + ll[(i64)ya * W2 + xb] * lx;
    const float bot = ll[(i64)yb * W2 + xa] * (1.f - lx) + ll[(i64)yb * W2 + xb] * lx;
    const float up = top * (1.f - ly) + bot * ly;
    float r = out[(b * 3 + 0) * total + i] + corr[0];
    float g = out[(b * 3 + 1) * total + i] + corr[1];
    float bl = out[(b * 3 + 2) * total + i] + corr[2];
    const float yres = (up - (0.299f * r + 0.587f * g + 0.114f * bl)) * 0.03f;
    out[(b * 3 + 0) * total + i] = r + yres;
    out[(b * 3 + 1) * total + i] = g + yres;
    out[(b * 3 + 2) * total + i] = bl + yres;
  }
}
void launch_tail_apply(Ctx& ctx, float* out, const float* sums, const float* LL2, int H2, int W2, int B, int h, int w_,
                       int y_off, int rows) {
  if (ctx.dry) return;
  if (rows < 0) { y_off = 0; rows = 2 * h; }
  i64 total = 2 * (i64)rows * w_;
  unsigned gx = (unsigned)(cdivl(total, 256) < 8 * num_sms() ? cdivl(total, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_TAIL_APPLY, 24.0 * B * total);
  k_tail_apply<<<dim3(gx, B), 256, 0, ctx.stream>>>(out, sums, LL2, H2, W2, h, w_, y_off, rows);
}

}  // namespace rf

// WithBias_LayerNorm / BiasFree_LayerNorm of the WFB variant (RawFomer_WFB_FFAB/model.py:89-122) on the [.., C] rows they
// are called with (to_3d: 'b (h w) c'), fp32.  mode 0: (x - mu) / sqrt(var + eps) * w + b; mode 1: x / sqrt(var + eps) * w.
extern "C" int rf_layernorm_rows(const float* in, const float* weight, const float* bias, float* out, float eps, int mode,
                                 long long rows, int C, void* stream) {
  using namespace rf;
  if (!in || !weight || !out || (mode == 0 && !bias)) return RF_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return RF_ERR_BAD_ARG;
  if (rows < 0 || C <= 0 || C % 8) return RF_ERR_BAD_SHAPE;
  if ((uintptr_t)in % 16 || (uintptr_t)out % 16) return RF_ERR_BAD_ARG;
  if (rows == 0) return RF_OK;
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  ctx.dtype = RF_F32;
  launch_layernorm(ctx, in, weight, bias, out, eps, mode, rows, C);
  return check_cuda(cudaGetLastError());
}

