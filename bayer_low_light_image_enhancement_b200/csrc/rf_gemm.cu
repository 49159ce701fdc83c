// Generic "pixel-row" contraction on CUDA cores (fp32 FFMA, fp32 accumulate):
//     Y[b][m][n] = act( sum_k A[b][m][k] * W[(b)][n][k] + bias[n] ) (+ R[b][m][n])
// with  A = [A1 | A2] (K-concatenation: skip/cat fusions), optional implicit 3x3 im2col gather on A1
// (zero padding), and three output addressings (plain rows, ConvTranspose2d 2x2 scatter, pixel-unshuffle scatter).
// This is the parity-mode (fp32) engine and the fallback shape coverage of the bf16 mode; the bf16 hot shapes are
// served by the tcgen05 kernels in rf_tc_gemm.cu.
#include "rf_kernels.cuh"

namespace rf {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <typename T>
__device__ __forceinline__ void ldg8(const T* p, bool ok, float (&v)[8]) {
  if (ok) {
    load8(p, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(GTHREADS)
k_gemm(GemmP p) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads: 4 cols x 8 rows each
  const i64 b = blockIdx.z;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int K = p.K1 + p.K2;
  const T* A1 = (const T*)p.A1;
  const T* A2 = (const T*)p.A2;
  const T* Wt = (const T*)p.Wt + b * p.w_img;

  // A loader: thread -> (row, 8-wide k vector)
  const int arow = tid >> 1, akv = (tid & 1) * 8;
  const int am = m0 + arow;
  const bool am_ok = am < p.M;
  int ay = 0, ax = 0;
  if (p.amode == AMODE_CONV3 && am_ok) { ay = am / p.W; ax = am % p.W; }
  const int Cin = p.amode == AMODE_CONV3 ? p.K1 / 9 : 0;
  // W loader: threads 0..127 -> (n row, 8-wide k vector)
  const int brow = tid >> 1, bkv = (tid & 1) * 8;
  const bool b_thr = tid < 2 * GBN;
  const int bn = n0 + brow;
  const bool bn_ok = b_thr && bn < p.N;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto fetch = [&](int k0) {
    const int k = k0 + akv;
    bool ok = am_ok && k < K;
    const T* src = nullptr;
    if (ok) {
      if (p.amode == AMODE_CONV3) {
        const int tap = k / Cin, c = k - tap * Cin;
        const int yy = ay + tap / 3 - 1, xx = ax + tap % 3 - 1;
        ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        src = A1 + ((b * p.M + (i64)yy * p.W + xx) * p.lda1 + c);
      } else if (k < p.K1) {
        src = A1 + ((b * p.M + am) * p.lda1 + k);
      } else {
        src = A2 + ((b * p.M + am) * p.lda2 + (k - p.K1));
      }
    }
    ldg8(src, ok, ra);
    const int kb = k0 + bkv;
    const bool okb = bn_ok && kb < K;
    ldg8(Wt + ((i64)bn * K + kb), okb, rb);
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[akv + j][arow] = ra[j];
    if (b_thr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[bkv + j][brow] = rb[j];
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += GBK) {
    stash();
    __syncthreads();
    if (k0 + GBK < K) fetch(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[8], bb[4];
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = p.bias[n + j];
  }
  T* Y = (T*)p.Y;
  const T* R = (const T*)p.R;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    float v[4];
    if (p.ln_stats) {
      float su = 0.f, sq = 0.f;
      const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + (b * p.M + m) * p.ln_npart;
      for (int q = 0; q < p.ln_npart; ++q) { su += st[q].x; sq += st[q].y; }
      const float mu = su / (float)p.ln_C;
      const float rs = rsqrtf(fmaxf(sq / (float)p.ln_C - mu * mu, 0.f) + p.ln_eps);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = rs * (acc[i][j] - mu * p.ln_cs[n + j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (p.act == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = lrelu_f(v[j]);
    } else if (p.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (p.act == ACT_TANH_RES) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = 0.2f * tanhf(v[j]);
    }
    if (p.omode == OMODE_ROWS) {
      if (R) {
        float r[4];
        load4(R + ((b * p.M + m) * p.ldr + n), r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += r[j];
      }
      store4(Y + ((b * p.M + m) * p.ldy + n), v);
    } else if (p.omode == OMODE_CONVT) {
      // N = 4*Co, n = (2i+j)*Co + co ; rows m = y*W + x at (H,W) -> out [B,2H,2W,Co]
      const int Co = p.N >> 2;
      const int ij = n / Co, co = n - ij * Co;
      const int y = m / p.W, x = m - y * p.W;
      const i64 orow = (b * 2 * p.H + 2 * y + (ij >> 1)) * (i64)(2 * p.W) + 2 * x + (ij & 1);
      store4(Y + (orow * p.ldy + co), v);
    } else {
      // pixel-unshuffle scatter: rows at (H,W), N = C/2 -> out [B,H/2,W/2,4N], channel = n*4 + 2*(y&1) + (x&1)
      const int y = m / p.W, x = m - y * p.W;
      const i64 orow = (b * (p.H >> 1) + (y >> 1)) * (i64)(p.W >> 1) + (x >> 1);
      T* o = Y + orow * p.ldy + 2 * (y & 1) + (x & 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) from_f(o[(n + j) * 4], v[j]);
    }
  }
}

void launch_gemm_cuda_core(Ctx& ctx, const GemmP& p) {
  const int K = p.K1 + p.K2;
  double es = (double)esize(ctx.dtype);
  double rows = (double)p.B * p.M;
  double abytes = rows * (p.amode == AMODE_CONV3 ? p.K1 / 9 : K) * es;
  double bytes = abytes + rows * p.N * es * (p.R ? 2.0 : 1.0) + (double)p.N * K * es * (p.w_img ? p.B : 1);
  ScopedLaunch sl(p.kernel_id, bytes, 2.0 * rows * p.N * K);
  dim3 grid(cdiv(p.M, GBM), cdiv(p.N, GBN), p.B);
  if (ctx.dtype == RF_BF16)
    k_gemm<bf16><<<grid, GTHREADS, 0, ctx.stream>>>(p);
  else
    k_gemm<float><<<grid, GTHREADS, 0, ctx.stream>>>(p);
}

// Implemented in rf_tc_gemm.cu: returns true when it launched a tcgen05 kernel for this problem.
// returns the number of stats_out partials per row it wrote (0 = stats_out not handled), < 0 when it did not launch
int launch_gemm_tcgen05(Ctx& ctx, const GemmP& p);

int launch_gemm(Ctx& ctx, const GemmP& p) {
  if (ctx.dry || p.M <= 0 || p.N <= 0 || p.B <= 0) return p.stats_out ? 1 : 0;
  if (ctx.dtype == RF_BF16) {
    const int np = launch_gemm_tcgen05(ctx, p);
    if (np > 0 || (np == 0 && !p.stats_out)) return np;
    if (np == 0) {  // the kernel ran but its epilogue could not emit the row statistics: one extra pass over Y
      launch_row_stats(ctx, p.Y, p.stats_out, (i64)p.B * p.M, p.N);
      return 1;
    }
  }
  if (p.omode == OMODE_ATOMIC_F32) {  // only the tcgen05 kernel implements the split-K atomic epilogue
    recorder().last_cuda_error = (int)cudaErrorNotSupported;
    return 0;
  }
  launch_gemm_cuda_core(ctx, p);
  if (p.stats_out) {
    launch_row_stats(ctx, p.Y, p.stats_out, (i64)p.B * p.M, p.N);
    return 1;
  }
  return 0;
}

}  // namespace rf
