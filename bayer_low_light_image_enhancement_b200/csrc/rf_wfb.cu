// WFB "WMB" block pieces: the Illumination_Estimator (RawFomer_WFB_FFAB/model.py:174-200) and the
// rFFT amplitude / phase blocks FEB / ProcessBlock / FFAB (RawFomer_WFB_FFAB/blocks.py:11-92) as fp32 NCHW CUDA kernels
// behind plain C-ABI calls.  The module mirrors (wfb.py) compose them exactly as the reference's forward does.
//
// 2-d real FFT, norm = 'ortho', any H x W (the LL band of a SID Sony frame is 712 x 1064 = 2^3*89 x 2^3*7*19): the
// transform is evaluated as dense DFT matrix products in fp32 -- rows: X[H x W] * T[W x Wf] (real -> Wf = W/2+1 complex
// bins, real and imaginary parts as two planes), columns: one [2H x 2H] real matrix [[C, S], [-S, C]] applied to the
// stacked (re; im) planes.  The twiddle matrices are made on the device in double precision with exact argument reduction
// ((k*n) mod N before sincospi) and cached by the caller as a "plan".  The inverse ignores the imaginary parts of the DC
// and Nyquist bins exactly like a complex-to-real FFT does (their rows in the synthesis matrix are zero).
#include <math.h>

#include "rf_kernels.cuh"

namespace rf {

// ---------------------------------------------------------------------------------------------
// batched SGEMM: C[z] (+)= A[z] * B[z], row-major, 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread
// ---------------------------------------------------------------------------------------------
struct SgemmP {
  const float* A; const float* B; float* C;
  int M, N, K;
  i64 lda, ldb, ldc, sA, sB, sC;
  int accumulate;
};

__global__ void __launch_bounds__(256)
k_sgemm(const SgemmP p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int z = blockIdx.z;
  const float* A = p.A + (i64)z * p.sA;
  const float* B = p.B + (i64)z * p.sB;
  float* C = p.C + (i64)z * p.sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    // A tile 64 x 16 (transposed into As[k][m]); B tile 16 x 64
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + it * 256;
      const int am = e >> 4, ak = e & 15;
      const int gm = m0 + am, gk = k0 + ak;
      As[ak][am] = (gm < p.M && gk < p.K) ? __ldg(A + (i64)gm * p.lda + gk) : 0.f;
      const int bk = e >> 6, bn = e & 63;
      const int gk2 = k0 + bk, gn = n0 + bn;
      Bs[bk][bn] = (gk2 < p.K && gn < p.N) ? __ldg(B + (i64)gk2 * p.ldb + gn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float* c = C + (i64)gm * p.ldc + gn;
      *c = p.accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

static void sgemm(cudaStream_t st, const float* A, const float* B, float* C, int M, int N, int K, i64 lda, i64 ldb, i64 ldc, i64 sA,
                  i64 sB, i64 sC, int batch, int accumulate) {
  SgemmP p{A, B, C, M, N, K, lda, ldb, ldc, sA, sB, sC, accumulate};
  for (int z0 = 0; z0 < batch; z0 += 65535) {           // gridDim.z limit
    const int nz = batch - z0 < 65535 ? batch - z0 : 65535;
    SgemmP q = p;
    q.A += (i64)z0 * sA; q.B += (i64)z0 * sB; q.C += (i64)z0 * sC;
    ScopedLaunch sl(RF_K_INDEX_OP);
    k_sgemm<<<dim3(cdiv(N, 64), cdiv(M, 64), nz), 256, 0, st>>>(q);
  }
}

// ---------------------------------------------------------------------------------------------
// DFT plan: [Tre W x Wf | Tim W x Wf | G 2H x 2H | Gi 2H x 2H | Cre Wf x W | Cim Wf x W]
// ---------------------------------------------------------------------------------------------
struct PlanOff { i64 tre, tim, g, gi, cre, cim, total; };
static PlanOff plan_off(int H, int W) {
  const i64 Wf = W / 2 + 1;
  PlanOff o;
  o.tre = 0; o.tim = o.tre + (i64)W * Wf; o.g = o.tim + (i64)W * Wf; o.gi = o.g + (i64)4 * H * H; o.cre = o.gi + (i64)4 * H * H;
  o.cim = o.cre + Wf * W; o.total = o.cim + Wf * W;
  return o;
}
__device__ __forceinline__ void unit_root(i64 a, i64 b, int n, double& c, double& s) {   // cos, sin of 2 pi a b / n
  const i64 r = (a * b) % n;
  sincospi(2.0 * (double)r / (double)n, &s, &c);
}
__global__ void k_dft_plan(float* plan, PlanOff o, int H, int W) {
  const i64 Wf = W / 2 + 1;
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= o.total) return;
  const double sw = 1.0 / sqrt((double)W), sh = 1.0 / sqrt((double)H);
  double c, s;
  if (i < o.g) {                                    // Tre / Tim [w][f]
    const i64 j = i < o.tim ? i : i - o.tim;
    const i64 w = j / Wf, f = j % Wf;
    unit_root(w, f, W, c, s);
    plan[i] = (float)(i < o.tim ? c * sw : -s * sw);
  } else if (i < o.cre) {                           // G / Gi [part_out*H + k][part_in*H + h]
    const bool inv = i >= o.gi;
    const i64 j = inv ? i - o.gi : i - o.g;
    const i64 row = j / (2 * H), col = j % (2 * H);
    const int po = row >= H, pi = col >= H;
    const i64 k = row - (i64)po * H, h = col - (i64)pi * H;
    unit_root(k, h, H, c, s);
    double v;
    if (po == pi) v = c;
    else if (po == 0) v = inv ? -s : s;             // forward: Zre += s*Yim; inverse: Yre -= s*Zim
    else v = inv ? s : -s;
    plan[i] = (float)(v * sh);
  } else {                                          // Cre / Cim [f][w]
    const bool im = i >= o.cim;
    const i64 j = im ? i - o.cim : i - o.cre;
    const i64 f = j / W, w = j % W;
    unit_root(f, w, W, c, s);
    const double cf = (f == 0 || (W % 2 == 0 && f == W / 2)) ? 1.0 : 2.0;
    plan[i] = (float)(im ? -cf * s * sw : cf * c * sw);
  }
}

// ---------------------------------------------------------------------------------------------
// spectrum <-> (magnitude, phase)      blocks.py:28-34
// ---------------------------------------------------------------------------------------------
__global__ void k_spec_abs_angle(const float* __restrict__ spec, float* __restrict__ mag, float* __restrict__ pha, i64 BC, int H,
                                 int Wf, int W) {
  const i64 n = (i64)H * Wf;
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * n) return;
  const i64 bc = i / n, r = i % n;
  const int k = (int)(r / Wf), f = (int)(r % Wf);
  const float re = spec[(bc * 2) * n + r];
  float im = spec[(bc * 2 + 1) * n + r];
  // the four self-conjugate bins of a real input are real: a real-to-complex FFT returns imag = +0 there, a DFT by
  // matrix products some +-0 / rounding residue whose sign would flip the phase between +pi and -pi
  const bool selfconj = (f == 0 || (W % 2 == 0 && f == W / 2)) && (k == 0 || (H % 2 == 0 && k == H / 2));
  if (selfconj) im = 0.f;
  mag[i] = hypotf(re, im) + 1e-6f;
  pha[i] = atan2f(im, re);
}
__global__ void k_spec_polar(const float* __restrict__ mag, const float* __restrict__ pha, float* __restrict__ spec, i64 BC, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * n) return;
  const i64 bc = i / n, r = i % n;
  float s, c;
  sincosf(pha[i], &s, &c);
  spec[(bc * 2) * n + r] = mag[i] * c;
  spec[(bc * 2 + 1) * n + r] = mag[i] * s;
}

// ---------------------------------------------------------------------------------------------
// nn.Conv2d(Cin (+Cin2), Cout, 1) on NCHW with the clamps / activation / residual the WFB blocks put around it
// ---------------------------------------------------------------------------------------------
constexpr int C1_NB = 16;   // outputs per pass
__global__ void __launch_bounds__(128)
k_conv1x1_nchw(const float* __restrict__ in, const float* __restrict__ in2, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ resid, float* __restrict__ out, int Cin, int Cin2, int Cout, float in_clamp, int act,
               float out_lo, float out_hi, i64 P) {
  extern __shared__ float ws[];                     // [Cin + Cin2][C1_NB] for the current output chunk
  const int b = blockIdx.y;
  const i64 pix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ct = Cin + Cin2;
  const float* x1 = in + (i64)b * Cin * P;
  const float* x2 = in2 ? in2 + (i64)b * Cin2 * P : nullptr;
  for (int n0 = 0; n0 < Cout; n0 += C1_NB) {
    __syncthreads();
    for (int e = threadIdx.x; e < Ct * C1_NB; e += blockDim.x) {
      const int c = e / C1_NB, j = e % C1_NB;
      ws[e] = n0 + j < Cout ? __ldg(w + (i64)(n0 + j) * Ct + c) : 0.f;
    }
    __syncthreads();
    if (pix >= P) continue;
    float acc[C1_NB];
#pragma unroll
    for (int j = 0; j < C1_NB; ++j) acc[j] = 0.f;
    for (int c = 0; c < Ct; ++c) {
      float v = c < Cin ? __ldg(x1 + (i64)c * P + pix) : __ldg(x2 + (i64)(c - Cin) * P + pix);
      if (in_clamp > 0.f) v = fminf(fmaxf(v, -in_clamp), in_clamp);
      const float4* wr = reinterpret_cast<const float4*>(ws + c * C1_NB);
#pragma unroll
      for (int j4 = 0; j4 < C1_NB / 4; ++j4) {
        const float4 w4 = wr[j4];
        acc[j4 * 4 + 0] = fmaf(w4.x, v, acc[j4 * 4 + 0]);
        acc[j4 * 4 + 1] = fmaf(w4.y, v, acc[j4 * 4 + 1]);
        acc[j4 * 4 + 2] = fmaf(w4.z, v, acc[j4 * 4 + 2]);
        acc[j4 * 4 + 3] = fmaf(w4.w, v, acc[j4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < C1_NB; ++j) {
      const int n = n0 + j;
      if (n >= Cout) break;
      float v = acc[j] + (bias ? __ldg(bias + n) : 0.f);
      if (act == 1) v = v >= 0.f ? v : 0.1f * v;
      if (out_lo < out_hi) v = fminf(fmaxf(v, out_lo), out_hi);
      const i64 o = ((i64)b * Cout + n) * P + pix;
      if (resid) v += __ldg(resid + o);
      out[o] = v;
    }
  }
}

// out = clamp(a + clamp(b, -lim, lim), -lim, lim)      (FEB tail, blocks.py:24,36-37)
__global__ void k_add_clamp(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, float lim, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = fminf(fmaxf(b[i], -lim), lim);
  out[i] = fminf(fmaxf(a[i] + x, -lim), lim);
}
// mean over the channels: in [B,C,P] -> out [B,1,P]      (model.py:192)
__global__ void k_channel_mean(const float* __restrict__ in, float* __restrict__ out, int C, i64 P) {
  const int b = blockIdx.y;
  const i64 pix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __ldg(in + ((i64)b * C + c) * P + pix);
  out[(i64)b * P + pix] = s / (float)C;
}
// nn.Conv2d(C, C, 5, padding=2, groups=C) on NCHW      (model.py:181-182)
__global__ void __launch_bounds__(256)
k_dw5x5_nchw(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int C,
             int H, int W) {
  const int c = blockIdx.y, b = blockIdx.z;
  const i64 pix = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (i64)H * W) return;
  const int y = (int)(pix / W), x = (int)(pix % W);
  const float* src = in + ((i64)b * C + c) * H * W;
  const float* wc = w + (i64)c * 25;
  float acc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
    const int yy = y + ky - 2;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int xx = x + kx - 2;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(__ldg(wc + ky * 5 + kx), __ldg(src + (i64)yy * W + xx), acc);
    }
  }
  out[((i64)b * C + c) * H * W + pix] = acc;
}

}  // namespace rf

using namespace rf;

extern "C" {

size_t rf_dft2_plan_floats(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  return (size_t)plan_off(H, W).total;
}

int rf_dft2_plan_init(float* plan, int H, int W, void* stream) {
  if (!plan || H <= 0 || W <= 0) return RF_ERR_BAD_ARG;
  const PlanOff o = plan_off(H, W);
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_dft_plan<<<(unsigned)cdivl(o.total, 256), 256, 0, (cudaStream_t)stream>>>(plan, o, H, W);
  return check_cuda(cudaGetLastError());
}

int rf_rfft2_ortho(const float* x, const float* plan, float* spec, float* tmp, long long BC, int H, int W, void* stream) {
  if (!x || !plan || !spec || !tmp || BC < 0 || H <= 0 || W <= 0 || BC > 0x7fffffff) return RF_ERR_BAD_ARG;
  if (BC == 0) return RF_OK;
  const PlanOff o = plan_off(H, W);
  const int Wf = W / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 n = (i64)H * Wf;
  // rows: tmp[bc][part][h][f] = x[bc][h][:] * T{re,im}
  sgemm(st, x, plan + o.tre, tmp, H, Wf, W, W, Wf, Wf, (i64)H * W, 0, 2 * n, (int)BC, 0);
  sgemm(st, x, plan + o.tim, tmp + n, H, Wf, W, W, Wf, Wf, (i64)H * W, 0, 2 * n, (int)BC, 0);
  // columns: spec[bc] (2H x Wf) = G (2H x 2H) * tmp[bc] (2H x Wf)
  sgemm(st, plan + o.g, tmp, spec, 2 * H, Wf, 2 * H, 2 * H, Wf, Wf, 0, 2 * n, 2 * n, (int)BC, 0);
  return check_cuda(cudaGetLastError());
}

int rf_irfft2_ortho(const float* spec, const float* plan, float* out, float* tmp, long long BC, int H, int W, void* stream) {
  if (!spec || !plan || !out || !tmp || BC < 0 || H <= 0 || W <= 0 || BC > 0x7fffffff) return RF_ERR_BAD_ARG;
  if (BC == 0) return RF_OK;
  const PlanOff o = plan_off(H, W);
  const int Wf = W / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const i64 n = (i64)H * Wf;
  sgemm(st, plan + o.gi, spec, tmp, 2 * H, Wf, 2 * H, 2 * H, Wf, Wf, 0, 2 * n, 2 * n, (int)BC, 0);
  sgemm(st, tmp, plan + o.cre, out, H, W, Wf, Wf, W, W, 2 * n, 0, (i64)H * W, (int)BC, 0);
  sgemm(st, tmp + n, plan + o.cim, out, H, W, Wf, Wf, W, W, 2 * n, 0, (i64)H * W, (int)BC, 1);
  return check_cuda(cudaGetLastError());
}

int rf_spec_abs_angle(const float* spec, float* mag, float* pha, long long BC, int H, int W, void* stream) {
  if (!spec || !mag || !pha || BC < 0 || H <= 0 || W <= 0) return RF_ERR_BAD_ARG;
  const int Wf = W / 2 + 1;
  const i64 total = BC * H * Wf;
  if (total == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_spec_abs_angle<<<(unsigned)cdivl(total, 256), 256, 0, (cudaStream_t)stream>>>(spec, mag, pha, BC, H, Wf, W);
  return check_cuda(cudaGetLastError());
}

int rf_spec_polar(const float* mag, const float* pha, float* spec, long long BC, int H, int W, void* stream) {
  if (!spec || !mag || !pha || BC < 0 || H <= 0 || W <= 0) return RF_ERR_BAD_ARG;
  const i64 n = (i64)H * (W / 2 + 1);
  if (BC * n == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_spec_polar<<<(unsigned)cdivl(BC * n, 256), 256, 0, (cudaStream_t)stream>>>(mag, pha, spec, BC, n);
  return check_cuda(cudaGetLastError());
}

int rf_conv1x1_nchw(const float* in, const float* in2, const float* weight, const float* bias, const float* resid, float* out, int Cin,
                    int Cin2, int Cout, float in_clamp, int act, float out_lo, float out_hi, int B, long long P, void* stream) {
  if (!in || !weight || !out || Cin <= 0 || Cin2 < 0 || (Cin2 > 0 && !in2) || Cout <= 0 || B < 0 || P < 0 || act < 0 || act > 1)
    return RF_ERR_BAD_ARG;
  if (B == 0 || P == 0) return RF_OK;
  const size_t smem = (size_t)(Cin + Cin2) * C1_NB * sizeof(float);
  if (smem > 48 * 1024) return RF_ERR_UNSUPPORTED;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_conv1x1_nchw<<<dim3((unsigned)cdivl(P, 128), B), 128, smem, (cudaStream_t)stream>>>(in, Cin2 > 0 ? in2 : nullptr, weight, bias,
                                                                                       resid, out, Cin, Cin2, Cout, in_clamp, act,
                                                                                       out_lo, out_hi, P);
  return check_cuda(cudaGetLastError());
}

int rf_add_clamp(const float* a, const float* b, float* out, float lim, long long n, void* stream) {
  if (!a || !b || !out || n < 0) return RF_ERR_BAD_ARG;
  if (n == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_add_clamp<<<(unsigned)cdivl(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, lim, n);
  return check_cuda(cudaGetLastError());
}

int rf_channel_mean(const float* in, float* out, int B, int C, long long P, void* stream) {
  if (!in || !out || B < 0 || C <= 0 || P < 0) return RF_ERR_BAD_ARG;
  if (B == 0 || P == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_channel_mean<<<dim3((unsigned)cdivl(P, 256), B), 256, 0, (cudaStream_t)stream>>>(in, out, C, P);
  return check_cuda(cudaGetLastError());
}

int rf_dwconv5x5_nchw(const float* in, const float* weight, const float* bias, float* out, int B, int C, int H, int W, void* stream) {
  if (!in || !weight || !out || B < 0 || C <= 0 || H <= 0 || W <= 0 || C > 65535 || B > 65535) return RF_ERR_BAD_ARG;
  if (B == 0) return RF_OK;
  ScopedLaunch sl(RF_K_INDEX_OP);
  k_dw5x5_nchw<<<dim3((unsigned)cdivl((i64)H * W, 256), C, B), 256, 0, (cudaStream_t)stream>>>(in, weight, bias, out, C, H, W);
  return check_cuda(cudaGetLastError());
}

}  // extern "C"
