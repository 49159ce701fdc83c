// Callers either side of the forward and the WFB gated-GELU FFN.
#include "rf_kernels.cuh"

namespace rf {

// test.py:117-118 -- clamp(0,1) * 255 -> uint8 (C-style truncation, like numpy astype) -> HWC
__global__ void __launch_bounds__(256) k_post_u8(const float* __restrict__ in, unsigned char* __restrict__ out, i64 hw) {
  const i64 b = blockIdx.y;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    unsigned char v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = in[(b * 3 + c) * hw + i];
      x = fminf(fmaxf(x, 0.f), 1.f);           // torch.clamp (NaN propagates; cast of NaN is 0 here)
      v[c] = (unsigned char)(int)__fmul_rn(x, 255.f);
    }
    unsigned char* o = out + (b * hw + i) * 3;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  }
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// test.py:17-40,117-120 -- the same conversion followed by correct_bayer_channels (out channel k = channel perm[k]) and the
// statistics of auto_correct_rb: sums[b] = {sum of out channel 0, sum of out channel 2} (exact integer sums; comparing them
// is comparing the means).  SRC_U8: the input already is a uint8 HWC image (the ground-truth side, test.py:111-113; in place).
template <bool SRC_U8>
__global__ void __launch_bounds__(256)
k_post_rgb_u8(const float* __restrict__ in, unsigned char* out, i64 hw, int p0, int p1, int p2, unsigned long long* sums) {
  const i64 b = blockIdx.y;
  unsigned long long s0 = 0, s2 = 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    unsigned char v[3];
    unsigned char* o = out + (b * hw + i) * 3;
    if (SRC_U8) {
      v[0] = o[0]; v[1] = o[1]; v[2] = o[2];
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float x = in[(b * 3 + c) * hw + i];
        x = fminf(fmaxf(x, 0.f), 1.f);
        v[c] = (unsigned char)(int)__fmul_rn(x, 255.f);
      }
    }
    const unsigned char r = v[p0], g = v[p1], bl = v[p2];
    o[0] = r; o[1] = g; o[2] = bl;
    s0 += r;
    s2 += bl;
  }
  s0 = warp_sum_u64(s0);
  s2 = warp_sum_u64(s2);
  if ((threadIdx.x & 31) == 0 && sums != nullptr) {
    atomicAdd(sums + b * 2, s0);
    atomicAdd(sums + b * 2 + 1, s2);
  }
}

// auto_correct_rb (test.py:29-38): swap R and B of every image whose red mean is below its blue mean
__global__ void __launch_bounds__(256) k_swap_rb_if(unsigned char* img, i64 hw, const unsigned long long* __restrict__ sums) {
  const i64 b = blockIdx.y;
  if (!(sums[b * 2] < sums[b * 2 + 1])) return;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (i64)gridDim.x * blockDim.x) {
    unsigned char* o = img + (b * hw + i) * 3;
    const unsigned char t = o[0];
    o[0] = o[2];
    o[2] = t;
  }
}

// sum of squared differences of two uint8 images (exact): the mean_squared_error inside skimage's PSNR (test.py:123)
__global__ void __launch_bounds__(256)
k_sse_u8(const unsigned char* __restrict__ a, const unsigned char* __restrict__ bimg, i64 n, unsigned long long* sse) {
  const i64 b = blockIdx.y;
  unsigned long long s = 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const int d = (int)a[b * n + i] - (int)bimg[b * n + i];
    s += (unsigned)(d * d);
  }
  s = warp_sum_u64(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(sse + b, s);
}

// skimage.metrics.structural_similarity(a, b, channel_axis=-1) as called in test.py:124 on uint8 HWC images: 7x7 uniform
// window, data_range 255, K1 = 0.01, K2 = 0.03, sample covariance (x 49/48), mean of S over the pixels whose window lies
// inside the image (crop by 3) and over the channels.  The five window sums are exact integers; S is evaluated in double.
// sum_out[b] += sum of S over the valid pixels and the 3 channels (the host divides by 3*(H-6)*(W-6)).
constexpr int SS_TW = 32, SS_TH = 16, SS_R = 3;
__global__ void __launch_bounds__(SS_TW * SS_TH)
k_ssim_u8(const unsigned char* __restrict__ a, const unsigned char* __restrict__ bimg, double* sum_out, int H, int W) {
  __shared__ unsigned char sa[SS_TH + 2 * SS_R][(SS_TW + 2 * SS_R) * 3];
  __shared__ unsigned char sb[SS_TH + 2 * SS_R][(SS_TW + 2 * SS_R) * 3];
  __shared__ double swarp[SS_TW * SS_TH / 32];
  const i64 img = (i64)blockIdx.z * H * W * 3;
  // outputs of this block: centres (y, x) with y in [y0, y0 + TH), x in [x0, x0 + TW); valid centres are [3, H-3) x [3, W-3)
  const int x0 = SS_R + blockIdx.x * SS_TW, y0 = SS_R + blockIdx.y * SS_TH;
  const int tid = threadIdx.y * SS_TW + threadIdx.x;
  const int row_bytes = (SS_TW + 2 * SS_R) * 3;
  for (int i = tid; i < (SS_TH + 2 * SS_R) * row_bytes; i += SS_TW * SS_TH) {
    const int r = i / row_bytes, cb = i - r * row_bytes;
    const int gy = y0 - SS_R + r, gxb = (x0 - SS_R) * 3 + cb;      // byte column inside the image row
    unsigned char va = 0, vb = 0;
    if (gy < H && gxb < W * 3) {
      va = a[img + (i64)gy * W * 3 + gxb];
      vb = bimg[img + (i64)gy * W * 3 + gxb];
    }
    sa[r][cb] = va;
    sb[r][cb] = vb;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  double acc = 0.0;
  if (x < W - SS_R && y < H - SS_R) {
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0), cov_norm = 49.0 / 48.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int s_a = 0, s_b = 0, s_aa = 0, s_bb = 0, s_ab = 0;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const int pa = sa[threadIdx.y + dy][(threadIdx.x + dx) * 3 + c], pb = sb[threadIdx.y + dy][(threadIdx.x + dx) * 3 + c];
          s_a += pa; s_b += pb; s_aa += pa * pa; s_bb += pb * pb; s_ab += pa * pb;
        }
      }
      const double ux = s_a / 49.0, uy = s_b / 49.0, uxx = s_aa / 49.0, uyy = s_bb / 49.0, uxy = s_ab / 49.0;
      const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
      const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
      acc += (A1 * A2) / (B1 * B2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) swarp[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < SS_TW * SS_TH / 32; ++i) t += swarp[i];
    atomicAdd(sum_out + blockIdx.z, t);
  }
}

// WFB/load_dataset.py:88-89 (clip to [black, white], scale by the exposure ratio) and, with `clamp`, correctdataloader.py:103
// (min(., 1)).  8 pixels per thread (one 16-byte load, two 16-byte stores) when n % 8 == 0.
__device__ __forceinline__ float pre_u16_one(unsigned v, float black, float white, float denom, float ratio, int clamp) {
  float x = fminf(fmaxf((float)v, black), white);
  x = __fmul_rn(__fdiv_rn(__fsub_rn(x, black), denom), ratio);
  return clamp ? fminf(x, 1.0f) : x;
}
__global__ void __launch_bounds__(256)
k_pre_u16(const unsigned short* __restrict__ raw, float* __restrict__ out, float black, float white, float denom, float ratio,
          int clamp, i64 n) {
  const i64 n8 = (n & 7) == 0 ? n >> 3 : 0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (i64)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(raw) + i);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = pre_u16_one(w[k] & 0xffffu, black, white, denom, ratio, clamp);
      f[2 * k + 1] = pre_u16_one(w[k] >> 16, black, white, denom, ratio, clamp);
    }
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  for (i64 i = n8 * 8 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    out[i] = pre_u16_one(raw[i], black, white, denom, ratio, clamp);
}

// gated-GELU core: x1 = dwA(t)+bA, x2 = dwB(t)+bB, g = gelu(x2)*x1 + gelu(x1)*x2     (WFB/model.py:61-63)
template <typename T>
__global__ void __launch_bounds__(256)
k_gated_dw(const T* __restrict__ t, const float* __restrict__ wA, const float* __restrict__ bA, const float* __restrict__ wB,
           const float* __restrict__ bB, T* __restrict__ out, int H, int W, int Cn) {
  const i64 b = blockIdx.y;
  const int cv = Cn >> 3;
  const i64 total = (i64)H * W * cv;
  const T* img = t + b * (i64)H * W * Cn;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    const i64 p = i / cv;
    const int y = (int)(p / W), x = (int)(p % W);
    float a1[8], a2[8];
    load8(bA + c0, a1);
    load8(bB + c0, a2);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        float v[8], ka[8], kb[8];
        load8(img + ((i64)yy * W + xx) * Cn + c0, v);
        load8(wA + ((dy + 1) * 3 + dx + 1) * Cn + c0, ka);
        load8(wB + ((dy + 1) * 3 + dx + 1) * Cn + c0, kb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] = fmaf(v[j], ka[j], a1[j]);
          a2[j] = fmaf(v[j], kb[j], a2[j]);
        }
      }
    }
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = gelu_erf_f(a2[j]) * a1[j] + gelu_erf_f(a1[j]) * a2[j];
    store8(out + (b * (i64)H * W + p) * Cn + c0, g);
  }
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_postprocess_u8(const float* in, unsigned char* out, int B, int H, int W, void* stream) {
  if (!in || !out) return RF_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0) return (B < 0 || H < 0 || W < 0) ? RF_ERR_BAD_SHAPE : RF_OK;
  i64 hw = (i64)H * W;
  unsigned gx = (unsigned)(cdivl(hw, 256) < 8 * num_sms() ? cdivl(hw, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_INDEX_OP, 15.0 * B * hw);
  k_post_u8<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(in, out, hw);
  return check_cuda(cudaGetLastError());
}

static int post_rgb(const float* in, unsigned char* img, const int* perm_host, int auto_rb, int B, int H, int W,
                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!img || !perm_host) return RF_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0) return (B < 0 || H < 0 || W < 0) ? RF_ERR_BAD_SHAPE : RF_OK;
  int seen = 0;
  for (int k = 0; k < 3; ++k) {
    if (perm_host[k] < 0 || perm_host[k] > 2) return RF_ERR_BAD_ARG;
    seen |= 1 << perm_host[k];
  }
  if (seen != 7) return RF_ERR_BAD_ARG;
  unsigned long long* sums = nullptr;
  if (auto_rb) {
    if (!workspace || (uintptr_t)workspace % 8) return RF_ERR_BAD_ARG;
    if (workspace_bytes < (size_t)B * 16) return RF_ERR_WORKSPACE;
    sums = (unsigned long long*)workspace;
    RF_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * 16, (cudaStream_t)stream));
  }
  const i64 hw = (i64)H * W;
  const unsigned gx = (unsigned)(cdivl(hw, 256) < 8 * num_sms() ? cdivl(hw, 256) : 8 * num_sms());
  {
    ScopedLaunch sl(RF_K_INDEX_OP, (in ? 15.0 : 6.0) * B * hw);
    if (in)
      k_post_rgb_u8<false><<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(in, img, hw, perm_host[0], perm_host[1], perm_host[2], sums);
    else
      k_post_rgb_u8<true><<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(nullptr, img, hw, perm_host[0], perm_host[1], perm_host[2], sums);
  }
  if (auto_rb) {
    ScopedLaunch sl(RF_K_INDEX_OP, 6.0 * B * hw);
    k_swap_rb_if<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(img, hw, sums);
  }
  return check_cuda(cudaGetLastError());
}

int rf_postprocess_rgb_u8(const float* in, unsigned char* out, const int* perm_host, int auto_rb, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!in) return RF_ERR_BAD_ARG;
  return post_rgb(in, out, perm_host, auto_rb, B, H, W, workspace, workspace_bytes, stream);
}

int rf_correct_rgb_u8(unsigned char* img, const int* perm_host, int auto_rb, int B, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return post_rgb(nullptr, img, perm_host, auto_rb, B, H, W, workspace, workspace_bytes, stream);
}

int rf_sse_u8(const unsigned char* a, const unsigned char* b, unsigned long long* sse, int B, long long n_per_image,
              void* stream) {
  if (!a || !b || !sse || (uintptr_t)sse % 8) return RF_ERR_BAD_ARG;
  if (B < 0 || n_per_image < 0) return RF_ERR_BAD_SHAPE;
  if (B == 0) return RF_OK;
  RF_CUDA(cudaMemsetAsync(sse, 0, (size_t)B * 8, (cudaStream_t)stream));
  if (n_per_image == 0) return RF_OK;
  const unsigned gx = (unsigned)(cdivl(n_per_image, 1024) < 8 * num_sms() ? cdivl(n_per_image, 1024) : 8 * num_sms());
  ScopedLaunch sl(RF_K_INDEX_OP, 2.0 * B * n_per_image);
  k_sse_u8<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(a, b, n_per_image, sse);
  return check_cuda(cudaGetLastError());
}

int rf_ssim_u8(const unsigned char* a, const unsigned char* b, double* sum_out, int B, int H, int W, void* stream) {
  if (!a || !b || !sum_out || (uintptr_t)sum_out % 8) return RF_ERR_BAD_ARG;
  if (B < 0) return RF_ERR_BAD_SHAPE;
  if (B == 0) return RF_OK;
  if (H < 7 || W < 7) return RF_ERR_BAD_SHAPE;          // the 7x7 window must fit (skimage raises for smaller images)
  if (B > 65535) return RF_ERR_UNSUPPORTED;
  RF_CUDA(cudaMemsetAsync(sum_out, 0, (size_t)B * 8, (cudaStream_t)stream));
  const dim3 grid(cdiv(W - 6, SS_TW), cdiv(H - 6, SS_TH), B);
  if (grid.y > 65535) return RF_ERR_UNSUPPORTED;
  ScopedLaunch sl(RF_K_INDEX_OP, 6.0 * B * H * W);
  k_ssim_u8<<<grid, dim3(SS_TW, SS_TH), 0, (cudaStream_t)stream>>>(a, b, sum_out, H, W);
  return check_cuda(cudaGetLastError());
}

int rf_preprocess_u16(const unsigned short* raw, float* out, float black, float white, float ratio, int clamp, int B, int H,
                      int W, void* stream) {
  if (!raw || !out || (uintptr_t)raw % 16 || (uintptr_t)out % 16) return RF_ERR_BAD_ARG;
  if (B <= 0 || H <= 0 || W <= 0) return (B < 0 || H < 0 || W < 0) ? RF_ERR_BAD_SHAPE : RF_OK;
  i64 n = (i64)B * H * W;
  const float denom = (float)((double)white - (double)black + 1e-6);
  const i64 work = (n & 7) == 0 ? n >> 3 : n;
  unsigned gx = (unsigned)(cdivl(work, 256) < 8 * num_sms() ? cdivl(work, 256) : 8 * num_sms());
  ScopedLaunch sl(RF_K_INDEX_OP, 6.0 * n);
  k_pre_u16<<<gx, 256, 0, (cudaStream_t)stream>>>(raw, out, black, white, denom, ratio, clamp ? 1 : 0, n);
  return check_cuda(cudaGetLastError());
}

int rf_feedforward_gated(const float* project_in_w, const float* project_in_b, const float* dwA_w, const float* dwA_b,
                         const float* dwB_w, const float* dwB_b, const float* project_out_w, const float* project_out_b,
                         int C, int hidden, int dtype, const float* x, float* out, int B, int H, int W, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!project_in_w || !dwA_w || !dwB_w || !project_out_w || !x || !out || !workspace) return RF_ERR_BAD_ARG;
  if (dtype != RF_F32 && dtype != RF_BF16) return RF_ERR_BAD_ARG;
  if (C <= 0 || C % 8 || hidden <= 0 || hidden % 8 || B <= 0 || H <= 0 || W <= 0) return RF_ERR_BAD_SHAPE;
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  ctx.dtype = dtype;
  ctx.arena.base = (char*)workspace;
  ctx.arena.cap = workspace_bytes;
  if ((uintptr_t)workspace % 256) {
    size_t adj = 256 - (uintptr_t)workspace % 256;
    ctx.arena.base += adj;
    ctx.arena.cap = workspace_bytes > adj ? workspace_bytes - adj : 0;
  }
  recorder().last_cuda_error = 0;
  Arena& A = ctx.arena;
  const i64 P = (i64)H * W;
  void* w_in = A.elems((size_t)hidden * C, dtype);
  void* w_out = A.elems((size_t)C * hidden, dtype);
  float* b_in = A.get<float>(hidden);
  float* b_out = A.get<float>(C);
  float* wA = A.get<float>(9 * (size_t)hidden);
  float* wB = A.get<float>(9 * (size_t)hidden);
  float* bA = A.get<float>(hidden);
  float* bB = A.get<float>(hidden);
  void* xT = A.elems((size_t)B * P * C, dtype);
  void* t = A.elems((size_t)B * P * hidden, dtype);
  void* g = A.elems((size_t)B * P * hidden, dtype);
  void* oT = A.elems((size_t)B * P * C, dtype);
  if (!ctx.fits()) return RF_ERR_WORKSPACE;
  launch_pack3(ctx, project_in_w, w_in, dtype, 1, 1, hidden * C, 0, 0, 1, 0, 0, 1, 0);
  launch_pack3(ctx, project_out_w, w_out, dtype, 1, 1, hidden * C, 0, 0, 1, 0, 0, 1, 0);
  launch_fill_f32(ctx, b_in, 0.f, hidden);
  launch_fill_f32(ctx, b_out, 0.f, C);
  launch_fill_f32(ctx, bA, 0.f, hidden);
  launch_fill_f32(ctx, bB, 0.f, hidden);
  if (project_in_b) launch_pack3(ctx, project_in_b, b_in, RF_F32, 1, 1, hidden, 0, 0, 1, 0, 0, 1, 0);
  if (project_out_b) launch_pack3(ctx, project_out_b, b_out, RF_F32, 1, 1, C, 0, 0, 1, 0, 0, 1, 0);
  if (dwA_b) launch_pack3(ctx, dwA_b, bA, RF_F32, 1, 1, hidden, 0, 0, 1, 0, 0, 1, 0);
  if (dwB_b) launch_pack3(ctx, dwB_b, bB, RF_F32, 1, 1, hidden, 0, 0, 1, 0, 0, 1, 0);
  launch_pack3(ctx, dwA_w, wA, RF_F32, 1, 9, hidden, 0, 1, 9, 0, hidden, 1, 0);
  launch_pack3(ctx, dwB_w, wB, RF_F32, 1, 9, hidden, 0, 1, 9, 0, hidden, 1, 0);
  launch_nchw_to_nhwc(ctx, x, xT, B, C, P);
  GemmP g1;
  g1.A1 = xT; g1.K1 = C; g1.lda1 = C; g1.Wt = w_in; g1.bias = b_in; g1.Y = t; g1.ldy = hidden;
  g1.M = (int)P; g1.N = hidden; g1.B = B; g1.kernel_id = RF_K_GEMM_PW1;
  launch_gemm(ctx, g1);
  {
    i64 total = P * (hidden / 8);
    unsigned gx = (unsigned)(cdivl(total, 256) < 16 * num_sms() ? cdivl(total, 256) : 16 * num_sms());
    ScopedLaunch sl(RF_K_DW_GELU, 2.0 * B * P * hidden * esize(dtype), 36.0 * B * P * hidden);
    if (dtype == RF_BF16)
      k_gated_dw<bf16><<<dim3(gx, B), 256, 0, ctx.stream>>>((const bf16*)t, wA, bA, wB, bB, (bf16*)g, H, W, hidden);
    else
      k_gated_dw<float><<<dim3(gx, B), 256, 0, ctx.stream>>>((const float*)t, wA, bA, wB, bB, (float*)g, H, W, hidden);
  }
  GemmP g2;
  g2.A1 = g; g2.K1 = hidden; g2.lda1 = hidden; g2.Wt = w_out; g2.bias = b_out; g2.Y = oT; g2.ldy = C;
  g2.R = xT; g2.ldr = C;
  g2.M = (int)P; g2.N = C; g2.B = B; g2.kernel_id = RF_K_GEMM_PW2;
  launch_gemm(ctx, g2);
  launch_nhwc_to_nchw(ctx, oT, out, B, C, P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e);
  return recorder().last_cuda_error ? RF_ERR_CUDA : RF_OK;
}

}  // extern "C"
