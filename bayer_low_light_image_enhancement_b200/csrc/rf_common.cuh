// Shared device/host helpers for the RawFormer sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rawformer_b200.h"

namespace rf {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// launch bookkeeping (per host thread): count, optional per-launch CUDA events for profiling
// ---------------------------------------------------------------------------------------------
enum KernelId {
  RF_K_PACK_LUMA = 0,
  RF_K_LUMA_NORM,
  RF_K_DWT_HIGH,
  RF_K_GUIDANCE,
  RF_K_FLCA_MOD,
  RF_K_SE_FINALIZE,
  RF_K_LN_QKV,
  RF_K_DW_QKV_GRAM,
  RF_K_ATTN_FINALIZE,
  RF_K_PROJ_RESID,
  RF_K_LN_PW1,
  RF_K_DW_GELU_PW2,
  RF_K_CAT_REDUCE,
  RF_K_CONV3X3_OUT,
  RF_K_DOWN_CONV,
  RF_K_UP_CONVT,
  RF_K_SKIP_REDUCE,
  RF_K_EMBED,
  RF_K_HEAD,
  RF_K_LAYOUT,
  RF_K_WEIGHT_PACK,
  RF_K_MISC,
  RF_K_PYR_GATES,
  RF_K_PYR_SPATIAL,
  RF_K_PYR_RES1,
  RF_K_PYR_RES2,
  RF_K_TAIL_STATS,
  RF_K_TAIL_APPLY,
  RF_K_COUNT
};

struct LaunchRecorder {
  long long count = 0;
  bool profiling = false;
  int cap = 0;
  int n = 0;
  cudaEvent_t* ev = nullptr;  // 2 per launch
  int* ids = nullptr;
  double* bytes = nullptr;
  double* flops = nullptr;
  cudaStream_t stream = 0;
  int last_cuda_error = 0;
};
LaunchRecorder& recorder();

// Call before / after a kernel launch.
void launch_begin(int kernel_id, double algo_bytes, double algo_flops);
void launch_end();

struct ScopedLaunch {
  ScopedLaunch(int id, double bytes = 0, double flops = 0) { launch_begin(id, bytes, flops); }
  ~ScopedLaunch() { launch_end(); }
};

int check_cuda(cudaError_t e);

#define RF_CUDA(x)                          \
  do {                                      \
    int _s = rf::check_cuda((x));           \
    if (_s != RF_OK) return _s;             \
  } while (0)
#define RF_TRY(x)                 \
  do {                            \
    int _s = (x);                 \
    if (_s != RF_OK) return _s;   \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// element conversion helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int VEC = 4;  // elements per 16-byte vector
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <>
struct Elem<bf16> {
  static constexpr int VEC = 8;
  __device__ __forceinline__ static float to_f(bf16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static bf16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector load/store of T as floats
__device__ __forceinline__ void load_vec(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_vec(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store_vec(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
// pair store
__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float2 load_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load_pair(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// float max via integer atomics (handles negatives)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// PyTorch bilinear (align_corners=False) source taps for one axis
__device__ __forceinline__ void bilinear_taps(int dst, int n_in, int n_out, int& i0, int& i1, float& lam) {
  if (n_in == n_out) { i0 = dst; i1 = dst; lam = 0.f; return; }
  float scale = (float)n_in / (float)n_out;
  float src = fmaxf(((float)dst + 0.5f) * scale - 0.5f, 0.f);
  i0 = min((int)floorf(src), n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  lam = src - (float)i0;
}

}  // namespace rf
