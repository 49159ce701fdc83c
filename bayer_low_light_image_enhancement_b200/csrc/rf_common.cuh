// Shared host/device helpers for the RawFormer sm_100a kernels.
//
// Internal conventions:
//   * activations are NHWC ("pixel rows"): [B, H, W, C] of T, T = float (parity mode) or bf16;
//   * per-channel parameters, statistics and guidance maps are always fp32;
//   * every launcher takes a Ctx (stream + bump arena over the caller's workspace); in `dry` mode the
//     launchers do nothing, which is how the *_workspace_bytes queries size the arena.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/rawformer_b200.h"

namespace rf {

typedef __nv_bfloat16 bf16;
typedef long long i64;

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
  RF_K_FOLD_REDUCE,
  RF_K_LAYERNORM,
  RF_K_GEMM_QKV,
  RF_K_DW_QKV_GRAM,
  RF_K_ATTN_FINALIZE,
  RF_K_GEMM_PROJ,
  RF_K_GEMM_PW1,
  RF_K_DW_GELU,
  RF_K_GEMM_PW2,
  RF_K_GEMM_CAT_REDUCE,
  RF_K_CONV3X3_OUT,
  RF_K_DOWN_CONV,
  RF_K_UP_CONVT,
  RF_K_SKIP_REDUCE,
  RF_K_EMBED,
  RF_K_HEAD,
  RF_K_LAYOUT,
  RF_K_WEIGHT_PACK,
  RF_K_MISC,
  RF_K_PYR_SPATIAL,
  RF_K_GEMM_PYR_RES1,
  RF_K_GEMM_PYR_RES2,
  RF_K_CHANNEL_SUMS,
  RF_K_TAIL_STATS,
  RF_K_TAIL_APPLY,
  RF_K_INDEX_OP,
  RF_K_GEMM_GRAM,
  RF_K_BAND_HALO,
  RF_K_BAND_ALLREDUCE,
  RF_K_FFN_FUSED,
  RF_K_QKV_FUSED,
  RF_K_CONV3X3_LC,     // Conv_out through the k_lnconv pipeline (C = 32 / 64)
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
  // info of the last profiled forward (kept after profiling ends)
  int last_n = 0;
  int* last_ids = nullptr;
  double* last_bytes = nullptr;
  double* last_flops = nullptr;
};
LaunchRecorder& recorder();

void launch_begin(int kernel_id, double algo_bytes, double algo_flops);
void launch_end();

struct ScopedLaunch {
  ScopedLaunch(int id, double bytes = 0, double flops = 0) { launch_begin(id, bytes, flops); }
  ~ScopedLaunch() { launch_end(); }
};

int check_cuda(cudaError_t e);
int num_sms();

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while the previous kernel of the
// stream is still draining; it runs its prologue (barrier init, tensor-memory allocation, descriptor prefetch) and then
// blocks in pdl_wait() until that kernel has completed and its memory is visible.  RULES: every kernel launched this
// way executes pdl_wait() before its first global-memory access (reads AND writes: the previous kernel may still be
// reading a buffer this one recycles) -- that also keeps the ordering transitive along the chain; pdl_trigger() comes
// after the kernel's tensor-memory allocation (a dependent that grabbed TMEM first would starve it while waiting for
// it).  RAWFORMER_B200_PDL=0 turns the launch attribute off (plain stream order; the device-side calls are no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

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
inline i64 cdivl(i64 a, i64 b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t esize(int dtype) { return dtype == RF_BF16 ? 2 : 4; }

// Bump allocator over the caller-owned workspace.
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  size_t peak = 0;
  void* alloc(size_t bytes) {
    off = align_up(off, 256);
    char* p = base ? base + off : nullptr;
    // debugging aid (RAWFORMER_B200_ARENA_LOG=1): offset and size of every allocation of a real forward, in call order
    static const bool log = getenv("RAWFORMER_B200_ARENA_LOG") != nullptr;
    if (log && base) fprintf(stderr, "[arena] #%d off %zu size %zu\n", nlog++, off, bytes);
    off += bytes;
    if (off > peak) peak = off;
    return p;
  }
  int nlog = 0;
  template <typename U>
  U* get(size_t n) { return reinterpret_cast<U*>(alloc(n * sizeof(U))); }
  void* elems(size_t n, int dtype) { return alloc(n * esize(dtype)); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// Row-tiled single frame (BASELINE config 4): this rank owns a band of whole rows of ONE frame.  Every
// activation of the band is a "band image" [ht + rows_in + hb][W][C]: the interior rows plus BAND_HALO halo rows towards
// each neighbour (none at the frame border, where TMA zero fill is the convolution padding).  The kernels run on the band
// image as if it were a frame; the halo rows of a block input are fetched from the neighbours once per Conv_Transformer
// and lose one row of validity per 3x3 convolution of the chain (3 in the block + 1 in Downsample / the head).
constexpr int BAND_MAX_RANKS = RF_BAND_MAX_RANKS;
constexpr int BAND_HALO = 4;
constexpr int BAND_MAX_SYNCS = 32;
constexpr size_t BAND_FRAME_OFF = 64;        // u32 frame counter (advanced by the first kernel of every real forward)
constexpr size_t BAND_FLAGS_OFF = 256;       // [BAND_MAX_SYNCS][BAND_MAX_RANKS] u32 arrival counters
constexpr size_t BAND_MAIL_OFF = 4096;       // mailboxes (bump-allocated in call order, identical on all ranks)
struct Band {
  int rank = 0, nranks = 1;
  char* comm[BAND_MAX_RANKS] = {};   // comm region of every rank as mapped in this process (comm[rank] is local)
  unsigned epoch = 0;                // 0 = rehearsal (nobody signals or waits), else a real forward
  int next_sync = 0;                 // sync points used so far in this forward
  size_t mail_off = BAND_MAIL_OFF;   // bump cursor inside the comm region
  int ht = 0, hb = 0;                // halo rows above / below the interior
  int rows_in = 0;                   // interior rows at the current stage
  i64 P_full = 0;                    // pixels of the WHOLE frame at the current stage (squeeze-excite mean)
  float* se_partial = nullptr;       // squeeze-excite partial sums of the current block, reduced with its attention statistics
  int se_slots = 0;
};

struct Ctx {
  Band* band = nullptr;              // non-null: row-tiled forward
  cudaStream_t stream = 0;
  Arena arena;
  bool dry = false;   // size the arena only, launch nothing
  int dtype = RF_F32;
  // region of the workspace that the forward clears with ONE memset up front; the per-block accumulators (Gram, squared
  // norms, squeeze-excite partial sums) are carved out of it instead of being zeroed by a kernel each
  char* zero_base = nullptr;
  size_t zero_cap = 0, zero_off = 0;
  bool fits() const { return dry || arena.peak <= arena.cap; }
};
// fork / join of the per-thread side stream (rf_runtime.cu): false = no fork (dry run, profiling, disabled, or no side stream
// could be made because the stream is being captured and none exists yet)
bool side_fork(Ctx& ctx, cudaStream_t* side);
void side_join(Ctx& ctx, cudaStream_t side);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// 8 consecutive elements of T as floats (16 B for bf16, 32 B for float)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  uint2 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(bf16& d, float v) { d = __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float lrelu_f(float x) { return x >= 0.f ? x : 0.2f * x; }
// bf16-mode transcendental forms (errors far below bf16 resolution): MUFU-based sigmoid / tanh, and the exact-erf
// GELU through Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7) instead of the branchy libdevice erff.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfc_z = p * t * __expf(-z * z);       // erfc(|x|/sqrt2)
  const float cdf = x >= 0.f ? 1.0f - 0.5f * erfc_z : 0.5f * erfc_z;
  return x * cdf;
}
template <typename T> struct FastMath { static constexpr bool value = false; };
template <> struct FastMath<__nv_bfloat16> { static constexpr bool value = true; };

// float max via integer atomics (handles negatives); *addr must be initialised to -inf
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
  i0 = min((int)src, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  lam = src - (float)i0;
}
#endif  // __CUDACC__

}  // namespace rf
