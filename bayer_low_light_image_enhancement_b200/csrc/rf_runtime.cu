// Runtime glue of the C ABI: status strings, device check, launch recorder (count + optional CUDA-event timing).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rf_common.cuh"

namespace rf {

void set_tcgen05_enabled(bool on);
bool tcgen05_enabled();

static thread_local LaunchRecorder g_rec;
LaunchRecorder& recorder() { return g_rec; }

// ---------------------------------------------------------------------------------------------
// side stream (per host thread and device): small per-image kernels that only the END of a block needs (the squeeze-excite
// MLP + channel_reduce fold) run next to the block's transformer branch instead of in front of it.  fork: the side stream
// waits for everything enqueued on the forward's stream so far; join: the forward's stream waits for the side stream.  Inside
// a stream capture the two become a fork / join of the CUDA graph.  Never created inside a capture (a forward captured
// before any eager forward of this thread simply does not fork).
// ---------------------------------------------------------------------------------------------
struct SideStream {
  int dev = -1;
  cudaStream_t s = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
};
static thread_local SideStream g_side[8];

static bool side_enabled() {
  static int on = -1;                     // debugging aid: RAWFORMER_B200_SIDE_STREAM=0 keeps everything on one stream
  if (on < 0) {
    const char* e = getenv("RAWFORMER_B200_SIDE_STREAM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

bool side_fork(Ctx& ctx, cudaStream_t* side) {
  if (ctx.dry || !side_enabled() || recorder().profiling) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  SideStream* ss = nullptr;
  for (auto& c : g_side)
    if (c.dev == dev) { ss = &c; break; }
  if (ss == nullptr) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx.stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return false;
    for (auto& c : g_side)
      if (c.dev < 0) { ss = &c; break; }
    if (ss == nullptr) return false;
    SideStream n;
    if (cudaStreamCreateWithFlags(&n.s, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&n.fork_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&n.join_ev, cudaEventDisableTiming) != cudaSuccess)
      return false;
    n.dev = dev;
    *ss = n;
  }
  if (cudaEventRecord(ss->fork_ev, ctx.stream) != cudaSuccess || cudaStreamWaitEvent(ss->s, ss->fork_ev, 0) != cudaSuccess) {
    recorder().last_cuda_error = (int)cudaGetLastError();
    return false;
  }
  *side = ss->s;
  return true;
}

void side_join(Ctx& ctx, cudaStream_t side) {
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& c : g_side)
    if (c.dev == dev && c.s == side) {
      if (cudaEventRecord(c.join_ev, side) != cudaSuccess || cudaStreamWaitEvent(ctx.stream, c.join_ev, 0) != cudaSuccess)
        recorder().last_cuda_error = (int)cudaGetLastError();
      return;
    }
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("RAWFORMER_B200_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

int check_cuda(cudaError_t e) {
  if (e == cudaSuccess) return RF_OK;
  g_rec.last_cuda_error = (int)e;
  return RF_ERR_CUDA;
}

static int g_debug_sync = -1;
static int g_cur_kernel = -1;
const char* kernel_name_of(int id);

void launch_begin(int kernel_id, double algo_bytes, double algo_flops) {
  LaunchRecorder& r = g_rec;
  r.count++;
  g_cur_kernel = kernel_id;
  if (r.profiling && r.n < r.cap) {
    r.ids[r.n] = kernel_id;
    r.bytes[r.n] = algo_bytes;
    r.flops[r.n] = algo_flops;
    cudaEventRecord(r.ev[2 * r.n], r.stream);
  }
}


void launch_end() {
  LaunchRecorder& r = g_rec;
  if (g_debug_sync < 0) {
    const char* e = getenv("RAWFORMER_B200_DEBUG");
    g_debug_sync = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_debug_sync) {  // debugging aid: synchronise after every launch and name the first failing kernel
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess && r.last_cuda_error == 0)
      fprintf(stderr, "[rawformer_b200] kernel '%s' (launch #%lld) failed: %s\n", kernel_name_of(g_cur_kernel), r.count,
              cudaGetErrorString(e));
  }
  if (r.profiling) {
    if (r.n < r.cap) cudaEventRecord(r.ev[2 * r.n + 1], r.stream);
    r.n++;
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) r.last_cuda_error = (int)e;
}

static const char* const kKernelNames[RF_K_COUNT] = {
    "pack_luma", "luma_norm", "dwt_high", "guidance", "flca_mod", "se_finalize", "fold_reduce", "layernorm",
    "gemm_qkv", "dw_qkv_gram", "attn_finalize", "gemm_proj_resid", "gemm_pw1", "dw_gelu", "gemm_pw2_resid",
    "gemm_cat_reduce", "conv3x3_out", "down_conv3x3", "up_convT", "skip_reduce", "embed", "head", "layout",
    "weight_pack", "misc", "pyr_spatial", "gemm_pyr_res1", "gemm_pyr_res2", "channel_sums", "tail_stats",
    "tail_apply", "index_op", "gemm_gram", "band_halo", "band_allreduce", "ffn_fused", "qkv_fused", "conv3x3_lc"};

const char* kernel_name_of(int id) { return (id >= 0 && id < RF_K_COUNT) ? kKernelNames[id] : "?"; }

}  // namespace rf

using namespace rf;

extern "C" {

const char* rf_strerror(int status) {
  switch (status) {
    case RF_OK: return "ok";
    case RF_ERR_BAD_SHAPE: return "bad shape (H,W multiples of 16; dim % 8 == 0; even DWT sizes)";
    case RF_ERR_BAD_ARG: return "bad argument (null/misaligned pointer or unknown enum)";
    case RF_ERR_ARCH: return "device is not sm_100 (B200); this library has no other code path";
    case RF_ERR_CUDA: return "CUDA runtime error (see rf_last_cuda_error)";
    case RF_ERR_WORKSPACE: return "workspace too small";
    case RF_ERR_UNSUPPORTED: return "not implemented in this build";
    default: return "unknown status";
  }
}

int rf_version(void) { return 100; }

int rf_last_cuda_error(void) {
  int e = g_rec.last_cuda_error;
  return e;
}

int rf_init(int device) {
  cudaDeviceProp prop;
  RF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return RF_ERR_ARCH;
  return RF_OK;
}

int rf_set_tcgen05(int enable) {
  int prev = rf::tcgen05_enabled() ? 1 : 0;
  if (enable >= 0) rf::set_tcgen05_enabled(enable != 0);
  return prev;
}

long long rf_launch_count(void) { return g_rec.count; }
void rf_reset_launch_count(void) { g_rec.count = 0; }

const char* rf_kernel_name(int kernel_id) {
  if (kernel_id < 0 || kernel_id >= RF_K_COUNT) return "?";
  return kKernelNames[kernel_id];
}

int rf_profiled_launch_info(int index, double* bytes_host, double* flops_host) {
  LaunchRecorder& r = g_rec;
  if (index < 0 || index >= r.last_n) return RF_ERR_BAD_ARG;
  if (bytes_host) *bytes_host = r.last_bytes[index];
  if (flops_host) *flops_host = r.last_flops[index];
  return RF_OK;
}

}  // extern "C"

namespace rf {

// Used by rf_rawformer_forward_profiled (rf_model.cu).
int profile_begin(cudaStream_t stream, int cap) {
  LaunchRecorder& r = g_rec;
  r.ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * 2 * cap);
  r.ids = (int*)malloc(sizeof(int) * cap);
  r.bytes = (double*)malloc(sizeof(double) * cap);
  r.flops = (double*)malloc(sizeof(double) * cap);
  for (int i = 0; i < 2 * cap; ++i) {
    if (cudaEventCreate(&r.ev[i]) != cudaSuccess) return RF_ERR_CUDA;
  }
  r.cap = cap;
  r.n = 0;
  r.stream = stream;
  r.profiling = true;
  return RF_OK;
}

int profile_end(float* ms_host, int* ids_host, int cap_out, int* n_host) {
  LaunchRecorder& r = g_rec;
  r.profiling = false;
  int st = check_cuda(cudaStreamSynchronize(r.stream));
  int n = r.n < r.cap ? r.n : r.cap;
  free(r.last_ids); free(r.last_bytes); free(r.last_flops);
  r.last_ids = (int*)malloc(sizeof(int) * (n + 1));
  r.last_bytes = (double*)malloc(sizeof(double) * (n + 1));
  r.last_flops = (double*)malloc(sizeof(double) * (n + 1));
  r.last_n = n;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    if (st == RF_OK) cudaEventElapsedTime(&ms, r.ev[2 * i], r.ev[2 * i + 1]);
    if (i < cap_out) {
      if (ms_host) ms_host[i] = ms;
      if (ids_host) ids_host[i] = r.ids[i];
    }
    r.last_ids[i] = r.ids[i];
    r.last_bytes[i] = r.bytes[i];
    r.last_flops[i] = r.flops[i];
  }
  if (n_host) *n_host = r.n;
  for (int i = 0; i < 2 * r.cap; ++i) cudaEventDestroy(r.ev[i]);
  free(r.ev); free(r.ids); free(r.bytes); free(r.flops);
  r.ev = nullptr; r.ids = nullptr; r.bytes = nullptr; r.flops = nullptr;
  r.cap = 0;
  return st;
}

}  // namespace rf
