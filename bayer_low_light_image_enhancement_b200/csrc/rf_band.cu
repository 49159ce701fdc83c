// Cross-GPU steps of the row-tiled single-frame forward (BASELINE config 4): halo exchange between band
// neighbours and all-reduce of the per-image reductions, both through peer-mapped "comm regions" over NVLink.
//
// Every rank owns one comm region (cudaMalloc'ed by this library, exported with CUDA IPC): arrival counters
// flags[sync point][source rank] and one mailbox per sync point.  A sync point is ONE kernel per rank:
//   1. push:   plain 16-byte / coalesced 4-byte stores of the local data into the peers' mailboxes (NVLink writes)
//   2. signal: per CTA  bar.sync -> fence.sys -> atomicAdd_system on the peer's counter of this rank
//   3. wait:   spin (bounded, ld.acquire.sys) until every source's counter reached frame * CTAs (frame = the frame
//               counter of the local region, advanced on the device by k_band_begin)
//   4. combine from the LOCAL mailbox (L1-bypassing loads): copy into the halo rows / sum over ranks in rank order,
//      so every rank computes bit-identical statistics.
// The mailbox of a sync point is reused every frame: a peer can only reach sync point j of frame f+1 after the last
// sync point of frame f, which needs this rank's push there, which is stream-ordered after its combine of j in frame f.
// No CTA depends on another CTA of the same grid, so nothing requires co-residency; a peer that never arrives makes the
// wait time out (sticky error word, rf_band_comm_status) instead of hanging the GPU.
// epoch 0 is a REHEARSAL: nothing is signalled or waited for.  It exists because the first launch of a kernel may have
// to load it (lazy module loading), which can wait for the device to go idle -- while a sync-point kernel spins for a
// peer that the same host thread has yet to launch.  One rehearsal forward per rank loads every kernel of its plan.
#include <string.h>

#include "rf_kernels.cuh"

namespace rf {

constexpr long long BAND_SPIN_CYCLES = 12000000000LL;   // ~6 s at 1.9 GHz

struct BandSyncP {
  unsigned* flags_local;                 // [BAND_MAX_RANKS] counters of this sync point in the local region
  unsigned* flags_peer[BAND_MAX_RANKS];  // the same counters in every rank's region
  unsigned* err;                         // local sticky error word
  const unsigned* frame;                 // local frame counter (k_band_begin); nullptr = rehearsal
  unsigned target;                       // frame * gridDim.x, filled in by the kernel (0 = rehearsal)
  int rank, nranks, sync_id;
};

// The frame counter lives in the comm region and is advanced by the first kernel of every real forward, so the launch
// sequence of a forward has no per-frame host argument and can be replayed as a CUDA graph.
__global__ void k_band_begin(unsigned* frame) { *frame += 1u; }

__device__ __forceinline__ void band_load_target(BandSyncP& s) {
  s.target = s.frame != nullptr ? *reinterpret_cast<const volatile unsigned*>(s.frame) * gridDim.x : 0u;
}

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all threads of the CTA call both; `mask` = ranks to signal / to wait for
__device__ __forceinline__ void band_signal(const BandSyncP& s, unsigned mask) {
  __syncthreads();
  if (s.target == 0u) return;
  if (threadIdx.x < (unsigned)s.nranks && ((mask >> threadIdx.x) & 1u)) {
    __threadfence_system();
    atomicAdd_system(s.flags_peer[threadIdx.x] + s.rank, 1u);
  }
}
__device__ __forceinline__ void band_wait(const BandSyncP& s, unsigned mask) {
  if (s.target == 0u) return;
  if (threadIdx.x < (unsigned)s.nranks && ((mask >> threadIdx.x) & 1u)) {
    const unsigned* f = s.flags_local + threadIdx.x;
    const long long t0 = clock64();
    // counters only grow; compare as a signed distance so that a wrapped counter still orders correctly
    while ((int)(ld_acquire_sys_u32(f) - s.target) < 0) {
      if (*reinterpret_cast<volatile unsigned*>(s.err) != 0u) break;
      if (clock64() - t0 > BAND_SPIN_CYCLES) {
        atomicCAS(s.err, 0u, (unsigned)s.sync_id + 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// ---- halo exchange of one band image -------------------------------------------------------------------------------
struct BandHaloP {
  BandSyncP s;
  uint4* x;                 // band image, row_u4 uint4 per row
  i64 row_u4;
  int ht, hb, rows_in;
  uint4* mail_local;        // [2][BAND_HALO rows]: slot 0 = rows from the rank above, slot 1 = from the rank below
  uint4* mail_above;        // the same mailbox in rank-1's region (we fill its slot 1)
  uint4* mail_below;        // ... in rank+1's region (we fill its slot 0)
};

__global__ void __launch_bounds__(512)
k_band_halo(BandHaloP p) {
  pdl_wait();           // (no pdl_trigger: the kernels behind a sync point must not occupy SMs while it waits for a peer)
  band_load_target(p.s);
  const i64 n = (i64)BAND_HALO * p.row_u4;
  const i64 i0 = (i64)blockIdx.x * blockDim.x + threadIdx.x, step = (i64)gridDim.x * blockDim.x;
  unsigned mask = 0;
  if (p.ht) {
    mask |= 1u << (p.s.rank - 1);
    const uint4* src = p.x + (i64)p.ht * p.row_u4;                          // first interior rows
    uint4* dst = p.mail_above + n;
    for (i64 i = i0; i < n; i += step) dst[i] = src[i];
  }
  if (p.hb) {
    mask |= 1u << (p.s.rank + 1);
    const uint4* src = p.x + (i64)(p.ht + p.rows_in - BAND_HALO) * p.row_u4;  // last interior rows
    uint4* dst = p.mail_below;
    for (i64 i = i0; i < n; i += step) dst[i] = src[i];
  }
  band_signal(p.s, mask);
  band_wait(p.s, mask);
  if (p.ht) {
    uint4* dst = p.x;
    for (i64 i = i0; i < n; i += step) dst[i] = __ldcg(p.mail_local + i);
  }
  if (p.hb) {
    uint4* dst = p.x + (i64)(p.ht + p.rows_in) * p.row_u4;
    for (i64 i = i0; i < n; i += step) dst[i] = __ldcg(p.mail_local + n + i);
  }
}

// ---- all-reduce (sum) of the per-image reductions of one Conv_Transformer ---------------------------------------------
// ONE exchange per block: the attention statistics {gram [C][C], |q|^2 [C], |k|^2 [C]} -- of the Gram only the 8 per-head
// diagonal blocks are used (FLCA_RF.py:230) and exchanged -- followed by the C squeeze-excite channel sums
// (FLCA_RF.py:160), which the FLCA kernel leaves as `se_slots` partial rows and which come back reduced into row 0.
struct BandReduceP {
  BandSyncP s;
  float* stats;                   // attention statistics of C channels
  float* norms;                   // [2C] squared norms of q,k when they are kept outside stats (else nullptr)
  float* se;                      // [se_slots][C] partial channel sums (nullptr: none)
  int C, se_slots;
  int n;                          // elements exchanged = C*C/8 + 2C (+ C)
  int n_pad;                      // slot pitch in floats
  float* mail_local;              // [nranks][n_pad]
  float* mail_peer[BAND_MAX_RANKS];
};

__device__ __forceinline__ i64 band_stats_idx(int e, int C) {
  const int c = C >> 3;
  if (e >= C * c) return (i64)C * C + (e - C * c);
  const int h = e / (c * c), r = e - h * c * c;
  const int i = r / c, j = r - i * c;
  return (i64)(h * c + i) * C + h * c + j;
}

__global__ void __launch_bounds__(256)
k_band_allreduce(BandReduceP p) {
  pdl_wait();
  band_load_target(p.s);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
  const int n_attn = p.C * (p.C >> 3) + 2 * p.C;
  for (int e = i0; e < p.n; e += step) {
    float v;
    if (e < n_attn) {
      v = (p.norms != nullptr && e >= n_attn - 2 * p.C) ? p.norms[e - (n_attn - 2 * p.C)] : p.stats[band_stats_idx(e, p.C)];
    } else {
      v = 0.f;
      for (int k = 0; k < p.se_slots; ++k) v += p.se[(i64)k * p.C + (e - n_attn)];
    }
    for (int r = 0; r < p.s.nranks; ++r) p.mail_peer[r][(i64)p.s.rank * p.n_pad + e] = v;
  }
  const unsigned mask = (1u << p.s.nranks) - 1u;
  band_signal(p.s, mask);
  band_wait(p.s, mask);
  for (int e = i0; e < p.n; e += step) {
    float sum = 0.f;
    for (int r = 0; r < p.s.nranks; ++r) sum += __ldcg(p.mail_local + (i64)r * p.n_pad + e);
    if (e >= n_attn) p.se[e - n_attn] = sum;
    else if (p.norms != nullptr && e >= n_attn - 2 * p.C) p.norms[e - (n_attn - 2 * p.C)] = sum;
    else p.stats[band_stats_idx(e, p.C)] = sum;
  }
}

// ---- launchers (advance the sync / mailbox cursors in dry mode too: that is how the comm region is sized) --------------
static bool band_sync_params(Ctx& ctx, BandSyncP& s, int grid) {
  Band& b = *ctx.band;
  if (b.next_sync >= BAND_MAX_SYNCS) {
    recorder().last_cuda_error = (int)cudaErrorInvalidValue;
    return false;
  }
  const int id = b.next_sync++;
  if (ctx.dry) return true;
  for (int r = 0; r < b.nranks; ++r)
    s.flags_peer[r] = reinterpret_cast<unsigned*>(b.comm[r] + BAND_FLAGS_OFF) + (size_t)id * BAND_MAX_RANKS;
  s.flags_local = s.flags_peer[b.rank];
  s.err = reinterpret_cast<unsigned*>(b.comm[b.rank]);
  s.frame = b.epoch != 0 ? reinterpret_cast<const unsigned*>(b.comm[b.rank] + BAND_FRAME_OFF) : nullptr;
  s.target = 0;
  (void)grid;
  s.rank = b.rank; s.nranks = b.nranks; s.sync_id = id;
  return true;
}

void band_halo_exchange(Ctx& ctx, void* x, int W, int C) {
  Band& b = *ctx.band;
  const i64 row_bytes = (i64)W * C * (i64)esize(ctx.dtype);
  const size_t slot = align_up((size_t)BAND_HALO * row_bytes, 256);
  const size_t off = b.mail_off;
  b.mail_off += 2 * slot;
  const int grid = 16;
  BandHaloP p;
  if (!band_sync_params(ctx, p.s, grid) || ctx.dry) return;
  if (b.nranks == 1) return;                 // no neighbours: nothing to fetch
  if (row_bytes % 16 || slot != (size_t)BAND_HALO * row_bytes) {   // W*C*2 is a multiple of 256 for every legal shape
    recorder().last_cuda_error = (int)cudaErrorInvalidValue;
    return;
  }
  p.x = reinterpret_cast<uint4*>(x);
  p.row_u4 = row_bytes / 16;
  p.ht = b.ht; p.hb = b.hb; p.rows_in = b.rows_in;
  p.mail_local = reinterpret_cast<uint4*>(b.comm[b.rank] + off);
  p.mail_above = b.rank > 0 ? reinterpret_cast<uint4*>(b.comm[b.rank - 1] + off) : nullptr;
  p.mail_below = b.rank + 1 < b.nranks ? reinterpret_cast<uint4*>(b.comm[b.rank + 1] + off) : nullptr;
  ScopedLaunch sl(RF_K_BAND_HALO, 2.0 * ((b.ht ? 1 : 0) + (b.hb ? 1 : 0)) * BAND_HALO * (double)row_bytes);
  launch_pdl(k_band_halo, dim3(grid), dim3(512), 0, ctx.stream, p);
}

void band_begin(Ctx& ctx) {
  Band& b = *ctx.band;
  if (ctx.dry || b.epoch == 0) return;
  ScopedLaunch sl(RF_K_MISC);
  k_band_begin<<<1, 1, 0, ctx.stream>>>(reinterpret_cast<unsigned*>(b.comm[b.rank] + BAND_FRAME_OFF));
}

void band_allreduce(Ctx& ctx, float* stats, int C, float* se, int se_slots, float* norms) {
  Band& b = *ctx.band;
  const int n = C * (C / 8) + 2 * C + (se != nullptr || ctx.dry ? C : 0);
  const int n_pad = (int)align_up((size_t)n, 64);
  const size_t off = b.mail_off;
  b.mail_off += (size_t)b.nranks * n_pad * sizeof(float);
  int grid = cdiv(n, 2048);
  if (grid > 16) grid = 16;
  BandReduceP p;
  if (!band_sync_params(ctx, p.s, grid) || ctx.dry) return;
  p.stats = stats; p.norms = norms; p.se = se; p.C = C; p.se_slots = se_slots; p.n = n; p.n_pad = n_pad;
  for (int r = 0; r < b.nranks; ++r) p.mail_peer[r] = reinterpret_cast<float*>(b.comm[r] + off);
  p.mail_local = p.mail_peer[b.rank];
  ScopedLaunch sl(RF_K_BAND_ALLREDUCE, 4.0 * n * (2.0 * b.nranks));
  launch_pdl(k_band_allreduce, dim3(grid), dim3(256), 0, ctx.stream, p);
}

// ---- all-reduce (sum) of a few floats (multi-level variant: the three output-channel sums of the colour anchor,
//      ML_RF.py:284-285); every rank adds the ranks' contributions in rank order
struct BandSmallP {
  BandSyncP s;
  float* v;
  int n, n_pad;
  float* mail_local;              // [nranks][n_pad]
  float* mail_peer[BAND_MAX_RANKS];
};

__global__ void __launch_bounds__(64)
k_band_allreduce_small(BandSmallP p) {
  pdl_wait();
  band_load_target(p.s);
  for (int e = threadIdx.x; e < p.n; e += blockDim.x) {
    const float v = p.v[e];
    for (int r = 0; r < p.s.nranks; ++r) p.mail_peer[r][(i64)p.s.rank * p.n_pad + e] = v;
  }
  const unsigned mask = (1u << p.s.nranks) - 1u;
  band_signal(p.s, mask);
  band_wait(p.s, mask);
  if (p.s.target == 0u) return;                       // rehearsal: nothing was exchanged
  for (int e = threadIdx.x; e < p.n; e += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < p.s.nranks; ++r) sum += __ldcg(p.mail_local + (i64)r * p.n_pad + e);
    p.v[e] = sum;
  }
}

void band_allreduce_small(Ctx& ctx, float* v, int n) {
  Band& b = *ctx.band;
  const int n_pad = (int)align_up((size_t)n, 64);
  const size_t off = b.mail_off;
  b.mail_off += (size_t)b.nranks * n_pad * sizeof(float);
  BandSmallP p;
  if (!band_sync_params(ctx, p.s, 1) || ctx.dry) return;
  if (b.nranks == 1) return;
  p.v = v; p.n = n; p.n_pad = n_pad;
  for (int r = 0; r < b.nranks; ++r) p.mail_peer[r] = reinterpret_cast<float*>(b.comm[r] + off);
  p.mail_local = p.mail_peer[b.rank];
  ScopedLaunch sl(RF_K_BAND_ALLREDUCE, 4.0 * n * (2.0 * b.nranks));
  launch_pdl(k_band_allreduce_small, dim3(1), dim3(64), 0, ctx.stream, p);
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_band_comm_alloc(size_t bytes, void** ptr_host, unsigned char* handle_host) {
  if (!ptr_host || bytes < BAND_MAIL_OFF) return RF_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == RF_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  RF_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess && handle_host) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) memcpy(handle_host, &h, sizeof(h));
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return check_cuda(e);
  }
  *ptr_host = p;
  return RF_OK;
}

int rf_band_comm_open(const unsigned char* handle_host, void** ptr_host) {
  if (!handle_host || !ptr_host) return RF_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  RF_CUDA(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
  return RF_OK;
}

int rf_band_comm_close(void* ptr) {
  if (!ptr) return RF_ERR_BAD_ARG;
  RF_CUDA(cudaIpcCloseMemHandle(ptr));
  return RF_OK;
}

int rf_band_comm_free(void* ptr) {
  if (!ptr) return RF_ERR_BAD_ARG;
  RF_CUDA(cudaFree(ptr));
  return RF_OK;
}

int rf_band_comm_reset(void* comm_own, void* stream) {
  // header of the region: sticky error word, frame counter, arrival counters (the mailboxes need no reset)
  if (!comm_own) return RF_ERR_BAD_ARG;
  RF_CUDA(cudaMemsetAsync(comm_own, 0, BAND_MAIL_OFF, (cudaStream_t)stream));
  RF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return RF_OK;
}

int rf_band_comm_status(const void* comm_own, int* err_host, void* stream) {
  if (!comm_own || !err_host) return RF_ERR_BAD_ARG;
  unsigned v = 0;
  RF_CUDA(cudaMemcpyAsync(&v, comm_own, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *err_host = (int)v;
  return RF_OK;
}

}  // extern "C"
