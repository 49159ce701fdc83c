// Microbenchmarks that size the fused kernels (B200, sm_100a):
//   (1) tcgen05.ld throughput per SM vs. number of reading warps / CTAs per SM
//   (2) tcgen05.mma issue rate from ONE thread for small N (block-diagonal depthwise form: M=128, N=16/32/64, K=16)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../bayer_low_light_image_enhancement_b200/csrc -o ubench_tc ubench_tc.cu
#include <cstdio>
#include <cstdlib>
#include "rf_tma.cuh"
using namespace rf;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---- (1) TMEM read -----------------------------------------------------------------------------------------------
template <int X>   // columns per tcgen05.ld (16 or 32)
__global__ void k_tmem_read(int iters, int cols_alloc, unsigned long long* cyc_out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), cols_alloc);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 2 * X) {
      uint32_t va[16], vb[16];
      tmem_ld16(taddr + (c % cols_alloc), va);
      tmem_ld16(taddr + ((c + 16) % cols_alloc), vb);
      if (X == 32) {
        uint32_t vc[16], vd[16];
        tmem_ld16(taddr + ((c + 32) % cols_alloc), vc);
        tmem_ld16(taddr + ((c + 48) % cols_alloc), vd);
        tmem_ld_wait();
        acc += __uint_as_float(vc[3]) + __uint_as_float(vd[5]);
      } else {
        tmem_ld_wait();
      }
      acc += __uint_as_float(va[0]) + __uint_as_float(vb[7]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc_out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, cols_alloc); }
}

// ---- (2) MMA cost vs N and operand layout ---------------------------------------------------------------------------
// one thread issues `n` MMAs (M=128, N, K=16), commit, wait.  layout 0: no-swizzle K-major (LBO 4096, SBO 128: the halo-conv
// form), 1: SWIZZLE_128B K-major (rows of 128 B), 2: SWIZZLE_64B (rows of 64 B)
__global__ void k_mma_issue(int n, int N, int reps, int layout, unsigned long long* cyc_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1 && lane == 0) {
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint64_t a0, b0;
    if (layout == 0) { a0 = desc_noswz(base, 4096, 128); b0 = desc_noswz(base + 32768, 4096, 128); }
    else if (layout == 1) { a0 = make_sw128_desc(base); b0 = make_sw128_desc(base + 32768); }
    else {
      a0 = ((uint64_t)((base & 0x3FFFF) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
      b0 = a0 + (32768 >> 4);
    }
    const uint32_t idesc = make_idesc_m128(N);
    uint32_t ph = 0;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const uint64_t ad = a0 + (uint64_t)((i & 3) * 2);       // k-step inside the atom / chunk pair
        const uint64_t bd = b0 + (uint64_t)((i & 3) * 2);
        umma_f16(tb + (uint32_t)((i & 1) * 128), ad, bd, idesc, i >= 2 ? 1u : 0u);
      }
      if (r == 0) t1 = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), ph);
      ph ^= 1;
    }
    t2 = clock64();
    cyc_out[2 * blockIdx.x] = (unsigned long long)(t2 - t0);
    cyc_out[2 * blockIdx.x + 1] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

int main() {
  unsigned long long* d; float* sink;
  cudaMalloc(&d, 1 << 16); cudaMalloc(&sink, 4);
  unsigned long long h[2048];
  int dev_sms = 0; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", dev_sms);
  // (1)
  if (getenv("UB_TMEM"))
  for (int x = 16; x <= 32; x += 16)
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int warps = 4; warps <= 16; warps *= 2) {
      const int iters = 2000;
      const int grid = dev_sms * ctas;
      if (x == 16) k_tmem_read<16><<<grid, warps * 32>>>(iters, 128, d, sink);
      else k_tmem_read<32><<<grid, warps * 32>>>(iters, 128, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("tmem_read error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)iters * 128 * 32 * 4 * warps * ctas;   // per SM
      printf("tmem_read x%d ctas/SM %d warps/CTA %2d: %.1f B/cyc/SM (%.0f cyc)\n", x, ctas, warps, bytes / mx, mx);
    }
  // (2)
  cudaFuncSetAttribute(k_mma_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int layout = 0; layout <= 2; ++layout)
      for (int N = 16; N <= 128; N *= 2) {
        const int n = 64, reps = 100;
        const int grid = dev_sms * ctas;
        k_mma_issue<<<grid, 64, 100 * 1024>>>(n, N, reps, layout, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma_issue error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
        double mx = 0, is = 0; for (int i = 0; i < grid; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; is = h[2 * i + 1] > is ? h[2 * i + 1] : is; }
        printf("mma ctas/SM %d layout %d N=%3d: %.1f cyc/MMA per CTA (issue-only first rep %.1f), formula floor %d\n", ctas, layout, N,
               mx / (reps * (double)n), is / n, 128 * N / 256);
      }
  return 0;
}
