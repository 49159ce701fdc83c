// Inner-loop helpers of the bf16 depthwise kernels (rf_dw_tma.cu, rf_ffn_fused.cu, rf_qkv_fused.cu): shared-memory 8-byte
// loads, bf16x4 -> packed fp32 pairs, and the erf-GELU in packed fp32.
#pragma once
#include "rf_common.cuh"

namespace rf {

#ifdef __CUDACC__
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
// 4 bf16 -> two packed fp32 pairs (channel pairs (0,1), (2,3)); the conv runs on packed FFMA2 (fma.rn.f32x2)
__device__ __forceinline__ void unpack_bf16x4(const uint2& t, float2 (&v)[2]) {
  // PRMT / LOP3 run on the ALU pipe (a plain shift is turned into IMAD.U32, which competes with the FFMA2s)
  v[0] = make_float2(__uint_as_float(__byte_perm(t.x, 0u, 0x1044u)), __uint_as_float(t.x & 0xffff0000u));
  v[1] = make_float2(__uint_as_float(__byte_perm(t.y, 0u, 0x1044u)), __uint_as_float(t.y & 0xffff0000u));
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU of two values in packed fp32:  gelu(x) = x * Phi(x),  Phi(x) = 1 / (1 + exp(-x * P(min(x^2, 64)))),
// P = degree-4 minimax fit of logit(Phi(x)) / x on |x| <= 8 (tools/fit_gelu.py): |gelu error| <= 1.2e-5 absolute and
// <= 7.5e-5 of max(|gelu|, 0.02), i.e. < 1/10 of a bf16 half-ulp; no cancellation in the negative tail (x -> -inf
// gives x * 0), exact identity for x >= 8.  8 packed FMA-pipe ops + 2 MUFU per value pair (the Abramowitz-Stegun
// erfc form needs 12 + 2); coefficients are pre-multiplied by -log2(e) so the exponential is a bare ex2.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float k0 = -1.4426950408889634f * 1.5954254501877632f, k1 = -1.4426950408889634f * 0.07325994800505611f,
              k2 = -1.4426950408889634f * -0.00036788829074290585f, k3 = -1.4426950408889634f * -4.5953771468195396e-05f,
              k4 = -1.4426950408889634f * 1.6250086403938804e-06f;
  float2 t = __fmul2_rn(x, x);
  t = make_float2(fminf(t.x, 64.f), fminf(t.y, 64.f));
  float2 pz = __ffma2_rn(make_float2(k4, k4), t, make_float2(k3, k3));
  pz = __ffma2_rn(pz, t, make_float2(k2, k2));
  pz = __ffma2_rn(pz, t, make_float2(k1, k1));
  pz = __ffma2_rn(pz, t, make_float2(k0, k0));
  const float2 w = __fmul2_rn(x, pz);
  const float2 d = __fadd2_rn(make_float2(ex2_approx(w.x), ex2_approx(w.y)), make_float2(1.f, 1.f));
  return __fmul2_rn(x, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
// The same function with ONE MUFU op per value: Phi(x) = 1/2 + 1/2 tanh(x * Q(min(x^2, 25))), Q = degree-2 minimax fit of
// atanh(2 Phi(x) - 1) / x: |gelu error| of the polynomial <= 6.2e-5 absolute, 3.7e-4 of max(|gelu|, 0.02); tanh.approx adds
// up to 2^-11 relative to tanh, i.e. <= 2.5e-4 |x| absolute -- about 1/4 of a bf16 half-ulp of the values that matter.  For
// epilogues that are MUFU-bound (rf_lnconv.cu: 64 GELUs per pixel between two tensor-core contractions).
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  const float q0 = 0.7971700425576733f, q1 = 0.03726032541872442f, q2 = -0.000384762947119491f;
  float2 t = __fmul2_rn(x, x);
  t = make_float2(fminf(t.x, 25.f), fminf(t.y, 25.f));
  float2 pz = __ffma2_rn(make_float2(q2, q2), t, make_float2(q1, q1));
  pz = __ffma2_rn(pz, t, make_float2(q0, q0));
  const float2 w = __fmul2_rn(x, pz);
  const float2 ph = __ffma2_rn(make_float2(0.5f, 0.5f), make_float2(tanh_fast(w.x), tanh_fast(w.y)), make_float2(0.5f, 0.5f));
  return __fmul2_rn(x, ph);
}
#endif  // __CUDACC__

}  // namespace rf
