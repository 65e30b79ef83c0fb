// Arithmetic shared by the epilogues of the tcgen05 kernels (frontend.cu, posconv.cu): packed fp32x2 helpers, the
// single-MUFU exact GELU and its derivative, bf16 packing, 256-bit global stores.
#pragma once

#include "common.cuh"

namespace nrse {
namespace {

// 256-bit global store (STG.256 on sm_100): one full 32-byte sector per thread per instruction; p 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction) --------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_make(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 f2_bits(uint32_t lo, uint32_t hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_split(f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Exact (erf) GELU of two values with ONE MUFU each, written in the HALVED argument w = x / 2 (a = |w|):
//   gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2) = w + a (1 - e),   e = erfc(sqrt2 a) = 2^-q(a),   q(a) = a (c0 + c1 a + ..)
// q is a fit of -log2 erfc(sqrt2 a) on [0, 2.8] that minimises the absolute error of gelu itself (weight a e ln 2):
//   NRSE_GELU_DEG 5 (default): quintic q, 5 packed instructions, max |gelu error| 8.7e-7;
//   NRSE_GELU_DEG 3:           cubic q, 3 packed instructions,  max |gelu error| 8.6e-5 (2 % of a bf16 ulp at 1) -- measured
//                              no faster on the B200 (the epilogues are not instruction-bound), so unused;
// tests/test_oracle_golden.py pins both against torch's erf GELU.  q keeps growing beyond the fit range (q(2.8) = 25.6,
// q(6) > 126), so e flushes to 0 by itself: no clamp.
// The epilogues are bound by instruction ISSUE (23.5 k warp instructions per 128 x 512 tile at IPC 1.9 before this form),
// so everything that is not arithmetic on the value is moved out of the per-element path: the factor 1/2 lives in the
// LayerNorm affine (shared memory holds gamma / 2 and beta / 2 -- exact, a power of two) or in the folded layer-0 operands,
// |w| is an operand modifier of the packed instructions, relu(x) is never formed (w + a h cancels to -a e for x < 0 with an
// absolute error below 1e-7 a).  NaN inputs propagate; +-inf is not expected behind a LayerNorm (-inf gives NaN).
#ifndef NRSE_STATS_SHIFT
#define NRSE_STATS_SHIFT 1  // 0 (timing experiments only): raw sums in the LayerNorm statistics pass
#endif
#ifndef NRSE_GELU_DEG
#define NRSE_GELU_DEG 5
#endif
__device__ __forceinline__ f2 gelu2h(f2 w) {
  float w0, w1;
  f2_split(w, w0, w1);
  const f2 a = f2_make(fabsf(w0), fabsf(w1));
#define NRSE_F2C(v) f2_make(v, v)
#if NRSE_GELU_DEG == 3
  f2 p = f2_fma(a, NRSE_F2C(-0.2210327833890915f), NRSE_F2C(-1.9530425071716309f));
  p = f2_fma(p, a, NRSE_F2C(-2.281832695007324f));
#elif NRSE_GELU_DEG == 5
  f2 p = f2_fma(a, NRSE_F2C(-0.015619270503520966f), NRSE_F2C(0.11517950147390366f));
  p = f2_fma(p, a, NRSE_F2C(-0.4171730577945709f));
  p = f2_fma(p, a, NRSE_F2C(-1.838383436203003f));
  p = f2_fma(p, a, NRSE_F2C(-2.3020009994506836f));
#else
#error "NRSE_GELU_DEG must be 3 or 5"
#endif
  float q0, q1;
  f2_split(f2_mul(p, a), q0, q1);  // = -q(a)
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const f2 h = f2_fma(f2_make(e0, e1), NRSE_F2C(-1.0f), NRSE_F2C(1.0f));
#undef NRSE_F2C
  return f2_fma(a, h, w);
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& a, float (&x)[8]) {
  const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[2 * j] = __uint_as_float(w[j] << 16);
    x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// gelu'(v) for two values with ONE MUFU each:  for u = |v|,  1 - gelu'(u) = phi(u) (mills(u) - u)  and
// gelu'(-u) = 1 - gelu'(u);  r(u) = 0.39894 (mills(u) - u) is a degree-6 fit on [0, 5.6] (max abs error of gelu'
// 2.6e-5, two orders below bf16 resolution), so gelu'(v) = 0.5 + copysign(0.5 - exp(-v^2/2) r(|v|), v).
__device__ __forceinline__ f2 gelu_grad2(f2 v) {
  float v0, v1;
  f2_split(v, v0, v1);
  const f2 u = f2_make(fminf(fabsf(v0), 5.6f), fminf(fabsf(v1), 5.6f));
  constexpr float k = 0.3989422804f;
  f2 r = f2_fma(u, f2_make(k * 0.0016475850716233253f, k * 0.0016475850716233253f),
                f2_make(k * -0.019207235425710678f, k * -0.019207235425710678f));
  r = f2_fma(r, u, f2_make(k * 0.09663444012403488f, k * 0.09663444012403488f));
  r = f2_fma(r, u, f2_make(k * -0.2893761098384857f, k * -0.2893761098384857f));
  r = f2_fma(r, u, f2_make(k * 0.6104238033294678f, k * 0.6104238033294678f));
  r = f2_fma(r, u, f2_make(k * -1.9976462125778198f, k * -1.9976462125778198f));
  r = f2_fma(r, u, f2_make(k * 1.2532488107681274f, k * 1.2532488107681274f));
  float q0, q1;
  f2_split(f2_mul(f2_mul(u, u), f2_make(-0.72134752f, -0.72134752f)), q0, q1);  // -u^2/2 * log2(e)
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  float h0, h1;
  f2_split(f2_fma(f2_mul(f2_make(e0, e1), r), f2_make(-1.f, -1.f), f2_make(0.5f, 0.5f)), h0, h1);  // 0.5 - (1 - gelu'(u))
  return f2_add(f2_make(0.5f, 0.5f), f2_make(copysignf(h0, v0), copysignf(h1, v1)));
}

}  // namespace
}  // namespace nrse
