// Table-driven FP64 log / exp / reciprocal / tanh for the staged model code
// of the fused kernel (oqg_fast.cuh).
//
// The CUDA math library's log(), exp(), tanh() and the IEEE divide cost
// 105 / 54 / 77 / 25 SASS instructions on sm_100a (ncu source view of the
// fused kernel, profiles/r1_v6_fused.md): they carry the special cases
// (zero, denormals, infinities, NaN, negative arguments) and table-free
// polynomials of degree 12+.  The model code only ever passes positive,
// finite, normal arguments, so the versions below drop the special cases
// and trade the long polynomials for a 2 KB + 512 B lookup table in shared
// memory:
//
//   fast_log(x)   x = 2^e m; i = top 7 bits of m; r = m invc_i - 1 (one
//                 FMA, |r| < 2^-8); log x = e ln2 + logc_i + log1p(r) with a
//                 degree-6 series.  18 instructions, <= 1 ulp-ish: absolute
//                 error <= 2.5e-16 max(1, |log x|).  Domain: x positive,
//                 finite, normal.
//   fast_exp(x)   k = round(64 x / ln2); r = x - k ln2/64 (two FMAs,
//                 |r| < 0.0055); e^x = 2^(k>>6) T[k&63] (1 + r + ... r^5/120).
//                 19 instructions, relative error <= 3e-16.  Domain
//                 |x| < 700.
//   fast_rcp(x)   MUFU.RCP64H + two Newton steps; relative error <= 2e-16.
//                 Domain: x finite, normal, non-zero.
//   fast_tanh_pos(x)  1 - 2 / (e^{2x} + 1), x >= 0 (absolute error 2e-16).
//   fast_cosh(x)      (e^x + 1 / e^x) / 2, relative error <= 5e-16.
//
// The unfused operators (k_mean_stds, oqg_gmpe.cuh) keep the CUDA library
// functions; the GPU parity tests compare the two (oqg_mean_stds_staged
// against the reference fixtures at rtol 1e-10 in tests/test_gpu_parity.py).
//
// A kernel that calls any of these must call math_tables_init() (all
// threads of the CTA) followed by __syncthreads() first.
#pragma once
#include "oqg_math_tables.h"

namespace oqg {

__shared__ double2 s_logtab[128];
__shared__ double s_exptab[64];

__device__ __forceinline__ void math_tables_init() {
  for (int i = threadIdx.x; i < 128; i += blockDim.x)
    s_logtab[i] = make_double2(g_logtab[i][0], g_logtab[i][1]);
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    s_exptab[i] = g_exptab[i];
}

__device__ __forceinline__ double fast_log(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const double2 t = s_logtab[(hi >> 13) & 127];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  // e as a double without I2F: 2^52 + 2^31 + (e ^ 2^31) trick
  const double de = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000)
                    - 4503601774854144.0;
  const double r = fma(m, t.x, -1.0);
  double p = fma(r, -1. / 6., 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1. / 3.);
  p = fma(p, r, -0.5);
  p = fma(p * r, r, r);            // r - r^2/2 + ... - r^6/6
  return fma(de, 0.6931471805599453094, t.y + p);
}

__device__ __forceinline__ double fast_exp(double x) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 0x1.71547652b82fep+6, MAGIC);  // 64 / ln2
  const int k = __double2loint(t);
  const double kd = t - MAGIC;
  // ln2 / 64 split in two (hi has its low 21 bits zero: kd * hi is exact)
  double r = fma(kd, -0x1.62e42fee00000p-7, x);
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  const double T = s_exptab[k & 63];
  double q = fma(r, 1. / 120., 1. / 24.);
  q = fma(q, r, 1. / 6.);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = q * r;                        // e^r - 1
  const double v = fma(T, q, T);
  // times 2^(k >> 6): add to the exponent field
  return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20),
                          __double2loint(v));
}

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// tanh for x >= 0 (x < 350)
__device__ __forceinline__ double fast_tanh_pos(double x) {
  return fma(-2.0, fast_rcp(fast_exp(x + x) + 1.0), 1.0);
}

// cosh (|x| < 350): (e^x + e^-x) / 2 with one exp and one reciprocal
__device__ __forceinline__ double fast_cosh(double x) {
  const double e = fast_exp(x);
  return 0.5 * (e + fast_rcp(e));
}

}  // namespace oqg
