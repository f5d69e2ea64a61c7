// Sampling side of the ground motion fields:
// what GmfComputer.compute_all (openquake/hazardlib/calc/gmf.py:394-474)
// does around the arithmetic of _compute (k_gmf, oqg_disagg.cuh).
//
//   k_truncnorm_rvs   the epsilons: stats.truncnorm(-tl, tl).rvs(size, rng)
//                     with rng = numpy.random.default_rng(seed), i.e. PCG64.
//                     scipy draws U = rng.uniform(size) and returns the
//                     inverse CDF; the kernel reproduces the SAME uniforms
//                     (PCG64 XSL-RR 128/64 with jump-ahead: every thread
//                     advances the generator to its own position of the
//                     stream) and evaluates ndtri(Phi(-tl) + U (Phi(tl) -
//                     Phi(-tl))) in float64, rounded to float32 as the
//                     reference stores them (.astype(F32), gmf.py:443-446)
//   k_corr_matrix     SpatialCorrelationModel.correlation_matrix for the
//                     models of the form exp(c * distance) (JB2009,
//                     correlation_models/spatial/jayaram_baker_2009.py:33-44)
//                     on the geodetic distance matrix of the sites
//                     (geo/mesh.py:401-429, geo/geodetic.py geodetic_distance)
//   k_gmf_max_min     set_max_min (gmf.py:71-91): values above max_iml capped
//                     at the median, (site, event) columns with every IMT
//                     below min_iml zeroed
#pragma once
#include <math.h>
#include <stdint.h>

namespace oqg {

typedef unsigned __int128 u128;

struct Pcg64Jump {
  // state_{n + 2^i} = mult[i] * state_n + plus[i]  (mod 2^128)
  uint64_t mult_hi[64], mult_lo[64], plus_hi[64], plus_lo[64];
};

__device__ __forceinline__ u128 mk128(uint64_t hi, uint64_t lo) {
  return ((u128)hi << 64) | lo;
}

// numpy/random/src/pcg64/pcg64.h: pcg_output_xsl_rr_128_64
__device__ __forceinline__ uint64_t pcg64_output(u128 s) {
  const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}

// samples i = first + t, first + t + T, ... of the stream: numpy steps the
// state and then outputs, so sample i comes from the state after i + 1 steps
__global__ void __launch_bounds__(256)
k_truncnorm_rvs(const Pcg64Jump *__restrict__ J, uint64_t state_hi,
                uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                uint64_t stride_mult_hi, uint64_t stride_mult_lo,
                uint64_t stride_plus_hi, uint64_t stride_plus_lo,
                int64_t n, double phi_a, double mass, float *__restrict__ out) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u128 s = mk128(state_hi, state_lo);
  {
    uint64_t delta = (uint64_t)i + 1;
    for (int b = 0; delta; b++, delta >>= 1)
      if (delta & 1)
        s = mk128(J->mult_hi[b], J->mult_lo[b]) * s +
            mk128(J->plus_hi[b], J->plus_lo[b]);
  }
  const u128 sm = mk128(stride_mult_hi, stride_mult_lo);
  const u128 sp = mk128(stride_plus_hi, stride_plus_lo);
  (void)inc_hi; (void)inc_lo;
  for (; i < n; i += T) {
    // Generator.uniform(0, 1): (next_uint64 >> 11) / 2^53
    const double u = (double)(pcg64_output(s) >> 11) * (1.0 / 9007199254740992.0);
    out[i] = (float)normcdfinv(fma(u, mass, phi_a));
    s = sm * s + sp;
  }
}

// corr[i][j] = exp(c * d_ij), d_ij the geodetic (haversine) distance in km
// of sites i and j (geodetic.py: 2 R asin(sqrt(sin^2(dlat / 2) + cos lat1
// cos lat2 sin^2(dlon / 2)))), the diagonal exactly 1
__global__ void __launch_bounds__(256)
k_corr_matrix(const double *__restrict__ lon, const double *__restrict__ lat,
              int N, double c, double *__restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)N * N) return;
  const int i = (int)(t / N), j = (int)(t - (int64_t)i * N);
  const double D2R = 0.017453292519943295, R = 6371.0;
  const double la1 = lat[i] * D2R, la2 = lat[j] * D2R;
  const double s1 = sin((la1 - la2) * 0.5);
  const double s2 = sin((lon[i] * D2R - lon[j] * D2R) * 0.5);
  const double d = (2.0 * R) * asin(sqrt(s1 * s1 + cos(la1) * cos(la2) * (s2 * s2)));
  out[t] = exp(c * d);
}

// gmf [M][N][E] float32, mean [M][N] float32 (log, or MMI), max_iml / min_iml
// [M] float64, not_mmi [M]: one thread per (site, event)
__global__ void __launch_bounds__(256)
k_gmf_max_min(float *__restrict__ gmf, const float *__restrict__ mean,
              const double *__restrict__ max_iml,
              const double *__restrict__ min_iml,
              const unsigned char *__restrict__ not_mmi, int M, int N, int E) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)N * E) return;
  const int n = (int)(t / E);
  bool all_below = true;
  for (int m = 0; m < M; m++) {
    const int64_t k = (int64_t)m * N * E + t;
    float v = gmf[k];
    if ((double)v > max_iml[m]) {    // capped at the median, gmf.py:76-84
      const float mu = mean[(int64_t)m * N + n];
      v = not_mmi[m] ? expf(mu) : mu;
      gmf[k] = v;
    }
    all_below &= (double)v < min_iml[m];
  }
  if (all_below)
    for (int m = 0; m < M; m++) gmf[(int64_t)m * N * E + t] = 0.f;
}

}  // namespace oqg
