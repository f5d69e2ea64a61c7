// Post-processing of the rates map on the device (SURVEY.md section 8f row
// 3): hazard curves per realization, their statistics over the logic-tree
// realizations, and hazard maps.
//
// Reference: MapGetter.get_hcurve (calculators/getters.py:351-364),
// build_stat_curve (calculators/getters.py:47-76), mean_curve / std_curve /
// quantile_curve (hazardlib/stats.py:156-214), to_rates / to_probs
// (hazardlib/calc/mean_rates.py:28-50), compute_hazard_maps
// (hazardlib/map_array.py:94-141), driven by postclassical
// (calculators/classical.py:258-325).
//
// HBM-bound: one thread per (site, level) reads Gt rates and writes R + S
// float32 values; the realization values of the thread live in local memory
// (R <= OQG_RMAX doubles), sorted once for all the quantiles.
#pragma once
#include <math.h>
#include "../../include/oqg.h"

namespace oqg {

#define OQG_RMAX 256     // realizations per call (local array per thread)
#define OQG_SMAX 16      // statistics per call

struct PostArgs {
  const float *rates;    // [N][L][Gt]
  const int32_t *g_off;  // [Gt + 1]
  const int32_t *g_rlz;  // realizations of each gsim column
  const double *weights; // [W][R], W = 1 or M
  double wsum[64];       // sum of the weights of column w (W <= 64)
  float *rlz_out;        // [R][N][L] or null
  float *stat_out;       // [S][N][L] or null
  int64_t n_elems;       // N * L
  int L, L1, Gt, R, W, S, use_rates;
  int kind[OQG_SMAX];
  double q[OQG_SMAX];
};

// mean_rates.py:28-41 with itime = 1, minrate = 0
__device__ __forceinline__ double to_rate(double prob) {
  double pne = 1. - prob;
  if (pne == 0.) pne = 1E-45;
  double rate = -log(pne);
  if (rate < 1E-12) rate = 0.;
  if (rate > 100.) rate = 100.;
  return rate;
}

__global__ void __launch_bounds__(128)
k_postclassical(const PostArgs A) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.n_elems) return;
  const int l = (int)(e % A.L);
  const int R = A.R;
  double v[OQG_RMAX];
  unsigned short idx[OQG_RMAX];
  for (int r = 0; r < R; r++) v[r] = 0.;
  // get_hcurve: the rates of gsim column g go to every realization using it
  const float *row = A.rates + e * A.Gt;
  for (int g = 0; g < A.Gt; g++) {
    const double rate = (double)row[g];
    if (rate == 0.) continue;
    for (int k = A.g_off[g]; k < A.g_off[g + 1]; k++) v[A.g_rlz[k]] += rate;
  }
  for (int r = 0; r < R; r++) v[r] = 1. - exp(-v[r]);      // to_probs
  if (A.rlz_out)
    for (int r = 0; r < R; r++) A.rlz_out[(int64_t)r * A.n_elems + e] = (float)v[r];
  if (!A.stat_out || A.S == 0) return;
  const int wcol = A.W > 1 ? l / A.L1 : 0;
  const double *w = A.weights + (int64_t)wcol * R;
  const double wsum = A.wsum[wcol];
  if (wsum == 0.) {   // no data expected for this IMT, getters.py:61-63
    for (int s = 0; s < A.S; s++) A.stat_out[(int64_t)s * A.n_elems + e] = 0.f;
    return;
  }
  if (A.use_rates)
    for (int r = 0; r < R; r++) v[r] = to_rate(v[r]);
  // mean (numpy.average) and standard deviation
  double mean = 0.;
  for (int r = 0; r < R; r++) mean += w[r] * v[r];
  mean /= wsum;
  bool sorted = false;
  for (int s = 0; s < A.S; s++) {
    double res;
    if (A.kind[s] == OQG_STAT_MEAN) {
      res = mean;
    } else if (A.kind[s] == OQG_STAT_STD) {
      double acc = 0.;
      for (int r = 0; r < R; r++) acc += w[r] * ((mean - v[r]) * (mean - v[r]));
      res = sqrt(acc);
    } else {
      if (!sorted) {
        // insertion sort of the realizations by (value, weight): numpy's
        // sort(order='c') breaks ties with the remaining field
        for (int r = 0; r < R; r++) {
          const double vr = v[r], wr = w[r];
          int j = r;
          while (j > 0) {
            const int k = idx[j - 1];
            if (v[k] < vr || (v[k] == vr && w[k] <= wr)) break;
            idx[j] = idx[j - 1];
            j--;
          }
          idx[j] = (unsigned short)r;
        }
        sorted = true;
      }
      // numpy.interp(q, cumsum(w sorted), v sorted)
      const double q = A.q[s];
      double cum = 0., prev_cum = 0., prev_v = 0.;
      res = v[idx[R - 1]];
      for (int j = 0; j < R; j++) {
        const int k = idx[j];
        cum += w[k];
        if (q < cum || (j == R - 1 && q == cum)) {
          if (j == 0 || q == cum) res = v[k];
          else res = prev_v + (q - prev_cum) * ((v[k] - prev_v) / (cum - prev_cum));
          break;
        }
        prev_cum = cum;
        prev_v = v[k];
      }
    }
    if (A.use_rates) res = 1. - exp(-res);
    A.stat_out[(int64_t)s * A.n_elems + e] = (float)res;
  }
}

// compute_hazard_maps, map_array.py:94-141: one thread per (curve, imt, poe)
// curves [K][M][L1] float32, imls [M][L1], out [K][M][P]
__global__ void k_hazard_maps(const float *__restrict__ curves, int64_t K,
                              int M, int L1, const double *__restrict__ imls,
                              int P, const double *__restrict__ poes,
                              float *__restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= K * M * P) return;
  const int p = (int)(e % P);
  const int64_t km = e / P;
  const int m = (int)(km % M);
  const float *c = curves + km * L1;
  const double *im = imls + (int64_t)m * L1;
  const double lp = log(poes[p]);
  // reversed curve: x[j] = log(max(c[L1 - 1 - j], EPSILON)), increasing in j
  const double EPS = 1E-30;
#define LC(j) log(fmax((double)c[L1 - 1 - (j)], EPS))
#define LI(j) log(im[L1 - 1 - (j)])
  double res = 0.;
  if (!(lp > LC(L1 - 1))) {
    // numpy.interp: the last j with x[j] <= lp
    if (lp < LC(0)) {
      res = exp(LI(0));
    } else {
      int j = 0;
      for (int t = 1; t < L1; t++)
        if (LC(t) <= lp) j = t;
      if (j == L1 - 1) {
        res = exp(LI(j));
      } else {
        const double x0 = LC(j), x1 = LC(j + 1);
        const double y0 = LI(j), y1 = LI(j + 1);
        res = x0 == lp ? exp(y0) : exp(y0 + (lp - x0) * ((y1 - y0) / (x1 - x0)));
      }
    }
  }
#undef LC
#undef LI
  out[e] = (float)res;
}

}  // namespace oqg
