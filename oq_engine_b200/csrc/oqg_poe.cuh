// Truncated-normal exceedance probabilities and probabilities of no
// exceedance as device functions.
//
// Reference: openquake/hazardlib/stats.py:27-73 (ndtr, truncnorm_sf),
// openquake/hazardlib/contexts.py:482-489 (_set_poes) and
// openquake/hazardlib/tom.py:185-217 (get_pnes).
#pragma once
#include <math.h>

namespace oqg {

#define OQG_SQRT05 0.70710678118654752440

// stats.py:31-33
__device__ __forceinline__ double ndtr(double z) {
  return 0.5 * (1.0 + erf(z * OQG_SQRT05));
}

// stats.py:38-73 with inv_norm = 1 / (2 phi_b - 1); numpy's clip keeps NaN
__device__ __forceinline__ double truncnorm_sf(double phi_b, double inv_norm,
                                               double z) {
  double sf = (phi_b - ndtr(z)) * inv_norm;
  if (sf < 0.) sf = 0.;
  if (sf > 1.) sf = 1.;
  return sf;
}

// same quantity through erfc, which keeps full relative accuracy in the
// upper tail: phi_b - ndtr(z) = (erfc(z/sqrt2) - erfc(tl/sqrt2)) / 2
// with half_erfc_tl = erfc(tl / sqrt 2) / 2 = 1 - phi_b
__device__ __forceinline__ double truncnorm_sf_erfc(double half_erfc_tl,
                                                    double inv_norm,
                                                    double z) {
  double sf = (0.5 * erfc(z * OQG_SQRT05) - half_erfc_tl) * inv_norm;
  if (sf < 0.) sf = 0.;
  if (sf > 1.) sf = 1.;
  return sf;
}

// tom.py:185-217 for a non-Poissonian rupture: sum_k p_k (1 - poe)^k clipped
__device__ __forceinline__ double pne_from_pmf(const double *probs, int K,
                                               double poe) {
  double q = 1. - poe, qk = 1., pne = probs[0];
  for (int k = 1; k < K; k++) {
    qk *= q;
    pne += probs[k] * qk;
  }
  if (pne < 0.) pne = 0.;
  if (pne > 1.) pne = 1.;
  return pne;
}

}  // namespace oqg
