// Disaggregation on the device: the function
// the reference marks "the crucial bit for performance", _disaggregate
// (openquake/hazardlib/calc/disagg.py:188-243, Poissonian lane) with
// _disagg_eps (:246-256) and _build_disagg_matrix (:262-296, _digitize_lons
// :311-337): for the U context rows of one site and one magnitude bin, the
// PoE of each (IMT, PoE level) is split over the epsilon bins, turned into a
// probability of no exceedance, and multiplied into the
// (distance, lon, lat, epsilon, IMT, level) matrix.
//
// One thread per (row, IMT, level): log pne = -rate * gp * poe * T is added
// to the bin with a float64 atomic (the product of the reference becomes a
// sum of logarithms; the order of the atomics moves the result by ~1e-16),
// a second kernel applies 1 - exp().  The matrix is small (10^3 - 10^5
// doubles, L2 resident); rows are many (10^4 - 10^6 per site and bin).
#pragma once
#include <math.h>
#include "../../include/oqg.h"
#include "oqg_poe.cuh"

namespace oqg {

#define OQG_DISAGG_EMAX 32

struct DisaggArgs {
  const double *rrup, *clon, *clat, *rate;   // [U]
  const double *mea, *std;                   // [M][U]
  const double *iml2;                        // [M][P] log intensities
  const double *dist_edges, *lon_edges, *lat_edges;
  int64_t U;
  int M, P, D, Lo, La, E;
  int lon_idl;                               // lon bins cross the date line
  int epsstar;
  double gp, itime, phi_b, inv_norm, min_eps, max_eps;
  double eps_edges[OQG_DISAGG_EMAX + 1];
  double bands[OQG_DISAGG_EMAX], cum[OQG_DISAGG_EMAX + 1];
  double *acc;                               // [D][Lo][La][E][M][P]
};

// numpy.digitize(x, bins) - 1 for increasing bins (right = False)
__device__ __forceinline__ int digitize_m1(double x, const double *bins, int n) {
  int c = 0;
  for (int i = 0; i < n; i++) c += bins[i] <= x;
  return c - 1;
}

// geo/utils.py:440-448
__device__ __forceinline__ double lon_extent(double lon1, double lon2) {
  double d = lon2 - lon1 + 180.;
  d = d - 360. * floor(d / 360.);           // python's %
  return d - 180.;
}

__global__ void k_disagg_accumulate(const DisaggArgs A) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int MP = A.M * A.P;
  if (t >= A.U * MP) return;
  const int64_t u = t / MP;
  const int mp = (int)(t - u * MP);
  const int m = mp / A.P;
  const double iml = A.iml2[mp];
  if (isinf(iml) && iml < 0.) return;        // zero hazard, disagg.py:217
  // bins: closed on the lower bound; the last edge belongs to the last bin;
  // an index of -1 (below the first edge) wraps to the last bin like the
  // numpy indexing of the reference
  int di = digitize_m1(A.rrup[u], A.dist_edges, A.D + 1);
  int la = digitize_m1(A.clat[u], A.lat_edges, A.La + 1);
  int lo;
  const double lon = A.clon[u];
  if (A.lon_idl) {
    lo = 0;
    for (int i = 0; i < A.Lo; i++) {
      bool ok = lon_extent(lon, A.lon_edges[i + 1]) > 0.;
      if (i != 0) ok = ok && lon_extent(A.lon_edges[i], lon) >= 0.;
      if (ok) lo = i;
    }
  } else {
    lo = digitize_m1(lon, A.lon_edges, A.Lo + 1);
  }
  if (di == A.D) di = A.D - 1;
  if (lo == A.Lo) lo = A.Lo - 1;
  if (la == A.La) la = A.La - 1;
  if (di < 0) di += A.D;
  if (lo < 0) lo += A.Lo;
  if (la < 0) la += A.La;
  const double lvl = (iml - A.mea[(int64_t)m * A.U + u]) /
                     A.std[(int64_t)m * A.U + u];
  int idx = 0;                               // numpy.searchsorted, side left
  for (int i = 0; i <= A.E; i++) idx += A.eps_edges[i] < lvl;
  const double sf = truncnorm_sf(A.phi_b, A.inv_norm, lvl);
  const double k = -A.rate[u] * A.itime * A.gp;
  double *dst = A.acc + ((((int64_t)di * A.Lo + lo) * A.La + la) * A.E) * MP + mp;
  if (A.epsstar) {
    if (lvl >= A.min_eps && lvl < A.max_eps) {
      const int e = idx - 1 < 0 ? idx - 1 + A.E : idx - 1;
      atomicAdd(dst + (int64_t)e * MP, k * sf);
    }
    return;
  }
  for (int e = 0; e < A.E; e++) {
    double poe = 0.;
    if (idx <= e) poe = A.bands[e];
    else if (idx == e + 1) poe = sf - A.cum[idx];
    if (poe != 0.) atomicAdd(dst + (int64_t)e * MP, k * poe);
  }
}

__global__ void k_disagg_finalize(double *acc, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) acc[i] = 1. - exp(acc[i]);
}

}  // namespace oqg
