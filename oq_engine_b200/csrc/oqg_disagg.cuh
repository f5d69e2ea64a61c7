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

// ---------------------------------------------------------------------------
// Conditional spectrum: _cs_out (openquake/hazardlib/calc/cond_spectra.py:
// 35-85, Lin et al. 2013 eq. 13-15) for one site and one gsim.  Block
// (p, j): j < M -> the weighted sums of IMT j (mean term and, given the
// previously computed mean spectrum, the variance term), j == M -> the sum
// of the weights.  Deterministic tree reduction over the U ruptures.
namespace oqg {

struct CsArgs {
  const double *mu, *sig;   // [M][U]
  const double *probs;      // [U] probability of one or more occurrences
  const double *rho;        // [M]
  const double *imls;       // [P] levels of the conditioning IMT
  const double *cs_poes;    // [P]
  const double *prev;       // [M][P] previous c[m, 1, p] or null
  double *out;              // [M][3][P]
  int64_t U;
  int M, P, imti;
  double phi_b, inv_norm, invtime;
};

__global__ void __launch_bounds__(256) k_cs_out(const CsArgs A) {
  __shared__ double r1[256], r2[256];
  const int p = blockIdx.x, j = blockIdx.y;
  const double liml = log(A.imls[p]);
  const double norm = -log(1. - A.cs_poes[p]) / A.invtime;
  const double rho = j < A.M ? A.rho[j] : 0.;
  const double prev = (j < A.M && A.prev) ? A.prev[j * A.P + p] : 0.;
  double s1 = 0., s2 = 0.;
  for (int64_t u = threadIdx.x; u < A.U; u += blockDim.x) {
    const double eps = (liml - A.mu[(int64_t)A.imti * A.U + u]) /
                       A.sig[(int64_t)A.imti * A.U + u];
    const double poe = truncnorm_sf(A.phi_b, A.inv_norm, eps);
    double ws = -log(pow(1. - A.probs[u], poe)) / A.invtime;
    ws /= norm;
    if (j == A.M) {
      s1 += ws;
    } else {
      const double sg = A.sig[(int64_t)j * A.U + u];
      const double term1 = A.mu[(int64_t)j * A.U + u] + rho * eps * sg;
      s1 += ws * term1;
      if (A.prev) {
        const double term2 = sg * sqrt(1. - rho * rho);
        const double term3 = term1 - prev;
        s2 += ws * (term2 * term2 + term3 * term3);
      }
    }
  }
  r1[threadIdx.x] = s1;
  r2[threadIdx.x] = s2;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      r1[threadIdx.x] += r1[threadIdx.x + k];
      r2[threadIdx.x] += r2[threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (j == A.M) {
      for (int m = 0; m < A.M; m++) A.out[(m * 3 + 0) * A.P + p] = r1[0];
    } else {
      A.out[(j * 3 + 1) * A.P + p] = r1[0];
      A.out[(j * 3 + 2) * A.P + p] = A.prev ? r2[0] : 0.;
    }
  }
}

}  // namespace oqg

// ---------------------------------------------------------------------------
// Ground motion fields: GmfComputer._compute
// (openquake/hazardlib/calc/gmf.py:476-525), float32 like the reference
// (mean_stds.astype(F32), float32 epsilons).  One thread per (IMT, site,
// event); HBM bound: reads the epsilon and writes the value.
//   mode 0  both truncation levels zero: exp(mean)
//   mode 1  gsim with total standard deviation only: exp(mean + sig * eps)
//   mode 2  exp(mean + phi * within_eps + tau * between_eps)
//   mode 3  exp(mean + within_res + tau * between_eps), within_res the
//           residuals after SpatialCorrelationModel.apply_correlation
namespace oqg {

__global__ void k_gmf(const float *__restrict__ ms,      // [4][M][N]
                      const float *__restrict__ within,  // [M][N][E]
                      const float *__restrict__ between, // [E][M]
                      const unsigned char *__restrict__ not_mmi,  // [M]
                      int M, int N, int E, int mode,
                      float *__restrict__ out) {         // [M][N][E]
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)M * N * E) return;
  const int e = (int)(t % E);
  const int64_t mn = t / E;
  const int m = (int)(mn / N);
  const int64_t MN = (int64_t)M * N;
  float v = ms[mn];
  if (mode == 1) {
    v = v + ms[MN + mn] * within[t];
  } else if (mode == 2) {
    // (mean + within_res) + between_res, each product rounded to float32
    v = (v + ms[3 * MN + mn] * within[t]) + ms[2 * MN + mn] * between[e * M + m];
  } else if (mode == 3) {
    // `within` holds the spatially correlated within-event residuals
    v = (v + within[t]) + ms[2 * MN + mn] * between[e * M + m];
  }
  out[t] = not_mmi[m] ? expf(v) : v;
}

}  // namespace oqg
