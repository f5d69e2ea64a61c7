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
// Kernels (oqg_postclassical picks one):
//   k_postclassical_reg     R <= 32: keys sorted in registers; 32 < R <= 256:
//                           the same network on keys in shared memory
//   k_postclassical_cta     quantiles of 256 < R <= 4096 realizations: one CTA
//                           per (site, level), cooperative sort and scan
//   k_postclassical_stream  no quantile and R > 256 (or tables that do not
//                           fit): one pass, nothing kept per realization
//   k_postclassical         generic fallback for R <= 256 when the tables of
//                           the register path do not fit in shared memory:
//                           one thread per (site, level), values in local
//                           memory, insertion sort
//   k_mean_rates            calc_mean_rates (calc/mean_rates.py:62-74)
//   k_hazard_maps           compute_hazard_maps
//   k_rate_rows_count /     MapArray.to_array + the chunk grouping of _store
//   k_rate_rows_write       (map_array.py:396-405, 520-540; calculators/
//                           classical.py:58-101): stream compaction of the
//                           rates into (sid, lid, gid, rate) records
#pragma once
#include <math.h>
#include "../../include/oqg.h"
#include "oqg_math.cuh"

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
    } else if (A.kind[s] == OQG_STAT_MAX) {   // stats.max_curve: no weights
      res = v[0];
      for (int r = 1; r < R; r++) res = fmax(res, v[r]);
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

// ---- register path: up to 256 GROUPS of realizations ----------------------
// Realizations that use the same set of gsim columns have the same curve
// (full enumeration with one tectonic region type, sampled logic trees: a few
// distinct sets, many realizations).  The host groups them (oqg_postclassical)
// and the kernel works on the D distinct groups: one exp per group, the
// realization curves are copies of the group values, and for the statistics
// a group d is an element of weight W_d = the sum of its weights whose first
// w_d = min weight matters for the quantiles: numpy sorts equal values by
// weight, so of the interpolation points (cum, value) of a group only the
// first one, at cum + w_d, is reached on a ramp from the previous value; the
// others repeat the value.  With one realization per group (w_d = W_d) this
// is the plain algorithm.
// The values of a thread never leave the SM: accumulated through the map
// group -> gsim columns, written out, folded into the mean / variance sums
// in float64, and packed into 64-bit sort keys
//     [ bits 63..10 of the (non negative) float64 value | rank of w_d (5) |
//       group (5) ]
// (staged through shared memory so that the value loop stays rolled) which
// a bitonic network with compile-time register indices sorts for the quantiles
// (ties of the value are broken by the weight like numpy's sort(order='c');
// values are compared, and used in the interpolation, to 2^-42 relative).
// All the quantiles are read off ONE scan of the sorted groups (the host
// passes them in increasing order).  The variance uses sums shifted by the
// first value, exact for equal curves.
struct PostRegArgs {
  const float *rates;      // [N][L][Gt]
  const int32_t *r_off;    // [R + 1] gsim columns of each realization (group)
  const int32_t *r_gs;
  const double *weights;   // [W][R] (register path: [W][D] group sums W_d)
  const int32_t *wrank;    // [W][R] rank of the weight (of w_d) among them
  // register path only (R = the number of groups D there)
  const double *wmin;      // [W][D] smallest weight of the group
  const int32_t *m_off;    // [D + 1] realizations of each group, as the
  const int64_t *m_pos;    // element offsets r * n_elems of their curves
  int nq;                  // quantiles in increasing order and their statistic
  int qstat[OQG_SMAX];
  double qsorted[OQG_SMAX];
  double wsum[64];
  float *rlz_out, *stat_out;
  int64_t n_elems;
  int L, L1, Gt, R, W, S, use_rates;
  int kind[OQG_SMAX];
  double q[OQG_SMAX];
};

// IB = bits of the realization index (5 for R <= 32, 8 for R <= 256); the
// key keeps bits 63 .. 2 IB of the value
template <int IB>
__device__ __forceinline__ unsigned long long post_key(double v, int wrank,
                                                        int r) {
  // v >= 0 (a probability or a rate): the bit pattern is order preserving
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & ~((1ull << (2 * IB)) - 1)) | ((unsigned)wrank << IB) | (unsigned)r;
}

template <int IB>
__device__ __forceinline__ double post_key_value(unsigned long long k) {
  return __longlong_as_double((long long)(k & ~((1ull << (2 * IB)) - 1)));
}

// RP <= 32: TPB = 128 threads, keys sorted in registers; RP = 64 .. 256: TPB =
// 64 threads, the same network on the keys in shared memory ([RP][TPB],
// conflict free: consecutive threads, consecutive 8-byte words).  A.R is the
// number of groups here.
template <int RP, int TPB>
__global__ void __launch_bounds__(TPB)
k_postclassical_reg(const PostRegArgs A) {
  constexpr int IB = RP <= 32 ? 5 : 8;
  extern __shared__ double sm_post[];
  math_tables_init();   // fast_exp of oqg_math.cuh
  // shared: keys [RP][TPB], group weights and first weights [W][D] (double),
  // wrank [W][D], r_off [D + 1], m_off [D + 1], r_gs, rates [TPB][Gt | 1]
  const int R = A.R, WR = A.W * R;
  unsigned long long *s_key = (unsigned long long *)sm_post + threadIdx.x;
  double2 *s_ww = (double2 *)(sm_post + RP * TPB);   // (W_d, w_d)
  int *s_rank = (int *)(s_ww + WR);
  int *s_off = s_rank + WR;
  int *s_moff = s_off + R + 1;
  int *s_gs = s_moff + R + 1;
  const int64_t e0 = blockIdx.x * (int64_t)blockDim.x;
  for (int i = threadIdx.x; i < WR; i += blockDim.x) {
    s_ww[i] = make_double2(A.weights[i], A.wmin[i]);
    s_rank[i] = A.wrank[i];
  }
  for (int i = threadIdx.x; i <= R; i += blockDim.x) {
    s_off[i] = A.r_off[i];
    s_moff[i] = A.m_off[i];
  }
  __syncthreads();
  const int ngs = s_off[R];
  for (int i = threadIdx.x; i < ngs; i += blockDim.x) s_gs[i] = A.r_gs[i];
  // the rates of the CTA's elements: one coalesced read of TPB * Gt floats,
  // rows padded to an odd stride (conflict-free row reads)
  const int gstride = A.Gt | 1;
  float *s_rates = (float *)(s_gs + ngs);
  {
    const int64_t base = e0 * A.Gt;
    const int64_t lim = min((int64_t)blockDim.x, A.n_elems - e0) * A.Gt;
    // (row, column) of element i advanced without a division
    const int dt = TPB / A.Gt, dg = TPB - dt * A.Gt;
    int t = threadIdx.x / A.Gt, g = threadIdx.x - t * A.Gt;
    for (int i = threadIdx.x; i < (int)lim; i += TPB) {
      s_rates[t * gstride + g] = A.rates[base + i];
      t += dt;
      g += dg;
      if (g >= A.Gt) { g -= A.Gt; t++; }
    }
  }
  __syncthreads();
  const int64_t e = e0 + threadIdx.x;
  if (e >= A.n_elems) return;
  const int l = (int)(e % A.L);
  const int wcol = A.W > 1 ? l / A.L1 : 0;
  const double2 *ww = s_ww + wcol * R;
  const int *rk = s_rank + wcol * R;
  const double wsum = A.wsum[wcol];
  const float *row = s_rates + threadIdx.x * gstride;
  const bool stats = A.stat_out && A.S > 0 && wsum != 0.;
  double sw1 = 0., sw2 = 0., v0 = 0., vmax = 0.;
  float *const out_e = A.rlz_out ? A.rlz_out + e : nullptr;
  int kc = 0, km = 0;      // running positions in the column / member lists
#pragma unroll 1
  for (int r = 0; r < R; r++) {
    double x = 0.;
    // (short trip counts: unrolled by the compiler these two loops cost more
    // in prologue / remainder code than their bodies)
#pragma unroll 1
    for (const int kend = s_off[r + 1]; kc < kend; kc++)
      x += (double)row[s_gs[kc]];
    double v = 1. - fast_exp(fmax(-x, -700.));   // to_probs
    if (out_e) {
      // the curve of every realization of the group (uniform offsets)
      const float vf = (float)v;
      const int mend = s_moff[r + 1];
      if (mend - km >= 8) {           // sampled logic trees: long lists
#pragma unroll 4
        for (; km + 4 <= mend; km += 4) {
          const int64_t p0 = __ldg(A.m_pos + km), p1 = __ldg(A.m_pos + km + 1),
                        p2 = __ldg(A.m_pos + km + 2), p3 = __ldg(A.m_pos + km + 3);
          out_e[p0] = vf; out_e[p1] = vf; out_e[p2] = vf; out_e[p3] = vf;
        }
      }
#pragma unroll 1
      for (; km < mend; km++) out_e[__ldg(A.m_pos + km)] = vf;
    }
    if (A.use_rates) v = to_rate(v);
    if (r == 0) v0 = v;
    vmax = fmax(vmax, v);                       // probabilities, rates: >= 0
    const double d = v - v0;
    const double wr = ww[r].x;
    sw1 = fma(wr, d, sw1);
    sw2 = fma(wr * d, d, sw2);
    s_key[r * TPB] = post_key<IB>(v, rk[r], r);
  }
  if (A.stat_out && A.S > 0 && wsum == 0.) {   // getters.py:61-63
    for (int s = 0; s < A.S; s++) A.stat_out[(int64_t)s * A.n_elems + e] = 0.f;
    return;
  }
  if (!stats) return;
  const double dm = sw1 / wsum;               // mean - v0
  for (int s = 0; s < A.S; s++) {
    double res;
    if (A.kind[s] == OQG_STAT_MEAN) {
      res = v0 + dm;
    } else if (A.kind[s] == OQG_STAT_STD) {
      // sum w (m - v)^2 = sum w d^2 - 2 dm sum w d + dm^2 sum w
      res = sqrt(fmax(sw2 - 2. * dm * sw1 + dm * dm * wsum, 0.));
    } else if (A.kind[s] == OQG_STAT_MAX) {
      res = vmax;
    } else {
      continue;
    }
    if (A.use_rates) res = 1. - exp(-res);
    A.stat_out[(int64_t)s * A.n_elems + e] = (float)res;
  }
  if (A.nq == 0) return;
  if (RP > 32) {
    for (int r = R; r < RP; r++) s_key[r * TPB] = ~0ull;   // padding
    for (int k = 2; k <= RP; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1)
        for (int i = 0; i < RP; i++) {
          const int p = i ^ j;
          if (p > i) {
            const unsigned long long a = s_key[i * TPB], b = s_key[p * TPB];
            if ((a > b) == ((i & k) == 0)) {
              s_key[i * TPB] = b;
              s_key[p * TPB] = a;
            }
          }
        }
  } else {
    // register network, RP <= 32.  The keys are bit patterns of non negative
    // doubles (value bits | rank | group: the order of the patterns is the
    // order of the doubles, denormals included), so a compare-exchange is
    // one DSETP on the FP64 pipe and four selects (sm_100 has no DMNMX:
    // fmin / fmax expand to more than that); the network is the bitonic
    // sorter in its direction-free form (first step of a merge of size k:
    // partner i ^ (k - 1), then half cleaners i ^ j), so no direction logic
    // is executed; fully unrolled: 240 compare-exchanges of 5 instructions
    // for 32 keys.
    constexpr int RR = RP <= 32 ? RP : 4;
    double key[RR];
#pragma unroll
    for (int r = 0; r < RR; r++)     // padding (+inf) sorts last
      key[r] = r < R ? __longlong_as_double((long long)s_key[r * TPB])
                     : __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
    for (int k = 2; k <= RR; k <<= 1) {
#pragma unroll
      for (int i = 0; i < RR; i++) {
        const int p = i ^ (k - 1);
        if (p > i) {
          const double a = key[i], b = key[p];
          const bool sw = a > b;
          key[i] = sw ? b : a;
          key[p] = sw ? a : b;
        }
      }
#pragma unroll
      for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < RR; i++) {
          const int p = i ^ j;
          if (p > i) {
            const double a = key[i], b = key[p];
            const bool sw = a > b;
            key[i] = sw ? b : a;
            key[p] = sw ? a : b;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RR; r++)
      s_key[r * TPB] = (unsigned long long)__double_as_longlong(key[r]);
  }
  // numpy.interp(q, cumsum(w sorted), v sorted) for all the quantiles in one
  // scan of the sorted groups: group j covers (cum, cum + W_j]; up to
  // cum + w_j the curve ramps from the previous value, then it is flat
  int qi = 0;
  double qcur = A.qsorted[0];
  double cum = 0., prev_v = 0.;
#pragma unroll 1
  for (int j = 0; j < R; j++) {
    const unsigned long long kj = s_key[j * TPB];
    const double vk = post_key_value<IB>(kj);
    const double2 wg = ww[(int)(kj & ((1u << IB) - 1))];
    const double c2 = cum + wg.x;
    const bool last = j == R - 1;
    if (qcur < c2 || last) {          // at most nq times in the whole scan
      const double c1 = cum + wg.y;
      do {
        double res = vk;
        if (qcur < c1 && j > 0)
          res = prev_v + (qcur - cum) * ((vk - prev_v) / (c1 - cum));
        if (A.use_rates) res = 1. - exp(-res);
        A.stat_out[(int64_t)A.qstat[qi] * A.n_elems + e] = (float)res;
        if (++qi == A.nq) return;     // the scan is the end of the kernel
        qcur = A.qsorted[qi];
      } while (qcur < c2 || last);
    }
    cum = c2;
    prev_v = vk;
  }
}

// ---- streaming path: no quantile asked, any number of realizations ---------
// mean and standard deviation need no sort: one pass over the realizations
// with the sums shifted by the first value; nothing is kept per realization,
// so R is unlimited (sampled logic trees with thousands of realizations).
__global__ void __launch_bounds__(128)
k_postclassical_stream(const PostRegArgs A) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= A.n_elems) return;
  const int l = (int)(e % A.L);
  const int wcol = A.W > 1 ? l / A.L1 : 0;
  const double *w = A.weights + (int64_t)wcol * A.R;
  const double wsum = A.wsum[wcol];
  const float *row = A.rates + e * A.Gt;
  double sw1 = 0., sw2 = 0., v0 = 0., vmax = 0.;
  for (int r = 0; r < A.R; r++) {
    double x = 0.;
    for (int k = A.r_off[r]; k < A.r_off[r + 1]; k++) x += (double)row[A.r_gs[k]];
    double v = 1. - exp(-x);
    if (A.rlz_out) A.rlz_out[(int64_t)r * A.n_elems + e] = (float)v;
    if (!A.stat_out) continue;
    if (A.use_rates) v = to_rate(v);
    if (r == 0) v0 = v;
    vmax = fmax(vmax, v);
    const double d = v - v0, wr = w[r];
    sw1 = fma(wr, d, sw1);
    sw2 = fma(wr * d, d, sw2);
  }
  if (!A.stat_out) return;
  const double dm = wsum != 0. ? sw1 / wsum : 0.;
  for (int s = 0; s < A.S; s++) {
    double res = A.kind[s] == OQG_STAT_MEAN ? v0 + dm
                 : A.kind[s] == OQG_STAT_MAX
                     ? vmax
                     : sqrt(fmax(sw2 - 2. * dm * sw1 + dm * dm * wsum, 0.));
    if (wsum == 0.) res = 0.;             // getters.py:61-63
    else if (A.use_rates) res = 1. - exp(-res);
    A.stat_out[(int64_t)s * A.n_elems + e] = (float)res;
  }
}

// ---- CTA-cooperative path: quantiles of 256 < R <= 4096 realizations -------
// (sampled logic trees).  One CTA of 256 threads per (site, level), looping
// over the elements: the realization values are packed into sort keys
//     [ bits 63..24 of the float64 value | weight rank (12) | realization (12) ]
// sorted by a cooperative bitonic network (keys in registers, see
// k_postclassical_cta), the weights in sorted order are scanned by the CTA
// and every quantile is found by the thread whose slot brackets it.  Values are
// compared, and used in the interpolation, to 2^-28 relative (the output is
// float32); ties are broken by the weight as in the other paths.
#define OQG_RMAX_CTA 4096
#define POST_CTA_TPB 256
#define POST_CTA_IB 12

__device__ __forceinline__ double post_block_sum(double v, double *s_red,
                                                 bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = s_red[0];
  for (int w = 1; w < POST_CTA_TPB / 32; w++)
    r = is_max ? fmax(r, s_red[w]) : r + s_red[w];
  return r;
}

// One compare-exchange stage of the bitonic network on the keys of a thread
// (sort position i = base + q): the element keeps the minimum when it is
// the lower one of its pair in an ascending run (or the upper one in a
// descending run)
__device__ __forceinline__ unsigned long long cta_pick(
    unsigned long long mine, unsigned long long other, int i, int j, int k) {
  const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
  return keep_min ? (mine < other ? mine : other)
                  : (mine > other ? mine : other);
}

template <int PER, int J>
__device__ __forceinline__ void cta_reg_stage(unsigned long long (&key)[PER],
                                              int base, int k) {
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int p = q ^ J;
    if (p > q) {
      const unsigned long long a = key[q], b = key[p];
      const bool sw = (a > b) == (((base + q) & k) == 0);
      key[q] = sw ? b : a;
      key[p] = sw ? a : b;
    }
  }
}

// PER = RP / 256 keys per thread, thread t owning the sort positions
// t PER .. t PER + PER - 1: the stages with j < PER exchange registers, those
// with j < 32 PER exchange between the lanes of a warp (shuffles), only the
// six stages with j >= 32 PER per sort go through shared memory.
template <int PER>
__global__ void __launch_bounds__(POST_CTA_TPB)
k_postclassical_cta(const PostRegArgs A) {
  constexpr int RP = PER * POST_CTA_TPB;
  extern __shared__ double sm_post[];
  __shared__ double s_red[POST_CTA_TPB / 32];
  __shared__ double s_part[POST_CTA_TPB];
  unsigned long long *s_key = (unsigned long long *)sm_post;   // [RP]
  double *s_cum = sm_post + RP;                                // [RP]
  const int R = A.R, tid = threadIdx.x;
  constexpr int per = PER;
  for (int64_t e = blockIdx.x; e < A.n_elems; e += gridDim.x) {
    const int l = (int)(e % A.L);
    const int wcol = A.W > 1 ? l / A.L1 : 0;
    const double *w = A.weights + (int64_t)wcol * R;
    const int *rk = A.wrank + (int64_t)wcol * R;
    const double wsum = A.wsum[wcol];
    const float *row = A.rates + e * A.Gt;
    // the value of realization 0 shifts the sums (exact for equal curves)
    double v0 = 0.;
    for (int k = A.r_off[0]; k < A.r_off[1]; k++) v0 += (double)row[A.r_gs[k]];
    v0 = 1. - exp(-v0);
    if (A.use_rates) v0 = to_rate(v0);
    double sw1 = 0., sw2 = 0., vmax = 0.;
    unsigned long long key[PER];
    __syncthreads();                       // the previous element is done
#pragma unroll
    for (int q = 0; q < PER; q++) {
      const int r = tid + q * POST_CTA_TPB;
      key[q] = ~0ull;                      // padding sorts last
      if (r >= R) continue;
      double x = 0.;
      for (int k = A.r_off[r]; k < A.r_off[r + 1]; k++)
        x += (double)row[A.r_gs[k]];
      double v = 1. - exp(-x);
      if (A.rlz_out) A.rlz_out[(int64_t)r * A.n_elems + e] = (float)v;
      if (A.use_rates) v = to_rate(v);
      vmax = fmax(vmax, v);
      const double d = v - v0, wr = w[r];
      sw1 = fma(wr, d, sw1);
      sw2 = fma(wr * d, d, sw2);
      key[q] = post_key<POST_CTA_IB>(v, rk[r], r);
    }
    if (!A.stat_out || A.S == 0) continue;
    sw1 = post_block_sum(sw1, s_red, false);
    sw2 = post_block_sum(sw2, s_red, false);
    vmax = post_block_sum(vmax, s_red, true);
    if (wsum == 0.) {                      // getters.py:61-63
      if (tid < A.S) A.stat_out[(int64_t)tid * A.n_elems + e] = 0.f;
      continue;
    }
    const double dm = sw1 / wsum;
    bool sorted = false;
    for (int s = 0; s < A.S; s++) {
      if (A.kind[s] != OQG_STAT_QUANTILE) {
        if (tid == 0) {
          double res = A.kind[s] == OQG_STAT_MEAN ? v0 + dm
                       : A.kind[s] == OQG_STAT_MAX
                           ? vmax
                           : sqrt(fmax(sw2 - 2. * dm * sw1 + dm * dm * wsum, 0.));
          if (A.use_rates) res = 1. - exp(-res);
          A.stat_out[(int64_t)s * A.n_elems + e] = (float)res;
        }
        continue;
      }
      if (!sorted) {
        const int base = tid * PER;
#pragma unroll 1
        for (int k = 2; k <= RP; k <<= 1) {
#pragma unroll 1
          for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32 * PER) {
              // partner in another warp: through shared memory, [q][thread]
              __syncthreads();
#pragma unroll
              for (int q = 0; q < PER; q++) s_key[q * POST_CTA_TPB + tid] = key[q];
              __syncthreads();
              const int pt = tid ^ (j / PER);
#pragma unroll
              for (int q = 0; q < PER; q++)
                key[q] = cta_pick(key[q], s_key[q * POST_CTA_TPB + pt],
                                  base + q, j, k);
            } else if (j >= PER) {
              const int lm = j / PER;      // partner lane
#pragma unroll
              for (int q = 0; q < PER; q++) {
                const unsigned long long other =
                    __shfl_xor_sync(0xffffffffu, key[q], lm);
                key[q] = cta_pick(key[q], other, base + q, j, k);
              }
            } else {
              switch (j) {
                case 8: if (PER > 8) cta_reg_stage<PER, (PER > 8 ? 8 : 1)>(key, base, k); break;
                case 4: if (PER > 4) cta_reg_stage<PER, (PER > 4 ? 4 : 1)>(key, base, k); break;
                case 2: if (PER > 2) cta_reg_stage<PER, (PER > 2 ? 2 : 1)>(key, base, k); break;
                default: cta_reg_stage<PER, 1>(key, base, k); break;
              }
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) s_key[base + q] = key[q];
        __syncthreads();
        // inclusive scan of the weights in sorted order: `per` consecutive
        // slots per thread, then the partial sums of the threads
        double acc = 0.;
        for (int j = tid * per; j < (tid + 1) * per; j++) {
          if (j < R) acc += w[(int)(s_key[j] & ((1u << POST_CTA_IB) - 1))];
          s_cum[j] = acc;
        }
        s_part[tid] = acc;
        __syncthreads();
        if (tid < 32) {          // exclusive scan of 256 partials by one warp
          double run = 0.;
          for (int c = 0; c < POST_CTA_TPB / 32; c++) {
            const int i = c * 32 + tid;
            double v = s_part[i];
            for (int o = 1; o < 32; o <<= 1) {
              const double t = __shfl_up_sync(0xffffffffu, v, o);
              if (tid >= o) v += t;
            }
            s_part[i] = run + v - s_part[i];
            run += __shfl_sync(0xffffffffu, v, 31);
          }
        }
        __syncthreads();
        const double before = s_part[tid];
        for (int j = tid * per; j < (tid + 1) * per; j++) s_cum[j] += before;
        __syncthreads();
        sorted = true;
      }
      // numpy.interp(q, cumsum(w sorted), v sorted): the first slot j with
      // q < cum[j], or the last one when q >= the total weight
      const double q = A.q[s];
      for (int j = tid; j < R; j += POST_CTA_TPB) {
        const double cum = s_cum[j];
        const bool lt = q < cum;
        const bool here = lt ? (j == 0 || !(q < s_cum[j - 1])) : j == R - 1;
        if (here) {
          const double vk = post_key_value<POST_CTA_IB>(s_key[j]);
          double res = vk;
          if (lt && j > 0) {
            const double pc = s_cum[j - 1],
                         pv = post_key_value<POST_CTA_IB>(s_key[j - 1]);
            res = pv + (q - pc) * ((vk - pv) / (cum - pc));
          }
          if (A.use_rates) res = 1. - exp(-res);
          A.stat_out[(int64_t)s * A.n_elems + e] = (float)res;
        }
      }
    }
  }
}

// calc_mean_rates, calc/mean_rates.py:62-74: rates[N][L][Gt] (float32) times
// the weights of the gsim columns, one weight vector or one per IMT;
// float64 out [N][L]: a pure HBM stream.  A CTA copies the rows of its
// elements to shared memory as one contiguous, coalesced range, then one
// thread per element sums its row; the column loop of thread r starts at
// column r % Gt and wraps, which spreads the rows (stride Gt words) over the
// banks.  (One thread reading its own row straight from global memory runs
// at 17 % of the HBM peak for Gt = 32: 32 different lines per request.)
#define MR_TPB 256
#define MR_FLOATS 12288           // 48 KB of rows per CTA
__global__ void __launch_bounds__(MR_TPB)
k_mean_rates(const float *__restrict__ rates, int64_t n_elems, int L, int L1,
             int Gt, int rows, const double *__restrict__ weights, int W,
             double *__restrict__ out) {
  __shared__ float s_row[MR_FLOATS];
  const int64_t e0 = blockIdx.x * (int64_t)rows;
  const int nr = (int)min((int64_t)rows, n_elems - e0);
  const float *src = rates + e0 * Gt;
  for (int i = threadIdx.x; i < nr * Gt; i += MR_TPB) s_row[i] = src[i];
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= nr) return;
  const int64_t e = e0 + r;
  const double *w = weights + (int64_t)(W > 1 ? (int)(e % L) / L1 : 0) * Gt;
  const float *row = s_row + r * Gt;
  double acc = 0.;
  int g = r % Gt;
  for (int k = 0; k < Gt; k++) {
    acc = fma((double)row[g], w[g], acc);
    if (++g == Gt) g = 0;
  }
  out[e] = acc;
}

// more gsim columns than a CTA can stage: one thread per element
__global__ void __launch_bounds__(256)
k_mean_rates_direct(const float *__restrict__ rates, int64_t n_elems, int L,
                    int L1, int Gt, const double *__restrict__ weights, int W,
                    double *__restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const double *w = weights + (int64_t)(W > 1 ? (int)(e % L) / L1 : 0) * Gt;
  const float *row = rates + e * Gt;
  double acc = 0.;
  for (int g = 0; g < Gt; g++) acc = fma((double)row[g], w[g], acc);
  out[e] = acc;
}

// compute_hazard_maps, map_array.py:94-141.  curves [K][M][L1] float32, imls
// [M][L1], out [K][M][P].  One thread per (curve, imt) row and all its P
// poes; a CTA copies its rows to shared memory as one contiguous, coalesced
// range (odd row stride: conflict free) and the logs of the levels once.
// The bracketing levels are found by comparing the values themselves
// (log is increasing and the curve values are float32: log(max(c, EPSILON))
// <= log(poe) exactly when max(c, EPSILON) <= poe), so only the two
// bracketing values go through log: 2 log + 1 exp per (row, poe) instead of
// 2 L1 logs, and the kernel is an HBM stream.
#define HM_TPB 256
__global__ void __launch_bounds__(HM_TPB)
k_hazard_maps(const float *__restrict__ curves, int64_t rows_total, int M,
              int L1, int rows, const double *__restrict__ imls, int P,
              const double *__restrict__ poes, float *__restrict__ out) {
  extern __shared__ double sm_hm[];
  double *s_li = sm_hm;                       // [M][L1] log of the levels
  double *s_p = s_li + M * L1;                // [P] poes, [P] log poes
  float *s_c = (float *)(s_p + 2 * P);        // [rows][L1 | 1]
  const int stride = L1 | 1;
  const int64_t r0 = blockIdx.x * (int64_t)rows;
  const int nr = (int)min((int64_t)rows, rows_total - r0);
  for (int i = threadIdx.x; i < M * L1; i += HM_TPB) s_li[i] = log(imls[i]);
  for (int i = threadIdx.x; i < P; i += HM_TPB) {
    s_p[i] = poes[i];
    s_p[P + i] = log(poes[i]);
  }
  const float *src = curves + r0 * L1;
  {
    // (row, level) of element i advanced without a division
    const int dr = HM_TPB / L1, dl = HM_TPB - dr * L1;
    int r = threadIdx.x / L1, l = threadIdx.x - r * L1;
    for (int i = threadIdx.x; i < nr * L1; i += HM_TPB) {
      s_c[r * stride + l] = src[i];
      r += dr;
      l += dl;
      if (l >= L1) { l -= L1; r++; }
    }
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= nr) return;
  const int64_t km = r0 + r;
  const int m = (int)(km % M);
  const float *c = s_c + r * stride;
  const double *li = s_li + m * L1;
  const double EPS = 1E-30;
  // reversed curve: x[j] = log(max(c[L1 - 1 - j], EPSILON)), increasing in j
#define CV(j) fmax((double)c[L1 - 1 - (j)], EPS)
  for (int p = 0; p < P; p++) {
    const double poe = s_p[p], lp = s_p[P + p];
    double res = 0.;
    if (!(poe > CV(L1 - 1))) {
      if (poe < CV(0)) {
        res = exp(li[L1 - 1]);
      } else {
        // numpy.interp: the last j with x[j] <= lp.  For a float32 value c,
        // max(c, EPSILON) <= poe is c <= pf with pf the largest float32 not
        // above poe (and EPSILON <= poe): a float32 comparison per level
        int j = 0;
        if (EPS <= poe) {
          const float pf = __double2float_rd(poe);
          for (int t = 1; t < L1; t++)
            if (c[L1 - 1 - t] <= pf) j = t;
        }
        const double y0 = li[L1 - 1 - j];
        if (j == L1 - 1) {
          res = exp(y0);
        } else {
          const double x0 = log(CV(j)), x1 = log(CV(j + 1));
          const double y1 = li[L1 - 2 - j];
          res = x0 == lp ? exp(y0) : exp(y0 + (lp - x0) * ((y1 - y0) / (x1 - x0)));
        }
      }
    }
    out[km * P + p] = (float)res;
  }
#undef CV
}

// curves too long for the shared memory of a CTA: one thread per
// (curve, imt, poe) reading global memory
__global__ void k_hazard_maps_direct(const float *__restrict__ curves, int64_t K,
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


// ---- rates map -> records (sid, lid, gid, rate) ------------------------------
// MapArray.to_array (map_array.py:396-405 with to_rates :520-540): for every
// gsim column g, site s and level l with rate > 0 one record of rates_dt
// (u4 sid, u2 lid, u2 gid, f4 rate: 12 bytes), ordered by (g, s, l); _store
// (calculators/classical.py:58-101) then regroups them by chunk = sid %
// num_chunks (gen_chunks, map_array.py:49-64) keeping that order inside a
// chunk.  Here a "row" is one (g, s) curve of L levels; the host passes the
// rows in output order (by chunk, then g, then s), one warp counts / writes
// one row with ballots, and an exclusive scan of the counts (CUB) gives the
// offsets: the records come out in the reference's order.
struct RateRecord {
  uint32_t sid;
  uint16_t lid, gid;
  float rate;
};
static_assert(sizeof(RateRecord) == 12, "rates_dt is 12 bytes");

__global__ void __launch_bounds__(128)
k_rate_rows_count(const float *__restrict__ rates, int N, int L, int G,
                  const int64_t *__restrict__ rows, int64_t R,
                  int64_t *__restrict__ counts) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int64_t row = rows ? rows[r] : r;       // g * N + s
  const int g = (int)(row / N);
  const int64_t sx = row - (int64_t)g * N;
  const float *c = rates + (sx * L) * G + g;
  int n = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const bool ok = l < L && c[(int64_t)l * G] > 0.f;
    n += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) counts[r] = n;
}

__global__ void __launch_bounds__(128)
k_rate_rows_write(const float *__restrict__ rates, int N, int L, int G,
                  const uint32_t *__restrict__ sids,
                  const int32_t *__restrict__ gid, int level0,
                  const int64_t *__restrict__ rows, int64_t R,
                  const int64_t *__restrict__ offsets,
                  RateRecord *__restrict__ out) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int64_t row = rows ? rows[r] : r;
  const int g = (int)(row / N);
  const int64_t sx = row - (int64_t)g * N;
  const float *c = rates + (sx * L) * G + g;
  int64_t at = offsets[r];
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const float v = l < L ? c[(int64_t)l * G] : 0.f;
    const bool ok = v > 0.f;
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      RateRecord rec;
      rec.sid = sids ? sids[sx] : (uint32_t)sx;
      rec.lid = (uint16_t)(level0 + l);
      rec.gid = (uint16_t)gid[g];
      rec.rate = v;
      out[at + __popc(b & ((1u << lane) - 1u))] = rec;
    }
    at += __popc(b);
  }
}

}  // namespace oqg
