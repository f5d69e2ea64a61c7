// liboqg: kernels and C ABI of the B200-native classical-PSHA hot path.
// See include/oqg.h for the boundary and DESIGN.md for the layout.
//
// Kernels (all FP64 arithmetic for distances / GMPE / PoE like the
// reference; float32 only where the reference stores float32):
//   k_prep_ruptures  planar records -> prepared rupture SoA [NF_RUP][U]
//   k_prep_sites     lon/lat/depth -> site SoA [NF_SITE][N]
//   k_distances      (U, N) distance planes (unfused API, get_ctxs)
//   k_mean_stds<F>   context rows -> (4, G, M, P) (unfused API)
//   k_poes           (4, G, M, P) -> (P, L, G) float32 (unfused API)
//   k_update         fold (P, L, G) PoEs into the map, one thread per map
//                    element looping over the rows of its site (no atomics)
//   k_fused<F,NB>    ruptures x sites -> partial sums of rate*poe*T per
//                    (site, level) with distances, GMPE and PoEs in
//                    registers / shared memory, nothing materialised
//   k_finalize       partial sums -> float32 pne / rate map
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <mutex>

#include "../../include/oqg.h"
#include "oqg_gmpe.cuh"
#include "oqg_geo.cuh"
#include "oqg_poe.cuh"
#include "oqg_fast.cuh"
#include "oqg_source.cuh"
#include "oqg_post.cuh"
#include "oqg_disagg.cuh"
#include "oqg_gmf.cuh"

using namespace oqg;

#define OQG_MAX_SLOTS 96
#define OQG_MAX_G 16
#define OQG_MAX_LEVELS 1024
#define NL 20          // level accumulators per thread in the fused kernel
#ifndef RT
#define RT 16          // ruptures per shared-memory tile
#endif
#ifndef NLR
#define NLR 12         // level accumulators kept in registers (fast path); the
                       // upper NL - NLR levels, which the level skipping
                       // reaches rarely, accumulate in shared memory
#endif
#ifndef OQG_ROLL_ALL
#define OQG_ROLL_ALL 0   // experiment: roll the rupture stage of every family
#endif
#ifndef LGRP
#define LGRP 4         // levels evaluated per level-skipping test (NLR % LGRP == 0)
#endif
static_assert(NLR % LGRP == 0 && NLR % 2 == 0 && NLR <= NL, "NLR, LGRP");
#define NWMAX 7        // warps (units) per CTA in the fused kernel
#ifndef MINB
#define MINB 4         // CTAs per SM the fused kernel is compiled for
#endif
#define MAXPLANES 6
#define OQG_GRID_Y 65535   // CUDA limit of gridDim.y: larger extents are launched in chunks

__constant__ double c_coef[OQG_MAX_SLOTS][OQG_KMAX];
// [0, MAX_LEVELS): loglevels[M][L1]; [MAX_LEVELS, 2 MAX_LEVELS): the same by
// fused-kernel unit, [unit][NL] padded with +inf
__constant__ double c_loglev[2 * OQG_MAX_LEVELS];
__constant__ oqg_gsim c_gsim[OQG_MAX_G];
// 1 where the PoEs of (gsim, IMT) are zeroed: the ImtWeight == 0 branch of
// set_poes (contexts.py:1469-1475); index g * M + m
__constant__ unsigned char c_imtmask[OQG_MAX_SLOTS];

// ------------------------------------------------------------------ handle
struct Buf {
  void *p = nullptr;
  size_t cap = 0;
};

struct oqg_handle {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::string err;
  int G = 0, M = 0, L1 = 0;
  std::vector<oqg_gsim> gsims;
  std::vector<double> coef, loglev, loglev_pad;
  std::vector<unsigned char> imtmask;
  double truncation_level = 0, phi_b = 0;
  bool const_dirty = true;
  unsigned long long launches = 0;
  Buf rup, site, part, ptab, conv, stage[12];
  bool has_conv = false;       // oqg_set_conversion
  bool fast_ok = false, exact_poe = false, sorted = false;
  int pt_nodes = 0;
  double ptab_err = 0., xscale = 0.;
  int *flag_dev = nullptr;
  cudaEvent_t ev[2 * OQG_MAX_G + 2];
  cudaEvent_t const_ev = nullptr;   // after the last launch reading the __constant__ tables
  int n_fused_ev = 0;
  int sm_count = 148;
  int fused_cfg[2] = {0, 0};   // oqg_set_fused_config
  bool count_evals = false;    // oqg_set_count_evals
  int last_subs = 1;
  // site grouping of the fused pass (oqg_set_site_grouping): scratch of the
  // sort and the permuted site parameters
  bool site_grouping = true;
  double site_cell_deg = 0.16;
  Buf sg_keys[2], sg_idx[2], sg_tmp, sg_par[4], sg_sidx;
  // oqg_postclassical: grouping of the realizations with the same gsim
  // columns (oqg_set_post_grouping), the kernel path of the last call
  bool post_grouping = true;
  int last_post_path = -1, last_post_groups = 0;
  bool wave_fit = true;        // OQG_WAVE_FIT=0: the first guess of the chunks
  int force_chunks = 0;        // OQG_CHUNKS=n: exactly n rupture chunks
  // oqg_host_begin / add / end: sites and map resident, two rupture slots
  Buf hs_site[7], hs_map, hs_slot[2][4];
  cudaStream_t hs_copy = nullptr;
  cudaEvent_t hs_copied[2], hs_used[2];
  bool hs_rec[2] = {false, false}, hs_open = false;
  int hs_N = 0, hs_rates = 0, hs_n = 0;
  oqg_sites hs_sd;
  // oqg_truncnorm_rvs: the jump table in stage[11] belongs to this increment
  bool pcg_tab_ok = false;
  uint64_t pcg_inc[2] = {0, 0};
};

static std::string g_create_err;

static int fail(oqg_handle *h, const std::string &msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return 1;
}

#define CK(call)                                                          \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess)                                                \
      return fail(h, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static int reserve(oqg_handle *h, Buf &b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr; b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}

// ---- ownership of the __constant__ tables -------------------------------
// The coefficient rows, levels, model descriptors and the IMT mask live in
// __constant__ memory (the fused kernel reads them as instruction operands),
// of which a device has ONE copy shared by all the handles of the process.
// A handle therefore OWNS the tables of its device while its kernels use
// them: ConstGuard, held by every entry point that launches such kernels,
//   * serialises the host threads that enqueue work using the tables,
//   * when the owner changes, makes the new owner's stream wait for the
//     event recorded after the previous owner's last such launch before it
//     overwrites the tables (stream-ordered: no host synchronisation), and
//   * records that event for this handle when the entry point returns.
// Two handles on one device with different models on different streams thus
// interleave correctly (tests/test_gpu_parity.py::test_two_handles_two_streams);
// they do not run concurrently.
#define OQG_MAX_DEVICES 64
static std::mutex g_const_mu;
static oqg_handle *g_const_owner[OQG_MAX_DEVICES];

static int upload_const(oqg_handle *h) {
  oqg_handle *&owner = g_const_owner[h->device];
  if (!h->const_dirty && owner == h) return 0;
  if (owner && owner != h)
    CK(cudaStreamWaitEvent(h->stream, owner->const_ev, 0));
  if (h->G && !h->coef.empty()) {
    CK(cudaMemcpyToSymbolAsync(c_coef, h->coef.data(),
                               h->coef.size() * sizeof(double), 0,
                               cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyToSymbolAsync(c_gsim, h->gsims.data(),
                               h->gsims.size() * sizeof(oqg_gsim), 0,
                               cudaMemcpyHostToDevice, h->stream));
    h->imtmask.resize(OQG_MAX_SLOTS, 0);
    CK(cudaMemcpyToSymbolAsync(c_imtmask, h->imtmask.data(), OQG_MAX_SLOTS, 0,
                               cudaMemcpyHostToDevice, h->stream));
  }
  if (!h->loglev.empty())
    CK(cudaMemcpyToSymbolAsync(c_loglev, h->loglev.data(),
                               h->loglev.size() * sizeof(double), 0,
                               cudaMemcpyHostToDevice, h->stream));
  if (!h->loglev_pad.empty())
    CK(cudaMemcpyToSymbolAsync(c_loglev, h->loglev_pad.data(),
                               h->loglev_pad.size() * sizeof(double),
                               OQG_MAX_LEVELS * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream));
  // the host vectors must stay alive until the copies ran: they are members
  // and only replaced by set_* calls, which synchronise first
  h->const_dirty = false;
  owner = h;
  return 0;
}

struct ConstGuard {
  oqg_handle *h;
  std::unique_lock<std::mutex> lock;
  explicit ConstGuard(oqg_handle *h_) : h(h_), lock(g_const_mu) {}
  // uploads the tables of h if it does not own them (or they changed)
  int acquire() { return upload_const(h); }
  ~ConstGuard() {
    // everything enqueued so far on the handle's stream may read the tables
    cudaEventRecord(h->const_ev, h->stream);
  }
};

// ----------------------------------------------------------- prep kernels
// the prepared ruptures are stored [NF_RUP][Us] with the stride Us = U rounded
// up to even, so that pairs of ruptures are 16-byte aligned in every field
// (the fused kernel fetches its rupture tiles with 16-byte cp.async)
static inline int64_t rup_stride(int64_t U) { return (U + 1) & ~(int64_t)1; }

__global__ void k_prep_ruptures(const double *__restrict__ planar,
                                const double *__restrict__ mag,
                                const double *__restrict__ magdist, int64_t U,
                                int64_t Us, double *__restrict__ out) {
  int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  double pl[OQG_PLANAR_LEN];
  for (int k = 0; k < OQG_PLANAR_LEN; k++) pl[k] = planar[u * OQG_PLANAR_LEN + k];
  prep_rupture(pl, mag[u], magdist ? magdist[u] : INFINITY, out + u, Us);
}

// ruptures with external distances (oqg_hazard_dist): par [U][8] = mag, rake,
// dip, ztor, width, hypo_depth, occurrence rate, integration distance ->
// the fields of the rupture SoA the stages after the distances read
__global__ void k_prep_xruptures(const double *__restrict__ par, int64_t U,
                                 int64_t Us, double *__restrict__ out) {
  int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  const double *p = par + u * 8;
  out[RF_MAG * Us + u] = p[0];
  out[RF_RAKE * Us + u] = p[1];
  out[RF_DIP * Us + u] = p[2];
  out[RF_ZTOR * Us + u] = p[3];
  out[RF_WIDTH * Us + u] = p[4];
  out[RF_HYPO_DEPTH * Us + u] = p[5];
  out[RF_RATE * Us + u] = p[6];
  out[RF_MAGDIST * Us + u] = p[7];
}

// one thread per rupture of a batch of point sources (oqg_source.cuh).
// Without the pointsource_distance approximation a source has M * N * D
// ruptures ordered (magnitude, nodal plane, depth).  With it, a source with
// N * D > 1 has M blocks of N * D + 1 ruptures: the N * D ruptures of the
// magnitude, seen by the sites within psdist of the source (mode 1), and
// the average rupture of PointSource.iruptures (point.py:311-318,
// calc_average :47-95), seen by the others (mode 2).
#define PR_TPB 64     // threads (ruptures) per block of k_point_ruptures
__global__ void k_point_ruptures(const oqg_point_sources P, int S, int64_t U,
                                 double *__restrict__ planar,
                                 double *__restrict__ mag_out,
                                 double *__restrict__ magdist_out,
                                 double *__restrict__ aux_out) {
  // the 42-double records of a block are contiguous in `planar`: they are
  // staged in shared memory (row stride 43: conflict free) and written out
  // by the whole block along the memory (as 336-byte records written by
  // their own thread the stores were 8 useful bytes per 32-byte sector:
  // 17.7 % of the HBM bandwidth)
  __shared__ double s_rec[PR_TPB][OQG_PLANAR_LEN + 1];
  const int64_t u0 = blockIdx.x * (int64_t)PR_TPB;
  const int64_t u = u0 + threadIdx.x;
  if (u < U) {
  int lo = 0, hi = S;   // last s with rup_off[s] <= u
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (P.rup_off[mid] <= u) lo = mid; else hi = mid;
  }
  const int s = lo;
  const int64_t r = u - P.rup_off[s];
  const int N = P.np_cnt[s], D = P.hd_cnt[s];
  const bool psd_on = P.pointsource_distance > 0. && N * D > 1;
  const int B = N * D + (psd_on ? 1 : 0);
  const int m = (int)(r / B), j = (int)(r % B);
  const int im = P.mag_beg[s] + m;
  const double mag = P.mag[im];
  double strike, dip, rake, dep, frac, weight;
  if (j < N * D) {
    const int in = P.np_beg[s] + j / D, id = P.hd_beg[s] + j % D;
    strike = P.np_strike[in]; dip = P.np_dip[in]; rake = P.np_rake[in];
    dep = P.hd_depth[id];
    frac = P.hd_frac ? P.hd_frac[id] : 0.5;
    weight = P.np_prob[in] * P.hd_prob[id];
  } else {
    // numpy.average over the nodal planes / the depths, weight 1
    double ws = 0., a = 0., b = 0., c = 0.;
    for (int k = 0; k < N; k++) {
      const int in = P.np_beg[s] + k;
      const double w = P.np_prob[in];
      ws += w; a += w * P.np_strike[in]; b += w * P.np_dip[in];
      c += w * P.np_rake[in];
    }
    strike = a / ws; dip = b / ws; rake = c / ws;
    ws = 0.; a = 0.; b = 0.;
    for (int k = 0; k < D; k++) {
      const int id = P.hd_beg[s] + k;
      const double w = P.hd_prob[id];
      ws += w; a += w * P.hd_depth[id];
      b += w * (P.hd_frac ? P.hd_frac[id] : 0.5);
    }
    dep = a / ws; frac = b / ws;
    weight = 1.;
  }
  double pl[OQG_PLANAR_LEN];
  build_planar_rupture(P.usd[s], P.lsd[s], P.rar[s],
                       median_area(P.msr, P.msr_par, mag, rake), strike, dip, rake,
                       P.lon[s], P.lat[s], dep, frac, P.mag_rate[im] * weight,
                       P.shift_hypo, pl);
  for (int k = 0; k < OQG_PLANAR_LEN; k++) s_rec[threadIdx.x][k] = pl[k];
  mag_out[u] = mag;
  if (magdist_out) magdist_out[u] = P.mag_dist[im];
  if (aux_out) {
    double *ax = aux_out + u * OQG_AUX_LEN;
    double sl, cl, sp, cp;
    sincos(P.lon[s] * 0.017453292519943295, &sl, &cl);
    sincos(P.lat[s] * 0.017453292519943295, &sp, &cp);
    ax[0] = OQG_EARTH_RADIUS * cp * cl;
    ax[1] = OQG_EARTH_RADIUS * cp * sl;
    ax[2] = OQG_EARTH_RADIUS * sp;
    double psdist = 0., mode = 0.;
    if (psd_on) {
      // PointSource.get_psdist / _get_max_rupture_projection_radius
      // (point.py:215-256): half the diagonal of the surface projection of
      // the rupture of the LAST nodal plane; 0 for PointMSR
      double radius = 0.;
      if (P.msr != OQG_MSR_POINT) {
        const int il = P.np_beg[s] + N - 1;
        double len, wproj;
        rupture_dims(P.usd[s], P.lsd[s], P.rar[s],
                     median_area(P.msr, P.msr_par, mag, P.np_rake[il]), P.np_dip[il], len,
                     wproj);
        radius = sqrt(len * len + wproj * wproj) / 2.0;
      }
      psdist = fmin(radius, P.mag_dist[im] / 2.) + P.pointsource_distance;
      mode = j < N * D ? 1. : 2.;
    }
    if (P.mode) mode = (double)P.mode[u];
    ax[3] = psdist;
    ax[4] = mode;
  }
  }   // u < U
  __syncthreads();
  const int64_t nrec = min((int64_t)PR_TPB, U - u0);
  for (int64_t e = threadIdx.x; e < nrec * OQG_PLANAR_LEN; e += PR_TPB)
    planar[u0 * OQG_PLANAR_LEN + e] =
        s_rec[e / OQG_PLANAR_LEN][e % OQG_PLANAR_LEN];
}

__global__ void k_prep_sites(const double *__restrict__ lon,
                             const double *__restrict__ lat,
                             const double *__restrict__ dep, int N,
                             double *__restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  prep_site(lon[n], lat[n], dep ? dep[n] : 0., out + n, N);
}

// -------------------------------------------------------------- distances
// out[k][U][N]; one thread per site, blockIdx.y strides over the ruptures
__global__ void k_distances(const double *__restrict__ rup, int U,
                            int64_t Us, const double *__restrict__ site,
                            int N, int mask, double *__restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double sx[3], sq[3], sdep;
  for (int k = 0; k < 3; k++) {
    sx[k] = site[(int64_t)(SF_XYZ + k) * N + n];
    sq[k] = site[(int64_t)(SF_Q + k) * N + n];
  }
  sdep = site[(int64_t)SF_DEPTH * N + n];
  for (int u = blockIdx.y; u < U; u += gridDim.y) {
    Dist d;
    pair_distances(rup + u, Us, sx, sq, sdep, mask, d);
    const double vals[OQG_NOUT] = {d.rrup, d.rjb, d.rx, d.ry0, d.repi,
                                   d.rhypo, d.azimuth, d.clon, d.clat};
    int plane = 0;
#pragma unroll
    for (int k = 0; k < OQG_NOUT; k++)
      if (mask & (1 << k)) {
        out[((int64_t)plane * U + u) * N + n] = vals[k];
        plane++;
      }
  }
}

// -------------------------------------------------------------- mean_stds
struct CtxDev {
  oqg_ctx c;
};

__device__ __forceinline__ double ld(const double *p, int64_t i) {
  return p ? p[i] : 0.;
}

template <int F>
__global__ void __launch_bounds__(128)
k_mean_stds(int g, int G, int M, CtxDev cd, int64_t P,
            double *__restrict__ out, int *__restrict__ flag,
            const double *__restrict__ conv) {
  typedef typename Fam<F>::C CT;
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  const oqg_ctx &c = cd.c;
  Row r;
  r.mag = ld(c.mag, p); r.rake = ld(c.rake, p); r.dip = ld(c.dip, p);
  r.ztor = ld(c.ztor, p); r.width = ld(c.width, p);
  r.hypo_depth = ld(c.hypo_depth, p);
  r.rrup = ld(c.rrup, p); r.rjb = ld(c.rjb, p); r.rx = ld(c.rx, p);
  r.ry0 = ld(c.ry0, p); r.repi = ld(c.repi, p); r.rhypo = ld(c.rhypo, p);
  r.vs30 = ld(c.vs30, p); r.z1pt0 = ld(c.z1pt0, p); r.z2pt5 = ld(c.z2pt5, p);
  r.vs30measured = c.vs30measured ? c.vs30measured[p] != 0 : false;
  const oqg_gsim &gs = c_gsim[g];
  const CT &Cp = *reinterpret_cast<const CT *>(c_coef[g * (M + 1) + M]);
  typename Fam<F>::R st = Fam<F>::row(gs, Cp, r);
  const int64_t plane = (int64_t)G * M * P;
  for (int m = 0; m < M; m++) {
    const CT &C = *reinterpret_cast<const CT *>(c_coef[g * (M + 1) + m]);
    Out o = Fam<F>::imt(gs, C, Cp, r, st);
    int64_t idx = ((int64_t)g * M + m) * P + p;
    // the zero-sigma check of get_4MN comes BEFORE the conversion
    // (contexts.py:1297-1302)
    if (o.sig == 0.) *flag = 1;  // benign race: every writer stores 1
    if (conv) {
      // horiz_comp_to_geom_mean (contexts.py:761-776): [log conv_median,
      // conv_sigma^2, 1 / rstd^2] per (gsim, IMT); tau and phi stay
      const double *cv = conv + (g * M + m) * 3;
      if (cv[2] != 0.) {
        o.mean -= cv[0];
        o.sig = sqrt((o.sig * o.sig - cv[1]) * cv[2]);
      }
    }
    out[idx] = o.mean;
    out[plane + idx] = o.sig;
    out[2 * plane + idx] = o.tau;
    out[3 * plane + idx] = o.phi;
  }
}

// the staged model code of the fused kernel (oqg_fast.cuh) on context rows:
// out[2][G][M][P] = mean, sig.  Verification entry point: the parity tests
// compare it with the reference fixtures to the same tolerance as
// k_mean_stds, so that the staging is pinned independently of the PoEs.
template <int F>
__global__ void __launch_bounds__(128)
k_mean_stds_staged(int g, int G, int M, CtxDev cd, int64_t P,
                   double *__restrict__ out) {
  typedef Fast<F> FF;
  typedef typename FF::C CT;
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  math_tables_init();
  __syncthreads();
  if (p >= P) return;
  const oqg_ctx &c = cd.c;
  RupIn ri;
  ri.mag = ld(c.mag, p); ri.rake = ld(c.rake, p); ri.dip = ld(c.dip, p);
  ri.ztor = ld(c.ztor, p); ri.width = ld(c.width, p);
  ri.hypo_depth = ld(c.hypo_depth, p);
  Dist dd;
  dd.rrup = ld(c.rrup, p); dd.rjb = ld(c.rjb, p); dd.rx = ld(c.rx, p);
  dd.ry0 = ld(c.ry0, p); dd.repi = ld(c.repi, p); dd.rhypo = ld(c.rhypo, p);
  dd.azimuth = 0.;
  SiteIn si;
  si.vs30 = ld(c.vs30, p);
  si.z1pt0 = c.z1pt0 ? c.z1pt0[p] : -999.;
  si.z2pt5 = c.z2pt5 ? c.z2pt5[p] : -999.;
  si.vs30measured = c.vs30measured ? c.vs30measured[p] != 0 : false;
  const oqg_gsim &gs = c_gsim[g];
  const CT &Cp = *reinterpret_cast<const CT *>(c_coef[g * (M + 1) + M]);
  const typename FF::S sp = FF::site(gs, Cp, si);
  const typename FF::Q qp = FF::rup(gs, Cp, ri);
  const typename FF::P pp = FF::pair(gs, Cp, sp, qp, ri, si, dd);
  const int64_t plane = (int64_t)G * M * P;
  for (int m = 0; m < M; m++) {
    const CT &C = *reinterpret_cast<const CT *>(c_coef[g * (M + 1) + m]);
    const typename FF::S ss = FF::site(gs, C, si);
    const typename FF::Q q = FF::rup(gs, C, ri);
    double mean, sig, inv;
    FF::eval(gs, C, Cp, ss, q, qp, pp, ri, si, dd, mean, sig, inv);
    int64_t idx = ((int64_t)g * M + m) * P + p;
    out[idx] = mean;
    out[plane + idx] = sig;
  }
}

// ------------------------------------------------------------------- poes
// one thread per output element poes[p][l][g] (coalesced float32 stores)
__global__ void k_poes(const double *__restrict__ ms, int64_t P, int G, int M,
                       int L1, double phi_b, double inv_norm,
                       float *__restrict__ out) {
  const int64_t LG = (int64_t)M * L1 * G;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= P * LG) return;
  int64_t p = e / LG;
  int rem = (int)(e - p * LG);
  int l = rem / G, g = rem - l * G;
  int m = l / L1;
  int64_t idx = ((int64_t)g * M + m) * P + p;
  double mean = ms[idx], sig = ms[(int64_t)G * M * P + idx];
  double z = (c_loglev[l] - mean) / sig;
  out[e] = c_imtmask[g * M + m] ? 0.f : (float)truncnorm_sf(phi_b, inv_norm, z);
}

// ----------------------------------------------------------------- update
// one thread per map element (s, l, g); loops over the rows of site row s
__global__ void k_update(float *__restrict__ map, int N, int LG, int G,
                         const float *__restrict__ poes,
                         const int64_t *__restrict__ order,
                         const int64_t *__restrict__ seg,
                         const double *__restrict__ rate,
                         const double *__restrict__ probs, int K, double itime,
                         int rates_mode,
                         const double *__restrict__ weights, int y0) {
  int s = blockIdx.y + y0;   // y0: the host launches grid.y in chunks
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= LG) return;
  int64_t a = seg[s], b = seg[s + 1];
  if (a == b) return;
  double sum = 0., prod = 1.;
  for (int64_t i = a; i < b; i++) {
    int64_t row = order[i];
    double poe = (double)poes[row * LG + e];
    if (weights) {
      // mutually exclusive ruptures, update_pmap_m (map_array.py:262-269):
      // arr += (1 - pne) * weight
      double pne = itime == 0. ? 1. - poe
                   : (K == 0 ? exp(-rate[row] * poe * itime)
                             : pne_from_pmf(probs + row * K, K, poe));
      sum += (1. - pne) * weights[row];
    } else if (K == 0) {
      sum += rate[row] * poe * itime;
    } else {
      double pne = itime == 0. ? 1. - poe
                               : pne_from_pmf(probs + row * K, K, poe);
      if (rates_mode) sum += -log(pne); else prod *= pne;
    }
  }
  int64_t at = (int64_t)s * LG + e;
  if (weights || rates_mode) map[at] = (float)((double)map[at] + sum);
  else map[at] = (float)((double)map[at] * (exp(-sum) * prod));
}

// ------------------------------------------------------------------ fused
// Fast PoE: poe as a function of x = (tl - z) / (2 tl) in [0, 1] is tabulated
// as one cubic per node k = round(x (nodes - 1)) in u = x (nodes - 1) - k in
// [-.5, .5] (float32 coefficients built in float64 by oqg_set_levels; 256 or
// 512 nodes, whichever reproduces the exact formula to 2e-6); node 0 has a
// zero constant term, so that the relative accuracy is kept as poe -> 0
// (z -> truncation level), where the textbook formula cancels.  x itself is
// computed in float64.  The table is [nodes][4] floats in shared memory, one
// 16-byte load per evaluation: lanes with the same node broadcast, the lanes
// of a quarter warp (neighbouring sites) hit nodes less than 8 apart, i.e.
// different banks.
//
// Shared-memory bandwidth is what bounds the evaluation: an LDS.128 whose 32
// lanes read 32 different nodes is served quarter warp by quarter warp, and
// the 8 lanes of a quarter conflict when their nodes fall into the same
// 16-byte bank group (k mod 8): 8.3 wavefronts per load were measured instead
// of the 4 that deliver 512 bytes.  With REP = 8 copies of the table (copy r
// of node k at float4 index 8 k + r, lane i reads copy i mod 8) every lane of
// a quarter warp has its own bank group and the load is conflict free.  A
// 32 KB table per CTA only pays when one CTA fills the SM, hence the
// "sub-CTAs" of k_fused: SUBS independent (site tile x rupture chunk) work
// items of 32 * nw threads each share one CTA, one table and nothing else,
// and synchronise among themselves with named barriers.
#define PT_DEG 3
#define PT_STRIDE 4       // floats per node: one LDS.128
#define PT_MAXNODES 512
#ifndef FLUSH_TILES
#define FLUSH_TILES 8
#endif
//    // float32 partial sums are flushed to FP64 every
                          // FLUSH_TILES * RT ruptures

struct FusedArgs {
  const double *rup;      // [NF_RUP][U]
  const double *probs;    // [U][K] or null
  const double *aux;      // [U][OQG_AUX_LEN] or null (pointsource_distance)
  const double *xd[6];    // external distance planes [U][N] (rrup, rjb, rx,
                          // ry0, repi, rhypo) of non-planar ruptures, or null
  const double *site;     // [NF_SITE][N]
  const double *vs30, *z1pt0, *z2pt5;
  const uint8_t *vs30measured;
  const float *ptab;      // [nodes][PT_STRIDE] or null (exact FP64 PoEs)
  int pt_nodes;
  unsigned pt_bias;
  double *part;           // [RC][G][L][N] (site fastest)
  unsigned long long *counters;
  int64_t U, Us;          // ruptures, stride of the fields of `rup`
  int N, K;
  int g, G, M, L1;
  int nlc;                // level chunks per IMT
  int lpc;                // levels per chunk (<= NL)
  int nunits;             // M * nlc
  int nw;                 // warps (units) per sub-CTA
  int subs;               // sub-CTAs per CTA
  int rup_per_chunk;
  int ntiles;             // site tiles of 32 sites
  int dmask;              // distances to compute
  int clampmask;          // distances clamped by minimum_distance
  int count_pairs;
  int sorted;             // the levels of every IMT are increasing
  double mindist, itime, phi_b, inv_norm, tl, xscale;
};

// tab32: shared-window address of the table minus (0x4b000000 * 16), so
// that tab32 + bits(k + 2^23) * 16 is the address of node k.  12 instructions
// per evaluation with the DFMA of x and the FFMA into the accumulator: F2F,
// FADD.SAT, FFMA (round), FADD, FFMA (u), IMAD (address), LDS.128, 3 FFMA.
#ifndef OQG_F2F_SAT
#define OQG_F2F_SAT 1     // one F2F.SAT instead of F2F + FADD.SAT
#endif
#ifndef OQG_POE_PAIR
#define OQG_POE_PAIR 0    // two levels per evaluation with packed f32x2 ops
#endif

// float64 -> float32 clamped to [0, 1], NaN -> 0 (poe 0)
__device__ __forceinline__ float sat_f32(double x64) {
#if OQG_F2F_SAT
  float x;
  asm("cvt.rn.sat.f32.f64 %0, %1;" : "=f"(x) : "d"(x64));
  return x;
#else
  return __saturatef((float)x64);
#endif
}

// Blackwell packed float32 pairs (FFMA2 / FADD2): one issue slot for two
// independent operations
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// two evaluations of the table at once: the rounding to the node and the
// offset inside it in packed arithmetic, the two cubics interleaved
template <int REP>
__device__ __forceinline__ f32x2 poe_table2(unsigned tab32, f32x2 nm1,
                                            double x64a, double x64b) {
  const f32x2 big = pack2(8388608.f, 8388608.f);
  const f32x2 mone = pack2(-1.f, -1.f);
  const f32x2 x = pack2(sat_f32(x64a), sat_f32(x64b));
  const f32x2 y = fma2(x, nm1, big);        // round to the nearest integer
  const f32x2 negk = fma2(y, mone, big);    // 2^23 - y = -k, exact
  const f32x2 u = fma2(x, nm1, negk);
  float ya, yb, ua, ub;
  unpack2(y, ya, yb);
  unpack2(u, ua, ub);
  float a0, a1, a2, a3, b0, b1, b2, b3;
  unsigned aa = tab32 + __float_as_uint(ya) * (PT_STRIDE * 4 * REP);
  unsigned ab = tab32 + __float_as_uint(yb) * (PT_STRIDE * 4 * REP);
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3) : "r"(aa));
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(ab));
  float pa = fmaf(a3, ua, a2), pb = fmaf(b3, ub, b2);
  pa = fmaf(pa, ua, a1); pb = fmaf(pb, ub, b1);
  pa = __saturatef(fmaf(pa, ua, a0)); pb = __saturatef(fmaf(pb, ub, b0));
  return pack2(pa, pb);
}

template <int REP>
__device__ __forceinline__ float poe_table(unsigned tab32, float nm1,
                                           double x64) {
  float x = sat_f32(x64);    // NaN -> 0 -> poe 0
  float y = fmaf(x, nm1, 8388608.f);    // 2^23: round to nearest integer
  float u = fmaf(x, nm1, -(y - 8388608.f));
  float c0, c1, c2, c3;
  unsigned a = tab32 + __float_as_uint(y) * (PT_STRIDE * 4 * REP);
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(c0), "=f"(c1), "=f"(c2), "=f"(c3) : "r"(a));
  float p = fmaf(c3, u, c2);
  p = fmaf(p, u, c1);
  return __saturatef(fmaf(p, u, c0));
}

template <typename T>
__device__ __forceinline__ void st_soa(double *base, int stride, const T &v) {
  const double *a = reinterpret_cast<const double *>(&v);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 8); k++) base[k * stride] = a[k];
}

template <typename T>
__device__ __forceinline__ T ld_soa(const double *base, int stride) {
  T v;
  double *a = reinterpret_cast<double *>(&v);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 8); k++) a[k] = base[k * stride];
  return v;
}

// Distances a family may read, as a compile-time superset of
// required_distances() (+ rrup for the filter, repi under rhypo): the planes
// of s_dist and every shared-memory offset of the fused kernel are then
// compile-time constants (immediates in the LDS / STS instead of address
// registers and IMADs)
// rupture parameters the stages after the distances read (a copy of those
// fields of the rupture tile: s_rup itself is refilled with the next tile
// while the tile is evaluated)
enum { PF_MAG, PF_RAKE, PF_DIP, PF_ZTOR, PF_WIDTH, PF_HYPO_DEPTH, PF_RATE,
       NPAR = 8 };

template <int F> struct DistMask {
  static constexpr int value =
      F == OQG_BSSA14 || F == OQG_BA08 || F == OQG_TORO02
          ? (1 << OQG_RRUP) | (1 << OQG_RJB)
      : F == OQG_CY14 || F == OQG_CB14 || F == OQG_CY08
          ? (1 << OQG_RRUP) | (1 << OQG_RJB) | (1 << OQG_RX)
      : F == OQG_ASK14
          ? (1 << OQG_RRUP) | (1 << OQG_RJB) | (1 << OQG_RX) | (1 << OQG_RY0)
      : F == OQG_AB06 || F == OQG_C03 || F == OQG_SADIGH97 ? (1 << OQG_RRUP)
      : (1 << OQG_RRUP) | (1 << OQG_RJB) | (1 << OQG_REPI) | (1 << OQG_RHYPO);
};
__host__ __device__ constexpr int popc_below(int mask, int k) {   // planes before distance k
  int n = 0;
  for (int j = 0; j < k; j++) n += (mask >> j) & 1;
  return n;
}

// shared memory of one sub-CTA, in doubles (the layout of k_fused)
template <int F, bool NEGBIN, bool FAST> struct FusedLayout {
  typedef Fast<F> FF;
  static constexpr int QW = FF::HAS_Q ? sizeof(typename FF::Q) / 8 : 0;
  static constexpr int PW = FF::HAS_P ? sizeof(typename FF::P) / 8 : 0;
  static constexpr int DM = DistMask<F>::value;
  static constexpr int NPL = popc_below(DM, OQG_NDIST);
  // rupture fields staged per tile: without repi / rhypo the hypocentre
  // fields stay in HBM
  static constexpr int NFR =
      (DM & ((1 << OQG_REPI) | (1 << OQG_RHYPO))) ? NF_RUP : NF_RUP_NOHYPO;
  static constexpr int RUP = 0;                          // [NFR][RT]
  static constexpr int DIST = RUP + NFR * RT;            // [NPL][RT][32]
  static constexpr int SITE = DIST + NPL * RT * 32;      // [NF_SITE][32]
  static constexpr int SQ = SITE + NF_SITE * 32;         // [NWMAX + 1][RT][QW]
  static constexpr int SP = SQ + (NWMAX + 1) * RT * QW;  // [PW][RT][32]
  static constexpr int PAR = SP + PW * RT * 32;          // [NPAR][RT]
  static constexpr int PROBS = PAR + NPAR * RT;          // [RT][OQG_MAX_OCC]
  static constexpr int VALID = PROBS + (NEGBIN ? RT * OQG_MAX_OCC : 0);  // u32 [RT]
  static constexpr int CNT = VALID + RT / 2;             // u64 evaluation counter (+ pad)
  static constexpr int HI = CNT + 2;                     // f32 [NL - NLR][32 NWMAX]
  static constexpr int SIZE =
      (HI + (FAST ? (NL - NLR) * 32 * NWMAX / 2 : 0) + 1) / 2 * 2;
};

template <int F, bool NEGBIN, bool FAST, int SUBS, int REP>
__global__ void __launch_bounds__(32 * NWMAX * SUBS, SUBS == 1 ? MINB : 1)
k_fused(const FusedArgs A) {
  typedef Fast<F> FF;
  typedef typename FF::C CT;
  typedef typename FF::S ST;
  typedef typename FF::Q QT;
  typedef typename FF::P PT;
  constexpr int QW = FF::HAS_Q ? sizeof(QT) / 8 : 0;   // doubles
  constexpr int PW = FF::HAS_P ? sizeof(PT) / 8 : 0;
  typedef FusedLayout<F, NEGBIN, FAST> LY;
  constexpr int DM = LY::DM;
  extern __shared__ __align__(16) double smem[];
  // a CTA holds `subs` sub-CTAs of nw warps each (subs = 1 when SUBS == 1):
  // sub-CTA `sub` works on site tile blockIdx.x * subs + sub and shares only
  // the read-only tables with the others.  Layout: the sub-CTA regions
  // (compile-time offsets, FusedLayout), then the PoE table.
  const int nw = A.nw;
  const int tb = nw << 5;                          // threads of a sub-CTA
  // REDUX leaves the warp index in a UNIFORM register: the compiler then
  // knows that everything derived from it (the IMT of this warp, its
  // coefficient row, its levels) is warp-uniform and reads the __constant__
  // tables through the uniform datapath (LDCU / UR operands) instead of
  // per-thread indexed LDC into vector registers
  const int wall = __reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
  const int sub = SUBS == 1 ? 0 : wall / nw;
  const int warp = wall - sub * nw;
  const int lane = threadIdx.x & 31;
  const int tid = (warp << 5) | lane;
  const int tile = SUBS == 1 ? blockIdx.x : blockIdx.x * A.subs + sub;
  double *const s_base = SUBS == 1 ? smem : smem + sub * LY::SIZE;
  float *s_tab = (float *)(smem + (SUBS == 1 ? 1 : A.subs) * LY::SIZE);  // [nodes][REP][4]
  double *const s_rup = s_base + LY::RUP;
  double *const s_dist = s_base + LY::DIST;
  double *const s_site = s_base + LY::SITE;
  double *const s_q = s_base + LY::SQ;
  double *const s_p = s_base + LY::SP;
  double *const s_par = s_base + LY::PAR;
  double *const s_probs = s_base + LY::PROBS;
  unsigned *const s_valid = (unsigned *)(s_base + LY::VALID);
  unsigned long long *const s_cnt = (unsigned long long *)(s_base + LY::CNT);
  float *const s_hi = (float *)(s_base + LY::HI) + tid;   // [NL - NLR][32 NWMAX]
  constexpr int hs = 32 * NWMAX;
  // barrier among the threads of this sub-CTA
#define SUB_SYNC()                                                         \
  do {                                                                     \
    if (SUBS == 1) __syncthreads();                                        \
    else asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "r"(tb) : "memory"); \
  } while (0)

  const int site0 = tile * 32;
  const int n = site0 + lane;
  const bool site_ok = n < A.N;
  const int unit = blockIdx.z * nw + warp;     // (m, level chunk)
  const bool unit_ok = unit < A.nunits;
  const int m = unit_ok ? unit / A.nlc : 0;
  const int lc = unit_ok ? unit - m * A.nlc : 0;
  const int l0 = lc * A.lpc;
  const int nl = min(A.lpc, A.L1 - l0);

  // the PoE table and the log / exp tables of oqg_math.cuh: whole CTA
  if (FAST)
    for (int i = threadIdx.x; i < A.pt_nodes * REP; i += blockDim.x)
      reinterpret_cast<float4 *>(s_tab)[i] =
          reinterpret_cast<const float4 *>(A.ptab)[i / REP];
  math_tables_init();
  if (SUBS > 1) {
    __syncthreads();
    if (tile >= A.ntiles) return;     // named barriers only from here on
  }
  // site tile to shared memory (geometry) and registers (site parameters)
  for (int i = tid; i < NF_SITE * 32; i += tb) {
    int f = i >> 5, j = i & 31;
    s_site[i] = site0 + j < A.N ? A.site[(int64_t)f * A.N + site0 + j] : 0.;
  }
  if (tid == 0) *s_cnt = 0ull;
  // pt_bias = -(0x4b000000 * 16 * REP) mod 2^32 comes in as a kernel
  // argument: as a literal the compiler splits it off again and adds it back
  // before every load
  const unsigned tab32 = (unsigned)__cvta_generic_to_shared(s_tab) + A.pt_bias +
                         (lane & (REP - 1)) * (PT_STRIDE * 4);
  const float nm1 = (float)(A.pt_nodes - 1);
  SiteIn si;
  si.vs30 = site_ok && A.vs30 ? A.vs30[n] : 760.;
  si.z1pt0 = site_ok && A.z1pt0 ? A.z1pt0[n] : -999.;
  si.z2pt5 = site_ok && A.z2pt5 ? A.z2pt5[n] : -999.;
  si.vs30measured = site_ok && A.vs30measured ? A.vs30measured[n] != 0 : false;

  const oqg_gsim &gs = c_gsim[A.g];
  const CT &C = *reinterpret_cast<const CT *>(c_coef[A.g * (A.M + 1) + m]);
  const CT &Cp = *reinterpret_cast<const CT *>(c_coef[A.g * (A.M + 1) + A.M]);
  // log levels of this unit, padded to NL with +inf (poe = 0)
  const double *ll = c_loglev + OQG_MAX_LEVELS + (unit_ok ? unit : 0) * NL;
  const ST ss = FF::site(gs, C, si);      // this IMT
  const ST sp = FF::site(gs, Cp, si);     // PGA, for the shared pair state

  // exact path: FP64 accumulators (sum of rate*poe*T, or product of pnes);
  // fast path: float32 accumulators per tile, flushed to the FP64 partial
  // sums of this thread in global memory (L2 resident)
  double acc[FAST ? 1 : NL];
#if OQG_POE_PAIR
  f32x2 facc2[FAST ? NLR / 2 : 1];
#define FACC_GET(l) __uint_as_float((unsigned)(facc2[(l) >> 1] >> (((l) & 1) * 32)))
#else
  float facc[FAST ? NLR : 1];
#define FACC_GET(l) facc[l]
#endif
  const int L = A.M * A.L1;
  // partial sums [chunk][gsim][level][site]: site fastest, so that the warp's
  // read-modify-write of one level is one contiguous 256-byte request
  const int64_t ds = A.N;
  const int64_t dat = (((int64_t)blockIdx.y * A.G + A.g) * L + m * A.L1 + l0) *
                          ds + (site_ok ? n : 0);
  // exact path: float64 results; fast path: float32 partial sums (the
  // float32 register accumulators of <= 128 ruptures are added to them: one
  // more rounding of 6e-8 per flush, against 128 in the registers)
  double *dst = A.part + dat;
  float *dstf = reinterpret_cast<float *>(A.part) + dat;
  const bool writer = unit_ok && site_ok;
  if (FAST) {
    if (writer)
      for (int l = 0; l < nl; l++) dstf[l * ds] = 0.f;
  } else {
#pragma unroll
    for (int l = 0; l < NL; l++) acc[l] = NEGBIN ? 1. : 0.;
  }
  unsigned long long kept = 0;
  if (FAST) {
#if OQG_POE_PAIR
#pragma unroll
    for (int l = 0; l < NLR / 2; l++) facc2[l] = 0ull;
#else
#pragma unroll
    for (int l = 0; l < NLR; l++) facc[l] = 0.f;
#endif
#pragma unroll
    for (int l = NLR; l < NL; l++) s_hi[(l - NLR) * hs] = 0.f;
  }
  int tiles_left = FLUSH_TILES;

  const int64_t r_begin = (int64_t)blockIdx.y * A.rup_per_chunk;
  const int64_t r_end = min(A.U, r_begin + A.rup_per_chunk);

#define DGET(k) (((DM >> (k)) & 1) ? s_dist[(popc_below(DM, k) * RT + i) * 32 + lane] : 0.)
#define LOAD_PAIR(i)                                                       \
  RupIn ri;                                                                \
  ri.mag = s_par[PF_MAG * RT + i];                                         \
  ri.rake = s_par[PF_RAKE * RT + i];                                       \
  ri.dip = s_par[PF_DIP * RT + i];                                         \
  ri.ztor = s_par[PF_ZTOR * RT + i];                                       \
  ri.width = s_par[PF_WIDTH * RT + i];                                     \
  ri.hypo_depth = s_par[PF_HYPO_DEPTH * RT + i];                           \
  Dist dd;                                                                 \
  dd.rrup = DGET(OQG_RRUP); dd.rjb = DGET(OQG_RJB);                        \
  dd.rx = DGET(OQG_RX); dd.ry0 = DGET(OQG_RY0);                            \
  dd.repi = DGET(OQG_REPI); dd.rhypo = DGET(OQG_RHYPO);

  // The rupture tile [NF_RUP][RT] is fetched with 16-byte cp.async (two
  // ruptures of one field), zero-filled past the end of the chunk.  The
  // stages after the distances only read the copy s_par of the rupture
  // parameters, so the fetch of the NEXT tile is issued as soon as the pair
  // stage is done and flies during the evaluation of this one (78 % of the
  // loop): its latency was 10-12 % of the kernel when it was a plain load
  // between two barriers.
  auto fetch_tile = [&](int64_t r0) {
    const int nr = (int)min((int64_t)RT, r_end - r0);
    for (int i = tid; i < LY::NFR * (RT / 2); i += tb) {
      const int f = i / (RT / 2), j = (i - f * (RT / 2)) * 2;
      const int nb = min(max(nr - j, 0), 2) * 8;          // bytes to read
      const double *src = A.rup + (int64_t)f * A.Us + r0 + (nb ? j : 0);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(s_rup + f * RT + j);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                   "l"(src), "r"(nb) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (r_begin < r_end) fetch_tile(r_begin);
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RT) {
    const int nr = (int)min((int64_t)RT, r_end - r0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    SUB_SYNC();   // tile r0 is in s_rup; everybody is done with the last one
    if (NEGBIN) {
      for (int i = tid; i < nr * A.K; i += tb)
        s_probs[(i / A.K) * OQG_MAX_OCC + i % A.K] = A.probs[r0 * A.K + i];
    }
    if (warp == 0 && lane < RT) {
      s_par[PF_MAG * RT + lane] = s_rup[RF_MAG * RT + lane];
      s_par[PF_RAKE * RT + lane] = s_rup[RF_RAKE * RT + lane];
      s_par[PF_DIP * RT + lane] = s_rup[RF_DIP * RT + lane];
      s_par[PF_ZTOR * RT + lane] = s_rup[RF_ZTOR * RT + lane];
      s_par[PF_WIDTH * RT + lane] = s_rup[RF_WIDTH * RT + lane];
      s_par[PF_HYPO_DEPTH * RT + lane] = s_rup[RF_HYPO_DEPTH * RT + lane];
      s_par[PF_RATE * RT + lane] = s_rup[RF_RATE * RT + lane];
    }
    // ---- distances: rupture i of the tile by warp i % nw, lane = site
    for (int i = warp; i < nr; i += nw) {
      double sx[3], sq[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        sx[k] = s_site[(SF_XYZ + k) * 32 + lane];
        sq[k] = s_site[(SF_Q + k) * 32 + lane];
      }
      double vals[OQG_NDIST];
      if (A.xd[0]) {
        // non-planar ruptures (oqg_hazard_dist): the distances were computed
        // by the mesh kernels into [U][N] planes; lane = site, coalesced
        const int64_t at = (r0 + i) * (int64_t)A.N + (site_ok ? n : 0);
#pragma unroll
        for (int k = 0; k < 6; k++)
          vals[k] = ((DM >> k) & 1) && A.xd[k] ? A.xd[k][at] : 0.;
        vals[6] = 0.;
      } else {
        Dist d;
        pair_distances(s_rup + i, RT, sx, sq, s_site[SF_DEPTH * 32 + lane],
                       A.dmask, d);
        vals[0] = d.rrup; vals[1] = d.rjb; vals[2] = d.rx; vals[3] = d.ry0;
        vals[4] = d.repi; vals[5] = d.rhypo; vals[6] = 0.;
      }
      // minimum_distance clamp of the required distances, contexts.py:391
      if (A.mindist != 0.) {
#pragma unroll
        for (int k = 0; k < 6; k++)
          if ((A.clampmask >> k) & 1)
            if (vals[k] < A.mindist) vals[k] = A.mindist;
      }
      // planar ruptures: rrup < magdist (contexts.py:472); the others:
      // rrup <= magdist (genctxs, contexts.py:974)
      bool ok = site_ok && (A.xd[0] ? vals[0] <= s_rup[RF_MAGDIST * RT + i]
                                    : vals[0] < s_rup[RF_MAGDIST * RT + i]);
      if (A.aux) {
        // pointsource_distance (contexts.py:347-370): [x, y, z of the source
        // location, psdist, mode] per rupture, uniform over the warp; near
        // sites (chord distance to the location <= psdist) see every nodal
        // plane and depth (mode 1), far sites one average rupture (mode 2)
        const double *ax = A.aux + (r0 + i) * OQG_AUX_LEN;
        const double mode = ax[4];
        if (mode != 0.) {
          const double dx = sx[0] - ax[0], dy = sx[1] - ax[1],
                       dz = sx[2] - ax[2];
          const bool close = fma(dx, dx, fma(dy, dy, dz * dz)) <= ax[3] * ax[3];
          ok = ok && (mode == 1. ? close : (mode == 2. ? !close : false));
        }
      }
#pragma unroll
      for (int k = 0; k < 6; k++)
        if ((DM >> k) & 1)
          s_dist[(popc_below(DM, k) * RT + i) * 32 + lane] = vals[k];
      unsigned b = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_valid[i] = b;
      if (A.count_pairs && lane == 0 && blockIdx.z == 0) kept += __popc(b);
    }
    // ---- rupture-only part of the model: lane = rupture, one IMT per
    // warp; the last warp also does the PGA row
    if (FF::HAS_Q && lane < nr) {
      const int i = lane;
      RupIn ri;
      ri.mag = s_rup[RF_MAG * RT + i];
      ri.rake = s_rup[RF_RAKE * RT + i];
      ri.dip = s_rup[RF_DIP * RT + i];
      ri.ztor = s_rup[RF_ZTOR * RT + i];
      ri.width = s_rup[RF_WIDTH * RT + i];
      ri.hypo_depth = s_rup[RF_HYPO_DEPTH * RT + i];
      if (FF::ROLL_Q || OQG_ROLL_ALL) {
        // one inlined copy of FF::rup for both rows: the CY14 kernel was
        // fetch bound (6592 SASS instructions, stall no_instruction 1.5 per
        // issue); measured -6 % time for CY14, +9 % for ASK14
        const int npass = warp == nw - 1 ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; pass++) {
          // indexed constant load (LDC), not a select of two references
          // (which turns the coefficient reads into generic loads)
          const CT &cc = *reinterpret_cast<const CT *>(
              c_coef[A.g * (A.M + 1) + (pass ? A.M : m)]);
          st_soa(s_q + ((pass ? nw : warp) * RT + i) * QW, 1,
                 FF::rup(gs, cc, ri));
        }
      } else {
        st_soa(s_q + (warp * RT + i) * QW, 1, FF::rup(gs, C, ri));
        if (warp == nw - 1)
          st_soa(s_q + (nw * RT + i) * QW, 1, FF::rup(gs, Cp, ri));
      }
    }
    SUB_SYNC();
    // s_rup is dead from here on: fetch the next tile into it
    if (r0 + RT < r_end) fetch_tile(r0 + RT);
    // ---- the IMT-independent part of the model, once per pair, shared by
    // the IMT warps through shared memory
    if (FF::HAS_P) {
      for (int i = warp; i < nr; i += nw) {
        if (!((s_valid[i] >> lane) & 1)) continue;
        LOAD_PAIR(i)
        QT qp;
        if (FF::HAS_Q) qp = ld_soa<QT>(s_q + (nw * RT + i) * QW, 1);
        st_soa(s_p + i * 32 + lane, RT * 32,
               FF::pair(gs, Cp, sp, qp, ri, si, dd));
      }
      SUB_SYNC();
    }
    // ---- GMPE + PoEs + accumulation: this warp's (IMT, level chunk)
    if (!unit_ok) continue;
    for (int i = 0; i < nr; i++) {
      unsigned vm = s_valid[i];
      if (!((vm >> lane) & 1)) continue;
      LOAD_PAIR(i)
      QT q, qp;
      PT pp;
      if (FF::HAS_Q) q = ld_soa<QT>(s_q + (warp * RT + i) * QW, 1);
      if (FF::EVAL_QP) qp = ld_soa<QT>(s_q + (nw * RT + i) * QW, 1);
      if (FF::HAS_P) pp = ld_soa<PT>(s_p + i * 32 + lane, RT * 32);
      else pp = FF::pair(gs, Cp, sp, q, ri, si, dd);
      double mean, sig, inv;
      FF::eval(gs, C, Cp, ss, q, qp, pp, ri, si, dd, mean, sig, inv);
      if (FAST) {
        // x = (tl - (l - mean) / sig) * xscale, xscale = 1 / (2 tl)
        const double xa = -inv * A.xscale;
        const double xb = (A.tl + mean * inv) * A.xscale;
        const float rt = (float)(s_par[PF_RATE * RT + i] * A.itime);
        // levels are sorted: from the first one at or above mean + tl * sig
        // for every lane of the warp on, all the PoEs are zero (the padding
        // levels are +inf).  The threshold is reduced over the warp in
        // float32, rounded up.
        // (one REDUX over the active lanes on the order-preserving integer
        // image of the float)
        int key = __float_as_int(__double2float_ru(fma(A.tl, sig, mean)));
        key ^= (key >> 31) & 0x7fffffff;
        key = __reduce_max_sync(vm, key);
        key ^= (key >> 31) & 0x7fffffff;
        const double thr_w = A.sorted ? (double)__int_as_float(key) : INFINITY;
        if (A.count_pairs > 1 && lane == __ffs(vm) - 1) {
          // statistics pass only (oqg_set_count_evals): the evaluations this
          // warp executes for this rupture, i.e. the levels the skipping
          // below does not cut off, times the active lanes
          int nlev = 0;
#pragma unroll
          for (int l = 0; l < NLR; l += LGRP) nlev += ll[l] < thr_w ? LGRP : 0;
#pragma unroll
          for (int l = NLR; l < NL; l++) nlev += ll[l] < thr_w ? 1 : 0;
          atomicAdd(s_cnt, (unsigned long long)(nlev * __popc(vm)));
        }
        // The test is made once per LGRP levels: the evaluations of a group
        // are independent instruction chains in one basic block (a level
        // past the threshold evaluates to exactly 0: x saturates to node 0)
#if OQG_POE_PAIR
        const f32x2 rt2 = pack2(rt, rt), nm2 = pack2(nm1, nm1);
#pragma unroll
        for (int l = 0; l < NLR; l += 2) {
          if (ll[l] >= thr_w) break;
          facc2[l >> 1] = fma2(rt2, poe_table2<REP>(tab32, nm2, fma(ll[l], xa, xb),
                                                    fma(ll[l + 1], xa, xb)),
                               facc2[l >> 1]);
        }
#else
#pragma unroll
        for (int l = 0; l < NLR; l += LGRP) {
          if (ll[l] >= thr_w) break;
#pragma unroll
          for (int j = 0; j < LGRP; j++)
            facc[l + j] = fmaf(rt, poe_table<REP>(tab32, nm1, fma(ll[l + j], xa, xb)),
                               facc[l + j]);
        }
#endif
        // the levels are sorted: if the loop above stopped early, this one
        // stops at once
#pragma unroll
        for (int l = NLR; l < NL; l++) {
          const double lv = ll[l];
          if (lv >= thr_w) break;
          s_hi[(l - NLR) * hs] = fmaf(rt, poe_table<REP>(tab32, nm1, fma(lv, xa, xb)),
                                      s_hi[(l - NLR) * hs]);
        }
      } else if (NEGBIN) {
        const double *pr = s_probs + i * OQG_MAX_OCC;
#pragma unroll
        for (int l = 0; l < NL; l++) {
          double poe = truncnorm_sf(A.phi_b, A.inv_norm, (ll[l] - mean) * inv);
          acc[l] *= A.itime == 0. ? 1. - poe : pne_from_pmf(pr, A.K, poe);
        }
      } else {
        const double rt = s_par[PF_RATE * RT + i] * A.itime;
#pragma unroll
        for (int l = 0; l < NL; l++) {
          double poe = truncnorm_sf(A.phi_b, A.inv_norm, (ll[l] - mean) * inv);
          acc[l] = fma(rt, poe, acc[l]);
        }
      }
    }
    if (FAST && (--tiles_left == 0 || r0 + RT >= r_end)) {
      tiles_left = FLUSH_TILES;
      if (site_ok) {
        // all the loads first: as `dst[l] += x` the twenty read-modify-write
        // round trips to L2 were serialised (47 % of the long-scoreboard
        // stalls of the kernel)
        float t[NL];
        // the stride goes through an opaque asm so that the twenty level
        // addresses are formed here and not hoisted out of the rupture
        // loop (as loop invariants they were spilled: 392 bytes of stack)
        int64_t dsv = ds;
        asm volatile("" : "+l"(dsv));
        float *pl = dstf;
#pragma unroll
        for (int l = 0; l < NL; l++, pl += dsv) t[l] = l < nl ? __ldcg(pl) : 0.f;
        pl = dstf;
#pragma unroll
        for (int l = 0; l < NLR; l++, pl += dsv)
          if (l < nl) __stcg(pl, t[l] + FACC_GET(l));
#pragma unroll
        for (int l = NLR; l < NL; l++, pl += dsv)
          if (l < nl) __stcg(pl, t[l] + s_hi[(l - NLR) * hs]);
      }
#if OQG_POE_PAIR
#pragma unroll
      for (int l = 0; l < NLR / 2; l++) facc2[l] = 0ull;
#else
#pragma unroll
      for (int l = 0; l < NLR; l++) facc[l] = 0.f;
#endif
#pragma unroll
      for (int l = NLR; l < NL; l++) s_hi[(l - NLR) * hs] = 0.f;
    }
  }
#undef DGET
#undef LOAD_PAIR
  if (A.count_pairs && lane == 0 && kept)
    atomicAdd(A.counters, kept);  // statistics only, not the map
  if (FAST && A.count_pairs > 1) {
    SUB_SYNC();
    if (tid == 0 && *s_cnt) atomicAdd(A.counters + 1, *s_cnt);
  }
#undef SUB_SYNC
  if (FAST || !writer) return;
#pragma unroll
  for (int l = 0; l < NL; l++)
    if (l < nl) dst[l * ds] = NEGBIN ? -log(acc[l]) : acc[l];
}

// partial sums -> map[sidx[n]][l][g]
// part [RC][G][L][N] (site fastest) -> map[sidx[n]][l][g]: a 32 x 32 tile of
// (site, level-gsim) elements goes through shared memory so that both the
// reads (along the sites) and the writes (along l, g) are coalesced
__global__ void __launch_bounds__(256)
k_finalize(const double *__restrict__ part, int part_f32, int RC, int G,
           int N, int L, int L1, const uint32_t *__restrict__ sidx,
           int rates_mode, float *__restrict__ map) {
  __shared__ double tile[32][33];
  const int LG = L * G;
  const int n0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;     // (32, 8)
  for (int j = ty; j < 32; j += 8) {
    const int e = e0 + j, n = n0 + tx;              // e = l * G + g
    double s = 0.;
    if (e < LG && n < N) {
      const int l = e / G, g = e - l * G;
      const float *partf = reinterpret_cast<const float *>(part);
      for (int rc = 0; rc < RC; rc++) {
        const int64_t at = (((int64_t)rc * G + g) * L + l) * N + n;
        s += part_f32 ? (double)partf[at] : part[at];
      }
      // ImtWeight == 0 (set_poes, contexts.py:1469-1475): the PoEs of this
      // (gsim, IMT) count as zero.  Applied here, once per map element, and
      // not in the fused kernel, where one more live value cost ASK14 15 %
      if (c_imtmask[g * (L / L1) + l / L1]) s = 0.;
    }
    tile[j][tx] = s;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j, e = e0 + tx;
    if (n >= N || e >= LG) continue;
    const double s = tile[tx][j];
    const int64_t at = (int64_t)(sidx ? sidx[n] : n) * LG + e;
    if (rates_mode) map[at] = (float)((double)map[at] + s);
    else map[at] = (float)((double)map[at] * exp(-s));
  }
}

// FP64 FMA microbenchmark (8 independent chains per thread)
__global__ void k_dfma(double *out, int iters) {
  double a[8];
  double x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = k * 0.1 + threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], x, y);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------- PoE table
// F(t) = (Q(tl - t) - Q(tl)) / (1 - 2 Q(tl)), Q the upper tail of the normal
static double poe_exact(double tl, double t) {
  double qtl = 0.5 * erfc(tl * sqrt(0.5));
  return (0.5 * erfc((tl - t) * sqrt(0.5)) - qtl) / (1. - 2. * qtl);
}

// least-squares-free polynomial interpolation: solve V c = f (n <= 6)
static void solve_vandermonde(int n, const long double *u,
                              const long double *f, long double *c) {
  long double a[6][7];
  for (int i = 0; i < n; i++) {
    long double p = 1;
    for (int j = 0; j < n; j++) { a[i][j] = p; p *= u[i]; }
    a[i][n] = f[i];
  }
  for (int col = 0; col < n; col++) {
    int piv = col;
    for (int r = col + 1; r < n; r++)
      if (fabsl(a[r][col]) > fabsl(a[piv][col])) piv = r;
    for (int j = 0; j <= n; j++) std::swap(a[col][j], a[piv][j]);
    for (int r = 0; r < n; r++) {
      if (r == col) continue;
      long double fct = a[r][col] / a[col][col];
      for (int j = col; j <= n; j++) a[r][j] -= fct * a[col][j];
    }
  }
  for (int i = 0; i < n; i++) c[i] = a[i][n] / a[i][i];
}

static float poe_table_host(const float *tab, int nodes, double x01) {
  float x = (float)x01;                 // the device sequence of poe_table()
  x = fminf(fmaxf(x, 0.f), 1.f);
  float nm1 = (float)(nodes - 1);
  float y = fmaf(x, nm1, 8388608.f);
  int k = (int)(y - 8388608.f);
  float u = fmaf(x, nm1, -(y - 8388608.f));
  const float *c = tab + k * PT_STRIDE;
  float p = c[PT_DEG];
  for (int d = PT_DEG - 1; d >= 0; d--) p = fmaf(p, u, c[d]);
  return fminf(fmaxf(p, 0.f), 1.f);
}

// builds the table for truncation level tl; returns the maximum relative
// error over poe >= 1e-9 (absolute error / 1e-9 below that), sampled densely
static double build_poe_table(double tl, int nodes, float *tab,
                              double *xscale) {
  const double pi = 3.14159265358979323846;
  const double s = (nodes - 1) / (2. * tl);
  *xscale = 1. / (2. * tl);
  for (int k = 0; k < nodes; k++) {
    long double u[6], f[6], c[6] = {0, 0, 0, 0, 0, 0};
    if (k == 0) {
      // F(u / s) = u * G(u) on [0, .5]: zero constant term
      for (int j = 0; j < PT_DEG; j++) {
        u[j] = 0.25L + 0.25L * cosl((2 * j + 1) * pi / (2 * PT_DEG));
        f[j] = poe_exact(tl, (double)u[j] / s) / u[j];
      }
      solve_vandermonde(PT_DEG, u, f, c + 1);
    } else {
      for (int j = 0; j <= PT_DEG; j++) {
        u[j] = 0.5L * cosl((2 * j + 1) * pi / (2 * (PT_DEG + 1)));
        f[j] = poe_exact(tl, (k + (double)u[j]) / s);
      }
      solve_vandermonde(PT_DEG + 1, u, f, c);
    }
    for (int d = 0; d < PT_STRIDE; d++)
      tab[k * PT_STRIDE + d] = d <= PT_DEG ? (float)c[d] : 0.f;
  }
  double worst = 0.;
  const int NS = 60000;
  for (int i = 0; i <= NS; i++) {
    double x = (nodes - 1) * (double)i / NS;
    // finer sampling near the truncation point
    if (i < 3000) x = 0.5 * pow(10., -8. + 8. * i / 3000.);
    double ex = poe_exact(tl, x / s);
    double got = poe_table_host(tab, nodes, x / (nodes - 1));
    double err = ex >= 1e-9 ? fabs(got / ex - 1.) : fabs(got - ex) / 1e-9;
    worst = std::max(worst, err);
  }
  return worst;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int oqg_create(oqg_handle **out, int device) {
  oqg_handle *h = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, std::string("no CUDA device: ") +
                             cudaGetErrorString(e) +
                             " (liboqg has no CPU fallback)");
  if (device < 0 || device >= n || device >= OQG_MAX_DEVICES)
    return fail(nullptr, "bad device index");
  h = new oqg_handle();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) !=
          cudaSuccess) {
    delete h;
    return fail(nullptr, "cannot initialise the device");
  }
  h->stream = h->own_stream;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  for (auto &ev : h->ev) cudaEventCreate(&ev);
  cudaEventCreateWithFlags(&h->const_ev, cudaEventDisableTiming);
  cudaMalloc((void **)&h->flag_dev, sizeof(int));
  cudaMemset(h->flag_dev, 0, sizeof(int));
  // tuning knobs for A/B measurements (oqg_set_site_grouping)
  if (const char *e = getenv("OQG_SITE_GROUPING")) h->site_grouping = atoi(e) != 0;
  if (const char *e = getenv("OQG_WAVE_FIT")) h->wave_fit = atoi(e) != 0;
  if (const char *e = getenv("OQG_CHUNKS")) h->force_chunks = atoi(e);
  if (const char *e = getenv("OQG_SITE_CELL"))
    if (atof(e) > 0.) h->site_cell_deg = atof(e);
  *out = h;
  return 0;
}

void oqg_destroy(oqg_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (Buf *b : {&h->rup, &h->site, &h->part, &h->ptab, &h->conv,
                 &h->sg_keys[0], &h->sg_keys[1], &h->sg_idx[0], &h->sg_idx[1],
                 &h->sg_tmp, &h->sg_par[0], &h->sg_par[1], &h->sg_par[2],
                 &h->sg_par[3], &h->sg_sidx})
    if (b->p) cudaFree(b->p);
  for (auto &b : h->stage) if (b.p) cudaFree(b.p);
  for (auto &b : h->hs_site) if (b.p) cudaFree(b.p);
  for (auto &sl : h->hs_slot) for (auto &b : sl) if (b.p) cudaFree(b.p);
  if (h->hs_map.p) cudaFree(h->hs_map.p);
  if (h->hs_copy) {
    cudaStreamDestroy(h->hs_copy);
    for (int i = 0; i < 2; i++) {
      cudaEventDestroy(h->hs_copied[i]);
      cudaEventDestroy(h->hs_used[i]);
    }
  }
  if (h->flag_dev) cudaFree(h->flag_dev);
  for (auto &ev : h->ev) cudaEventDestroy(ev);
  {
    // the stream is idle: nothing of this handle reads the tables any more
    std::lock_guard<std::mutex> lk(g_const_mu);
    if (g_const_owner[h->device] == h) g_const_owner[h->device] = nullptr;
  }
  cudaEventDestroy(h->const_ev);
  cudaStreamDestroy(h->own_stream);
  delete h;
}

const char *oqg_last_error(oqg_handle *h) {
  return h ? h->err.c_str() : g_create_err.c_str();
}

int oqg_set_stream(oqg_handle *h, void *s) {
  h->stream = (cudaStream_t)s;  // NULL is the legacy default stream
  return 0;
}

int oqg_use_own_stream(oqg_handle *h) {
  h->stream = h->own_stream;
  return 0;
}

int oqg_synchronize(oqg_handle *h) {
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

const char *oqg_family_fields(int family) {
  switch (family) {
    case OQG_BSSA14: return Fam<OQG_BSSA14>::fields();
    case OQG_CY14: return Fam<OQG_CY14>::fields();
    case OQG_ASK14: return Fam<OQG_ASK14>::fields();
    case OQG_CB14: return Fam<OQG_CB14>::fields();
    case OQG_BA08: return Fam<OQG_BA08>::fields();
    case OQG_AB06: return Fam<OQG_AB06>::fields();
    case OQG_AKKAR14: return Fam<OQG_AKKAR14>::fields();
    case OQG_CY08: return Fam<OQG_CY08>::fields();
    case OQG_TORO02: return Fam<OQG_TORO02>::fields();
    case OQG_C03: return Fam<OQG_C03>::fields();
    case OQG_SADIGH97: return Fam<OQG_SADIGH97>::fields();
  }
  return nullptr;
}

int oqg_required_distances(const oqg_gsim *g) { return required_distances(*g); }

int oqg_set_gsims(oqg_handle *h, int G, const oqg_gsim *gsims, int M,
                  const double *coeffs) {
  if (G < 1 || G > OQG_MAX_G) return fail(h, "G must be in 1..16");
  if (M < 1 || G * (M + 1) > OQG_MAX_SLOTS)
    return fail(h, "too many (gsim, IMT) pairs for the coefficient table "
                   "(G * (M + 1) <= 96)");
  for (int g = 0; g < G; g++) {
    const oqg_gsim &s = gsims[g];
    if (s.family < 0 || s.family >= OQG_NFAMILIES)
      return fail(h, "unsupported GMPE family " + std::to_string(s.family) +
                         " (no fallback)");
    int reg = s.flags & 3;
    bool bad = false;
    if (s.family == OQG_BSSA14 && reg > 2) bad = true;
    if (s.family == OQG_AB06 && reg > 2) bad = true;
    if (s.family == OQG_AKKAR14 && reg > 2) bad = true;
    if (bad) return fail(h, "unsupported flags for GMPE family");
  }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->G = G; h->M = M;
  h->has_conv = false;
  h->gsims.assign(gsims, gsims + G);
  h->imtmask.assign(OQG_MAX_SLOTS, 0);
  h->coef.assign(coeffs, coeffs + (size_t)G * (M + 1) * OQG_KMAX);
  // derived per-IMT constants of the staged model code (oqg_fast.cuh)
  for (int g = 0; g < G; g++)
    for (int m = 0; m <= M; m++) {
      double *row = h->coef.data() + ((size_t)g * (M + 1) + m) * OQG_KMAX;
      if (gsims[g].family == OQG_CY14) {
        const Cy14C *c = reinterpret_cast<const Cy14C *>(row);
        row[OQG_DER + CY_LOG_PHI4] = log(c->phi4);
        row[OQG_DER + CY_INV_C9B] = 1. / c->c9b;
      }
    }
  h->const_dirty = true;
  return 0;
}

int oqg_set_conversion(oqg_handle *h, const double *conv) {
  if (!h->G || !h->M) return fail(h, "set_conversion: set_gsims not called");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->has_conv = false;
  if (!conv) return 0;
  std::vector<double> t((size_t)h->G * h->M * 3, 0.);
  bool any = false;
  for (int i = 0; i < h->G * h->M; i++) {
    const double cm = conv[3 * i], cs = conv[3 * i + 1], rs = conv[3 * i + 2];
    if (cm == 0. && cs == 0. && rs == 0.) continue;   // no conversion
    if (!(cm > 0.) || !(rs > 0.))
      return fail(h, "set_conversion: conv_median and rstd must be positive");
    t[3 * i] = log(cm); t[3 * i + 1] = cs * cs; t[3 * i + 2] = 1. / (rs * rs);
    any = true;
  }
  if (!any) return 0;
  if (reserve(h, h->conv, t.size() * 8)) return 1;
  CK(cudaMemcpy(h->conv.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
  h->has_conv = true;
  return 0;
}

int oqg_set_imt_mask(oqg_handle *h, const unsigned char *mask) {
  if (!h->G || !h->M) return fail(h, "set_imt_mask: set_gsims not called");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->imtmask.assign(OQG_MAX_SLOTS, 0);
  if (mask)
    for (int i = 0; i < h->G * h->M; i++) h->imtmask[i] = mask[i] ? 1 : 0;
  h->const_dirty = true;
  return 0;
}

int oqg_set_levels(oqg_handle *h, int M, int L1, const double *loglevels,
                   double truncation_level) {
  if (M < 1 || L1 < 1 || M * L1 > OQG_MAX_LEVELS)
    return fail(h, "M * L1 must be in 1..1024");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->M && h->M != M) return fail(h, "set_levels: M differs from set_gsims");
  h->M = M; h->L1 = L1;
  h->loglev.assign(loglevels, loglevels + (size_t)M * L1);
  h->sorted = true;
  for (int m = 0; m < M; m++)
    for (int l = 1; l < L1; l++)
      if (!(loglevels[m * L1 + l] > loglevels[m * L1 + l - 1]))
        h->sorted = false;
  {
    int nlc = (L1 + NL - 1) / NL, lpc = (L1 + nlc - 1) / nlc;
    if (M * nlc * NL > OQG_MAX_LEVELS)
      return fail(h, "too many levels for the fused kernel");
    h->loglev_pad.assign((size_t)M * nlc * NL, INFINITY);
    for (int m = 0; m < M; m++)
      for (int l = 0; l < L1; l++)
        h->loglev_pad[(size_t)(m * nlc + l / lpc) * NL + l % lpc] =
            loglevels[m * L1 + l];
  }
  h->truncation_level = truncation_level;
  h->phi_b = 0.5 * (1.0 + erf(truncation_level * sqrt(0.5)));
  h->const_dirty = true;
  // fast PoE table: used when it reproduces the exact formula to 2e-6
  h->fast_ok = false;
  if (truncation_level >= 0.5 && truncation_level <= 8.) {
    float tab[PT_MAXNODES * PT_STRIDE];
    for (int nodes = 256; nodes <= PT_MAXNODES && !h->fast_ok; nodes *= 2) {
      h->ptab_err = build_poe_table(truncation_level, nodes, tab, &h->xscale);
      if (h->ptab_err <= 2e-6) {
        if (reserve(h, h->ptab, sizeof(tab))) return 1;
        CK(cudaMemcpy(h->ptab.p, tab, sizeof(tab), cudaMemcpyHostToDevice));
        h->fast_ok = true;
        h->pt_nodes = nodes;
      }
    }
  }
  return 0;
}

static int prep_sites(oqg_handle *h, const oqg_sites *s, int N) {
  if (!s->lon || !s->lat) return fail(h, "sites need lon and lat");
  if (reserve(h, h->site, (size_t)NF_SITE * N * 8)) return 1;
  k_prep_sites<<<(N + 127) / 128, 128, 0, h->stream>>>(
      s->lon, s->lat, s->depth, N, (double *)h->site.p);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- site grouping ----------------------------------------------------------
// The fused kernel skips, for a whole warp of 32 sites, the levels above the
// largest mean + truncation_level * sigma of the warp (their PoEs are exactly
// 0).  How often that fires depends on how alike the 32 sites are: on a grid
// handed over row by row, with unrelated vs30 values, 32 % of the levels are
// skipped on the cfg3 workload; with the sites of a warp taken from one
// ~18 km cell and neighbouring in vs30 it is 43 % (a thread alone would skip
// 50 %).  The sites are therefore sorted by (cell row, cell column, vs30)
// before the pass -- a radix sort of N 64-bit keys on the device -- and the
// map rows go through the permuted site index.  Per site nothing changes:
// the same ruptures in the same order, so the map is bit-identical.
__global__ void k_site_keys(const double *__restrict__ lon,
                            const double *__restrict__ lat,
                            const double *__restrict__ vs30, int N,
                            double inv_cell, unsigned long long *keys,
                            int *idx) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double x = (lon[n] + 180.) * inv_cell, y = (lat[n] + 90.) * inv_cell;
  const unsigned long long cx = (unsigned long long)fmin(fmax(x, 0.), 65535.);
  const unsigned long long cy = (unsigned long long)fmin(fmax(y, 0.), 65535.);
  const float v = vs30 ? (float)vs30[n] : 0.f;
  const unsigned vb = v > 0.f ? __float_as_uint(v) : 0u;   // NaN, <= 0 -> 0
  keys[n] = (cy << 48) | (cx << 32) | vb;
  idx[n] = n;
}

__global__ void k_prep_sites_perm(const double *__restrict__ lon,
                                  const double *__restrict__ lat,
                                  const double *__restrict__ dep,
                                  const double *__restrict__ vs30,
                                  const double *__restrict__ z1pt0,
                                  const double *__restrict__ z2pt5,
                                  const uint8_t *__restrict__ measured,
                                  const uint32_t *__restrict__ sidx,
                                  const int *__restrict__ perm, int N,
                                  double *__restrict__ out,
                                  double *__restrict__ o_vs30,
                                  double *__restrict__ o_z1pt0,
                                  double *__restrict__ o_z2pt5,
                                  uint8_t *__restrict__ o_measured,
                                  uint32_t *__restrict__ o_sidx) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int j = perm[n];
  prep_site(lon[j], lat[j], dep ? dep[j] : 0., out + n, N);
  if (vs30) o_vs30[n] = vs30[j];
  if (z1pt0) o_z1pt0[n] = z1pt0[j];
  if (z2pt5) o_z2pt5[n] = z2pt5[j];
  if (measured) o_measured[n] = measured[j];
  o_sidx[n] = sidx ? sidx[j] : (uint32_t)j;
}

// prepared sites in grouped order: `view` gets the permuted parameter arrays
// and *sidx_out the permuted map rows
static int prep_sites_grouped(oqg_handle *h, const oqg_sites *s,
                              const uint32_t *sidx, int N, oqg_sites *view,
                              const uint32_t **sidx_out) {
  if (!s->lon || !s->lat) return fail(h, "sites need lon and lat");
  if (reserve(h, h->site, (size_t)NF_SITE * N * 8)) return 1;
  for (int k = 0; k < 2; k++) {
    if (reserve(h, h->sg_keys[k], (size_t)N * 8)) return 1;
    if (reserve(h, h->sg_idx[k], (size_t)N * 4)) return 1;
  }
  for (int k = 0; k < 3; k++)
    if (reserve(h, h->sg_par[k], (size_t)N * 8)) return 1;
  if (reserve(h, h->sg_par[3], (size_t)N)) return 1;
  if (reserve(h, h->sg_sidx, (size_t)N * 4)) return 1;
  auto *k0 = (unsigned long long *)h->sg_keys[0].p;
  auto *k1 = (unsigned long long *)h->sg_keys[1].p;
  int *i0 = (int *)h->sg_idx[0].p, *i1 = (int *)h->sg_idx[1].p;
  k_site_keys<<<(N + 255) / 256, 256, 0, h->stream>>>(
      s->lon, s->lat, s->vs30, N, 1. / h->site_cell_deg, k0, i0);
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, i0, i1, N, 0, 64,
                                     h->stream));
  if (reserve(h, h->sg_tmp, tmp)) return 1;
  CK(cub::DeviceRadixSort::SortPairs(h->sg_tmp.p, tmp, k0, k1, i0, i1, N, 0,
                                     64, h->stream));
  k_prep_sites_perm<<<(N + 127) / 128, 128, 0, h->stream>>>(
      s->lon, s->lat, s->depth, s->vs30, s->z1pt0, s->z2pt5, s->vs30measured,
      sidx, i1, N, (double *)h->site.p, (double *)h->sg_par[0].p,
      (double *)h->sg_par[1].p, (double *)h->sg_par[2].p,
      (uint8_t *)h->sg_par[3].p, (uint32_t *)h->sg_sidx.p);
  h->launches += 3;
  CK(cudaGetLastError());
  *view = *s;
  if (s->vs30) view->vs30 = (const double *)h->sg_par[0].p;
  if (s->z1pt0) view->z1pt0 = (const double *)h->sg_par[1].p;
  if (s->z2pt5) view->z2pt5 = (const double *)h->sg_par[2].p;
  if (s->vs30measured) view->vs30measured = (const uint8_t *)h->sg_par[3].p;
  *sidx_out = (const uint32_t *)h->sg_sidx.p;
  return 0;
}

static int prep_ruptures(oqg_handle *h, const double *planar,
                         const double *mag, const double *magdist, int64_t U) {
  if (reserve(h, h->rup, (size_t)NF_RUP * rup_stride(U) * 8)) return 1;
  k_prep_ruptures<<<(unsigned)((U + 127) / 128), 128, 0, h->stream>>>(
      planar, mag, magdist, U, rup_stride(U), (double *)h->rup.p);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_mesh_min_distance(oqg_handle *h, const double *lon, const double *lat,
                          const double *depth, int64_t P, const int64_t *off,
                          int S, const oqg_sites *sites, int N, double *out) {
  CK(cudaSetDevice(h->device));
  if (S <= 0 || N <= 0) return 0;
  if (P <= 0 || !lon || !lat || !off || !out)
    return fail(h, "mesh_min_distance: missing array");
  if (prep_sites(h, sites, N)) return 1;
  if (reserve(h, h->stage[9], (size_t)P * 3 * 8)) return 1;
  double *pts = (double *)h->stage[9].p;
  k_mesh_xyz<<<(unsigned)((P + 255) / 256), 256, 0, h->stream>>>(lon, lat, depth,
                                                                P, pts);
  h->launches++;
  for (int y0 = 0; y0 < S; y0 += OQG_GRID_Y) {   // gridDim.y <= 65535
    dim3 grid((N + 128 * MESH_SPT - 1) / (128 * MESH_SPT),
              std::min(S - y0, OQG_GRID_Y));
    k_mesh_min_distance<<<grid, 128, 0, h->stream>>>(
        pts, P, off, S, (const double *)h->site.p, N, out, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_mesh_rx_ry0(oqg_handle *h, const double *top_lon,
                    const double *top_lat, const int64_t *off,
                    const double *strike, int S, const oqg_sites *sites, int N,
                    double *rx, double *ry0) {
  CK(cudaSetDevice(h->device));
  if (S <= 0 || N <= 0) return 0;
  if (!top_lon || !top_lat || !off || (!rx && !ry0))
    return fail(h, "mesh_rx_ry0: missing array");
  if (ry0 && !strike) return fail(h, "mesh_rx_ry0: ry0 needs the mean strike");
  if (prep_sites(h, sites, N)) return 1;
  for (int y0 = 0; y0 < S; y0 += OQG_GRID_Y) {
    dim3 grid((N + 127) / 128, std::min(S - y0, OQG_GRID_Y));
    k_mesh_rx_ry0<<<grid, 128, 0, h->stream>>>(top_lon, top_lat, off, strike,
                                              (const double *)h->site.p, N,
                                              rx, ry0, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_mesh_rjb(oqg_handle *h, const double *lon, const double *lat,
                 int64_t P, const int64_t *off, const int32_t *shape,
                 const int32_t *owner, const double *frame, int S,
                 const oqg_sites *sites, int N, double *out) {
  CK(cudaSetDevice(h->device));
  if (S <= 0 || N <= 0) return 0;
  if (P <= 0 || !lon || !lat || !off || !shape || !owner || !frame || !out)
    return fail(h, "mesh_rjb: missing array");
  if (prep_sites(h, sites, N)) return 1;
  if (reserve(h, h->stage[9], (size_t)P * 3 * 8)) return 1;
  double *pts = (double *)h->stage[9].p;
  k_mesh_frame<<<(unsigned)((P + 255) / 256), 256, 0, h->stream>>>(
      lon, lat, P, owner, frame, pts);
  h->launches++;
  for (int y0 = 0; y0 < S; y0 += OQG_GRID_Y) {   // gridDim.y <= 65535
    dim3 grid((N + 127) / 128, std::min(S - y0, OQG_GRID_Y));
    k_mesh_rjb<<<grid, 128, 0, h->stream>>>(pts, P, off, shape, frame,
                                           (const double *)h->site.p, N, out,
                                           y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_line_tuw(oqg_handle *h, const double *seg, const int64_t *loff,
                 const double *proj, int S2, const oqg_sites *sites, int N,
                 float *out) {
  CK(cudaSetDevice(h->device));
  if (S2 <= 0 || N <= 0) return 0;
  if (!seg || !loff || !proj || !out || !sites->lon || !sites->lat)
    return fail(h, "line_tuw: missing array");
  if (S2 > OQG_GRID_Y) return fail(h, "line_tuw: at most 65535 lines per call");
  dim3 grid((N + 127) / 128, S2);
  k_line_tuw<<<grid, 128, 0, h->stream>>>(seg, loff, proj, S2, sites->lon,
                                          sites->lat, N, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_multi_compose(oqg_handle *h, const double *sec_rrup,
                      const double *sec_rjb, const float *tuw, int S, int N,
                      const int64_t *roff, const int32_t *rsec,
                      const uint8_t *rflip, const float *rshift,
                      const float *umax, int U, double *rrup, double *rjb,
                      double *rx, double *ry0) {
  CK(cudaSetDevice(h->device));
  if (U <= 0 || N <= 0) return 0;
  if (S <= 0 || !roff || !rsec) return fail(h, "multi_compose: missing array");
  if ((rx || ry0) && (!tuw || !rflip || !rshift))
    return fail(h, "multi_compose: rx / ry0 need tuw, flips and shifts");
  if (ry0 && !umax) return fail(h, "multi_compose: ry0 needs u_max");
  if ((rrup && !sec_rrup) || (rjb && !sec_rjb))
    return fail(h, "multi_compose: section plane missing");
  for (int y0 = 0; y0 < U; y0 += OQG_GRID_Y) {
    dim3 grid((N + 127) / 128, std::min(U - y0, OQG_GRID_Y));
    k_multi_compose<<<grid, 128, 0, h->stream>>>(
        rrup ? sec_rrup : nullptr, rjb ? sec_rjb : nullptr,
        (rx || ry0) ? tuw : nullptr, N, roff, rsec, rflip, rshift, umax, U,
        rrup, rjb, rx, ry0, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_mesh_dip_strike_width(oqg_handle *h, const double *lon,
                              const double *lat, const double *depth,
                              const int64_t *off, const int32_t *shape, int S,
                              double *out) {
  CK(cudaSetDevice(h->device));
  if (S <= 0) return 0;
  if (!lon || !lat || !off || !shape || !out)
    return fail(h, "mesh_dip_strike_width: missing array");
  k_mesh_dip_strike_width<<<S, MESH_DIP_TPB, 0, h->stream>>>(lon, lat, depth,
                                                            off, shape, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_section_min(oqg_handle *h, const double *sec, int S, int N,
                    const int64_t *rup_off, const int32_t *rup_sec, int U,
                    double *out) {
  CK(cudaSetDevice(h->device));
  if (U <= 0 || N <= 0) return 0;
  if (!sec || !rup_off || !rup_sec || !out || S <= 0)
    return fail(h, "section_min: missing array");
  for (int y0 = 0; y0 < U; y0 += OQG_GRID_Y) {
    dim3 grid((N + 127) / 128, std::min(U - y0, OQG_GRID_Y));
    k_section_min<<<grid, 128, 0, h->stream>>>(sec, N, rup_off, rup_sec, U,
                                              out, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_disaggregate(oqg_handle *h, const double *rrup, const double *clon,
                     const double *clat, const double *rate, int64_t U,
                     const double *mea, const double *std, int M, int P,
                     const double *iml2, const double *dist_edges, int D,
                     const double *lon_edges, int Lo, const double *lat_edges,
                     int La, const double *eps_edges, int E,
                     double truncation_level, double itime, int epsstar,
                     double gp, double *out) {
  CK(cudaSetDevice(h->device));
  if (M < 1 || P < 1 || D < 1 || Lo < 1 || La < 1 || E < 1)
    return fail(h, "disaggregate: bad shape");
  if (E > OQG_DISAGG_EMAX) return fail(h, "disaggregate: at most 32 epsilon bins");
  if (M * P > 4096) return fail(h, "disaggregate: M * P too large");
  if (!out || !iml2 || !dist_edges || !lon_edges || !lat_edges || !eps_edges)
    return fail(h, "disaggregate: missing array");
  const int64_t n = (int64_t)D * Lo * La * E * M * P;
  CK(cudaMemsetAsync(out, 0, (size_t)n * 8, h->stream));
  if (U > 0) {
    if (!rrup || !clon || !clat || !rate || !mea || !std)
      return fail(h, "disaggregate: missing column");
    // small host tables -> one staging buffer
    size_t nd = (size_t)M * P + D + 1 + Lo + 1 + La + 1;
    if (reserve(h, h->stage[8], nd * 8)) return 1;
    std::vector<double> tab;
    tab.insert(tab.end(), iml2, iml2 + (size_t)M * P);
    tab.insert(tab.end(), dist_edges, dist_edges + D + 1);
    tab.insert(tab.end(), lon_edges, lon_edges + Lo + 1);
    tab.insert(tab.end(), lat_edges, lat_edges + La + 1);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->stage[8].p, tab.data(), nd * 8, cudaMemcpyHostToDevice));
    DisaggArgs A;
    memset(&A, 0, sizeof(A));
    A.rrup = rrup; A.clon = clon; A.clat = clat; A.rate = rate;
    A.mea = mea; A.std = std;
    A.iml2 = (const double *)h->stage[8].p;
    A.dist_edges = A.iml2 + (size_t)M * P;
    A.lon_edges = A.dist_edges + D + 1;
    A.lat_edges = A.lon_edges + Lo + 1;
    A.U = U; A.M = M; A.P = P; A.D = D; A.Lo = Lo; A.La = La; A.E = E;
    // cross_idl, geo/utils.py:802-827
    A.lon_idl = lon_edges[0] * lon_edges[Lo] < 0 &&
                fabs(lon_edges[0] - lon_edges[Lo]) > 180;
    A.epsstar = epsstar; A.gp = gp; A.itime = itime;
    A.phi_b = 0.5 * (1.0 + erf(truncation_level * sqrt(0.5)));
    A.inv_norm = 1. / (2. * A.phi_b - 1.);
    // get_eps4, disagg.py:176-186: scipy.stats.truncnorm(-tl, tl).cdf
    auto cdf = [&](double x) {
      double c = (0.5 * (1.0 + erf(x * sqrt(0.5))) - (1. - A.phi_b)) * A.inv_norm;
      return c < 0. ? 0. : (c > 1. ? 1. : c);
    };
    A.min_eps = eps_edges[0]; A.max_eps = eps_edges[0];
    for (int e = 0; e <= E; e++) {
      A.eps_edges[e] = eps_edges[e];
      A.min_eps = std::min(A.min_eps, eps_edges[e]);
      A.max_eps = std::max(A.max_eps, eps_edges[e]);
    }
    for (int e = 0; e < E; e++) A.bands[e] = cdf(eps_edges[e + 1]) - cdf(eps_edges[e]);
    for (int e = 0; e <= E; e++) {
      double c = 0.;
      for (int k = e; k < E; k++) c += A.bands[k];
      A.cum[e] = c;
    }
    A.acc = out;
    int64_t nt = U * M * P;
    k_disagg_accumulate<<<(unsigned)((nt + 255) / 256), 256, 0, h->stream>>>(A);
    h->launches++;
  }
  k_disagg_finalize<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(out, n);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_cs_out(oqg_handle *h, const double *mu, const double *sig,
               const double *probs, int64_t U, int M, const double *rho,
               int imti, const double *imls, const double *cs_poes, int P,
               double truncation_level, double invtime, const double *prev,
               double *out) {
  CK(cudaSetDevice(h->device));
  if (M < 1 || P < 1 || imti < 0 || imti >= M) return fail(h, "cs_out: bad shape");
  if (!rho || !imls || !cs_poes || !out) return fail(h, "cs_out: missing array");
  const size_t nout = (size_t)M * 3 * P;
  if (U <= 0) {
    for (size_t i = 0; i < nout; i++) out[i] = 0.;
    return 0;
  }
  if (!mu || !sig || !probs) return fail(h, "cs_out: missing column");
  size_t nd = (size_t)M + 2 * P + (prev ? (size_t)M * P : 0) + nout;
  if (reserve(h, h->stage[8], nd * 8)) return 1;
  std::vector<double> tab;
  tab.insert(tab.end(), rho, rho + M);
  tab.insert(tab.end(), imls, imls + P);
  tab.insert(tab.end(), cs_poes, cs_poes + P);
  if (prev) tab.insert(tab.end(), prev, prev + (size_t)M * P);
  CK(cudaStreamSynchronize(h->stream));
  double *d = (double *)h->stage[8].p;
  CK(cudaMemcpy(d, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
  CsArgs A;
  memset(&A, 0, sizeof(A));
  A.mu = mu; A.sig = sig; A.probs = probs;
  A.rho = d; A.imls = d + M; A.cs_poes = A.imls + P;
  A.prev = prev ? A.cs_poes + P : nullptr;
  A.out = d + tab.size();
  A.U = U; A.M = M; A.P = P; A.imti = imti;
  A.phi_b = 0.5 * (1.0 + erf(truncation_level * sqrt(0.5)));
  A.inv_norm = 1. / (2. * A.phi_b - 1.);
  A.invtime = invtime;
  k_cs_out<<<dim3(P, M + 1), 256, 0, h->stream>>>(A);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, A.out, nout * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int oqg_gmf(oqg_handle *h, const float *ms, const float *within,
            const float *between, const unsigned char *not_mmi, int M, int N,
            int E, int mode, float *out) {
  CK(cudaSetDevice(h->device));
  if (M < 1 || N < 0 || E < 0 || mode < 0 || mode > 3)
    return fail(h, "gmf: bad shape or mode");
  if (N == 0 || E == 0) return 0;
  if (!ms || !out || !not_mmi) return fail(h, "gmf: missing array");
  if (mode >= 1 && !within) return fail(h, "gmf: missing within epsilons");
  if (mode >= 2 && !between) return fail(h, "gmf: missing between epsilons");
  if (M > 256) return fail(h, "gmf: at most 256 IMTs");
  if (reserve(h, h->stage[8], 256)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(h->stage[8].p, not_mmi, M, cudaMemcpyHostToDevice));
  int64_t n = (int64_t)M * N * E;
  k_gmf<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
      ms, within, between, (const unsigned char *)h->stage[8].p, M, N, E, mode,
      out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---- sampling side of the ground motion fields (oqg_gmf.cuh) ---------------
typedef unsigned __int128 h128;
static const h128 PCG_MULT =
    ((h128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;

// coefficients of `delta` steps of state -> mult * state + inc
static void pcg_advance_coeffs(h128 inc, uint64_t delta, h128 &am, h128 &ap) {
  h128 cm = PCG_MULT, cp = inc;
  am = 1; ap = 0;
  while (delta) {
    if (delta & 1) { am *= cm; ap = ap * cm + cp; }
    cp = (cm + 1) * cp;
    cm *= cm;
    delta >>= 1;
  }
}

int oqg_truncnorm_rvs(oqg_handle *h, const uint64_t *state, const uint64_t *inc,
                      uint64_t offset, int64_t n, double truncation_level,
                      float *out) {
  CK(cudaSetDevice(h->device));
  if (n <= 0) return 0;
  if (!state || !inc || !out) return fail(h, "truncnorm_rvs: missing array");
  if (!(truncation_level > 0.)) return fail(h, "truncnorm_rvs: truncation_level must be positive");
  const h128 vinc = ((h128)inc[0] << 64) | inc[1];
  h128 s0 = ((h128)state[0] << 64) | state[1], am, ap;
  pcg_advance_coeffs(vinc, offset, am, ap);
  s0 = am * s0 + ap;
  if (reserve(h, h->stage[11], sizeof(Pcg64Jump))) return 1;
  if (!h->pcg_tab_ok || h->pcg_inc[0] != inc[0] || h->pcg_inc[1] != inc[1]) {
    Pcg64Jump J;
    h128 cm = PCG_MULT, cp = vinc;
    for (int i = 0; i < 64; i++) {
      J.mult_hi[i] = (uint64_t)(cm >> 64); J.mult_lo[i] = (uint64_t)cm;
      J.plus_hi[i] = (uint64_t)(cp >> 64); J.plus_lo[i] = (uint64_t)cp;
      cp = (cm + 1) * cp;
      cm *= cm;
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->stage[11].p, &J, sizeof(J), cudaMemcpyHostToDevice));
    h->pcg_inc[0] = inc[0]; h->pcg_inc[1] = inc[1];
    h->pcg_tab_ok = true;
  }
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
  pcg_advance_coeffs(vinc, (uint64_t)(blocks * 256), am, ap);
  const double pa = 0.5 * erfc(truncation_level * M_SQRT1_2);
  const double pb = 0.5 * erfc(-truncation_level * M_SQRT1_2);
  k_truncnorm_rvs<<<(unsigned)blocks, 256, 0, h->stream>>>(
      (const Pcg64Jump *)h->stage[11].p, (uint64_t)(s0 >> 64), (uint64_t)s0,
      inc[0], inc[1], (uint64_t)(am >> 64), (uint64_t)am, (uint64_t)(ap >> 64),
      (uint64_t)ap, n, pa, pb - pa, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_corr_matrix(oqg_handle *h, const double *lon, const double *lat, int N,
                    double coeff, double *out) {
  CK(cudaSetDevice(h->device));
  if (N <= 0) return 0;
  if (!lon || !lat || !out) return fail(h, "corr_matrix: missing array");
  const int64_t n = (int64_t)N * N;
  k_corr_matrix<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(lon, lat, N,
                                                                  coeff, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_gmf_max_min(oqg_handle *h, float *gmf, const float *mean,
                    const double *max_iml, const double *min_iml,
                    const unsigned char *not_mmi, int M, int N, int E) {
  CK(cudaSetDevice(h->device));
  if (M < 1 || M > 256) return fail(h, "gmf_max_min: 1 .. 256 IMTs");
  if (N <= 0 || E <= 0) return 0;
  if (!gmf || !mean || !max_iml || !min_iml || !not_mmi)
    return fail(h, "gmf_max_min: missing array");
  if (reserve(h, h->stage[8], 16 * 256 + 256)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  double *d = (double *)h->stage[8].p;
  CK(cudaMemcpy(d, max_iml, M * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d + 256, min_iml, M * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d + 512, not_mmi, M, cudaMemcpyHostToDevice));
  const int64_t n = (int64_t)N * E;
  k_gmf_max_min<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
      gmf, mean, d, d + 256, (const unsigned char *)(d + 512), M, N, E);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_postclassical(oqg_handle *h, const float *rates, int N, int L, int Gt,
                      int M, const int32_t *g_off, const int32_t *g_rlz, int R,
                      const double *weights, int W, const double *weight_sums,
                      int S, const int32_t *kinds, const double *qs,
                      int use_rates, float *rlz_out, float *stat_out) {
  CK(cudaSetDevice(h->device));
  if (N <= 0 || L <= 0) return 0;
  if (R < 1) return fail(h, "postclassical: R must be positive");
  if (S < 0 || S > OQG_SMAX) return fail(h, "postclassical: at most 16 statistics");
  if (M < 1 || L % M) return fail(h, "postclassical: L must be a multiple of M");
  if (W != 1 && W != M) return fail(h, "postclassical: W must be 1 or M");
  if (W > 64) return fail(h, "postclassical: at most 64 weight columns");
  if (!rates || !g_off || !g_rlz) return fail(h, "postclassical: missing input");
  if (S && (!weights || !weight_sums || !kinds || !qs || !stat_out))
    return fail(h, "postclassical: missing weights / statistics / output");
  PostArgs A;
  memset(&A, 0, sizeof(A));
  A.rates = rates; A.g_off = g_off; A.g_rlz = g_rlz; A.weights = weights;
  A.rlz_out = rlz_out; A.stat_out = stat_out;
  A.n_elems = (int64_t)N * L; A.L = L; A.L1 = L / M; A.Gt = Gt; A.R = R; A.W = W;
  A.S = S; A.use_rates = use_rates;
  for (int w = 0; w < W && S; w++) A.wsum[w] = weight_sums[w];
  for (int s = 0; s < S; s++) {
    if (kinds[s] < OQG_STAT_MEAN || kinds[s] > OQG_STAT_MAX)
      return fail(h, "postclassical: unknown statistic");
    A.kind[s] = kinds[s];
    A.q[s] = qs[s];
  }
  bool any_quantile = false;
  for (int s = 0; s < S; s++) any_quantile |= kinds[s] == OQG_STAT_QUANTILE;
  // the gsim columns of every realization (inverse of g_off / g_rlz), and the
  // GROUPS of realizations with the same columns, hence the same curve
  std::vector<int32_t> goff(Gt + 1);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(goff.data(), g_off, (Gt + 1) * 4, cudaMemcpyDeviceToHost));
  if (goff[0] != 0 || goff[Gt] < 0)
    return fail(h, "postclassical: bad g_off");
  std::vector<int32_t> grlz(goff[Gt]);
  if (goff[Gt])
    CK(cudaMemcpy(grlz.data(), g_rlz, (size_t)goff[Gt] * 4, cudaMemcpyDeviceToHost));
  std::vector<std::vector<int32_t>> inv(R);
  for (int g = 0; g < Gt; g++)
    for (int k = goff[g]; k < goff[g + 1]; k++) {
      if (grlz[k] < 0 || grlz[k] >= R)
        return fail(h, "postclassical: realization index out of range");
      inv[grlz[k]].push_back(g);
    }
  std::vector<int32_t> grp(R);
  std::vector<std::vector<int32_t>> members;
  if (h->post_grouping) {
    std::map<std::vector<int32_t>, int32_t> seen;
    for (int r = 0; r < R; r++) {
      auto it = seen.find(inv[r]);
      if (it == seen.end()) {
        it = seen.emplace(inv[r], (int32_t)members.size()).first;
        members.emplace_back();
      }
      grp[r] = it->second;
      members[it->second].push_back(r);
    }
  } else {
    for (int r = 0; r < R; r++) { grp[r] = r; members.push_back({r}); }
  }
  const int D = (int)members.size();
  int rp_ = 4;
  while (rp_ < D) rp_ <<= 1;
  const int Wn = std::max(W, 1);
  const size_t smem_reg = (size_t)rp_ * (rp_ <= 32 ? 128 : 64) * 8 +
                          (size_t)Wn * D * 20 + (size_t)(D + 1) * 8 +
                          (size_t)goff[Gt] * 4 + 16 +
                          (size_t)(rp_ <= 32 ? 128 : 64) * (Gt | 1) * 4;
  const bool reg = D <= 256 && smem_reg <= 200 * 1024;
  if (any_quantile && !reg && R > OQG_RMAX_CTA)
    return fail(h, "postclassical: quantiles need at most 256 distinct sets of "
                   "gsim columns or R <= 4096 (mean, std and max have no limit)");
  const bool stream = !reg && !any_quantile;
  const bool cta = !reg && any_quantile && R > 256;
  h->last_post_path = reg ? 0 : stream ? 1 : cta ? 2 : 3;
  h->last_post_groups = reg ? D : R;
  std::vector<double> wh((size_t)Wn * R, 0.);
  if (S && (reg || cta))
    CK(cudaMemcpy(wh.data(), weights, wh.size() * 8, cudaMemcpyDeviceToHost));
  // ranks of n weights per weight column, equal weights get equal ranks
  auto ranks_of = [&](const std::vector<double> &x, int n) {
    std::vector<int32_t> rank((size_t)Wn * n, 0);
    for (int c = 0; c < Wn; c++) {
      std::vector<int> ord(n);
      for (int r = 0; r < n; r++) ord[r] = r;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
        return x[(size_t)c * n + a] < x[(size_t)c * n + b]; });
      int rk = 0;
      for (int i = 0; i < n; i++) {
        if (i && x[(size_t)c * n + ord[i]] != x[(size_t)c * n + ord[i - 1]]) rk = i;
        rank[(size_t)c * n + ord[i]] = rk;
      }
    }
    return rank;
  };
  PostRegArgs B;
  memset(&B, 0, sizeof(B));
  B.rates = rates; B.rlz_out = rlz_out; B.stat_out = stat_out;
  B.n_elems = A.n_elems; B.L = L; B.L1 = L / M; B.Gt = Gt;
  B.W = Wn; B.S = S; B.use_rates = use_rates;
  memcpy(B.wsum, A.wsum, sizeof(B.wsum));
  memcpy(B.kind, A.kind, sizeof(B.kind));
  memcpy(B.q, A.q, sizeof(B.q));
  if (reg) {
    // tables of the groups: columns, member realizations, W_d, w_d, rank of w_d
    std::vector<int32_t> roff(D + 1, 0), rgs, moff(D + 1, 0), mrlz;
    std::vector<double> gw((size_t)Wn * D, 0.), gmin((size_t)Wn * D, 0.);
    for (int d = 0; d < D; d++) {
      const auto &cols = inv[members[d][0]];
      rgs.insert(rgs.end(), cols.begin(), cols.end());
      roff[d + 1] = (int32_t)rgs.size();
      mrlz.insert(mrlz.end(), members[d].begin(), members[d].end());
      moff[d + 1] = (int32_t)mrlz.size();
      for (int c = 0; c < Wn && S; c++) {
        // summed in increasing order, like the cumsum over the sorted weights
        std::vector<double> ws;
        for (int r : members[d]) ws.push_back(wh[(size_t)c * R + r]);
        std::sort(ws.begin(), ws.end());
        double tot = 0.;
        for (double x : ws) tot += x;
        gw[(size_t)c * D + d] = tot;
        gmin[(size_t)c * D + d] = ws[0];
      }
    }
    std::vector<int32_t> rank = ranks_of(gmin, D);
    std::vector<int64_t> mpos(mrlz.size());
    for (size_t i = 0; i < mrlz.size(); i++) mpos[i] = (int64_t)mrlz[i] * A.n_elems;
    const size_t nd = gw.size() + gmin.size();
    const size_t ni = rank.size() + roff.size() + moff.size() + rgs.size();
    if (reserve(h, h->stage[10], (nd + mpos.size()) * 8 + ni * 4 + 16)) return 1;
    double *dd = (double *)h->stage[10].p;
    int64_t *dl = (int64_t *)(dd + nd);
    int32_t *di = (int32_t *)(dl + mpos.size());
    std::vector<int32_t> ints;
    ints.insert(ints.end(), rank.begin(), rank.end());
    ints.insert(ints.end(), roff.begin(), roff.end());
    ints.insert(ints.end(), moff.begin(), moff.end());
    ints.insert(ints.end(), rgs.begin(), rgs.end());
    std::vector<double> dbl(gw);
    dbl.insert(dbl.end(), gmin.begin(), gmin.end());
    CK(cudaMemcpy(dd, dbl.data(), nd * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dl, mpos.data(), mpos.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(di, ints.data(), ni * 4, cudaMemcpyHostToDevice));
    B.weights = dd; B.wmin = dd + gw.size(); B.m_pos = dl;
    B.wrank = di; B.r_off = di + rank.size(); B.m_off = B.r_off + roff.size();
    B.r_gs = B.m_off + moff.size();
    B.R = D;
    // the quantiles in increasing order: one scan of the sorted groups
    std::vector<int> qord;
    for (int s = 0; s < S; s++)
      if (kinds[s] == OQG_STAT_QUANTILE) qord.push_back(s);
    std::stable_sort(qord.begin(), qord.end(),
                     [&](int a, int b) { return qs[a] < qs[b]; });
    B.nq = (int)qord.size();
    for (int i = 0; i < B.nq; i++) {
      B.qstat[i] = qord[i];
      B.qsorted[i] = qs[qord[i]];
    }
    const int RP = rp_;
    const int TPB = RP <= 32 ? 128 : 64;
    const size_t smem = (size_t)RP * TPB * 8 + (size_t)Wn * D * 20 +
                        (size_t)(D + 1) * 8 + rgs.size() * 4 + 16 +
                        (size_t)TPB * (Gt | 1) * 4;
    unsigned grid = (unsigned)((A.n_elems + TPB - 1) / TPB);
#define POST_LAUNCH(rp, tpb)                                                  \
  {                                                                           \
    CK(cudaFuncSetAttribute(k_postclassical_reg<rp, tpb>,                     \
                            cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                            (int)smem));                                      \
    k_postclassical_reg<rp, tpb><<<grid, tpb, smem, h->stream>>>(B);          \
  }
    if (RP == 4) POST_LAUNCH(4, 128)
    else if (RP == 8) POST_LAUNCH(8, 128)
    else if (RP == 16) POST_LAUNCH(16, 128)
    else if (RP == 32) POST_LAUNCH(32, 128)
    else if (RP == 64) POST_LAUNCH(64, 64)
    else if (RP == 128) POST_LAUNCH(128, 64)
    else POST_LAUNCH(256, 64)
#undef POST_LAUNCH
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  if (stream || cta) {
    // one table entry per realization
    std::vector<int32_t> roff(R + 1, 0), rgs;
    for (int r = 0; r < R; r++) {
      rgs.insert(rgs.end(), inv[r].begin(), inv[r].end());
      roff[r + 1] = (int32_t)rgs.size();
    }
    std::vector<int32_t> rank((size_t)Wn * R, 0);
    if (S && cta) rank = ranks_of(wh, R);
    size_t nb = (roff.size() + rgs.size() + rank.size()) * 4;
    if (reserve(h, h->stage[10], nb + 16)) return 1;
    int32_t *d = (int32_t *)h->stage[10].p;
    CK(cudaMemcpy(d, roff.data(), roff.size() * 4, cudaMemcpyHostToDevice));
    if (!rgs.empty())
      CK(cudaMemcpy(d + roff.size(), rgs.data(), rgs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d + roff.size() + rgs.size(), rank.data(), rank.size() * 4,
                  cudaMemcpyHostToDevice));
    B.r_off = d; B.r_gs = d + roff.size();
    B.wrank = d + roff.size() + rgs.size();
    B.weights = weights; B.R = R;
    if (!S) { B.W = 1; B.weights = (const double *)d; }  // never read
    if (cta) {
      // one CTA per (site, level): keys and scanned weights in shared memory
      int RPc = 512;
      while (RPc < R) RPc <<= 1;
      const size_t smem_c = (size_t)RPc * 16;
#define CTA_ATTR(per)                                                         \
  CK(cudaFuncSetAttribute(k_postclassical_cta<per>,                           \
                          cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                          (int)smem_c))
      if (RPc == 512) CTA_ATTR(2);
      else if (RPc == 1024) CTA_ATTR(4);
      else if (RPc == 2048) CTA_ATTR(8);
      else CTA_ATTR(16);
#undef CTA_ATTR
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(
          8, (200 * 1024) / (smem_c + 4096)));
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
      const unsigned grid_c = (unsigned)std::min<int64_t>(
          A.n_elems, (int64_t)sms * per_sm);
      if (B.rlz_out) {
        // the realization curves are written by the streaming kernel, whose
        // consecutive threads write consecutive (site, level) elements: a
        // CTA per element would scatter 4-byte writes n_elems apart
        PostRegArgs C = B;
        C.stat_out = nullptr;
        C.S = 0;
        k_postclassical_stream<<<(unsigned)((A.n_elems + 127) / 128), 128, 0,
                                 h->stream>>>(C);
        h->launches++;
        B.rlz_out = nullptr;
      }
      if (B.stat_out && B.S) {
        if (RPc == 512)
          k_postclassical_cta<2><<<grid_c, POST_CTA_TPB, smem_c, h->stream>>>(B);
        else if (RPc == 1024)
          k_postclassical_cta<4><<<grid_c, POST_CTA_TPB, smem_c, h->stream>>>(B);
        else if (RPc == 2048)
          k_postclassical_cta<8><<<grid_c, POST_CTA_TPB, smem_c, h->stream>>>(B);
        else
          k_postclassical_cta<16><<<grid_c, POST_CTA_TPB, smem_c, h->stream>>>(B);
        h->launches++;
      }
      CK(cudaGetLastError());
      return 0;
    }
    k_postclassical_stream<<<(unsigned)((A.n_elems + 127) / 128), 128, 0,
                             h->stream>>>(B);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  // generic kernel (local arrays, insertion sort): quantiles of R <= 256
  // realizations whose tables do not fit in shared memory (many
  // IMT-dependent weight columns x many groups)
  k_postclassical<<<(unsigned)((A.n_elems + 127) / 128), 128, 0, h->stream>>>(A);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_set_post_grouping(oqg_handle *h, int on) {
  h->post_grouping = on != 0;
  return 0;
}

int oqg_last_post_path(oqg_handle *h, int *groups) {
  if (groups) *groups = h->last_post_groups;
  return h->last_post_path;
}

int oqg_mean_rates(oqg_handle *h, const float *rates, int64_t N, int L, int Gt,
                   int M, const double *weights, int W, double *out) {
  CK(cudaSetDevice(h->device));
  if (N <= 0 || L <= 0) return 0;
  if (!rates || !weights || !out) return fail(h, "mean_rates: missing array");
  if (Gt < 1 || M < 1 || L % M) return fail(h, "mean_rates: bad shape");
  if (W != 1 && W != M)
    return fail(h, "mean_rates: one weight vector, or one per IMT");
  const int64_t n = N * L;
  const int rows = std::min(MR_TPB, MR_FLOATS / Gt);
  // rows of <= 32 bytes are read well enough by neighbouring threads
  // (measured on B200, Gt = 8: 4.1 TB/s direct, 3.0 TB/s staged)
  if (Gt > 8 && rows >= 32)
    k_mean_rates<<<(unsigned)((n + rows - 1) / rows), MR_TPB, 0, h->stream>>>(
        rates, n, L, L / M, Gt, rows, weights, W, out);
  else
    k_mean_rates_direct<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        rates, n, L, L / M, Gt, weights, W, out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_hazard_maps(oqg_handle *h, const float *curves, int64_t K, int M,
                    int L1, const double *imls, int P, const double *poes,
                    float *out) {
  CK(cudaSetDevice(h->device));
  if (K <= 0 || P <= 0) return 0;
  if (M < 1 || L1 < 1) return fail(h, "hazard_maps: bad shape");
  if (!curves || !imls || !poes || !out) return fail(h, "hazard_maps: missing array");
  // rows of (curve, imt) per CTA: what 64 KB of shared memory hold
  const size_t fixed = ((size_t)M * L1 + 2 * (size_t)P) * 8;
  int rows = fixed < 32 * 1024
                 ? (int)std::min<size_t>(HM_TPB, (64 * 1024 - fixed) / (4 * (size_t)(L1 | 1)))
                 : 0;
  rows -= rows % 32;
  if (rows >= 64) {
    const size_t smem = fixed + (size_t)rows * (L1 | 1) * 4;
    CK(cudaFuncSetAttribute(k_hazard_maps,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t total_rows = K * M;
    k_hazard_maps<<<(unsigned)((total_rows + rows - 1) / rows), HM_TPB, smem,
                    h->stream>>>(curves, total_rows, M, L1, rows, imls, P, poes, out);
  } else {
    int64_t tot = K * M * P;
    k_hazard_maps_direct<<<(unsigned)((tot + 127) / 128), 128, 0, h->stream>>>(
        curves, K, M, L1, imls, P, poes, out);
  }
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_rates_to_records(oqg_handle *h, const float *rates, int N, int L,
                         int G, const uint32_t *sids, const int32_t *gid,
                         int level0, const int64_t *rows, int64_t R,
                         void *out, int64_t capacity, int64_t *total) {
  CK(cudaSetDevice(h->device));
  if (total) *total = 0;
  if (N <= 0 || L <= 0 || G <= 0 || R <= 0) return 0;
  if (!rates || !gid || !total) return fail(h, "rates_to_records: missing array");
  if (R > (int64_t)N * G) return fail(h, "rates_to_records: too many rows");
  if (level0 < 0 || level0 + L > 65536)
    return fail(h, "rates_to_records: the level index does not fit in 16 bits");
  // counts [R], offsets [R + 1] (the last one is the total)
  if (reserve(h, h->stage[8], (size_t)(2 * R + 2) * 8)) return 1;
  int64_t *counts = (int64_t *)h->stage[8].p, *offs = counts + R + 1;
  CK(cudaMemsetAsync(counts + R, 0, 8, h->stream));
  const unsigned blocks = (unsigned)((R * 32 + 127) / 128);
  k_rate_rows_count<<<blocks, 128, 0, h->stream>>>(rates, N, L, G, rows, R,
                                                   counts);
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, offs, R + 1,
                                   h->stream));
  if (reserve(h, h->sg_tmp, tmp)) return 1;
  CK(cub::DeviceScan::ExclusiveSum(h->sg_tmp.p, tmp, counts, offs, R + 1,
                                   h->stream));
  h->launches += 2;
  int64_t tot = 0;
  CK(cudaMemcpyAsync(&tot, offs + R, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *total = tot;
  if (!out) return 0;                  // counting pass only
  if (tot > capacity)
    return fail(h, "rates_to_records: " + std::to_string(tot) +
                       " records, room for " + std::to_string(capacity));
  k_rate_rows_write<<<blocks, 128, 0, h->stream>>>(
      rates, N, L, G, sids, gid, level0, rows, R, offs, (RateRecord *)out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_point_ruptures(oqg_handle *h, const oqg_point_sources *src, int S,
                       int64_t U, double *planar, double *mag,
                       double *magdist, double *aux) {
  CK(cudaSetDevice(h->device));
  if (S <= 0 || U <= 0) return 0;
  if (src->msr < OQG_MSR_WC1994 || src->msr > OQG_MSR_CSCALING)
    return fail(h, "unsupported magnitude-scaling relationship (no fallback)");
  if (!src->lon || !src->lat || !src->usd || !src->lsd || !src->rar ||
      !src->mag_beg || !src->mag_cnt || !src->np_beg || !src->np_cnt ||
      !src->hd_beg || !src->hd_cnt || !src->rup_off || !src->mag ||
      !src->mag_rate || !src->np_strike || !src->np_dip || !src->np_rake ||
      !src->np_prob || !src->hd_depth || !src->hd_prob)
    return fail(h, "point sources: missing array");
  if (!planar || !mag) return fail(h, "point sources: missing output");
  if (magdist && !src->mag_dist)
    return fail(h, "point sources: magdist output without mag_dist");
  if (src->pointsource_distance > 0. && (!src->mag_dist || !aux))
    return fail(h, "point sources: pointsource_distance needs mag_dist and "
                   "the aux output");
  if (src->mode && !aux)
    return fail(h, "point sources: mode needs the aux output");
  k_point_ruptures<<<(unsigned)((U + PR_TPB - 1) / PR_TPB), PR_TPB, 0,
                     h->stream>>>(
      *src, S, U, planar, mag, magdist, aux);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_distances(oqg_handle *h, const double *planar, int U,
                  const oqg_sites *sites, int N, int mask, double *out) {
  CK(cudaSetDevice(h->device));
  if (U <= 0 || N <= 0) return 0;
  if (mask <= 0 || mask >= (1 << OQG_NOUT)) return fail(h, "bad distance mask");
  // mag is not used by the distances: pass the planar pointer as a dummy
  if (reserve(h, h->stage[11], (size_t)U * 8)) return 1;
  CK(cudaMemsetAsync(h->stage[11].p, 0, (size_t)U * 8, h->stream));
  if (prep_ruptures(h, planar, (const double *)h->stage[11].p, nullptr, U))
    return 1;
  if (prep_sites(h, sites, N)) return 1;
  dim3 grid((N + 127) / 128, std::min(U, 4096));
  k_distances<<<grid, 128, 0, h->stream>>>((const double *)h->rup.p, U,
                                           rup_stride(U),
                                           (const double *)h->site.p, N, mask,
                                           out);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"

template <int F>
static void launch_mean_stds(oqg_handle *h, int g, const oqg_ctx *ctx,
                             int64_t P, double *out) {
  CtxDev cd;
  cd.c = *ctx;
  k_mean_stds<F><<<(unsigned)((P + 127) / 128), 128, 0, h->stream>>>(
      g, h->G, h->M, cd, P, out, h->flag_dev,
      h->has_conv ? (const double *)h->conv.p : nullptr);
}

template <int F>
static void launch_mean_stds_staged(oqg_handle *h, int g, const oqg_ctx *ctx,
                                    int64_t P, double *out) {
  CtxDev cd;
  cd.c = *ctx;
  k_mean_stds_staged<F><<<(unsigned)((P + 127) / 128), 128, 0, h->stream>>>(
      g, h->G, h->M, cd, P, out);
}

extern "C" {

static const char *missing_column(const oqg_gsim &g, const oqg_ctx *c) {
  int d = required_distances(g);
  if ((d >> OQG_RRUP & 1) && !c->rrup) return "rrup";
  if ((d >> OQG_RJB & 1) && !c->rjb) return "rjb";
  if ((d >> OQG_RX & 1) && !c->rx) return "rx";
  if ((d >> OQG_RY0 & 1) && !c->ry0) return "ry0";
  if ((d >> OQG_REPI & 1) && !c->repi) return "repi";
  if ((d >> OQG_RHYPO & 1) && !c->rhypo) return "rhypo";
  if (!c->mag) return "mag";
  int f = g.family;
  const bool plain = f == OQG_TORO02 || f == OQG_C03;   // mag + one distance
  if (!plain && !c->vs30) return "vs30";
  if (f != OQG_AB06 && !plain && !c->rake) return "rake";
  if (plain && (g.flags & 1) && !c->rake) return "rake";   // SHARE variants
  if (f == OQG_CY14 || f == OQG_ASK14 || f == OQG_CB14 || f == OQG_CY08) {
    if (!c->dip) return "dip";
    if (!c->ztor) return "ztor";
  }
  if ((f == OQG_ASK14 || f == OQG_CB14) && !c->width) return "width";
  if (f == OQG_CB14 && !c->hypo_depth) return "hypo_depth";
  if (f == OQG_CB14 && !c->z2pt5) return "z2pt5";
  if ((f == OQG_CY14 || f == OQG_ASK14 || f == OQG_CY08) &&
      (!c->z1pt0 || !c->vs30measured))
    return "z1pt0/vs30measured";
  if (f == OQG_BSSA14 && (g.flags & 3) && !c->z1pt0) return "z1pt0";
  if (g.par[0] != 0. && f <= OQG_CB14 && !c->rrup) return "rrup";
  return nullptr;
}

int oqg_mean_stds(oqg_handle *h, const oqg_ctx *ctx, int64_t P, double *out) {
  CK(cudaSetDevice(h->device));
  if (!h->G) return fail(h, "oqg_set_gsims not called");
  if (P <= 0) return 0;
  for (int g = 0; g < h->G; g++)
    if (const char *m = missing_column(h->gsims[g], ctx))
      return fail(h, std::string("context column missing: ") + m);
  ConstGuard guard(h);
  if (guard.acquire()) return 1;
  CK(cudaMemsetAsync(h->flag_dev, 0, sizeof(int), h->stream));
  for (int g = 0; g < h->G; g++) {
    switch (h->gsims[g].family) {
      case OQG_BSSA14: launch_mean_stds<OQG_BSSA14>(h, g, ctx, P, out); break;
      case OQG_CY14: launch_mean_stds<OQG_CY14>(h, g, ctx, P, out); break;
      case OQG_ASK14: launch_mean_stds<OQG_ASK14>(h, g, ctx, P, out); break;
      case OQG_CB14: launch_mean_stds<OQG_CB14>(h, g, ctx, P, out); break;
      case OQG_BA08: launch_mean_stds<OQG_BA08>(h, g, ctx, P, out); break;
      case OQG_AB06: launch_mean_stds<OQG_AB06>(h, g, ctx, P, out); break;
      case OQG_AKKAR14: launch_mean_stds<OQG_AKKAR14>(h, g, ctx, P, out); break;
      case OQG_CY08: launch_mean_stds<OQG_CY08>(h, g, ctx, P, out); break;
      case OQG_TORO02: launch_mean_stds<OQG_TORO02>(h, g, ctx, P, out); break;
      case OQG_C03: launch_mean_stds<OQG_C03>(h, g, ctx, P, out); break;
      case OQG_SADIGH97: launch_mean_stds<OQG_SADIGH97>(h, g, ctx, P, out); break;
    }
    h->launches++;
    CK(cudaGetLastError());
  }
  // zero total stddev with a real truncation is an error in the reference
  // (contexts.py:1297)
  double tl = h->truncation_level;
  if (h->L1 && !(tl == 0. || tl == 1e-9 || tl == 99.)) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, h->flag_dev, sizeof(int), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (flag) return fail(h, "Total StdDev is zero");
  }
  return 0;
}

int oqg_mean_stds_staged(oqg_handle *h, const oqg_ctx *ctx, int64_t P,
                         double *out) {
  CK(cudaSetDevice(h->device));
  if (!h->G) return fail(h, "oqg_set_gsims not called");
  if (P <= 0) return 0;
  for (int g = 0; g < h->G; g++)
    if (const char *m = missing_column(h->gsims[g], ctx))
      return fail(h, std::string("context column missing: ") + m);
  ConstGuard guard(h);
  if (guard.acquire()) return 1;
  for (int g = 0; g < h->G; g++) {
    switch (h->gsims[g].family) {
      case OQG_BSSA14: launch_mean_stds_staged<OQG_BSSA14>(h, g, ctx, P, out); break;
      case OQG_CY14: launch_mean_stds_staged<OQG_CY14>(h, g, ctx, P, out); break;
      case OQG_ASK14: launch_mean_stds_staged<OQG_ASK14>(h, g, ctx, P, out); break;
      case OQG_CB14: launch_mean_stds_staged<OQG_CB14>(h, g, ctx, P, out); break;
      case OQG_BA08: launch_mean_stds_staged<OQG_BA08>(h, g, ctx, P, out); break;
      case OQG_AB06: launch_mean_stds_staged<OQG_AB06>(h, g, ctx, P, out); break;
      case OQG_AKKAR14: launch_mean_stds_staged<OQG_AKKAR14>(h, g, ctx, P, out); break;
      case OQG_CY08: launch_mean_stds_staged<OQG_CY08>(h, g, ctx, P, out); break;
      case OQG_TORO02: launch_mean_stds_staged<OQG_TORO02>(h, g, ctx, P, out); break;
      case OQG_C03: launch_mean_stds_staged<OQG_C03>(h, g, ctx, P, out); break;
      case OQG_SADIGH97: launch_mean_stds_staged<OQG_SADIGH97>(h, g, ctx, P, out); break;
    }
    h->launches++;
    CK(cudaGetLastError());
  }
  return 0;
}

int oqg_poes(oqg_handle *h, const double *ms, int64_t P, float *poes) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  if (P <= 0) return 0;
  ConstGuard guard(h);
  if (guard.acquire()) return 1;
  int64_t tot = P * h->M * h->L1 * h->G;
  double inv_norm = 1. / (h->phi_b * 2. - 1.);
  k_poes<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(
      ms, P, h->G, h->M, h->L1, h->phi_b, inv_norm, poes);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

int oqg_update(oqg_handle *h, float *map, int N, const float *poes, int64_t P,
               const int64_t *order, const int64_t *seg, const double *rate,
               const double *probs, int K, double itime, int rates_mode) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  if (P <= 0 || N <= 0) return 0;
  if (K < 0 || K > OQG_MAX_OCC) return fail(h, "bad probs_occur length");
  int LG = h->M * h->L1 * h->G;
  for (int y0 = 0; y0 < N; y0 += OQG_GRID_Y) {   // gridDim.y <= 65535
    dim3 grid((LG + 127) / 128, std::min(N - y0, OQG_GRID_Y));
    k_update<<<grid, 128, 0, h->stream>>>(map, N, LG, h->G, poes, order, seg,
                                          rate, K ? probs : nullptr, K, itime,
                                          rates_mode, nullptr, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int oqg_update_mutex(oqg_handle *h, float *map, int N, const float *poes,
                     int64_t P, const int64_t *order, const int64_t *seg,
                     const double *rate, const double *probs, int K,
                     double itime, const double *weights) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  if (P <= 0 || N <= 0) return 0;
  if (K < 0 || K > OQG_MAX_OCC) return fail(h, "bad probs_occur length");
  if (!weights) return fail(h, "update_mutex: missing weights");
  int LG = h->M * h->L1 * h->G;
  for (int y0 = 0; y0 < N; y0 += OQG_GRID_Y) {
    dim3 grid((LG + 127) / 128, std::min(N - y0, OQG_GRID_Y));
    k_update<<<grid, 128, 0, h->stream>>>(map, N, LG, h->G, poes, order, seg,
                                          rate, K ? probs : nullptr, K, itime,
                                          0, weights, y0);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"

template <int F, bool NB, bool FAST, int SUBS, int REP>
static cudaError_t launch_fused2(const FusedArgs &A, dim3 grid, int threads,
                                 size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(
      k_fused<F, NB, FAST, SUBS, REP>,
      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_fused<F, NB, FAST, SUBS, REP><<<grid, threads, smem, st>>>(A);
  return cudaGetLastError();
}

#ifndef OQG_SUBS_MAX
#define OQG_SUBS_MAX 4
#endif
#define OQG_SMEM_SM (227 * 1024 - 4096)   // per SM, minus the static tables

// Launch of one model.  Fast path with a 256-node table: `subs` sub-CTAs per
// CTA around one 8-fold replicated PoE table (conflict-free LDS.128), as
// many as fit into the shared memory of an SM; otherwise the one-work-item
// CTAs with the plain table (4 per SM).  cfg[0] forces subs (0 = automatic),
// cfg[1] != 0 forces the plain table.
// sub-CTAs per CTA of a model (1 = the one-work-item CTAs)
template <int F>
static int fused_subs(const FusedArgs &A, const int *cfg) {
  const bool fast = A.ptab != nullptr && A.K == 0;
  if (!(fast && !cfg[1] && A.pt_nodes == 256)) return 1;
  const size_t tab1 = (size_t)A.pt_nodes * PT_STRIDE * 4;
  const size_t sub = FusedLayout<F, false, true>::SIZE * 8;
  int subs = (int)((OQG_SMEM_SM - 8 * tab1) / sub);
  subs = std::min(subs, OQG_SUBS_MAX);
  if (cfg[0]) subs = std::min(subs, cfg[0]);
  if (subs * A.nw > 32) subs = 32 / A.nw;
  return std::max(subs, 1);
}

static int fused_subs_of(int family, const FusedArgs &A, const int *cfg) {
  switch (family) {
    case OQG_BSSA14: return fused_subs<OQG_BSSA14>(A, cfg);
    case OQG_CY14: return fused_subs<OQG_CY14>(A, cfg);
    case OQG_ASK14: return fused_subs<OQG_ASK14>(A, cfg);
    case OQG_CB14: return fused_subs<OQG_CB14>(A, cfg);
    case OQG_BA08: return fused_subs<OQG_BA08>(A, cfg);
    case OQG_AB06: return fused_subs<OQG_AB06>(A, cfg);
    case OQG_AKKAR14: return fused_subs<OQG_AKKAR14>(A, cfg);
    case OQG_CY08: return fused_subs<OQG_CY08>(A, cfg);
    case OQG_TORO02: return fused_subs<OQG_TORO02>(A, cfg);
    case OQG_C03: return fused_subs<OQG_C03>(A, cfg);
    case OQG_SADIGH97: return fused_subs<OQG_SADIGH97>(A, cfg);
  }
  return 1;
}

template <int F>
static cudaError_t launch_fused(FusedArgs &A, int tiles, int rc, int ugroups,
                                const int *cfg, cudaStream_t st) {
  const bool fast = A.ptab != nullptr && A.K == 0;
  A.ntiles = tiles;
  const size_t tab1 = fast ? (size_t)A.pt_nodes * PT_STRIDE * 4 : 0;
  int subs = fused_subs<F>(A, cfg);
  if (subs > 1) {
    const size_t sub = FusedLayout<F, false, true>::SIZE * 8;
    {
      A.subs = subs;
      A.pt_bias = 0u - 0x4b000000u * (unsigned)(PT_STRIDE * 4 * 8);
      dim3 grid((tiles + subs - 1) / subs, rc, ugroups);
      return launch_fused2<F, false, true, OQG_SUBS_MAX, 8>(
          A, grid, 32 * A.nw * subs, 8 * tab1 + subs * sub, st);
    }
  }
  A.subs = 1;
  A.pt_bias = 0u - 0x4b000000u * (unsigned)(PT_STRIDE * 4);
  dim3 grid(tiles, rc, ugroups);
  const int threads = 32 * A.nw;
  if (A.K)
    return launch_fused2<F, true, false, 1, 1>(
        A, grid, threads, FusedLayout<F, true, false>::SIZE * 8, st);
  if (fast)
    return launch_fused2<F, false, true, 1, 1>(
        A, grid, threads, FusedLayout<F, false, true>::SIZE * 8 + tab1, st);
  return launch_fused2<F, false, false, 1, 1>(
      A, grid, threads, FusedLayout<F, false, false>::SIZE * 8, st);
}

static int hazard_fused(oqg_handle *h, const double *planar,
                        const double *mag, const double *magdist,
                        const double *xpar, const double *const *xplanes,
                        const double *aux, const double *probs, int K, int U,
                        const oqg_sites *sites, const uint32_t *sidx, int N,
                        double mindist, double itime, int rates_mode,
                        float *map, unsigned long long *counters);

extern "C" {

int oqg_hazard_planar(oqg_handle *h, const double *planar, const double *mag,
                      const double *magdist, const double *probs, int K,
                      int U, const oqg_sites *sites, const uint32_t *sidx,
                      int N, double mindist, double itime, int rates_mode,
                      float *map, unsigned long long *counters) {
  return oqg_hazard_planar_aux(h, planar, mag, magdist, nullptr, probs, K, U,
                               sites, sidx, N, mindist, itime, rates_mode, map,
                               counters);
}

int oqg_fused_planar(oqg_handle *h, const double *planar, const double *mag,
                     int R, const oqg_sites *sites, int N,
                     const double *magdist, double mindist, double itime,
                     int rates_mode, float *map,
                     unsigned long long *counters) {
  return oqg_hazard_planar_aux(h, planar, mag, magdist, nullptr, nullptr, 0, R,
                               sites, nullptr, N, mindist, itime, rates_mode,
                               map, counters);
}

int oqg_hazard_planar_aux(oqg_handle *h, const double *planar,
                          const double *mag, const double *magdist,
                          const double *aux, const double *probs, int K, int U,
                          const oqg_sites *sites, const uint32_t *sidx, int N,
                          double mindist, double itime, int rates_mode,
                          float *map, unsigned long long *counters) {
  return hazard_fused(h, planar, mag, magdist, nullptr, nullptr, aux, probs, K,
                      U, sites, sidx, N, mindist, itime, rates_mode, map,
                      counters);
}

int oqg_hazard_dist(oqg_handle *h, const double *rup_params,
                    const double *const *planes, const double *probs, int K,
                    int U, const oqg_sites *sites, const uint32_t *sidx, int N,
                    double mindist, double itime, int rates_mode, float *map,
                    unsigned long long *counters) {
  if (!rup_params || !planes || !planes[OQG_RRUP])
    return fail(h, "hazard_dist: rupture parameters and the rrup plane are "
                   "required");
  return hazard_fused(h, nullptr, nullptr, nullptr, rup_params, planes,
                      nullptr, probs, K, U, sites, sidx, N, mindist, itime,
                      rates_mode, map, counters);
}

}  // extern "C"

// The fused pass over planar ruptures (planar, mag, magdist) or over ruptures
// with external distance planes (xpar, xplanes)
static int hazard_fused(oqg_handle *h, const double *planar,
                        const double *mag, const double *magdist,
                        const double *xpar, const double *const *xplanes,
                        const double *aux, const double *probs, int K, int U,
                        const oqg_sites *sites, const uint32_t *sidx, int N,
                        double mindist, double itime, int rates_mode,
                        float *map, unsigned long long *counters) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  if (U <= 0 || N <= 0) return 0;
  if (K < 0 || K > OQG_MAX_OCC) return fail(h, "bad probs_occur length");
  if (K && !probs) return fail(h, "probs_occur missing");
  if (h->has_conv)
    return fail(h, "the fused pass has no horizontal-component conversion: "
                   "use the context operators (oqg_mean_stds, oqg_poes, "
                   "oqg_update)");
  const int G = h->G, M = h->M, L1 = h->L1, L = M * L1;
  ConstGuard guard(h);
  if (guard.acquire()) return 1;
  if (xpar) {
    const int64_t Us = rup_stride(U);
    if (reserve(h, h->rup, (size_t)NF_RUP * Us * 8)) return 1;
    CK(cudaMemsetAsync(h->rup.p, 0, (size_t)NF_RUP * Us * 8, h->stream));
    k_prep_xruptures<<<(unsigned)((U + 127) / 128), 128, 0, h->stream>>>(
        xpar, U, Us, (double *)h->rup.p);
    h->launches++;
    CK(cudaGetLastError());
  } else if (prep_ruptures(h, planar, mag, magdist, U)) {
    return 1;
  }
  // planar ruptures: sites grouped by cell and vs30 (the external distance
  // planes of the other ruptures are in the caller's site order)
  oqg_sites grouped;
  if (!xpar && h->site_grouping && N >= 256) {
    if (prep_sites_grouped(h, sites, sidx, N, &grouped, &sidx)) return 1;
    sites = &grouped;
  } else if (prep_sites(h, sites, N)) {
    return 1;
  }

  FusedArgs A;
  memset(&A, 0, sizeof(A));
  if (xpar)
    for (int k = 0; k < 6; k++) A.xd[k] = xplanes[k];
  A.rup = (const double *)h->rup.p;
  A.probs = K ? probs : nullptr;
  A.aux = aux;
  A.site = (const double *)h->site.p;
  A.vs30 = sites->vs30; A.z1pt0 = sites->z1pt0; A.z2pt5 = sites->z2pt5;
  A.vs30measured = sites->vs30measured;
  A.counters = counters;
  A.U = U; A.Us = rup_stride(U); A.N = N; A.K = K; A.G = G; A.M = M; A.L1 = L1;
  A.nlc = (L1 + NL - 1) / NL;
  A.lpc = (L1 + A.nlc - 1) / A.nlc;
  A.nunits = M * A.nlc;
  A.ptab = (h->fast_ok && !h->exact_poe) ? (const float *)h->ptab.p : nullptr;
  A.tl = h->truncation_level; A.xscale = h->xscale;
  A.pt_nodes = h->pt_nodes;
  A.sorted = h->sorted;
  A.mindist = mindist; A.itime = itime; A.phi_b = h->phi_b;
  A.inv_norm = 1. / (h->phi_b * 2. - 1.);

  // CTA shape: nw warps (units) per CTA, unit groups on blockIdx.z
  int ugroups = (A.nunits + NWMAX - 1) / NWMAX;
  int nw = (A.nunits + ugroups - 1) / ugroups;
  int tiles = (N + 31) / 32;
  // rupture chunks: enough CTAs for ~8 waves of 2 CTAs per SM, rounded to
  // whole shared-memory tiles, so that the tail wave is small
  int64_t ctas_wanted = (int64_t)h->sm_count * MINB * 8;
  int64_t rc = (ctas_wanted + (int64_t)tiles * ugroups - 1) /
               ((int64_t)tiles * ugroups);
  int64_t max_rc = (U + RT - 1) / RT;
  rc = std::max<int64_t>(1, std::min<int64_t>(rc, std::min<int64_t>(max_rc, 64)));
  if (h->wave_fit) {
    // Wave quantisation: the CTAs of a launch take about the same time (same
    // ruptures, 32 sites each), so a grid of 8.5 waves costs 9.  Among the
    // chunk counts around the first guess pick the one with the smallest
    // waves x tiles-per-chunk summed over the models (cfg3: 16 chunks = 1264
    // CTAs = 8.54 waves of 148, 303.9 ms -> 28 chunks = 2212 CTAs = 14.95
    // waves, 290.6 ms).
    A.nw = nw;
    const int64_t lo = std::max<int64_t>(1, rc / 2);
    const int64_t hi = std::min<int64_t>(std::min<int64_t>(max_rc, 64), rc * 2);
    int64_t best_rc = rc;
    double best_cost = 1e300;
    for (int64_t c = lo; c <= hi; c++) {
      const int64_t tiles_per_chunk = ((U + c - 1) / c + RT - 1) / RT;
      double cost = 0.;
      for (int g = 0; g < G; g++) {
        const int subs = fused_subs_of(h->gsims[g].family, A, h->fused_cfg);
        const int64_t ctas = (int64_t)((tiles + subs - 1) / subs) * c * ugroups;
        const int64_t slots = (int64_t)h->sm_count * (subs > 1 ? 1 : MINB);
        // + half a wave: the CTAs do not finish in lockstep, the tail of a
        // launch costs about half a CTA time whatever the wave count
        // (measured on cfg3: 13 chunks 293.8 ms, 28: 290.6, 43: 288.8)
        cost += ((double)((ctas + slots - 1) / slots) + 0.5) *
                (double)tiles_per_chunk;
      }
      // prefer the first guess, then fewer chunks, on (near) ties
      if (cost < best_cost * (1. - 1e-3)) { best_cost = cost; best_rc = c; }
    }
    rc = best_rc;
  }
  if (h->force_chunks > 0)     // OQG_CHUNKS: experiments
    rc = std::max<int64_t>(1, std::min<int64_t>(h->force_chunks, max_rc));
  int64_t rpc = (U + rc - 1) / rc;
  rpc = (rpc + RT - 1) / RT * RT;
  rc = (U + rpc - 1) / rpc;
  A.rup_per_chunk = (int)rpc;
  if (reserve(h, h->part, (size_t)rc * G * N * L * 8)) return 1;
  A.part = (double *)h->part.p;

  for (int g = 0; g < G; g++) {
    const oqg_gsim &gs = h->gsims[g];
    if (gs.family == OQG_CB14 && !sites->z2pt5) return fail(h, "sites need z2pt5");
    if (!sites->vs30 && gs.family != OQG_TORO02 && gs.family != OQG_C03)
      return fail(h, "sites need vs30");
    if ((gs.family == OQG_CY14 || gs.family == OQG_ASK14 ||
         gs.family == OQG_CY08) &&
        (!sites->z1pt0 || !sites->vs30measured))
      return fail(h, "sites need z1pt0 and vs30measured");
    A.g = g;
    A.clampmask = required_distances(gs);
    if (xpar)
      for (int k = 0; k < 6; k++)
        if (((A.clampmask >> k) & 1) && !xplanes[k])
          return fail(h, std::string("hazard_dist: the model needs the "
                                     "distance plane ") + std::to_string(k));
    A.dmask = A.clampmask | (1 << OQG_RRUP);
    if (A.dmask & (1 << OQG_RHYPO)) A.dmask |= 1 << OQG_REPI;
    // pairs: model 0, blockIdx.z == 0 only; executed evaluations (statistics
    // pass, oqg_set_count_evals): every model and unit group
    A.count_pairs = !counters ? 0 : (h->count_evals ? 2 : (g == 0 ? 1 : 0));
    // only blockIdx.z == 0 counts pairs
    cudaError_t e = cudaSuccess;
    A.nw = nw;
    CK(cudaEventRecord(h->ev[2 * g], h->stream));
#define LF(FAM) e = launch_fused<FAM>(A, tiles, (int)rc, ugroups, h->fused_cfg, h->stream)
    switch (gs.family) {
      case OQG_BSSA14: LF(OQG_BSSA14); break;
      case OQG_CY14: LF(OQG_CY14); break;
      case OQG_ASK14: LF(OQG_ASK14); break;
      case OQG_CB14: LF(OQG_CB14); break;
      case OQG_BA08: LF(OQG_BA08); break;
      case OQG_AB06: LF(OQG_AB06); break;
      case OQG_AKKAR14: LF(OQG_AKKAR14); break;
      case OQG_CY08: LF(OQG_CY08); break;
      case OQG_TORO02: LF(OQG_TORO02); break;
      case OQG_C03: LF(OQG_C03); break;
      case OQG_SADIGH97: LF(OQG_SADIGH97); break;
    }
#undef LF
    h->last_subs = A.subs;
    CK(cudaEventRecord(h->ev[2 * g + 1], h->stream));
    h->launches++;
    if (e != cudaSuccess)
      return fail(h, std::string("fused launch: ") + cudaGetErrorString(e));
  }
  h->n_fused_ev = G;
  {
    dim3 fgrid((N + 31) / 32, (L * G + 31) / 32), fblock(32, 8);
    const int part_f32 = A.ptab != nullptr && A.K == 0;   // the fast path
    k_finalize<<<fgrid, fblock, 0, h->stream>>>(A.part, part_f32, (int)rc, G,
                                                N, L, L1, sidx, rates_mode,
                                                map);
  }
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

static int h2d(oqg_handle *h, Buf &b, const void *src, size_t bytes,
               const void **dev) {
  if (!src) { *dev = nullptr; return 0; }
  if (reserve(h, b, bytes)) return 1;
  CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  *dev = b.p;
  return 0;
}

int oqg_hazard_planar_host(oqg_handle *h, const double *planar,
                           const double *mag, const double *magdist,
                           const double *probs, int K, int U,
                           const oqg_sites *s, int N, double mindist,
                           double itime, int rates_mode, float *map,
                           unsigned long long *pairs_kept) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  const void *d[11];
  oqg_sites sd;
  if (h2d(h, h->stage[0], planar, (size_t)U * OQG_PLANAR_LEN * 8, &d[0])) return 1;
  if (h2d(h, h->stage[1], mag, (size_t)U * 8, &d[1])) return 1;
  if (h2d(h, h->stage[2], magdist, (size_t)U * 8, &d[2])) return 1;
  if (h2d(h, h->stage[3], K ? probs : nullptr, (size_t)U * K * 8, &d[3])) return 1;
  if (h2d(h, h->stage[4], s->lon, (size_t)N * 8, &d[4])) return 1;
  if (h2d(h, h->stage[5], s->lat, (size_t)N * 8, &d[5])) return 1;
  if (h2d(h, h->stage[6], s->depth, (size_t)N * 8, &d[6])) return 1;
  if (h2d(h, h->stage[7], s->vs30, (size_t)N * 8, &d[7])) return 1;
  if (h2d(h, h->stage[8], s->z1pt0, (size_t)N * 8, &d[8])) return 1;
  if (h2d(h, h->stage[9], s->z2pt5, (size_t)N * 8, &d[9])) return 1;
  if (h2d(h, h->stage[10], s->vs30measured, (size_t)N, &d[10])) return 1;
  sd.lon = (const double *)d[4]; sd.lat = (const double *)d[5];
  sd.depth = (const double *)d[6]; sd.vs30 = (const double *)d[7];
  sd.z1pt0 = (const double *)d[8]; sd.z2pt5 = (const double *)d[9];
  sd.vs30measured = (const uint8_t *)d[10];
  size_t mbytes = (size_t)N * h->M * h->L1 * h->G * 4;
  // the map and the pair counter share one staging buffer
  if (reserve(h, h->stage[11], mbytes + 16)) return 1;
  float *map_dev = (float *)h->stage[11].p;
  unsigned long long *cnt =
      (unsigned long long *)((char *)h->stage[11].p + (mbytes + 7) / 8 * 8);
  CK(cudaMemcpyAsync(map_dev, map, mbytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(cnt, 0, 8, h->stream));
  if (oqg_hazard_planar(h, (const double *)d[0], (const double *)d[1],
                        (const double *)d[2], (const double *)d[3], K, U, &sd,
                        nullptr, N, mindist, itime, rates_mode, map_dev,
                        pairs_kept ? cnt : nullptr))
    return 1;
  CK(cudaMemcpyAsync(map, map_dev, mbytes, cudaMemcpyDeviceToHost, h->stream));
  if (pairs_kept)
    CK(cudaMemcpyAsync(pairs_kept, cnt, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

__global__ void k_fill_f32(float *p, size_t n, float v) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- a stream of sources from host memory: begin / add / end ----------------
// oqg_hazard_planar_host pays, per call, the upload of the sites and of the
// map, the download of the map and a synchronisation.  For a loop over many
// small sources (the reference's `for src in sources: cmaker.update(pmap,
// ...)`, contexts.py:1548-1579) the session keeps sites and map on the device
// and double-buffers the ruptures: the copy of source k + 1 (copy stream)
// overlaps the kernels of source k (compute stream); `add` returns as soon
// as ITS copy is done -- the caller may then reuse its buffers -- not when
// the kernels are.
int oqg_host_begin(oqg_handle *h, const oqg_sites *s, int N, int rates_mode,
                   const float *map) {
  CK(cudaSetDevice(h->device));
  if (!h->G || !h->L1) return fail(h, "set_gsims / set_levels not called");
  if (h->hs_open) return fail(h, "host_begin: a session is already open");
  if (!s || N <= 0) return fail(h, "host_begin: no sites");
  if (!h->hs_copy) {
    CK(cudaStreamCreateWithFlags(&h->hs_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CK(cudaEventCreateWithFlags(&h->hs_copied[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->hs_used[i], cudaEventDisableTiming));
    }
  }
  const void *d[7];
  if (h2d(h, h->hs_site[0], s->lon, (size_t)N * 8, &d[0])) return 1;
  if (h2d(h, h->hs_site[1], s->lat, (size_t)N * 8, &d[1])) return 1;
  if (h2d(h, h->hs_site[2], s->depth, (size_t)N * 8, &d[2])) return 1;
  if (h2d(h, h->hs_site[3], s->vs30, (size_t)N * 8, &d[3])) return 1;
  if (h2d(h, h->hs_site[4], s->z1pt0, (size_t)N * 8, &d[4])) return 1;
  if (h2d(h, h->hs_site[5], s->z2pt5, (size_t)N * 8, &d[5])) return 1;
  if (h2d(h, h->hs_site[6], s->vs30measured, (size_t)N, &d[6])) return 1;
  h->hs_sd.lon = (const double *)d[0]; h->hs_sd.lat = (const double *)d[1];
  h->hs_sd.depth = (const double *)d[2]; h->hs_sd.vs30 = (const double *)d[3];
  h->hs_sd.z1pt0 = (const double *)d[4]; h->hs_sd.z2pt5 = (const double *)d[5];
  h->hs_sd.vs30measured = (const uint8_t *)d[6];
  const size_t n = (size_t)N * h->M * h->L1 * h->G, mbytes = n * 4;
  if (reserve(h, h->hs_map, mbytes + 16)) return 1;
  float *map_dev = (float *)h->hs_map.p;
  if (map) {
    CK(cudaMemcpyAsync(map_dev, map, mbytes, cudaMemcpyHostToDevice, h->stream));
  } else if (rates_mode) {
    CK(cudaMemsetAsync(map_dev, 0, mbytes, h->stream));
  } else {
    k_fill_f32<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(map_dev, n, 1.f);
    h->launches++;
  }
  CK(cudaMemsetAsync((char *)h->hs_map.p + (mbytes + 7) / 8 * 8, 0, 8, h->stream));
  h->hs_N = N; h->hs_rates = rates_mode; h->hs_n = 0; h->hs_open = true;
  h->hs_rec[0] = h->hs_rec[1] = false;
  return 0;
}

int oqg_host_add(oqg_handle *h, const double *planar, const double *mag,
                 const double *magdist, const double *probs, int K, int U,
                 double mindist, double itime) {
  CK(cudaSetDevice(h->device));
  if (!h->hs_open) return fail(h, "host_add: no open session");
  if (U <= 0) return 0;
  if (!planar || !mag || !magdist) return fail(h, "host_add: missing array");
  if (K && !probs) return fail(h, "probs_occur missing");
  const int slot = h->hs_n++ & 1;
  Buf *b = h->hs_slot[slot];
  // the kernels that read this slot two calls ago must be done
  if (h->hs_rec[slot]) CK(cudaStreamWaitEvent(h->hs_copy, h->hs_used[slot], 0));
  const void *src[4] = {planar, mag, magdist, K ? probs : nullptr};
  const size_t nb[4] = {(size_t)U * OQG_PLANAR_LEN * 8, (size_t)U * 8,
                        (size_t)U * 8, (size_t)U * K * 8};
  for (int i = 0; i < 4; i++) {
    if (!src[i]) continue;
    if (reserve(h, b[i], nb[i])) return 1;
    CK(cudaMemcpyAsync(b[i].p, src[i], nb[i], cudaMemcpyHostToDevice, h->hs_copy));
  }
  CK(cudaEventRecord(h->hs_copied[slot], h->hs_copy));
  CK(cudaStreamWaitEvent(h->stream, h->hs_copied[slot], 0));
  const size_t mbytes = (size_t)h->hs_N * h->M * h->L1 * h->G * 4;
  unsigned long long *cnt =
      (unsigned long long *)((char *)h->hs_map.p + (mbytes + 7) / 8 * 8);
  if (hazard_fused(h, (const double *)b[0].p, (const double *)b[1].p,
                   (const double *)b[2].p, nullptr, nullptr, nullptr,
                   K ? (const double *)b[3].p : nullptr, K, U, &h->hs_sd,
                   nullptr, h->hs_N, mindist, itime, h->hs_rates,
                   (float *)h->hs_map.p, cnt))
    return 1;
  CK(cudaEventRecord(h->hs_used[slot], h->stream));
  h->hs_rec[slot] = true;
  // the caller's buffers are free again once the copy is done
  CK(cudaEventSynchronize(h->hs_copied[slot]));
  return 0;
}

int oqg_host_end(oqg_handle *h, float *map, unsigned long long *pairs_kept) {
  CK(cudaSetDevice(h->device));
  if (!h->hs_open) return fail(h, "host_end: no open session");
  h->hs_open = false;
  const size_t mbytes = (size_t)h->hs_N * h->M * h->L1 * h->G * 4;
  if (map)
    CK(cudaMemcpyAsync(map, h->hs_map.p, mbytes, cudaMemcpyDeviceToHost, h->stream));
  if (pairs_kept)
    CK(cudaMemcpyAsync(pairs_kept, (char *)h->hs_map.p + (mbytes + 7) / 8 * 8, 8,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

unsigned long long oqg_launch_count(oqg_handle *h) { return h->launches; }

int oqg_set_exact_poe(oqg_handle *h, int on) {
  h->exact_poe = on != 0;
  return 0;
}

int oqg_set_fused_config(oqg_handle *h, int subs, int plain_table) {
  if (subs < 0 || subs > OQG_SUBS_MAX) return fail(h, "subs must be in 0..4");
  h->fused_cfg[0] = subs;
  h->fused_cfg[1] = plain_table != 0;
  return 0;
}

int oqg_last_fused_subs(oqg_handle *h) { return h->last_subs; }

int oqg_set_site_grouping(oqg_handle *h, int on, double cell_deg) {
  if (on && !(cell_deg > 0.)) cell_deg = 0.16;
  h->site_grouping = on != 0;
  if (on) h->site_cell_deg = cell_deg;
  return 0;
}

int oqg_set_count_evals(oqg_handle *h, int on) {
  h->count_evals = on != 0;
  return 0;
}

int oqg_poe_table_info(oqg_handle *h, int *in_use, double *max_rel_err) {
  *in_use = h->fast_ok && !h->exact_poe;
  *max_rel_err = h->ptab_err;
  return 0;
}

int oqg_last_fused_ms(oqg_handle *h, float *ms, int cap) {
  if (cudaSetDevice(h->device) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
  int n = std::min(cap, h->n_fused_ev);
  for (int g = 0; g < n; g++)
    if (cudaEventElapsedTime(&ms[g], h->ev[2 * g], h->ev[2 * g + 1]) !=
        cudaSuccess)
      return -1;
  return n;
}

int oqg_measure_fp64_peak(oqg_handle *h, double *tflops) {
  CK(cudaSetDevice(h->device));
  double *out = nullptr;
  CK(cudaMalloc((void **)&out, 8));
  const int iters = 4096, threads = 256, blocks = h->sm_count * 8;
  cudaEvent_t a = h->ev[2 * OQG_MAX_G], b = h->ev[2 * OQG_MAX_G + 1];
  k_dfma<<<blocks, threads, 0, h->stream>>>(out, 64);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(a, h->stream));
    k_dfma<<<blocks, threads, 0, h->stream>>>(out, iters);
    CK(cudaEventRecord(b, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = std::min(best, ms);
  }
  h->launches += 6;
  CK(cudaFree(out));
  double flops = 2.0 * 8 * iters * (double)threads * blocks;
  *tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
