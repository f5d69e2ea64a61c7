// Ground-motion models of the classical-PSHA hot path as device functions.
//
// One function per model family evaluates, for ONE context row and ONE
// intensity measure type, what GMPE.compute fills into mean/sig/tau/phi
// (reference: openquake/hazardlib/gsim/base.py:389-401).  Everything is
// float64 as in the reference.  The part of a model that does not depend on
// the IMT (rock PGA and the like) is computed once per row by <fam>_row() and
// passed to <fam>_imt() for each IMT.
//
//   BSSA14   gsim/boore_2014.py:61-287, 335-359
//   CY14     gsim/chiou_youngs_2014.py:54-99, 175-213, 252-353, 436-635, 767-805
//   ASK14    gsim/abrahamson_2014.py:61-438, 517-560
//   CB14     gsim/campbell_bozorgnia_2014.py:54-463, 520-567
//   BA08     gsim/boore_atkinson_2008.py:100-151 + atkinson_boore_2006.py:63-195
//   AB06     gsim/atkinson_boore_2006.py:109-124, 213-221, 421-448
//   AKKAR14  gsim/akkar_2014.py:33-126, 177-198
//   CY08     gsim/chiou_youngs_2008.py:29-167, 204-228
//   TORO02   gsim/toro_2002.py:31-57, 103-128
//   C03      gsim/campbell_2003.py:39-47, 71-106, 178-189
//   SADIGH97 gsim/sadigh_1997.py:47-108, 134-161
//   epistemic sigma: gsim/abrahamson_2014.py:440-464
//
// Coefficient rows are arrays of OQG_KMAX doubles laid out by the *_FIELDS
// lists below; the host packs them by name (oqg_family_fields).
#pragma once
#include <math.h>
#include "../../include/oqg.h"

namespace oqg {

// one context row (the columns a model may read, contexts.py:660-678)
struct Row {
  double mag, rake, dip, ztor, width, hypo_depth;
  double rrup, rjb, rx, ry0, repi, rhypo;
  double vs30, z1pt0, z2pt5;
  bool vs30measured;
};

struct Out {
  double mean, sig, tau, phi;
};

#define OQG_DECL(n) double n;
#define OQG_NAME(n) #n ","

// ---------------------------------------------------------------- fields --
#define BSSA14_FIELDS(X)                                                     \
  X(e0) X(e1) X(e2) X(e3) X(e4) X(e5) X(e6) X(Mh) X(c1) X(c2) X(c3) X(h)     \
  X(Dc3) X(c) X(Vc) X(f4) X(f5) X(f6) X(f7) X(R1) X(R2) X(DfR) X(DfV) X(f1)  \
  X(f2) X(tau1) X(tau2) X(period)
#define CY14_FIELDS(X)                                                       \
  X(c1) X(c1a) X(c1b) X(c1c) X(c1d) X(cn) X(cm) X(c3) X(c5) X(chm) X(c6)     \
  X(c7) X(c7b) X(c9) X(c9a) X(c9b) X(c11b) X(cg1) X(cg2) X(cg3) X(phi1)      \
  X(phi2) X(phi3) X(phi4) X(phi5) X(gjpit) X(gwn) X(phi1jp) X(phi5jp)        \
  X(tau1) X(tau2) X(sig1) X(sig2) X(sig3) X(period)
#define ASK14_FIELDS(X)                                                      \
  X(m1) X(vlin) X(b) X(c) X(c4) X(a1) X(a2) X(a3) X(a4) X(a5) X(a6) X(a7)    \
  X(a8) X(a10) X(a11) X(a12) X(a13) X(a15) X(a17) X(a43) X(a44) X(a45)       \
  X(a46) X(a25) X(a28) X(a29) X(a31) X(a36) X(a37) X(a38) X(a39) X(a40)      \
  X(a41) X(a42) X(s1e) X(s2e) X(s3) X(s4) X(s1m) X(s2m) X(s5) X(s6) X(v1)    \
  X(period)
#define CB14_FIELDS(X)                                                       \
  X(c0) X(c1) X(c2) X(c3) X(c4) X(c5) X(c6) X(c7) X(c8) X(c9) X(c10) X(c11)  \
  X(c12) X(c13) X(c14) X(c15) X(c16) X(c17) X(c18) X(c19) X(c20) X(Dc20)     \
  X(a2) X(h1) X(h2) X(h3) X(h5) X(h6) X(k1) X(k2) X(k3) X(phi1) X(phi2)      \
  X(tau1) X(tau2) X(philnAF) X(rholny) X(period)
#define BA08_FIELDS(X)                                                       \
  X(c1) X(c2) X(c3) X(h) X(e1) X(e2) X(e3) X(e4) X(e5) X(e6) X(e7) X(Mh)     \
  X(sigma) X(tau) X(std) X(so_blin) X(so_b1) X(so_b2) X(is_pga) X(period)
#define AB06_FIELDS(X)                                                       \
  X(bc_c1) X(bc_c2) X(bc_c3) X(bc_c4) X(bc_c5) X(bc_c6) X(bc_c7) X(bc_c8)    \
  X(bc_c9) X(bc_c10) X(hr_c1) X(hr_c2) X(hr_c3) X(hr_c4) X(hr_c5) X(hr_c6)   \
  X(hr_c7) X(hr_c8) X(hr_c9) X(hr_c10) X(st_delta) X(st_M1) X(st_Mh)         \
  X(so_blin) X(so_b1) X(so_b2) X(is_pgv) X(period)
#define AKKAR14_FIELDS(X)                                                    \
  X(a1) X(a2) X(a3) X(a4) X(a5) X(a6) X(a7) X(a8) X(a9) X(c) X(n) X(b1)      \
  X(b2) X(sigma) X(tau) X(period)

#define CY08_FIELDS(X)                                                       \
  X(c2) X(c3) X(c4) X(c4a) X(crb) X(chm) X(cg3) X(c1) X(c1a) X(c1b) X(cn)    \
  X(cm) X(c5) X(c6) X(c7) X(c7a) X(c9) X(c9a) X(c10) X(cg1) X(cg2) X(phi1)   \
  X(phi2) X(phi3) X(phi4) X(phi5) X(phi6) X(phi7) X(phi8) X(tau1) X(tau2)    \
  X(sig1) X(sig2) X(sig3) X(sig4) X(period)
#define TORO02_FIELDS(X)                                                     \
  X(c1) X(c2) X(c3) X(c4) X(c5) X(c6) X(c7) X(m50) X(m55) X(m80) X(r5)       \
  X(r20) X(period) X(adj_ss) X(adj_rv) X(adj_nm)
#define C03_FIELDS(X)                                                        \
  X(c1) X(c2) X(c3) X(c4) X(c5) X(c6) X(c7) X(c8) X(c9) X(c10) X(c11)        \
  X(c12) X(c13) X(period) X(fs_ss) X(fs_rv) X(fs_nm) X(afrock)

#define SADIGH97_FIELDS(X)                                                   \
  X(lo_c1) X(lo_c2) X(lo_c3) X(lo_c4) X(lo_c5) X(lo_c6) X(lo_c7) X(hi_c1)    \
  X(hi_c2) X(hi_c3) X(hi_c4) X(hi_c5) X(hi_c6) X(hi_c7) X(sd_sigma0)         \
  X(sd_magfactor) X(sd_maxsigma) X(sd_maxmag) X(so_c6r) X(so_c6ss) X(so_c7)  \
  X(so_sigma0) X(so_magfactor) X(period)

struct Bssa14C { BSSA14_FIELDS(OQG_DECL) };
struct Cy14C { CY14_FIELDS(OQG_DECL) };
struct Ask14C { ASK14_FIELDS(OQG_DECL) };
struct Cb14C { CB14_FIELDS(OQG_DECL) };
struct Ba08C { BA08_FIELDS(OQG_DECL) };
struct Ab06C { AB06_FIELDS(OQG_DECL) };
struct Akkar14C { AKKAR14_FIELDS(OQG_DECL) };
struct Cy08C { CY08_FIELDS(OQG_DECL) };
struct Toro02C { TORO02_FIELDS(OQG_DECL) };
struct C03C { C03_FIELDS(OQG_DECL) };
struct Sadigh97C { SADIGH97_FIELDS(OQG_DECL) };

// the last OQG_KMAX - OQG_DER doubles of a row hold constants derived on the
// host from the packed coefficients (oqg_set_gsims; read by oqg_fast.cuh)
#define OQG_DER 48
static_assert(sizeof(Ask14C) <= OQG_DER * 8 && sizeof(Cb14C) <= OQG_DER * 8 &&
              sizeof(Cy14C) <= OQG_DER * 8, "derived slots overlap the row");
static_assert(sizeof(Bssa14C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Cy14C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Ask14C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Cb14C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Ba08C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Ab06C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Akkar14C) <= OQG_KMAX * 8, "row too long");
static_assert(sizeof(Cy08C) <= OQG_KMAX * 8, "row too long");

#define OQG_D2R 0.017453292519943295  // pi / 180, numpy.radians

// ------------------------------------------------- epistemic sigma (USGS) --
// abrahamson_2014.py:440-464
__device__ __forceinline__ double epistemic_sigma(double rrup, double mag) {
  const double n = 2.;
  double a, b, c;
  if (mag >= 5. && mag < 6.) { a = 11.; b = 38.; c = 94.; }
  else if (mag >= 6. && mag < 7.) { a = 2.; b = 7.; c = 13.; }
  else { a = 2.; b = 2.; c = 4.; }
  double d = rrup <= 10. ? a : (rrup < 30. ? b : c);
  return 0.4 * sqrt(n / d);
}

// =================================================================== BSSA14
struct Bssa14Row { double pga_rock; };

__device__ __forceinline__ double bssa_fm(const Bssa14C &C, const Row &r,
                                          bool sof) {
  double dm = r.mag - C.Mh;
  double fm = r.mag <= C.Mh ? C.e4 * dm + C.e5 * (dm * dm) : C.e6 * dm;
  if (!sof) return C.e0 + fm;
  double ar = fabs(r.rake);
  double s = C.e2;
  if (ar <= 30. || (180. - ar) <= 30.) s = C.e1;
  if (r.rake > 30. && r.rake < 150.) s = C.e3;
  return s + fm;
}

__device__ __forceinline__ double bssa_fp(const Bssa14C &C, const Row &r) {
  double rr = sqrt(r.rjb * r.rjb + C.h * C.h);
  return (C.c1 + C.c2 * (r.mag - 4.5)) * log(rr / 1.0) +
         (C.c3 + C.Dc3) * (rr - 1.0);
}

__device__ __forceinline__ Bssa14Row bssa14_row(const oqg_gsim &g,
                                                const Bssa14C &Cp,
                                                const Row &r) {
  Bssa14Row s;
  s.pga_rock = exp(bssa_fm(Cp, r, (g.flags >> 4) & 1) + bssa_fp(Cp, r));
  return s;
}

__device__ __forceinline__ double bssa_mu_z1(int region, double vs30) {
  double ln;
  if (region == 2) {  // JPN
    double k = 412.39 * 412.39;
    ln = -5.23 / 2.0 * log((vs30 * vs30 + k) / (1360.0 * 1360.0 + k)) -
         log(1000.);
  } else {
    double k = pow(570.94, 4.0);
    ln = -7.15 / 4.0 * log((pow(vs30, 4.0) + k) / (pow(1360.0, 4.0) + k)) -
         log(1000.);
  }
  return exp(ln);
}

__device__ __forceinline__ Out bssa14_imt(const oqg_gsim &g, const Bssa14C &C,
                                          const Row &r, const Bssa14Row &s) {
  Out o;
  const int region = g.flags & 3;
  const bool sof = (g.flags >> 4) & 1;
  // site term, boore_2014.py:160-166, 183-192, 244-256
  double flin = r.vs30 > C.Vc ? C.Vc / 760.0 : r.vs30 / 760.0;
  flin = C.c * log(flin);
  double vs = fmin(r.vs30, 760.);
  double f2 = C.f4 * (exp(C.f5 * (vs - 360.)) - exp(C.f5 * 400.));
  double fnl = 0.0 + f2 * log((s.pga_rock + 0.1) / 0.1);
  double fbd = 0.;
  if (region != 0 && C.period >= 0.65) {
    double mu = bssa_mu_z1(region, r.vs30);
    double z1 = r.z1pt0 == -999. ? mu * 1000. : r.z1pt0;
    double dz1 = (z1 / 1000.) - mu;
    fbd = dz1 <= C.f7 / C.f6 ? C.f6 * dz1 : C.f7;
  }
  o.mean = (bssa_fm(C, r, sof) + bssa_fp(C, r) + (flin + fnl + fbd));
  if (g.par[0] != 0.) o.mean += g.par[0] * epistemic_sigma(r.rrup, r.mag);
  // stddevs, boore_2014.py:107-145, 259-266
  double tau = C.tau1 + (C.tau2 - C.tau1) * (r.mag - 4.5);
  double phi = C.f1 + (C.f2 - C.f1) * (r.mag - 4.5);
  if (r.mag <= 4.5) { tau = C.tau1; phi = C.f1; }
  if (r.mag >= 5.5) { tau = C.tau2; phi = C.f2; }
  if (r.rjb > C.R2) phi += C.DfR;
  else if (r.rjb > C.R1) phi += C.DfR * (log(r.rjb / C.R1) / log(C.R2 / C.R1));
  if (r.vs30 <= 225.) phi -= C.DfV;
  if (r.vs30 >= 225. && r.vs30 <= 300.)
    phi -= C.DfV * (log(300. / r.vs30) / log(300. / 225.));
  o.tau = tau;
  o.phi = phi;
  o.sig = sqrt(tau * tau + phi * phi);
  return o;
}

// ===================================================================== CY14
struct Cy14Row { double pga_mean; };

__device__ __forceinline__ double cy_mean_z1(int region, double vs30) {
  if (region == 1)  // JPN
    return (-5.23 / 2.) * log((vs30 * vs30 + 412.39 * 412.39) /
                              (1360. * 1360. + 412.39 * 412.39));
  return (-7.15 / 4.) * log((pow(vs30, 4.) + pow(570.94, 4.)) /
                            (pow(1360., 4.) + pow(570.94, 4.)));
}

__device__ __forceinline__ Out cy14_one(const oqg_gsim &g, const Cy14C &C,
                                        const Row &r) {
  const int region = g.flags & 3;
  const bool use_hw = (g.flags >> 4) & 1, peer = (g.flags >> 5) & 1;
  const double alpha_nm = g.par[1];
  const double C2 = 1.06, C4 = -2.1, C4A = -0.5, CRB = 50.0, REFVS = 1130.;
  const double mag = r.mag, rrup = r.rrup;
  // delta ztor, chiou_youngs_2014.py:87-99
  bool rev = 30. <= r.rake && r.rake <= 150.;
  double mz;
  if (rev) mz = fmax(2.704 - 1.226 * fmax(mag - 5.849, 0.), 0.);
  else mz = fmax(2.673 - 1.136 * fmax(mag - 4.970, 0.), 0.);
  double dz = r.ztor - mz * mz;
  // magnitude scaling :489-498
  double fm = log(1.0 + exp(C.cn * (C.cm + 0. - mag)));
  fm = (C2 * (mag - 6.0) + ((C2 - C.c3) / C.cn) * fm - (C2 - C.c3) * 0.);
  // source terms :534-557
  double coshm = cosh(2.0 * fmax(mag - 4.5, 0.));
  double fsrc = 0.;
  if (rev) fsrc += C.c1a + (C.c1c / coshm);
  if (-120. <= r.rake && r.rake <= -60.)
    fsrc += (C.c1b + (C.c1d / coshm)) * alpha_nm;
  fsrc += (C.c7 + (C.c7b / coshm)) * dz;
  double cd = cos(r.dip * OQG_D2R);
  fsrc += (0.0 + (C.c11b / coshm)) * (cd * cd);
  // geometric spreading :331-338
  double fgeo = C4 * log(rrup + C.c5 * cosh(C.c6 * fmax(mag - C.chm, 0.0)));
  // far field :252-328
  double fr = (C4A - C4) * log(sqrt(rrup * rrup + CRB * CRB));
  double gamma = C.cg1 + C.cg2 / cosh(fmax(mag - C.cg3, 0.0));
  double frm = (gamma + 0.) * rrup;
  if (region == 1 || region == 2) {
    if (mag > 6.0 && mag < 6.9) frm = frm * C.gjpit;
  } else if (region == 3) {
    frm = frm * C.gwn;
  }
  double ffar = fr + frm;
  double ln_yref = C.c1 + fm + fsrc + fgeo + ffar + 0.;
  if (use_hw) {  // :341-353
    double fhw = 0.;
    if (r.rx >= 0.0) {
      double fd = 1.0 - (sqrt(r.rjb * r.rjb + r.ztor * r.ztor) / (rrup + 1.0));
      fd *= C.c9a + (1.0 - C.c9a) * tanh(r.rx / C.c9b);
      fhw += C.c9 * cd * fd;
    }
    ln_yref += fhw;
  }
  double yref = exp(ln_yref);
  // basin :175-213
  double mz1 = exp(cy_mean_z1(region, r.vs30));
  double z1 = r.z1pt0 == -999. ? mz1 : r.z1pt0;
  double dz1 = z1 - mz1;
  if (z1 <= 0.0) dz1 = 0.0;
  double fz1;
  if (region == 1) fz1 = C.phi5jp * (1.0 - exp(-dz1 / 800.0));
  else fz1 = (C.phi5 * (1.0 - exp(-dz1 / 300.0))) + 0.;
  double k1 = region == 1 ? C.phi1jp : C.phi1;
  double flin = k1 * fmin(log(r.vs30 / REFVS), 0.0);
  double vs = fmin(r.vs30, REFVS);
  double fnl_scale = C.phi2 * (exp(C.phi3 * (vs - 360.)) -
                               exp(C.phi3 * (REFVS - 360.)));
  double fnl = log((yref + C.phi4) / C.phi4) * fnl_scale;
  Out o;
  o.mean = ln_yref + flin + fnl + fz1;
  if (peer) { o.sig = 0.65; o.tau = 0.; o.phi = 0.; return o; }
  // stddevs :560-592, 515-531
  double nl0 = fnl_scale * (yref / (yref + C.phi4));
  double mclip = fmin(fmax(mag - 5.0, 0.), 1.5);
  double tau = C.tau1 + (C.tau2 - C.tau1) / 1.5 * mclip;
  double phi = r.vs30measured ? 0.7 : C.sig3;
  double nl1 = (1.0 + nl0) * (1.0 + nl0);
  phi = sqrt(phi + nl1);
  double mdep = C.sig1 + (C.sig2 - C.sig1) * mclip / 1.5;
  phi = mdep * phi;
  o.sig = sqrt(nl1 * (tau * tau) + phi * phi);
  o.tau = fabs((1. + nl0) * tau);
  o.phi = phi;
  return o;
}

__device__ __forceinline__ Cy14Row cy14_row(const oqg_gsim &g, const Cy14C &Cp,
                                            const Row &r) {
  Cy14Row s;
  s.pga_mean = cy14_one(g, Cp, r).mean;
  return s;
}

__device__ __forceinline__ Out cy14_imt(const oqg_gsim &g, const Cy14C &C,
                                        const Row &r, const Cy14Row &s) {
  Out o = cy14_one(g, C, r);
  // SA(T <= 0.3) >= PGA, chiou_youngs_2014.py:799-801 (period < 0 codes a
  // non-SA IMT, period == 0 the PGA)
  if (C.period > 0. && C.period <= 0.3 && o.mean < s.pga_mean)
    o.mean = s.pga_mean;
  o.mean = o.mean + g.par[0] * epistemic_sigma(r.rrup, r.mag);
  return o;
}

// ==================================================================== ASK14
struct Ask14Row { int dummy; };

// scipy interp1d(kind='linear') on a sorted grid (interval by searchsorted
// clipped to [1, n-1]); abrahamson_2014.py:225-230, 304-307
template <int N>
__device__ __forceinline__ double interp_lin(double x, const double (&xp)[N],
                                             const double (&fp)[N]) {
  double x0 = xp[N - 2], x1 = xp[N - 1], f0 = fp[N - 2], f1 = fp[N - 1];
#pragma unroll
  for (int i = N - 1; i >= 1; i--)
    if (xp[i] >= x) { x0 = xp[i - 1]; x1 = xp[i]; f0 = fp[i - 1]; f1 = fp[i]; }
  double slope = (f1 - f0) / (x1 - x0);
  return slope * (x - x0) + f0;
}

__device__ __forceinline__ double ask_basic(const Ask14C &C, const Row &r) {
  const double mag = r.mag, rrup = r.rrup;
  double c4m = C.c4 - (C.c4 - 1.) * (5. - mag);
  if (mag > 5.) c4m = C.c4;
  if (mag < 4.) c4m = 1.;
  double R = sqrt(rrup * rrup + c4m * c4m);
  double base = C.a1 * 1.0 + C.a17 * rrup;
  const double m1 = C.m1, m2 = 5.0;
  double lr = log(R), t;
  if (mag >= m1)
    t = C.a5 * (mag - m1) + C.a8 * ((8.5 - mag) * (8.5 - mag)) +
        (C.a2 + C.a3 * (mag - m1)) * lr;
  else if (mag >= m2)
    t = C.a4 * (mag - m1) + C.a8 * ((8.5 - mag) * (8.5 - mag)) +
        (C.a2 + C.a3 * (mag - m1)) * lr;
  else
    t = C.a4 * (m2 - m1) + C.a8 * ((8.5 - m2) * (8.5 - m2)) +
        C.a6 * (mag - m2) + C.a7 * ((mag - m2) * (mag - m2)) +
        (C.a2 + C.a3 * (m2 - m1)) * lr;
  return base + t;
}

__device__ __forceinline__ double ask_sof(const Ask14C &C, const Row &r) {
  double cl = fmin(fmax(r.mag - 4., 0.), 1.);
  double f7 = C.a11 * cl, f8 = C.a12 * cl;
  return f7 * ((r.rake > 30. && r.rake < 150.) ? 1. : 0.) +
         f8 * ((r.rake > -150. && r.rake < -30.) ? 1. : 0.);
}

__device__ __forceinline__ double ask_hw(const Ask14C &C, const Row &r) {
  const double mag = r.mag, rx = r.rx, dip = r.dip;
  double fhw = rx > 0. ? 1. : 0.;
  double t1 = dip <= 30. ? 60. / 45. : (90. - dip) / 45.0;
  const double a2hw = 0.2;
  double m65 = mag - 6.5;
  double t2 = (1. + a2hw * m65 - (1. - a2hw) * (m65 * m65));
  if (mag > 6.5) t2 = 1. + a2hw * m65;
  if (mag <= 5.5) t2 = 0.;
  double r1 = r.width * cos(dip * OQG_D2R);
  double r2 = 3. * r1;
  double t3;
  if (rx < r1) {
    double q = rx / r1;
    t3 = 0.25 + 1.50 * q + (-0.75) * (q * q);
  } else if (rx <= r2) {
    t3 = 1. - (rx - r1) / (r2 - r1);
  } else {
    t3 = 0.;
  }
  double t4 = r.ztor <= 10. ? 1. - (r.ztor * r.ztor) / 100. : 0.;
  double ry1 = rx * tan(20. * OQG_D2R);
  double d = r.ry0 - ry1;
  double t5 = d <= 0.0 ? 1. : (d < 5.0 ? 1. - d / 5.0 : 0.);
  if (dip == 90.) fhw = 0.;
  return fhw * C.a13 * t1 * t2 * t3 * t4 * t5;
}

__device__ __forceinline__ double ask_site(const Ask14C &C, double vs30,
                                           double sa1180) {
  double vstar = vs30 >= C.v1 ? C.v1 : vs30;
  double rat = vstar / C.vlin;
  if (vs30 >= C.vlin) return (C.a10 + C.b * 1.5) * log(rat);
  return (C.a10 * log(rat) - C.b * log(sa1180 + C.c) +
          C.b * log(sa1180 + C.c * pow(rat, 1.5)));
}

__device__ __forceinline__ double ask_z1ref(int region, double vs30) {
  if (region == 1)
    return exp(-5.23 / 2. * log((vs30 * vs30 + 412. * 412.) /
                                (1360. * 1360. + 412. * 412.))) / 1000.;
  return exp((-7.67 / 4.) * log((pow(vs30, 4.) + pow(610., 4.)) /
                                (pow(1360., 4.) + pow(610., 4.)))) / 1000.;
}

__device__ __forceinline__ double ask_basin_soil(const Ask14C &C, const Row &r,
                                                 int region) {
  double z1ref = ask_z1ref(region, r.vs30);
  double z10 = r.z1pt0 / 1000.;
  if (z10 < 0.) z10 = z1ref;
  double factor = log((z10 + 0.01) / (z1ref + 0.01));
  const double xp[7] = {0.0, 150., 250., 400., 700., 1000., 6000.};
  const double fp[7] = {C.a43, C.a43, C.a44, C.a45, C.a46, C.a46, C.a46};
  double f2 = interp_lin<7>(r.vs30, xp, fp);
  return (f2 * factor) + 0.;
}

__device__ __forceinline__ double ask_regional(const Ask14C &C, int region,
                                               double vs30, double rrup) {
  if (region == 2) {  // TWN
    double vstar = vs30 >= C.v1 ? C.v1 : vs30;
    return C.a31 * log(vstar / C.vlin) + C.a25 * rrup;
  } else if (region == 3) {  // CHN
    return C.a28 * rrup;
  } else if (region == 1) {  // JPN
    const double xp[8] = {150., 250., 350., 450., 600., 850., 1150., 2000.};
    const double fp[8] = {C.a36, C.a37, C.a38, C.a39, C.a40, C.a41, C.a42,
                          C.a42};
    return interp_lin<8>(vs30, xp, fp) + C.a29 * rrup;
  }
  return 0.;
}

__device__ __forceinline__ Ask14Row ask14_row(const oqg_gsim &, const Ask14C &,
                                              const Row &) {
  return Ask14Row();
}

__device__ __forceinline__ Out ask14_imt(const oqg_gsim &g, const Ask14C &C,
                                         const Row &r, const Ask14Row &) {
  const int region = g.flags & 3;
  double basic = ask_basic(C, r), sof = ask_sof(C, r), hw = ask_hw(C, r);
  double ztor_t = C.a15 * fmin(r.ztor / 20.0, 1.);
  // _get_sa_at_1180, abrahamson_2014.py:419-438 (rock basin term is 0)
  double sa1180 = exp(basic + sof + ask_site(C, 1180., 0.) + hw + ztor_t +
                      0. + ask_regional(C, region, 1180., r.rrup));
  double mean = (basic + hw + ask_site(C, r.vs30, sa1180) + ztor_t + sof +
                 ask_basin_soil(C, r, region));
  mean = mean + ask_regional(C, region, r.vs30, r.rrup);
  mean = mean + g.par[0] * epistemic_sigma(r.rrup, r.mag);
  // stddevs :61-90, 180-199, 306-316
  const double mag = r.mag;
  double phi_al;
  if (region == 1) {
    double rr = r.rrup;
    phi_al = rr < 30. ? C.s5
                      : (rr <= 80. ? C.s5 + (C.s6 - C.s5) / 50. * (rr - 30.)
                                   : C.s6);
  } else {
    double s1 = r.vs30measured ? C.s1m : C.s1e;
    double s2 = r.vs30measured ? C.s2m : C.s2e;
    phi_al = s1 + (s2 - s1) / 2. * (mag - 4.);
    if (mag < 4.) phi_al = s1;
    if (mag >= 6.) phi_al = s2;
  }
  double der = 0.;
  if (r.vs30 < C.vlin)
    der = (C.b * sa1180 *
           (-1. / (sa1180 + C.c) +
            1. / (sa1180 + C.c * pow(r.vs30 / C.vlin, 1.5))));
  double phi_amp = phi_al < 0.4 ? 0.99 * phi_al : 0.4;
  double phi_b = sqrt(phi_al * phi_al - phi_amp * phi_amp);
  double phi = sqrt((phi_b * phi_b) * ((1. + der) * (1. + der)) +
                    phi_amp * phi_amp);
  double tau_al = C.s3 + (C.s4 - C.s3) / 2. * (mag - 5.);
  if (mag < 5.) tau_al = C.s3;
  if (mag >= 7.) tau_al = C.s4;
  double tau = tau_al * (1. + der);
  Out o;
  o.mean = mean;
  o.sig = sqrt(phi * phi + tau * tau);
  o.tau = tau;
  o.phi = phi;
  return o;
}

// ===================================================================== CB14
struct Cb14Row { double pga1100, tau_pga, phi_pga, pga_mean; };

__device__ __forceinline__ double cb_z25ref(int SJ, double vs30) {
  if (SJ) return exp(5.359 - 1.102 * log(vs30));
  return exp(7.089 - 1.144 * log(vs30));
}

// get_mean_values, campbell_bozorgnia_2014.py:349-370; rock pass: vs30=1100
__device__ __forceinline__ double cb_mean(int SJ, const Cb14C &C, const Row &r,
                                          bool rock, double a1100) {
  const double CB_C = 1.88, CB_N = 1.18;
  const double mag = r.mag, rrup = r.rrup, rx = r.rx;
  double vs30 = rock ? 1100.0 : r.vs30;
  double fmag = C.c0 + C.c1 * mag;
  double t;
  if (mag > 6.5)
    t = C.c2 * (mag - 4.5) + C.c3 * (mag - 5.5) + C.c4 * (mag - 6.5);
  else if (mag > 5.5) t = C.c2 * (mag - 4.5) + C.c3 * (mag - 5.5);
  else if (mag > 4.5) t = C.c2 * (mag - 4.5);
  else t = 0.;
  fmag = fmag + t;
  double fdis = (C.c5 + C.c6 * mag) * log(sqrt(rrup * rrup + C.c7 * C.c7));
  double frv = (r.rake > 30. && r.rake < 150.) ? 1. : 0.;
  double fnm = (r.rake > -150. && r.rake < -30.) ? 1. : 0.;
  double fflt_m = mag <= 4.5 ? 0. : (mag > 5.5 ? 1. : mag - 4.5);
  double fflt = (C.c8 * frv + C.c9 * fnm) * fflt_m;
  // hanging wall :179-253
  double r1 = r.width * cos(r.dip * OQG_D2R);
  double r2 = 62.0 * mag - 350.0;
  double fhng_rx;
  if (rx >= r1) {
    double d = (rx - r1) / (r2 - r1);
    fhng_rx = fmax(1.0 + C.h5 * d + C.h6 * (d * d), 0.0);
  } else if (rx >= 0.) {
    double q = rx / r1;
    fhng_rx = C.h1 + C.h2 * q + C.h3 * (q * q);
  } else {
    fhng_rx = 0.;
  }
  double fhng_rrup = rrup > 0.0 ? (rrup - r.rjb) / rrup : 1.;
  double fhng_m = (mag - 5.5) * (1.0 + C.a2 * (mag - 6.5));
  if (mag < 5.5) fhng_m = 0.0;
  if (mag > 6.5) fhng_m = 1.0 + C.a2 * (mag - 6.5);
  double fhng_z = r.ztor > 16.66 ? 0. : 1. - 0.06 * r.ztor;
  double fhng_d = (90.0 - r.dip) / 45.0;
  double fhng = C.c10 * fhng_rx * fhng_rrup * fhng_m * fhng_z * fhng_d;
  // shallow site response :297-323
  double vmod = vs30 / C.k1;
  double lv = log(vmod);
  double fsite = C.c11 * lv;
  bool stiff = vs30 > C.k1;
  if (stiff) fsite = fsite + (C.k2 * CB_N * lv);
  else if (!rock) {
    double nl = C.k2 * (log(a1100 + CB_C * pow(vmod, CB_N)) -
                        log(a1100 + CB_C));
    fsite = fsite + nl;
  }
  if (SJ) {
    double fj = (C.c13 + C.k2 * CB_N) * lv;
    double soft = (C.c12 + C.k2 * CB_N) * (lv - log(200.0 / C.k1));
    fsite = fsite + (vs30 <= 200.0 ? fj + soft : fj);
  }
  // basin :116-150, 80-99
  double z25;
  if (rock) z25 = cb_z25ref(SJ, 1100.0);
  else z25 = r.z2pt5 == -999. ? cb_z25ref(SJ, r.vs30) : r.z2pt5;
  double fsed = 0.;
  if (z25 < 1.0) fsed = (C.c14 + C.c15 * SJ) * (z25 - 1.0);
  if (z25 > 3.0)
    fsed = (C.c16 * C.k3) * exp(-0.75) * (1. - exp(-0.25 * (z25 - 3.)));
  // hypocentral depth :256-265
  double fh = fmin(fmax(r.hypo_depth - 7.0, 0.), 13.);
  double fm = C.c17 + (C.c18 - C.c17) * (mag - 5.5);
  if (mag <= 5.5) fm = C.c17;
  if (mag > 6.5) fm = C.c18;
  double fhyp = fh * fm;
  // dip :160-168
  double fdip = C.c19 * (5.5 - mag) * r.dip;
  if (mag < 4.5) fdip = C.c19 * r.dip;
  if (mag > 5.5) fdip = 0.0;
  // anelastic :69-77
  double fatn = rrup >= 80.0 ? (C.c20 + C.Dc20) * (rrup - 80.0) : 0.;
  return fmag + fdis + fflt + fhng + fsite + fsed + fhyp + fdip + fatn;
}

__device__ __forceinline__ double cb_lin(double mag, double a, double b) {
  double res = b + (a - b) * (5.5 - mag);
  if (mag <= 4.5) res = a;
  if (mag >= 5.5) res = b;
  return res;
}

__device__ __forceinline__ Cb14Row cb14_row(const oqg_gsim &g, const Cb14C &Cp,
                                            const Row &r) {
  const int SJ = g.flags & 1;
  Cb14Row s;
  s.pga1100 = exp(cb_mean(SJ, Cp, r, true, 0.));
  s.tau_pga = cb_lin(r.mag, Cp.tau1, Cp.tau2);
  double p = cb_lin(r.mag, Cp.phi1, Cp.phi2);
  s.phi_pga = sqrt(p * p - Cp.philnAF * Cp.philnAF);
  s.pga_mean = cb_mean(SJ, Cp, r, false, s.pga1100);
  return s;
}

__device__ __forceinline__ Out cb14_imt(const oqg_gsim &g, const Cb14C &C,
                                        const Row &r, const Cb14Row &s) {
  const double CB_C = 1.88, CB_N = 1.18;
  const int SJ = g.flags & 1;
  Out o;
  double eps = g.par[0] * epistemic_sigma(r.rrup, r.mag);
  double mean = cb_mean(SJ, C, r, false, s.pga1100);
  mean = mean + eps;
  if (C.period > 0. && C.period < 0.25) {  // :554-564
    if (mean <= s.pga_mean) mean = s.pga_mean;
    mean = mean + eps;  // added a second time in the reference
  }
  o.mean = mean;
  // compute_sigma :432-463
  double tau_b = cb_lin(r.mag, C.tau1, C.tau2);
  double pl = cb_lin(r.mag, C.phi1, C.phi2);
  double phi_b = sqrt(pl * pl - C.philnAF * C.philnAF);
  double alpha = 0.;
  if (r.vs30 < C.k1) {
    double af1 = s.pga1100 + CB_C * pow(r.vs30 / C.k1, CB_N);
    double af2 = s.pga1100 + CB_C;
    alpha = C.k2 * s.pga1100 * ((1.0 / af1) - (1.0 / af2));
  }
  double t = sqrt(tau_b * tau_b + (alpha * alpha) * (s.tau_pga * s.tau_pga) +
                  2.0 * alpha * C.rholny * tau_b * s.tau_pga);
  double p = sqrt(phi_b * phi_b + C.philnAF * C.philnAF +
                  (alpha * alpha) * (s.phi_pga * s.phi_pga) +
                  2.0 * alpha * C.rholny * phi_b * s.phi_pga);
  o.sig = sqrt(t * t + p * p);
  o.tau = t;
  o.phi = p;
  return o;
}

// ============================================================ BA08 and AB06
__device__ __forceinline__ double ba_bnl(double vs30, double b1, double b2) {
  // atkinson_boore_2006.py:140-163
  if (vs30 <= 180.0) return b1;
  if (vs30 <= 300.0)
    return (b1 - b2) * log(vs30 / 300.0) / log(180.0 / 300.0) + b2;
  if (vs30 < 760.0) return b2 * log(vs30 / 760.0) / log(300.0 / 760.0);
  return 0.;
}

__device__ __forceinline__ double ba_fnl(double pga4nl, double bnl) {
  // atkinson_boore_2006.py:166-195
  const double a1 = 0.03, a2 = 0.09, pga_low = 0.06;
  double dx = log(a2 / a1);
  double dy = bnl * log(a2 / pga_low);
  double c = (3. * dy - bnl * dx) / (dx * dx);
  double d = -(2. * dy - bnl * dx) / (dx * dx * dx);
  if (pga4nl <= a1) return bnl * log(pga_low / 0.1);
  if (pga4nl <= a2) {
    double lp = log(pga4nl / a1);
    return bnl * log(pga_low / 0.1) + c * (lp * lp) + d * (lp * lp * lp);
  }
  return bnl * log(pga4nl / 0.1);
}

struct Ba08Row { double pga4nl; };

__device__ __forceinline__ double ba_fm(const Ba08C &C, const Row &r) {
  double ar = fabs(r.rake);
  double ss = (ar <= 30. || 180. - ar <= 30.) ? 1. : 0.;
  double rs = (r.rake > 30. && r.rake < 150.) ? 1. : 0.;
  double ns = (r.rake > -150. && r.rake < -30.) ? 1. : 0.;
  double dm = r.mag - C.Mh;
  if (r.mag <= C.Mh)
    return C.e2 * ss + C.e3 * ns + C.e4 * rs + C.e5 * dm + C.e6 * (dm * dm);
  return C.e2 * ss + C.e3 * ns + C.e4 * rs + C.e7 * dm;
}

__device__ __forceinline__ double ba_fd(const Ba08C &C, const Row &r) {
  double R = sqrt(r.rjb * r.rjb + C.h * C.h);
  return (C.c1 + C.c2 * (r.mag - 4.5)) * log(R / 1.0) + C.c3 * (R - 1.0);
}

__device__ __forceinline__ Ba08Row ba08_row(const oqg_gsim &, const Ba08C &Cp,
                                            const Row &r) {
  Ba08Row s;
  s.pga4nl = exp(ba_fm(Cp, r) + ba_fd(Cp, r));
  return s;
}

__device__ __forceinline__ Out ba08_imt(const oqg_gsim &, const Ba08C &C,
                                        const Row &r, const Ba08Row &s) {
  double lin = C.so_blin * log(r.vs30 / 760.0);
  double nl = ba_fnl(s.pga4nl, ba_bnl(r.vs30, C.so_b1, C.so_b2));
  Out o;
  if (C.is_pga != 0.) o.mean = log(s.pga4nl) + lin + nl;
  else o.mean = ba_fm(C, r) + ba_fd(C, r) + lin + nl;
  o.sig = C.std;
  o.tau = C.tau;
  o.phi = C.sigma;
  return o;
}

struct Ab06Row { double mag, rrup, f0, f1, f2; };

__device__ __forceinline__ Ab06Row ab06_row(const oqg_gsim &g, const Ab06C &,
                                            const Row &r) {
  Ab06Row s;
  const int mag_eq = g.flags & 3;
  double mag = r.mag;
  if (mag_eq == 1) mag = 2.715 - 0.277 * mag + 0.127 * mag * mag;
  else if (mag_eq == 2) mag = 1.14 + 0.24 * mag + 0.0933 * mag * mag;
  s.mag = mag;
  s.rrup = r.rrup < 1. ? 1. : r.rrup;
  s.f0 = fmax(log10(10.0 / s.rrup), 0.0);
  s.f1 = fmin(log10(s.rrup), log10(70.0));
  s.f2 = fmax(log10(s.rrup / 140.0), 0.0);
  return s;
}

__device__ __forceinline__ double ab06_rock(const double *c, const Ab06C &SC,
                                            const Ab06Row &s,
                                            double scale_fac) {
  // c[0..9] = c1..c10; atkinson_boore_2006.py:109-124, 213-221
  double cm = fmax(s.mag - SC.st_M1, 0.);
  double sda = scale_fac * fmin(0.05 + SC.st_delta * cm / (SC.st_Mh - SC.st_M1),
                                SC.st_delta + 0.05);
  const double mag = s.mag;
  return (c[0] + c[1] * mag + c[2] * (mag * mag) + (c[3] + c[4] * mag) * s.f1 +
          (c[5] + c[6] * mag) * s.f2 + (c[7] + c[8] * mag) * s.f0 +
          c[9] * s.rrup + sda);
}

__device__ __forceinline__ Out ab06_imt(const oqg_gsim &g, const Ab06C &C,
                                        const Ab06C &Cp, const Row &r,
                                        const Ab06Row &s) {
  const double G9 = 9.80665;
  const int mag_eq = g.flags & 3;
  double scale_fac = mag_eq == 0 ? 0. : g.par[0];
  bool soft = r.vs30 < 2000.0;
  double mean;
  if (soft) {
    double pga_bc = ab06_rock(&Cp.bc_c1, C, s, scale_fac);
    pga_bc = pow(10., pga_bc) * 1e-2 / G9;
    mean = ab06_rock(&C.bc_c1, C, s, scale_fac);
    double lin = C.so_blin * log(r.vs30 / 760.0);
    double nl = ba_fnl(pga_bc, ba_bnl(r.vs30, C.so_b1, C.so_b2));
    mean = mean + log10(exp(lin)) + log10(exp(nl));
  } else {
    mean = ab06_rock(&C.hr_c1, C, s, scale_fac);
  }
  if (C.is_pgv != 0.) mean = log(pow(10., mean));
  else mean = log(pow(10., mean) * 1e-2 / G9);
  if (mag_eq != 0) {  // clip_mean, gsim/utils.py:45-57
    if (C.period == 0.) mean = fmin(mean, 0.405);
    if (C.period > 0.02 && C.period < 0.55) mean = fmin(mean, 1.099);
  }
  Out o;
  o.mean = mean;
  o.sig = log(pow(10., 0.30));
  o.tau = 0.;
  o.phi = 0.;
  return o;
}

// ================================================================== AKKAR14
struct Akkar14Row { double pga, dist; };

__device__ __forceinline__ double akkar_rock(const Akkar14C &C, const Row &r,
                                             double dist) {
  const double c1 = 6.75, mag = r.mag;
  double fn = (r.rake > -135.0 && r.rake < -45.0) ? 1. : 0.;
  double fr = (r.rake > 45.0 && r.rake < 135.0) ? 1. : 0.;
  return (C.a1 + (mag <= c1 ? C.a2 * (mag - c1) : C.a7 * (mag - c1)) +
          C.a3 * ((8.5 - mag) * (8.5 - mag)) +
          (C.a4 + C.a5 * (mag - c1)) * log(sqrt(dist * dist + C.a6 * C.a6)) +
          (C.a8 * fn + C.a9 * fr));
}

__device__ __forceinline__ Akkar14Row akkar14_row(const oqg_gsim &g,
                                                  const Akkar14C &Cp,
                                                  const Row &r) {
  Akkar14Row s;
  const int kind = g.flags & 3;
  s.dist = kind == 0 ? r.rjb : (kind == 1 ? r.repi : r.rhypo);
  s.pga = exp(akkar_rock(Cp, r, s.dist));
  return s;
}

__device__ __forceinline__ Out akkar14_imt(const oqg_gsim &g,
                                           const Akkar14C &C, const Row &r,
                                           const Akkar14Row &s) {
  const double vref = 750.0, vcon = 1000.0;
  double site;
  if (r.vs30 < vref) {
    double vr = pow(r.vs30 / vref, C.n);
    site = C.b1 * log(r.vs30 / vref) +
           C.b2 * log((s.pga + C.c * vr) / ((s.pga + C.c) * vr));
  } else if (r.vs30 <= vcon) {
    site = C.b1 * log(r.vs30 / vref);
  } else {
    site = C.b1 * log(vcon / vref);
  }
  Out o;
  o.mean = akkar_rock(C, r, s.dist) + site + g.par[0];
  o.sig = sqrt(C.sigma * C.sigma + C.tau * C.tau);
  o.tau = C.tau;
  o.phi = C.sigma;
  return o;
}

// ===================================================================== CY08
// chiou_youngs_2008.py: no rock-PGA pass, every IMT has its own reference
// motion ln_y_ref (eq. 13a), site response (eq. 13b) and sigma (eq. 19-21)
struct NoRow {};

__device__ __forceinline__ Out cy08_imt(const oqg_gsim &, const Cy08C &C,
                                        const Row &r) {
  // :56-99
  const double frv = (30. <= r.rake && r.rake <= 150.) ? 1. : 0.;
  const double fnm = (-120. <= r.rake && r.rake <= -60.) ? 1. : 0.;
  const double fhw = r.rx >= 0. ? 1. : 0.;
  const double cd = cos(r.dip * OQG_D2R);
  double ln_y_ref =
      C.c1 + (C.c1a * frv + C.c1b * fnm + C.c7 * (r.ztor - 4.)) +
      C.c2 * (r.mag - 6.) +
      ((C.c2 - C.c3) / C.cn) * log(1. + exp(C.cn * (C.cm - r.mag))) +
      C.c4 * log(r.rrup + C.c5 * cosh(C.c6 * fmax(r.mag - C.chm, 0.))) +
      (C.c4a - C.c4) * log(sqrt(r.rrup * r.rrup + C.crb * C.crb)) +
      (C.cg1 + C.cg2 / cosh(fmax(r.mag - C.cg3, 0.))) * r.rrup +
      C.c9 * fhw * tanh(r.rx * (cd * cd) / C.c9a) *
          (1. - sqrt(r.rjb * r.rjb + r.ztor * r.ztor) / (r.rrup + 0.001));
  // :218-219
  const double exp1 = exp(C.phi3 * (fmin(r.vs30, 1130.) - 360.));
  const double exp2 = exp(C.phi3 * (1130. - 360.));
  // basin term :37-53 (z1pt0 = -999: the model's own vs30 -> z1pt0, :29-34)
  double z1 = r.z1pt0;
  if (z1 == -999.)
    z1 = exp(28.5 - (3.82 / 8.) * log(pow(r.vs30, 8.) + pow(378.7, 8.)));
  const double deep =
      C.phi5 * (1.0 - 1.0 / cosh(C.phi6 * fmax(z1 - C.phi7, 0.)));
  const double shallow = C.phi8 / cosh(0.15 * fmax(z1 - 15., 0.));
  const double y_ref = exp(ln_y_ref);
  Out o;
  // :102-124
  o.mean = ln_y_ref + C.phi1 * fmin(log(r.vs30 / 1130.), 0.) +
           C.phi2 * (exp1 - exp2) * log((y_ref + C.phi4) / C.phi4) +
           (deep + shallow);
  // :127-164
  const double mag_test = fmin(fmax(r.mag, 5.0), 7.0) - 5.0;
  const double t = C.tau1 + (C.tau2 - C.tau1) / 2. * mag_test;
  const double b = C.phi2 * (exp1 - exp2);
  const double nl = b * y_ref / (y_ref + C.phi4);
  const double fmeas = r.vs30measured ? 1. : 0.;
  const double sigma = (C.sig1 + 0.5 * (C.sig2 - C.sig1) * mag_test) *
                       sqrt((C.sig3 * (1. - fmeas) + 0.7 * fmeas) +
                            (1. + nl) * (1. + nl));
  o.sig = sqrt(((1. + nl) * (1. + nl)) * (t * t) + sigma * sigma);
  o.tau = fabs((1. + nl) * t);
  o.phi = sigma;
  return o;
}

// =================================================================== TORO02
// numpy.interp on increasing nodes: constant outside, slope * (x - x0) + y0
// inside (numpy/_core/src/multiarray/compiled_base.c arr_interp)
__device__ __forceinline__ double np_interp3(double x, double x0, double x1,
                                             double x2, double y0, double y1,
                                             double y2) {
  if (x <= x0) return y0;
  if (x >= x2) return y2;
  if (x < x1) return (y1 - y0) / (x1 - x0) * (x - x0) + y0;
  return (y2 - y1) / (x2 - x1) * (x - x1) + y1;
}

// _compute_faulting_style_term, campbell_2003.py:272-281: the value of the
// rake class (reverse (30, 150], normal (-120, -60], otherwise strike slip)
__device__ __forceinline__ double share_pick(double rake, double ss, double rv,
                                             double nm) {
  if (rake > 30.0 && rake <= 150.0) return rv;
  if (rake > -120.0 && rake <= -60.0) return nm;
  return ss;
}

__device__ __forceinline__ Out toro02_imt(const oqg_gsim &g, const Toro02C &C,
                                          const Row &r) {
  // toro_2002.py:31-57
  const double md = r.mag - 6.;
  const double e = exp(-1.25 + 0.227 * r.mag);
  const double RM = sqrt(r.rjb * r.rjb + (C.c7 * C.c7) * (e * e));
  const double term1 = C.c2 * md + C.c3 * (md * md);
  const double term2 = -C.c4 * log(RM) -
                       (C.c5 - C.c4) * fmax(log(RM / 100.), 0.) - C.c6 * RM;
  Out o;
  o.mean = C.c1 + term1 + term2;
  // :110-128
  const double am = np_interp3(r.mag, 5.0, 5.5, 8.0, C.m50, C.m55, C.m80);
  double ar;
  if (r.rjb <= 5.0) ar = C.r5;
  else if (r.rjb >= 20.0) ar = C.r20;
  else ar = (C.r20 - C.r5) / (20.0 - 5.0) * (r.rjb - 5.0) + C.r5;
  const double ale = sqrt(am * am + ar * ar);
  const double epi = C.period < 1. ? 0.36 + 0.07 * md : 0.34 + 0.06 * md;
  o.sig = sqrt(ale * ale + epi * epi);
  if (C.period == 3.0) o.mean /= 0.612;
  if (C.period == 4.0) o.mean /= 0.559;
  if (g.flags & 1) {
    // ToroEtAl2002SHARE, toro_2002.py:134-142: style of faulting and rock
    // adjustment, log(exp(mean) * (term(rake) * AFrock)); the three products
    // come with the coefficient row (adj_*)
    o.mean = log(exp(o.mean) * share_pick(r.rake, C.adj_ss, C.adj_rv, C.adj_nm));
  }
  o.tau = 0.;
  o.phi = 0.;
  return o;
}

// ====================================================================== C03
__device__ __forceinline__ Out c03_imt(const oqg_gsim &g, const C03C &C,
                                       const Row &r) {
  // campbell_2003.py:39-47, 71-106 (kind "base")
  const double term1 = (C.c2 * r.mag) + C.c3 * ((8.5 - r.mag) * (8.5 - r.mag));
  const double c78 = C.c7 * exp(C.c8 * r.mag);
  const double R = sqrt(r.rrup * r.rrup + c78 * c78);
  const double term2 = C.c4 * log(R) + (C.c5 + C.c6 * r.mag) * r.rrup;
  double f3 = 0.;
  if (r.rrup > 70. && r.rrup <= 130.)
    f3 = C.c9 * (log(r.rrup) - log(70.));
  else if (r.rrup > 130.)
    f3 = C.c9 * (log(r.rrup) - log(70.)) + C.c10 * (log(r.rrup) - log(130.));
  Out o;
  o.mean = C.c1 + term1 + term2 + f3;
  o.sig = r.mag < 7.16 ? C.c11 + C.c12 * r.mag : C.c13;
  if (g.flags & 1)   // Campbell2003SHARE, campbell_2003.py:236-245
    o.mean = log(exp(o.mean) * share_pick(r.rake, C.fs_ss, C.fs_rv, C.fs_nm) *
                 C.afrock);
  o.tau = 0.;
  o.phi = 0.;
  return o;
}

// ================================================================= SADIGH97
__device__ __forceinline__ Out sadigh97_imt(const oqg_gsim &,
                                            const Sadigh97C &C, const Row &r) {
  // sadigh_1997.py:134-161: rock above 750 m/s, deep soil otherwise; total
  // standard deviation only
  const bool reverse = 45. <= r.rake && r.rake <= 135.;
  const bool lowmag = r.mag <= 6.5;             // NEAR_FIELD_SATURATION_MAG
  const double mag = fmin(r.mag, 8.5);
  Out o;
  if (r.vs30 > 750.) {
    // get_mean_rock :71-87, get_stddev_rock :90-97
    const double c1 = lowmag ? C.lo_c1 : C.hi_c1, c2 = lowmag ? C.lo_c2 : C.hi_c2,
                 c3 = lowmag ? C.lo_c3 : C.hi_c3, c4 = lowmag ? C.lo_c4 : C.hi_c4,
                 c5 = lowmag ? C.lo_c5 : C.hi_c5, c6 = lowmag ? C.lo_c6 : C.hi_c6,
                 c7 = lowmag ? C.lo_c7 : C.hi_c7;
    o.mean = c1 + c2 * mag + c3 * pow(8.5 - mag, 2.5) +
             c4 * log(r.rrup + exp(c5 + c6 * mag)) + c7 * log(r.rrup + 2.);
    if (reverse) o.mean += 0.1823215567939546;   // log(1.2)
    o.sig = r.mag > C.sd_maxmag ? C.sd_maxsigma
                                : C.sd_sigma0 + C.sd_magfactor * r.mag;
  } else {
    // get_mean_deep_soil :47-68 (COEFFS_SOIL_IMT_INDEPENDENT),
    // get_stddev_deep_soil :100-107
    const double c1 = reverse ? -1.92 : -2.17;
    const double c4 = lowmag ? 2.1863 : 0.3825, c5 = lowmag ? 0.32 : 0.5882;
    const double c6 = reverse ? C.so_c6r : C.so_c6ss;
    o.mean = c1 + 1.0 * mag + c6 + C.so_c7 * pow(8.5 - mag, 2.5) -
             1.7 * log(r.rrup + c4 * exp(c5 * mag));
    o.sig = C.so_sigma0 + C.so_magfactor * fmin(r.mag, 7.);
  }
  o.tau = 0.;
  o.phi = 0.;
  return o;
}

// ================================================================ dispatch
// Fam<F>: the coefficient row type, the per-row state and the two stages
template <int F> struct Fam;

#define OQG_FAM(ID, CT, RT, ROWF, IMTF, FIELDS)                              \
  template <> struct Fam<ID> {                                               \
    typedef CT C;                                                            \
    typedef RT R;                                                            \
    static __device__ __forceinline__ R row(const oqg_gsim &g, const C &Cp,  \
                                            const Row &r) {                  \
      return ROWF(g, Cp, r);                                                 \
    }                                                                        \
    static __device__ __forceinline__ Out imt(const oqg_gsim &g, const C &c, \
                                              const C &Cp, const Row &r,     \
                                              const R &s) {                  \
      (void)Cp;                                                              \
      return IMTF;                                                           \
    }                                                                        \
    static const char *fields() { return FIELDS(OQG_NAME); }                 \
  };

OQG_FAM(OQG_BSSA14, Bssa14C, Bssa14Row, bssa14_row, bssa14_imt(g, c, r, s),
        BSSA14_FIELDS)
OQG_FAM(OQG_CY14, Cy14C, Cy14Row, cy14_row, cy14_imt(g, c, r, s), CY14_FIELDS)
OQG_FAM(OQG_ASK14, Ask14C, Ask14Row, ask14_row, ask14_imt(g, c, r, s),
        ASK14_FIELDS)
OQG_FAM(OQG_CB14, Cb14C, Cb14Row, cb14_row, cb14_imt(g, c, r, s), CB14_FIELDS)
OQG_FAM(OQG_BA08, Ba08C, Ba08Row, ba08_row, ba08_imt(g, c, r, s), BA08_FIELDS)
OQG_FAM(OQG_AB06, Ab06C, Ab06Row, ab06_row, ab06_imt(g, c, Cp, r, s),
        AB06_FIELDS)
OQG_FAM(OQG_AKKAR14, Akkar14C, Akkar14Row, akkar14_row,
        akkar14_imt(g, c, r, s), AKKAR14_FIELDS)
template <typename CT>
__device__ __forceinline__ NoRow no_row_t(const oqg_gsim &, const CT &,
                                          const Row &) { return NoRow(); }
OQG_FAM(OQG_CY08, Cy08C, NoRow, no_row_t<Cy08C>, ((void)s, cy08_imt(g, c, r)),
        CY08_FIELDS)
OQG_FAM(OQG_TORO02, Toro02C, NoRow, no_row_t<Toro02C>,
        ((void)s, toro02_imt(g, c, r)), TORO02_FIELDS)
OQG_FAM(OQG_C03, C03C, NoRow, no_row_t<C03C>, ((void)s, c03_imt(g, c, r)),
        C03_FIELDS)
OQG_FAM(OQG_SADIGH97, Sadigh97C, NoRow, no_row_t<Sadigh97C>,
        ((void)s, sadigh97_imt(g, c, r)), SADIGH97_FIELDS)

// distances a model reads (REQUIRES_DISTANCES of the reference classes,
// SURVEY.md appendix B) as a mask of 1 << OQG_R*
__host__ __device__ inline int required_distances(const oqg_gsim &g) {
  const int RRUP = 1 << OQG_RRUP, RJB = 1 << OQG_RJB, RX = 1 << OQG_RX,
            RY0 = 1 << OQG_RY0;
  switch (g.family) {
    case OQG_BSSA14: return RJB;
    case OQG_CY14: return RRUP | RJB | RX;
    case OQG_ASK14: return RRUP | RJB | RX | RY0;
    case OQG_CB14: return RRUP | RJB | RX;
    case OQG_BA08: return RJB;
    case OQG_AB06: return RRUP;
    case OQG_AKKAR14: {
      int k = g.flags & 3;
      return k == 0 ? RJB : (k == 1 ? (1 << OQG_REPI) : (1 << OQG_RHYPO));
    }
    case OQG_CY08: return RRUP | RJB | RX;
    case OQG_TORO02: return RJB;
    case OQG_C03: return RRUP;
    case OQG_SADIGH97: return RRUP;
  }
  return 0;
}

}  // namespace oqg
