// Staged evaluation of the ground-motion models for the fused kernel.
//
// The fused kernel visits every (rupture, site, IMT) triple; most of a
// model does not depend on all three.  Fast<F> splits a model into
//
//   S  site(C, site)            per (site, IMT): vs30 / basin terms,
//                               registers, once per thread
//   Q  rup(C, rupture)          per (rupture, IMT): magnitude / mechanism
//                               terms, shared memory, once per tile
//   P  pair(Cp, Sp, Qp, ...)    per (rupture, site): the part shared by all
//                               IMTs (rock PGA ...), shared memory, computed
//                               once per pair by one of the IMT warps
//   eval(C, S, Q, P, dists)     per (rupture, site, IMT): mean and sigma
//
// so that the innermost stage keeps only the transcendental calls that
// really depend on the triple.  The arithmetic is the same as in
// oqg_gmpe.cuh (which follows the reference term by term) up to the
// association of a few additions and the table-driven log / exp / tanh /
// reciprocal of oqg_math.cuh in the per-pair and per-IMT stages; results
// agree to ~1e-15.  Per-IMT constants that would cost a transcendental or a
// divide per evaluation (log phi4, 1 / c9b ...) are precomputed on the host
// into the tail of the coefficient row (OQG_DER, filled by oqg_set_gsims).  Families
// without a specialisation go through the generic Fast<F>, which calls the
// unstaged functions (P = the row state of oqg_gmpe.cuh).
#pragma once
#include "oqg_gmpe.cuh"
#include "oqg_geo.cuh"
#include "oqg_math.cuh"

#ifndef OQG_CY_NOINLINE
#define OQG_CY_NOINLINE 0
#endif
#if OQG_CY_NOINLINE
#define OQG_CY_INLINE __noinline__
#else
#define OQG_CY_INLINE __forceinline__
#endif

namespace oqg {

struct SiteIn {
  double vs30, z1pt0, z2pt5;
  bool vs30measured;
};

struct RupIn {
  double mag, rake, dip, ztor, width, hypo_depth;
};

struct Empty {};

// derived per-IMT constants in the tail of a coefficient row (filled on the
// host by oqg_set_gsims from the packed coefficients)
enum { CY_LOG_PHI4 = 0, CY_INV_C9B = 1 };
template <typename CT>
__device__ __forceinline__ double der(const CT &c, int k) {
  return reinterpret_cast<const double *>(&c)[OQG_DER + k];
}

__device__ __forceinline__ Row make_row(const RupIn &r, const SiteIn &s,
                                        const Dist &d) {
  Row w;
  w.mag = r.mag; w.rake = r.rake; w.dip = r.dip; w.ztor = r.ztor;
  w.width = r.width; w.hypo_depth = r.hypo_depth;
  w.rrup = d.rrup; w.rjb = d.rjb; w.rx = d.rx; w.ry0 = d.ry0;
  w.repi = d.repi; w.rhypo = d.rhypo;
  w.vs30 = s.vs30; w.z1pt0 = s.z1pt0; w.z2pt5 = s.z2pt5;
  w.vs30measured = s.vs30measured;
  return w;
}

// ------------------------------------------------------------- generic --
template <int F> struct Fast {
  typedef typename Fam<F>::C C;
  typedef Empty S;
  typedef Empty Q;
  typedef typename Fam<F>::R P;
  static constexpr bool HAS_S = false, HAS_Q = false, EVAL_QP = false;
  static constexpr bool ROLL_Q = false;
  static constexpr bool HAS_P = sizeof(P) >= 8;
  static __device__ __forceinline__ S site(const oqg_gsim &, const C &,
                                           const SiteIn &) { return S(); }
  static __device__ __forceinline__ Q rup(const oqg_gsim &, const C &,
                                          const RupIn &) { return Q(); }
  static __device__ __forceinline__ P pair(const oqg_gsim &g, const C &Cp,
                                           const S &, const Q &,
                                           const RupIn &r, const SiteIn &s,
                                           const Dist &d) {
    return Fam<F>::row(g, Cp, make_row(r, s, d));
  }
  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &Cp, const S &,
                                              const Q &, const Q &, const P &p,
                                              const RupIn &r, const SiteIn &s,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    Out o = Fam<F>::imt(g, c, Cp, make_row(r, s, d), p);
    mean = o.mean;
    sig = o.sig;
    inv = 1. / o.sig;
  }
};

// -------------------------------------------------------------- BSSA14 --
template <> struct Fast<OQG_BSSA14> {
  typedef Bssa14C C;
  struct S { double lin, f2, dphi; };
  struct Q { double fm, c12, tau, phi0, lR1, ilR; };
  struct P { double lnp, lrjb; };
  static constexpr bool HAS_S = true, HAS_Q = true, HAS_P = true;
  static constexpr bool EVAL_QP = false;
  static constexpr bool ROLL_Q = true;   // measured: BSSA14 -2 %

  static __device__ __forceinline__ S site(const oqg_gsim &g, const C &c,
                                           const SiteIn &s) {
    S o;
    const int region = g.flags & 3;
    double flin = s.vs30 > c.Vc ? c.Vc / 760.0 : s.vs30 / 760.0;
    flin = c.c * log(flin);
    double vs = fmin(s.vs30, 760.);
    o.f2 = c.f4 * (exp(c.f5 * (vs - 360.)) - exp(c.f5 * 400.));
    double fbd = 0.;
    if (region != 0 && c.period >= 0.65) {
      double mu = bssa_mu_z1(region, s.vs30);
      double z1 = s.z1pt0 == -999. ? mu * 1000. : s.z1pt0;
      double dz1 = (z1 / 1000.) - mu;
      fbd = dz1 <= c.f7 / c.f6 ? c.f6 * dz1 : c.f7;
    }
    o.lin = flin + fbd;
    double dphi = 0.;
    if (s.vs30 <= 225.) dphi -= c.DfV;
    if (s.vs30 >= 225. && s.vs30 <= 300.)
      dphi -= c.DfV * (log(300. / s.vs30) / log(300. / 225.));
    o.dphi = dphi;
    return o;
  }

  static __device__ __forceinline__ Q rup(const oqg_gsim &g, const C &c,
                                          const RupIn &r) {
    Q q;
    Row w;
    w.mag = r.mag; w.rake = r.rake;
    q.fm = bssa_fm(c, w, (g.flags >> 4) & 1);
    q.c12 = c.c1 + c.c2 * (r.mag - 4.5);
    double tau = c.tau1 + (c.tau2 - c.tau1) * (r.mag - 4.5);
    double phi = c.f1 + (c.f2 - c.f1) * (r.mag - 4.5);
    if (r.mag <= 4.5) { tau = c.tau1; phi = c.f1; }
    if (r.mag >= 5.5) { tau = c.tau2; phi = c.f2; }
    q.tau = tau;
    q.phi0 = phi;
    q.lR1 = log(c.R1);
    q.ilR = 1. / log(c.R2 / c.R1);
    return q;
  }

  static __device__ __forceinline__ P pair(const oqg_gsim &, const C &Cp,
                                           const S &, const Q &qp,
                                           const RupIn &, const SiteIn &,
                                           const Dist &d) {
    P p;
    double rr = sqrt(d.rjb * d.rjb + Cp.h * Cp.h);
    double pga_rock = fast_exp(qp.fm + (qp.c12 * fast_log(rr) +
                                        (Cp.c3 + Cp.Dc3) * (rr - 1.0)));
    p.lnp = fast_log((pga_rock + 0.1) * 10.);
    p.lrjb = log(d.rjb);   // rjb may be 0: the library function
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const Q &, const P &p,
                                              const RupIn &r, const SiteIn &,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    double rr2 = d.rjb * d.rjb + c.h * c.h;
    double rr = sqrt(rr2);
    double fp = q.c12 * (0.5 * fast_log(rr2)) + (c.c3 + c.Dc3) * (rr - 1.0);
    mean = (q.fm + fp) + (s.lin + s.f2 * p.lnp);
    if (g.par[0] != 0.) mean += g.par[0] * epistemic_sigma(d.rrup, r.mag);
    double phi = q.phi0;
    if (d.rjb > c.R2) phi += c.DfR;
    else if (d.rjb > c.R1) phi += c.DfR * ((p.lrjb - q.lR1) * q.ilR);
    phi += s.dphi;
    double var = q.tau * q.tau + phi * phi;
    inv = rsqrt(var);
    sig = var * inv;
  }
};

// ---------------------------------------------------------------- CY14 --
template <> struct Fast<OQG_CY14> {
  typedef Cy14C C;
  struct S { double lin, nls, phis; };
  struct Q { double a, b, gam, hw, tau, mdep; };
  struct P { double fr, hwd, pga_mean; };
  static constexpr bool HAS_S = true, HAS_Q = true, HAS_P = true;
  static constexpr bool EVAL_QP = false;
  static constexpr bool ROLL_Q = true;

  static __device__ __forceinline__ S site(const oqg_gsim &g, const C &c,
                                           const SiteIn &s) {
    const int region = g.flags & 3;
    const double REFVS = 1130.;
    S o;
    double mz1 = exp(cy_mean_z1(region, s.vs30));
    double z1 = s.z1pt0 == -999. ? mz1 : s.z1pt0;
    double dz1 = z1 - mz1;
    if (z1 <= 0.0) dz1 = 0.0;
    double fz1;
    if (region == 1) fz1 = c.phi5jp * (1.0 - exp(-dz1 / 800.0));
    else fz1 = (c.phi5 * (1.0 - exp(-dz1 / 300.0))) + 0.;
    double k1 = region == 1 ? c.phi1jp : c.phi1;
    double flin = k1 * fmin(log(s.vs30 / REFVS), 0.0);
    double vs = fmin(s.vs30, REFVS);
    o.nls = c.phi2 * (exp(c.phi3 * (vs - 360.)) -
                      exp(c.phi3 * (REFVS - 360.)));
    o.lin = flin + fz1;
    o.phis = s.vs30measured ? 0.7 : c.sig3;
    return o;
  }

  static __device__ __forceinline__ Q rup(const oqg_gsim &g, const C &c,
                                          const RupIn &r) {
    const int region = g.flags & 3;
    const double alpha_nm = g.par[1];
    const double C2 = 1.06;
    const double mag = r.mag;
    Q q;
    bool rev = 30. <= r.rake && r.rake <= 150.;
    double mz;
    if (rev) mz = fmax(2.704 - 1.226 * fmax(mag - 5.849, 0.), 0.);
    else mz = fmax(2.673 - 1.136 * fmax(mag - 4.970, 0.), 0.);
    double dz = r.ztor - mz * mz;
    double fm = fast_log(1.0 + fast_exp(c.cn * (c.cm + 0. - mag)));
    fm = (C2 * (mag - 6.0) + ((C2 - c.c3) / c.cn) * fm - (C2 - c.c3) * 0.);
    double coshm = fast_cosh(2.0 * fmax(mag - 4.5, 0.));
    double fsrc = 0.;
    if (rev) fsrc += c.c1a + (c.c1c / coshm);
    if (-120. <= r.rake && r.rake <= -60.)
      fsrc += (c.c1b + (c.c1d / coshm)) * alpha_nm;
    fsrc += (c.c7 + (c.c7b / coshm)) * dz;
    double cd = cos(r.dip * OQG_D2R);
    fsrc += (0.0 + (c.c11b / coshm)) * (cd * cd);
    q.a = c.c1 + fm + fsrc;
    q.b = c.c5 * fast_cosh(c.c6 * fmax(mag - c.chm, 0.0));
    double gamma = c.cg1 + c.cg2 / fast_cosh(fmax(mag - c.cg3, 0.0));
    if (region == 1 || region == 2) {
      if (mag > 6.0 && mag < 6.9) gamma = gamma * c.gjpit;
    } else if (region == 3) {
      gamma = gamma * c.gwn;
    }
    q.gam = gamma;
    q.hw = c.c9 * cd;
    double mclip = fmin(fmax(mag - 5.0, 0.), 1.5);
    q.tau = c.tau1 + (c.tau2 - c.tau1) / 1.5 * mclip;
    q.mdep = c.sig1 + (c.sig2 - c.sig1) * mclip / 1.5;
    return q;
  }

  // mean (before floor and epistemic term) and sigma for one IMT; not
  // inlined: shared by the pair stage (PGA) and the IMT stage to keep the
  // kernel within the instruction cache
  static __device__ OQG_CY_INLINE void one(const oqg_gsim &g, const C &c,
                                             const S &s, const Q &q,
                                             double fr, double hwd,
                                             const Dist &d, double &mean,
                                             double &sig, double &inv) {
    const double C4 = -2.1;
    double ln_yref = q.a + C4 * fast_log(d.rrup + q.b) +
                     (fr + q.gam * d.rrup) + 0.;
    if (((g.flags >> 4) & 1) && d.rx >= 0.0) {
      double fd = hwd * (c.c9a + (1.0 - c.c9a) *
                                     fast_tanh_pos(d.rx * der(c, CY_INV_C9B)));
      ln_yref += q.hw * fd;
    }
    double yref = fast_exp(ln_yref);
    double yp = yref + c.phi4;
    double fnl = (fast_log(yp) - der(c, CY_LOG_PHI4)) * s.nls;
    mean = ln_yref + fnl + s.lin;
    if ((g.flags >> 5) & 1) { sig = 0.65; inv = 1. / 0.65; return; }  // peer
    double nl0 = s.nls * (yref * fast_rcp(yp));
    double nl1 = (1.0 + nl0) * (1.0 + nl0);
    double phi2 = (q.mdep * q.mdep) * (s.phis + nl1);
    double var = nl1 * (q.tau * q.tau) + phi2;
    inv = rsqrt(var);
    sig = var * inv;
  }

  static __device__ __forceinline__ P pair(const oqg_gsim &g, const C &Cp,
                                           const S &sp, const Q &qp,
                                           const RupIn &r, const SiteIn &,
                                           const Dist &d) {
    const double C4 = -2.1, C4A = -0.5, CRB = 50.0;
    P p;
    p.fr = (C4A - C4) * (0.5 * fast_log(d.rrup * d.rrup + CRB * CRB));
    p.hwd = 1.0 - (sqrt(d.rjb * d.rjb + r.ztor * r.ztor) / (d.rrup + 1.0));
    double sig, inv;
    one(g, Cp, sp, qp, p.fr, p.hwd, d, p.pga_mean, sig, inv);
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const Q &, const P &p,
                                              const RupIn &r, const SiteIn &,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    one(g, c, s, q, p.fr, p.hwd, d, mean, sig, inv);
    if (c.period > 0. && c.period <= 0.3 && mean < p.pga_mean)
      mean = p.pga_mean;
    if (g.par[0] != 0.) mean = mean + g.par[0] * epistemic_sigma(d.rrup, r.mag);
  }
};

// --------------------------------------------------------------- ASK14 --
// abrahamson_2014.py:62-438 split by dependency.  Rearrangements with
// respect to oqg_gmpe.cuh (all exact identities, ~1e-16 relative):
//   b log(sa + c rat^1.5) - b log(sa + c) = b log((sa + c rat^1.5)/(sa + c))
//   sig^2 = (phi_al^2 - phi_amp^2 + tau_al^2) (1 + der)^2 + phi_amp^2
// and the rock spectral acceleration sa1180 is only evaluated for the sites
// with vs30 < vlin, the only ones whose mean or sigma depend on it.
template <> struct Fast<OQG_ASK14> {
  typedef Ask14C C;
  struct S { double lin, crat, add, s1180, rcoef; };
  struct Q { double a0, klr, c4m2, sofz, hwq, r1, tau, pm, pe; };
  struct P { double hwp; };
  static constexpr bool HAS_S = true, HAS_Q = true, HAS_P = true;
  static constexpr bool EVAL_QP = false;
  static constexpr bool ROLL_Q = false;  // measured: ASK14 +6 % when rolled

  static __device__ __forceinline__ S site(const oqg_gsim &g, const C &c,
                                           const SiteIn &s) {
    const int region = g.flags & 3;
    S o;
    double vstar = s.vs30 >= c.v1 ? c.v1 : s.vs30;
    double rat = vstar / c.vlin;
    if (s.vs30 >= c.vlin) {
      o.lin = (c.a10 + c.b * 1.5) * log(rat);
      o.crat = -1.;                       // linear site: no sa1180 needed
    } else {
      o.lin = c.a10 * log(rat);
      o.crat = c.c * pow(rat, 1.5);
    }
    Row w;
    w.vs30 = s.vs30; w.z1pt0 = s.z1pt0;
    // the rrup-independent part of the regional term, for the site and for
    // the 1180 m/s rock pass
    o.add = ask_basin_soil(c, w, region) + ask_regional(c, region, s.vs30, 0.);
    o.s1180 = ask_site(c, 1180., 0.) + ask_regional(c, region, 1180., 0.);
    o.rcoef = region == 2 ? c.a25 : (region == 3 ? c.a28
                                                 : (region == 1 ? c.a29 : 0.));
    return o;
  }

  static __device__ __forceinline__ Q rup(const oqg_gsim &, const C &c,
                                          const RupIn &r) {
    Q q;
    const double mag = r.mag, m1 = c.m1, m2 = 5.0;
    double c4m = c.c4 - (c.c4 - 1.) * (5. - mag);
    if (mag > 5.) c4m = c.c4;
    if (mag < 4.) c4m = 1.;
    q.c4m2 = c4m * c4m;
    double t;
    if (mag >= m1) {
      t = c.a5 * (mag - m1) + c.a8 * ((8.5 - mag) * (8.5 - mag));
      q.klr = c.a2 + c.a3 * (mag - m1);
    } else if (mag >= m2) {
      t = c.a4 * (mag - m1) + c.a8 * ((8.5 - mag) * (8.5 - mag));
      q.klr = c.a2 + c.a3 * (mag - m1);
    } else {
      t = c.a4 * (m2 - m1) + c.a8 * ((8.5 - m2) * (8.5 - m2)) +
          c.a6 * (mag - m2) + c.a7 * ((mag - m2) * (mag - m2));
      q.klr = c.a2 + c.a3 * (m2 - m1);
    }
    q.a0 = c.a1 * 1.0 + t;
    Row w;
    w.mag = mag; w.rake = r.rake;
    q.sofz = ask_sof(c, w) + c.a15 * fmin(r.ztor / 20.0, 1.);
    // rupture-only factors of the hanging-wall term :201-257
    const double dip = r.dip;
    double t1 = dip <= 30. ? 60. / 45. : (90. - dip) / 45.0;
    double m65 = mag - 6.5;
    double t2 = (1. + 0.2 * m65 - (1. - 0.2) * (m65 * m65));
    if (mag > 6.5) t2 = 1. + 0.2 * m65;
    if (mag <= 5.5) t2 = 0.;
    double t4 = r.ztor <= 10. ? 1. - (r.ztor * r.ztor) / 100. : 0.;
    q.hwq = dip == 90. ? 0. : c.a13 * t1 * t2 * t4;
    q.r1 = r.width * cos(dip * OQG_D2R);
    // aleatory terms :61-90, 180-199
    double tau_al = c.s3 + (c.s4 - c.s3) / 2. * (mag - 5.);
    if (mag < 5.) tau_al = c.s3;
    if (mag >= 7.) tau_al = c.s4;
    q.tau = tau_al;
    double pm = c.s1m + (c.s2m - c.s1m) / 2. * (mag - 4.);
    double pe = c.s1e + (c.s2e - c.s1e) / 2. * (mag - 4.);
    if (mag < 4.) { pm = c.s1m; pe = c.s1e; }
    if (mag >= 6.) { pm = c.s2m; pe = c.s2e; }
    q.pm = pm; q.pe = pe;
    return q;
  }

  // distance tapers of the hanging-wall term (T3 and T5), IMT independent
  static __device__ __forceinline__ P pair(const oqg_gsim &, const C &,
                                           const S &, const Q &qp,
                                           const RupIn &, const SiteIn &,
                                           const Dist &d) {
    P p;
    const double rx = d.rx, r1 = qp.r1, r2 = 3. * r1;
    double t3;
    if (rx < r1) {
      double q = rx / r1;
      t3 = 0.25 + 1.50 * q + (-0.75) * (q * q);
    } else if (rx <= r2) {
      t3 = 1. - (rx - r1) / (r2 - r1);
    } else {
      t3 = 0.;
    }
    double ry1 = rx * tan(20. * OQG_D2R);
    double dd = d.ry0 - ry1;
    double t5 = dd <= 0.0 ? 1. : (dd < 5.0 ? 1. - dd / 5.0 : 0.);
    p.hwp = rx > 0. ? t3 * t5 : 0.;
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const Q &, const P &p,
                                              const RupIn &r, const SiteIn &si,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    const int region = g.flags & 3;
    const double rrup = d.rrup;
    double basic = (q.a0 + c.a17 * rrup) +
                   q.klr * (0.5 * fast_log(rrup * rrup + q.c4m2));
    double common = basic + q.sofz + q.hwq * p.hwp;
    double rr = s.rcoef * rrup;
    double site = s.lin, der = 0.;
    if (s.crat >= 0.) {
      double sa1180 = fast_exp(common + s.s1180 + rr);
      double t1 = sa1180 + s.crat, t2 = sa1180 + c.c;
      double w = fast_rcp(t1 * t2);
      site += c.b * fast_log((t1 * t1) * w);
      der = c.b * sa1180 * ((t2 - t1) * w);
    }
    mean = common + site + s.add + rr;
    if (g.par[0] != 0.) mean = mean + g.par[0] * epistemic_sigma(rrup, r.mag);
    double phi_al;
    if (region == 1)
      phi_al = rrup < 30. ? c.s5
                          : (rrup <= 80. ? c.s5 + (c.s6 - c.s5) / 50. * (rrup - 30.)
                                         : c.s6);
    else
      phi_al = si.vs30measured ? q.pm : q.pe;
    double phi_amp = phi_al < 0.4 ? 0.99 * phi_al : 0.4;
    double pa2 = phi_amp * phi_amp;
    double d1 = 1. + der;
    double var = ((phi_al * phi_al - pa2) + q.tau * q.tau) * (d1 * d1) + pa2;
    inv = rsqrt(var);
    sig = var * inv;
  }
};

// ---------------------------------------------------------------- CB14 --
// campbell_bozorgnia_2014.py:54-463 split by dependency.  Rearrangements
// (exact identities): log(sqrt(x)) = log(x) / 2;
//   k2 (log(a + c v^n) - log(a + c)) = k2 log((a + c v^n) / (a + c));
//   sig^2 = tau^2 + phi^2 expanded, no intermediate square roots.
template <> struct Fast<OQG_CB14> {
  typedef Cb14C C;
  struct S { double lin, cvn, rock; };
  struct Q { double a0, kdis, hngq, r1, r2, taub, phib, varb; };
  // the hanging-wall coordinates hx, hrr of a pair are two divisions: they
  // are recomputed per IMT instead of being shared through shared memory
  // (4 KB per field and sub-CTA), so that 4 sub-CTAs fit on an SM (3 with
  // them: CB14 146 -> 135 ms on the cfg4 sample)
  struct P { double a1100, pga_mean; };
  struct H { double hx, hrr; };
  static constexpr bool HAS_S = true, HAS_Q = true, HAS_P = true;
  static constexpr bool EVAL_QP = true;
  static constexpr bool ROLL_Q = true;

  static __device__ __forceinline__ H hang(const Q &q, const Dist &d) {
    H h;
    if (d.rx >= q.r1) h.hx = (d.rx - q.r1) * fast_rcp(q.r2 - q.r1);
    else h.hx = d.rx * fast_rcp(q.r1);
    h.hrr = d.rrup > 0.0 ? (d.rrup - d.rjb) * fast_rcp(d.rrup) : 1.;
    return h;
  }

  static __device__ __forceinline__ double site_lin(const C &c, int SJ,
                                                    double vs30, bool rock,
                                                    double z2pt5) {
    const double CB_N = 1.18;
    double lv = log(vs30 / c.k1);
    double fsite = c.c11 * lv;
    if (vs30 > c.k1) fsite = fsite + (c.k2 * CB_N * lv);
    if (SJ) {
      double fj = (c.c13 + c.k2 * CB_N) * lv;
      double soft = (c.c12 + c.k2 * CB_N) * (lv - log(200.0 / c.k1));
      fsite = fsite + (vs30 <= 200.0 ? fj + soft : fj);
    }
    double z25;
    if (rock) z25 = cb_z25ref(SJ, 1100.0);
    else z25 = z2pt5 == -999. ? cb_z25ref(SJ, vs30) : z2pt5;
    double fsed = 0.;
    if (z25 < 1.0) fsed = (c.c14 + c.c15 * SJ) * (z25 - 1.0);
    if (z25 > 3.0)
      fsed = (c.c16 * c.k3) * exp(-0.75) * (1. - exp(-0.25 * (z25 - 3.)));
    return fsite + fsed;
  }

  static __device__ __forceinline__ S site(const oqg_gsim &g, const C &c,
                                           const SiteIn &s) {
    const int SJ = g.flags & 1;
    S o;
    o.lin = site_lin(c, SJ, s.vs30, false, s.z2pt5);
    o.rock = site_lin(c, SJ, 1100., true, 0.);
    o.cvn = s.vs30 > c.k1 ? -1. : 1.88 * pow(s.vs30 / c.k1, 1.18);
    return o;
  }

  static __device__ __forceinline__ Q rup(const oqg_gsim &, const C &c,
                                          const RupIn &r) {
    Q q;
    const double mag = r.mag;
    double fmag = c.c0 + c.c1 * mag, t;
    if (mag > 6.5)
      t = c.c2 * (mag - 4.5) + c.c3 * (mag - 5.5) + c.c4 * (mag - 6.5);
    else if (mag > 5.5) t = c.c2 * (mag - 4.5) + c.c3 * (mag - 5.5);
    else if (mag > 4.5) t = c.c2 * (mag - 4.5);
    else t = 0.;
    fmag = fmag + t;
    double frv = (r.rake > 30. && r.rake < 150.) ? 1. : 0.;
    double fnm = (r.rake > -150. && r.rake < -30.) ? 1. : 0.;
    double fflt_m = mag <= 4.5 ? 0. : (mag > 5.5 ? 1. : mag - 4.5);
    double fflt = (c.c8 * frv + c.c9 * fnm) * fflt_m;
    double fh = fmin(fmax(r.hypo_depth - 7.0, 0.), 13.);
    double fm = c.c17 + (c.c18 - c.c17) * (mag - 5.5);
    if (mag <= 5.5) fm = c.c17;
    if (mag > 6.5) fm = c.c18;
    double fdip = c.c19 * (5.5 - mag) * r.dip;
    if (mag < 4.5) fdip = c.c19 * r.dip;
    if (mag > 5.5) fdip = 0.0;
    q.a0 = fmag + fflt + fh * fm + fdip;
    q.kdis = 0.5 * (c.c5 + c.c6 * mag);
    double fhng_m = (mag - 5.5) * (1.0 + c.a2 * (mag - 6.5));
    if (mag < 5.5) fhng_m = 0.0;
    if (mag > 6.5) fhng_m = 1.0 + c.a2 * (mag - 6.5);
    double fhng_z = r.ztor > 16.66 ? 0. : 1. - 0.06 * r.ztor;
    double fhng_d = (90.0 - r.dip) / 45.0;
    q.hngq = c.c10 * fhng_m * fhng_z * fhng_d;
    q.r1 = r.width * cos(r.dip * OQG_D2R);
    q.r2 = 62.0 * mag - 350.0;
    q.taub = cb_lin(mag, c.tau1, c.tau2);
    double pl = cb_lin(mag, c.phi1, c.phi2);
    q.phib = sqrt(pl * pl - c.philnAF * c.philnAF);
    q.varb = q.taub * q.taub + pl * pl;
    return q;
  }

  // the part of the mean without the nonlinear soil term
  static __device__ __forceinline__ double base(const C &c, const Q &q,
                                                const H &p, const Dist &d) {
    const double rrup = d.rrup;
    double fdis = q.kdis * fast_log(rrup * rrup + c.c7 * c.c7);
    double fhng_rx;
    if (d.rx >= q.r1) fhng_rx = fmax(1.0 + c.h5 * p.hx + c.h6 * (p.hx * p.hx), 0.0);
    else if (d.rx >= 0.) fhng_rx = c.h1 + c.h2 * p.hx + c.h3 * (p.hx * p.hx);
    else fhng_rx = 0.;
    double fatn = rrup >= 80.0 ? (c.c20 + c.Dc20) * (rrup - 80.0) : 0.;
    return (q.a0 + fdis) + (q.hngq * fhng_rx) * p.hrr + fatn;
  }

  static __device__ __forceinline__ P pair(const oqg_gsim &, const C &Cp,
                                           const S &sp, const Q &qp,
                                           const RupIn &, const SiteIn &,
                                           const Dist &d) {
    P p;
    double b = base(Cp, qp, hang(qp, d), d);
    p.a1100 = fast_exp(b + sp.rock);
    double nl = 0.;
    if (sp.cvn >= 0.)
      nl = Cp.k2 * fast_log((p.a1100 + sp.cvn) * fast_rcp(p.a1100 + 1.88));
    p.pga_mean = b + sp.lin + nl;
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const Q &qp,
                                              const P &p, const RupIn &r,
                                              const SiteIn &, const Dist &d,
                                              double &mean, double &sig,
                                              double &inv) {
    double eps = 0.;
    if (g.par[0] != 0.) eps = g.par[0] * epistemic_sigma(d.rrup, r.mag);
    double nl = 0., alpha = 0.;
    if (s.cvn >= 0.) {
      double af1 = p.a1100 + s.cvn, af2 = p.a1100 + 1.88;
      double w = fast_rcp(af1 * af2);
      nl = c.k2 * fast_log((af1 * af1) * w);
      alpha = c.k2 * p.a1100 * ((af2 - af1) * w);
    }
    mean = base(c, q, hang(q, d), d) + s.lin + nl;
    mean = mean + eps;
    if (c.period > 0. && c.period < 0.25) {  // :554-564
      if (mean <= p.pga_mean) mean = p.pga_mean;
      mean = mean + eps;  // added a second time in the reference
    }
    double var = q.varb +
                 (alpha * alpha) * (qp.taub * qp.taub + qp.phib * qp.phib) +
                 (2.0 * alpha * c.rholny) * (q.taub * qp.taub + q.phib * qp.phib);
    inv = rsqrt(var);
    sig = var * inv;
  }
};

}  // namespace oqg
