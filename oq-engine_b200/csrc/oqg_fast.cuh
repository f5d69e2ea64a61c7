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
// association of a few additions; results agree to ~1e-15.  Families
// without a specialisation go through the generic Fast<F>, which calls the
// unstaged functions (P = the row state of oqg_gmpe.cuh).
#pragma once
#include "oqg_gmpe.cuh"
#include "oqg_geo.cuh"

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
  static constexpr bool HAS_S = false, HAS_Q = false;
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
                                              const Q &, const P &p,
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
    double pga_rock = exp(qp.fm + (qp.c12 * log(rr / 1.0) +
                                   (Cp.c3 + Cp.Dc3) * (rr - 1.0)));
    p.lnp = log((pga_rock + 0.1) / 0.1);
    p.lrjb = log(d.rjb);
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const P &p,
                                              const RupIn &r, const SiteIn &,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    double rr = sqrt(d.rjb * d.rjb + c.h * c.h);
    double fp = q.c12 * log(rr / 1.0) + (c.c3 + c.Dc3) * (rr - 1.0);
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
    double fm = log(1.0 + exp(c.cn * (c.cm + 0. - mag)));
    fm = (C2 * (mag - 6.0) + ((C2 - c.c3) / c.cn) * fm - (C2 - c.c3) * 0.);
    double coshm = cosh(2.0 * fmax(mag - 4.5, 0.));
    double fsrc = 0.;
    if (rev) fsrc += c.c1a + (c.c1c / coshm);
    if (-120. <= r.rake && r.rake <= -60.)
      fsrc += (c.c1b + (c.c1d / coshm)) * alpha_nm;
    fsrc += (c.c7 + (c.c7b / coshm)) * dz;
    double cd = cos(r.dip * OQG_D2R);
    fsrc += (0.0 + (c.c11b / coshm)) * (cd * cd);
    q.a = c.c1 + fm + fsrc;
    q.b = c.c5 * cosh(c.c6 * fmax(mag - c.chm, 0.0));
    double gamma = c.cg1 + c.cg2 / cosh(fmax(mag - c.cg3, 0.0));
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
    double ln_yref = q.a + C4 * log(d.rrup + q.b) + (fr + q.gam * d.rrup) + 0.;
    if (((g.flags >> 4) & 1) && d.rx >= 0.0) {
      double fd = hwd * (c.c9a + (1.0 - c.c9a) * tanh(d.rx / c.c9b));
      ln_yref += q.hw * fd;
    }
    double yref = exp(ln_yref);
    double fnl = log((yref + c.phi4) / c.phi4) * s.nls;
    mean = ln_yref + fnl + s.lin;
    if ((g.flags >> 5) & 1) { sig = 0.65; inv = 1. / 0.65; return; }  // peer
    double nl0 = s.nls * (yref / (yref + c.phi4));
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
    p.fr = (C4A - C4) * log(sqrt(d.rrup * d.rrup + CRB * CRB));
    p.hwd = 1.0 - (sqrt(d.rjb * d.rjb + r.ztor * r.ztor) / (d.rrup + 1.0));
    double sig, inv;
    one(g, Cp, sp, qp, p.fr, p.hwd, d, p.pga_mean, sig, inv);
    return p;
  }

  static __device__ __forceinline__ void eval(const oqg_gsim &g, const C &c,
                                              const C &, const S &s,
                                              const Q &q, const P &p,
                                              const RupIn &r, const SiteIn &,
                                              const Dist &d, double &mean,
                                              double &sig, double &inv) {
    one(g, c, s, q, p.fr, p.hwd, d, mean, sig, inv);
    if (c.period > 0. && c.period <= 0.3 && mean < p.pga_mean)
      mean = p.pga_mean;
    mean = mean + g.par[0] * epistemic_sigma(d.rrup, r.mag);
  }
};

}  // namespace oqg
