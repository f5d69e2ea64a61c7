"""
ORACLE (test infrastructure, never imported by the product): numpy
restatement of the ground-motion models on the classical-PSHA hot path.

Each ``<model>(ctx, imts, **kw)`` returns four (M, N) float64 arrays
``mean, sig, tau, phi`` exactly like ``GMPE.compute`` fills them in the
reference (openquake/hazardlib/gsim/base.py:389-401).  ``ctx`` is anything
with array attributes (a numpy.recarray works), ``imts`` a list of strings
('PGA', 'PGV', 'SA(0.2)').  The arithmetic follows, term by term and in the
same order of operations, the reference modules cited in each docstring so
that agreement with the reference is at the 1e-15 level; the pin is
tests/golden/gmpe_*.npz (produced by oracle/make_golden.py running the
reference itself) plus the reference's own CSV verification tables
(tests/golden/kat_*.csv).

Coefficients come from oq_engine_b200/data/coeffs.json (published tables,
extracted by tools/extract_coeffs.py); interpolation between tabulated
periods is linear in log-period as in
openquake/hazardlib/gsim/coeffs_table.py:231-286.
"""
import os
import re
import json
import math
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..',
                     'oq_engine_b200', 'data', 'coeffs.json')
_TABLES = None


# --------------------------------------------------------------------- #
#                          coefficient look-up                           #
# --------------------------------------------------------------------- #

def parse_imt(s):
    """'PGA' -> ('PGA', 0.0); 'SA(0.2)' -> ('SA', 0.2)"""
    mo = re.fullmatch(r'SA\(([^)]+)\)', s)
    if mo:
        return 'SA', float(mo.group(1))
    return s, 0.0


def coeffs(table, imt):
    """
    Row of `table` for `imt` as a dict; SA periods not tabulated are
    interpolated in log-period between the neighbouring rows, outside the
    tabulated range a KeyError is raised (coeffs_table.py:231-277).
    """
    global _TABLES
    if _TABLES is None:
        with open(_DATA) as f:
            _TABLES = json.load(f)
    tab = _TABLES[table]
    cols = tab['columns']
    kind, period = parse_imt(imt)
    if kind != 'SA':
        return dict(zip(cols, tab['non_sa'][imt]))
    lo = hi = None
    for row in tab['sa']:
        if row[0] == period:
            return dict(zip(cols, row[1:]))
        if row[0] < period:
            lo = row
        elif hi is None:
            hi = row
    if lo is None or hi is None:
        raise KeyError(imt)
    w = ((math.log(period) - math.log(lo[0])) /
         (math.log(hi[0]) - math.log(lo[0])))
    return {c: (b - a) * w + a for c, a, b in zip(cols, lo[1:], hi[1:])}


def epistemic_sigma(ctx):
    """USGS-2014 style epistemic sigma, abrahamson_2014.py:440-464"""
    n = 2
    r, mag = ctx.rrup, ctx.mag
    near, mid = r <= 10, (r > 10) & (r < 30)

    def pick(a, b, c):
        return np.where(near, 0.4 * np.sqrt(n / a),
                        np.where(mid, 0.4 * np.sqrt(n / b),
                                 0.4 * np.sqrt(n / c)))
    return np.where((mag >= 5) & (mag < 6), pick(11, 38, 94),
                    np.where((mag >= 6) & (mag < 7), pick(2, 7, 13),
                             pick(2, 2, 4)))


def _out(M, N):
    return np.zeros((4, M, N))


# --------------------------------------------------------------------- #
#                 Boore, Stewart, Seyhan, Atkinson (2014)                #
# --------------------------------------------------------------------- #

def _bssa_sof(C, rake):
    # boore_2014.py:269-287
    out = np.full_like(rake, C['e2'])
    out[(np.abs(rake) <= 30.) | ((180. - np.abs(rake)) <= 30.)] = C['e1']
    out[(rake > 30.) & (rake < 150.)] = C['e3']
    return out


def _bssa_fm(C, ctx, sof):
    # boore_2014.py:169-180
    dm = ctx.mag - C['Mh']
    fm = np.where(ctx.mag <= C['Mh'],
                  C['e4'] * dm + C['e5'] * dm ** 2.0, C['e6'] * dm)
    if not sof:
        return C['e0'] + fm
    return _bssa_sof(C, ctx.rake) + fm


def _bssa_fp(C, ctx):
    # boore_2014.py:198-207 (Mref=4.5, Rref=1)
    r = np.sqrt(ctx.rjb ** 2.0 + C['h'] ** 2.0)
    return ((C['c1'] + C['c2'] * (ctx.mag - 4.5)) * np.log(r / 1.0) +
            (C['c3'] + C['Dc3']) * (r - 1.0))


def _bssa_mu_z1(region, vs30):
    # boore_2014.py:711-730
    if region == 'JPN':
        k = 412.39 ** 2.0
        ln = -5.23 / 2.0 * np.log(
            ((vs30 ** 2.0) + k) / ((1360.0 ** 2.0) + k)) - np.log(1000.)
    else:
        k = 570.94 ** 4.0
        ln = -7.15 / 4.0 * np.log(
            (vs30 ** 4.0 + k) / (1360.0 ** 4.0 + k)) - np.log(1000.)
    return np.exp(ln)


def _bssa_basin(C, ctx, region, period):
    # boore_2014.py:61-104 without the USGS / CyberShake options
    if region == 'nobasin' or period < 0.65:
        return np.zeros(len(ctx.vs30))
    mu = _bssa_mu_z1(region, ctx.vs30)
    z1 = np.array(ctx.z1pt0, dtype=float)
    unknown = z1 == -999
    z1[unknown] = mu[unknown] * 1000.
    dz1 = (z1 / 1000.) - mu
    return np.where(dz1 <= C['f7'] / C['f6'], C['f6'] * dz1, C['f7'])


def _bssa_site(C, ctx, pga_rock, region, period):
    # boore_2014.py:160-166, 183-192, 244-256
    flin = ctx.vs30 / 760.0
    flin = np.where(ctx.vs30 > C['Vc'], C['Vc'] / 760.0, flin)
    flin = C['c'] * np.log(flin)
    vs = np.minimum(ctx.vs30, 760.)
    f2 = C['f4'] * (np.exp(C['f5'] * (vs - 360.)) - np.exp(C['f5'] * 400.))
    fnl = 0.0 + f2 * np.log((pga_rock + 0.1) / 0.1)
    return flin + fnl + _bssa_basin(C, ctx, region, period)


def _bssa_std(C, ctx):
    # boore_2014.py:107-145, 259-266
    mag, rjb, vs30 = ctx.mag, ctx.rjb, ctx.vs30
    tau = np.full_like(mag, C['tau1'] + (C['tau2'] - C['tau1']) * (mag - 4.5))
    tau[mag <= 4.5] = C['tau1']
    tau[mag >= 5.5] = C['tau2']
    phi = np.full_like(rjb, C['f1'] + (C['f2'] - C['f1']) * (mag - 4.5))
    phi[mag <= 4.5] = C['f1']
    phi[mag >= 5.5] = C['f2']
    far = rjb > C['R2']
    phi[far] += C['DfR']
    mid = (rjb > C['R1']) & (rjb <= C['R2'])
    phi[mid] += C['DfR'] * (np.log(rjb[mid] / C['R1']) /
                            np.log(C['R2'] / C['R1']))
    soft = vs30 <= 225.
    phi[soft] -= C['DfV']
    ramp = (vs30 >= 225.) & (vs30 <= 300.)
    phi[ramp] -= C['DfV'] * (np.log(300. / vs30[ramp]) / np.log(300. / 225.))
    return np.sqrt(tau ** 2 + phi ** 2), tau, phi


def bssa14(ctx, imts, region='nobasin', sof=True, sigma_mu_epsilon=0.0,
           table='BSSA14'):
    """BooreEtAl2014.compute, boore_2014.py:335-359"""
    out = _out(len(imts), len(ctx.mag))
    Cp = coeffs(table, 'PGA')
    for m, imt in enumerate(imts):
        C = coeffs(table, imt)
        pga_rock = np.exp(_bssa_fm(Cp, ctx, sof) + _bssa_fp(Cp, ctx))
        out[0, m] = (_bssa_fm(C, ctx, sof) + _bssa_fp(C, ctx) +
                     _bssa_site(C, ctx, pga_rock, region, parse_imt(imt)[1]))
        if sigma_mu_epsilon:
            out[0, m] += sigma_mu_epsilon * epistemic_sigma(ctx)
        out[1, m], out[2, m], out[3, m] = _bssa_std(C, ctx)
    return out


# --------------------------------------------------------------------- #
#                         Chiou & Youngs (2014)                          #
# --------------------------------------------------------------------- #

_CY_C2, _CY_C4, _CY_C4A, _CY_CRB = 1.06, -2.1, -0.5, 50.0
_CY_REFVS = 1130.


def _cy_mean_z1(region, vs30):
    # chiou_youngs_2014.py:54-69
    if region == 'JPN':
        return (-5.23 / 2.) * np.log(
            ((vs30 ** 2.) + 412.39 ** 2.) / (1360 ** 2. + 412.39 ** 2.))
    return (-7.15 / 4.) * np.log(
        ((vs30) ** 4. + 570.94 ** 4.) / (1360 ** 4. + 570.94 ** 4.))


def _cy_delta_ztor(ctx):
    # chiou_youngs_2014.py:87-99
    mz = np.clip(2.673 - 1.136 * np.clip(ctx.mag - 4.970, 0., None),
                 0., None) ** 2
    rev = (30. <= ctx.rake) & (ctx.rake <= 150.)
    mz[rev] = np.clip(2.704 - 1.226 * np.clip(
        ctx.mag[rev] - 5.849, 0.0, None), 0., None) ** 2
    return ctx.ztor - mz


def _cy_ln_yref(region, C, ctx, use_hw, alpha_nm):
    # get_ln_y_ref chiou_youngs_2014.py:436-486 with delta_cm = delta_g = 0,
    # no directivity (no rcdpp in the context) and add_delta_c1 False
    mag, rrup = ctx.mag, ctx.rrup
    dz = _cy_delta_ztor(ctx)
    # magnitude scaling, :489-498
    fm = np.log(1.0 + np.exp(C['cn'] * (C['cm'] + 0 - mag)))
    fm = (_CY_C2 * (mag - 6.0) + ((_CY_C2 - C['c3']) / C['cn']) * fm -
          (_CY_C2 - C['c3']) * 0)
    # source terms, :534-557
    coshm = np.cosh(2.0 * np.clip(mag - 4.5, 0., None))
    fsrc = np.zeros_like(mag)
    pos = (30 <= ctx.rake) & (ctx.rake <= 150)
    neg = (-120 <= ctx.rake) & (ctx.rake <= -60)
    fsrc[pos] += C['c1a'] + (C['c1c'] / coshm[pos])
    fsrc[neg] += (C['c1b'] + (C['c1d'] / coshm[neg])) * alpha_nm
    fsrc += (C['c7'] + (C['c7b'] / coshm)) * dz
    fsrc += ((0.0 + (C['c11b'] / coshm)) *
             np.cos(np.radians(ctx.dip)) ** 2.0)
    # near-field geometric spreading, :331-338
    fgeo = _CY_C4 * np.log(
        rrup + C['c5'] * np.cosh(C['c6'] * np.clip(mag - C['chm'], 0.0, None)))
    # far field, :252-328
    fr = (_CY_C4A - _CY_C4) * np.log(np.sqrt(rrup ** 2. + _CY_CRB ** 2.))
    gamma = C['cg1'] + C['cg2'] / np.cosh(np.clip(mag - C['cg3'], 0.0, None))
    frm = (gamma + 0) * rrup
    if region in ('JPN', 'ITA'):
        frm = np.where((mag > 6.0) & (mag < 6.9), frm * C['gjpit'], frm)
        ffar = fr + frm
    elif region == 'WEN':
        ffar = fr + (frm * C['gwn'])
    else:
        ffar = fr + frm
    out = C['c1'] + fm + fsrc + fgeo + ffar + 0.
    if use_hw:  # :341-353
        fhw = np.zeros(rrup.shape)
        hw = ctx.rx >= 0.0
        if hw.any():
            fd = 1.0 - (np.sqrt(ctx.rjb[hw] ** 2. + ctx.ztor[hw] ** 2.) /
                        (rrup[hw] + 1.0))
            fd *= C['c9a'] + (1.0 - C['c9a']) * np.tanh(ctx.rx[hw] / C['c9b'])
            fhw[hw] += C['c9'] * np.cos(np.radians(ctx.dip[hw])) * fd
        out += fhw
    return out


def _cy_basin(C, ctx, region):
    # chiou_youngs_2014.py:175-213 (no USGS / CyberShake)
    z1 = np.array(ctx.z1pt0, dtype=float)
    unknown = z1 == -999
    z1[unknown] = np.exp(_cy_mean_z1(region, ctx.vs30[unknown]))
    dz1 = z1 - np.exp(_cy_mean_z1(region, ctx.vs30))
    dz1[z1 <= 0.0] = 0.0
    if region == 'JPN':
        return C['phi5jp'] * (1.0 - np.exp(-dz1 / 800.0))
    return (C['phi5'] * (1.0 - np.exp(-dz1 / 300.0))) + 0.


def _cy_one(region, C, ctx, use_hw, alpha_nm, peer):
    # get_mean_stddevs chiou_youngs_2014.py:595-635
    ln_yref = _cy_ln_yref(region, C, ctx, use_hw, alpha_nm)
    yref = np.exp(ln_yref)
    fz1 = _cy_basin(C, ctx, region)
    k1 = C['phi1jp'] if region == 'JPN' else C['phi1']
    flin = k1 * np.clip(np.log(ctx.vs30 / _CY_REFVS), -np.inf, 0.0)
    vs = np.clip(ctx.vs30, -np.inf, _CY_REFVS)
    fnl_scale = C['phi2'] * (np.exp(C['phi3'] * (vs - 360.)) -
                             np.exp(C['phi3'] * (_CY_REFVS - 360.)))
    fnl = np.log((yref + C['phi4']) / C['phi4']) * fnl_scale
    mean = ln_yref + flin + fnl + fz1
    if peer:
        return mean, 0.65 * np.ones_like(ctx.vs30), 0, 0
    # stddevs, :560-592 and :515-531
    nl0 = fnl_scale * (yref / (yref + C['phi4']))
    tau = C['tau1'] + (C['tau2'] - C['tau1']) / 1.5 * np.clip(
        ctx.mag - 5.0, 0., 1.5)
    phi = C['sig3'] * np.ones(ctx.vs30.shape)
    phi[np.asarray(ctx.vs30measured, bool)] = 0.7
    phi = np.sqrt(phi + ((1.0 + nl0) ** 2.))
    mdep = C['sig1'] + (C['sig2'] - C['sig1']) * np.clip(
        ctx.mag - 5., 0., 1.5) / 1.5
    phi = mdep * phi
    sig = np.sqrt(((1.0 + nl0) ** 2.) * (tau ** 2.) + phi ** 2.)
    return mean, sig, np.abs((1 + nl0) * tau), phi


def cy14(ctx, imts, region='CAL', sigma_mu_epsilon=0.0, use_hw=True,
         alpha_nm=1.0, peer=False):
    """ChiouYoungs2014.compute, chiou_youngs_2014.py:767-805"""
    out = _out(len(imts), len(ctx.mag))
    pga = _cy_one(region, coeffs('CY14', 'PGA'), ctx, use_hw, alpha_nm, peer)
    for m, imt in enumerate(imts):
        kind, period = parse_imt(imt)
        if imt == 'PGA':
            mea, sig, tau, phi = pga
            mea = mea.copy()
        else:
            mea, sig, tau, phi = _cy_one(
                region, coeffs('CY14', imt), ctx, use_hw, alpha_nm, peer)
            if kind == 'SA' and period <= 0.3:
                mea = np.where(mea < pga[0], pga[0], mea)
        out[0, m] = mea + sigma_mu_epsilon * epistemic_sigma(ctx)
        out[1, m], out[2, m], out[3, m] = sig, tau, phi
    return out


# --------------------------------------------------------------------- #
#                   Abrahamson, Silva & Kamai (2014)                     #
# --------------------------------------------------------------------- #

def _interp_lin(x, xp, fp, extrapolate):
    """scipy.interpolate.interp1d(kind='linear') on a sorted grid"""
    x = np.asarray(x, float)
    xp, fp = np.asarray(xp, float), np.asarray(fp, float)
    if not extrapolate and ((x < xp[0]) | (x > xp[-1])).any():
        raise ValueError('A value in x_new is outside the interpolation range')
    # interp1d picks the interval with searchsorted and clips to [1, n-1]
    hi = np.clip(np.searchsorted(xp, x), 1, len(xp) - 1)
    lo = hi - 1
    slope = (fp[hi] - fp[lo]) / (xp[hi] - xp[lo])
    return slope * (x - xp[lo]) + fp[lo]


def _ask_v1(imt):
    # abrahamson_2014.py:345-365
    kind, t = parse_imt(imt)
    if kind == 'SA':
        if t <= 0.50:
            return 1500.0
        elif t < 3.0:
            return np.exp(-0.35 * np.log(t / 0.5) + np.log(1500.))
        return 800.0
    return 1500.0


def _ask_vs30star(vs30, imt):
    return np.where(vs30 >= _ask_v1(imt), _ask_v1(imt), vs30)


def _ask_basic(C, ctx):
    # abrahamson_2014.py:93-123
    mag, rrup = ctx.mag, ctx.rrup
    c4m = C['c4'] - (C['c4'] - 1.) * (5. - mag)
    c4m = np.where(mag > 5., C['c4'], c4m)
    c4m = np.where(mag < 4, 1., c4m)
    R = np.sqrt(rrup ** 2. + c4m ** 2.)
    base = C['a1'] * np.ones_like(rrup) + C['a17'] * rrup
    m1, m2 = C['m1'], 5.0
    hi = (C['a5'] * (mag - m1) + C['a8'] * (8.5 - mag) ** 2. +
          (C['a2'] + C['a3'] * (mag - m1)) * np.log(R))
    mid = (C['a4'] * (mag - m1) + C['a8'] * (8.5 - mag) ** 2. +
           (C['a2'] + C['a3'] * (mag - m1)) * np.log(R))
    lo = (C['a4'] * (m2 - m1) + C['a8'] * (8.5 - m2) ** 2. +
          C['a6'] * (mag - m2) + C['a7'] * (mag - m2) ** 2. +
          (C['a2'] + C['a3'] * (m2 - m1)) * np.log(R))
    return base + np.where(mag >= m1, hi, np.where(mag >= m2, mid, lo))


def _ask_sof(C, ctx):
    # abrahamson_2014.py:141-154
    f7 = C['a11'] * np.clip(ctx.mag - 4., 0., 1.)
    f8 = C['a12'] * np.clip(ctx.mag - 4., 0., 1.)
    return (f7 * ((ctx.rake > 30) & (ctx.rake < 150)) +
            f8 * ((ctx.rake > -150) & (ctx.rake < -30)))


def _ask_hw(C, ctx):
    # abrahamson_2014.py:157-177 and tapers :385-437
    mag, rx, dip = ctx.mag, ctx.rx, ctx.dip
    fhw = np.where(rx > 0, 1., 0.)
    t1 = np.where(dip <= 30., 60. / 45., (90. - dip) / 45.0)
    a2hw = 0.2
    t2 = (1. + a2hw * (mag - 6.5) - (1. - a2hw) * (mag - 6.5) ** 2)
    t2 = np.where(mag > 6.5, 1. + a2hw * (mag - 6.5), t2)
    t2 = np.where(mag <= 5.5, 0., t2)
    r1 = ctx.width * np.cos(np.radians(dip))
    r2 = 3. * r1
    with np.errstate(all='ignore'):
        t3a = 0.25 + 1.50 * (rx / r1) + (-0.75) * (rx / r1) ** 2
        t3b = 1. - (rx - r1) / (r2 - r1)
    t3 = np.where(rx < r1, t3a, np.where(rx <= r2, t3b, 0.))
    t4 = np.where(ctx.ztor <= 10., 1. - ctx.ztor ** 2. / 100., 0.)
    ry1 = rx * np.tan(np.radians(20.))
    d = ctx.ry0 - ry1
    t5 = np.where(d <= 0.0, 1., np.where(d < 5.0, 1. - d / 5.0, 0.))
    fhw = np.where(dip == 90., 0., fhw)
    return fhw * C['a13'] * t1 * t2 * t3 * t4 * t5


def _ask_site(C, imt, vs30, sa1180):
    # abrahamson_2014.py:226-248
    rat = _ask_vs30star(vs30, imt) / C['vlin']
    above = (C['a10'] + C['b'] * 1.5) * np.log(rat)
    below = (C['a10'] * np.log(rat) - C['b'] * np.log(sa1180 + C['c']) +
             C['b'] * np.log(sa1180 + C['c'] * rat ** 1.5))
    return np.where(vs30 >= C['vlin'], above, below)


def _ask_z1ref(region, vs30):
    # abrahamson_2014.py:368-380
    if region == 'JPN':
        return np.exp(-5.23 / 2. * np.log(
            (vs30 ** 2 + 412. ** 2.) / (1360. ** 2 + 412 ** 2.))) / 1000.
    return np.exp((-7.67 / 4.) * np.log(
        (vs30 ** 4 + 610. ** 4) / (1360. ** 4 + 610. ** 4))) / 1000.


def _ask_basin(C, ctx, region, rock):
    # abrahamson_2014.py:251-303 (no USGS / CyberShake)
    if rock:
        vs30 = np.ones_like(ctx.vs30) * 1180.
        z10 = np.ones_like(vs30) * -1
    else:
        vs30 = ctx.vs30
        z10 = np.array(ctx.z1pt0, dtype=float)
    z1ref = _ask_z1ref(region, vs30)
    z10 = z10 / 1000.
    z10 = np.where(z10 < 0, z1ref, z10)
    factor = np.log((z10 + 0.01) / (z1ref + 0.01))
    f2 = _interp_lin(vs30, [0.0, 150, 250, 400, 700, 1000, 6000],
                     [C['a43'], C['a43'], C['a44'], C['a45'], C['a46'],
                      C['a46'], C['a46']], extrapolate=False)
    return (f2 * factor) + 0.


def _ask_regional(region, C, imt, vs30, rrup):
    # abrahamson_2014.py:202-223
    if region == 'TWN':
        return (C['a31'] * np.log(_ask_vs30star(vs30, imt) / C['vlin']) +
                C['a25'] * rrup)
    elif region == 'CHN':
        return C['a28'] * rrup
    elif region == 'JPN':
        f3 = _interp_lin(
            vs30, [150, 250, 350, 450, 600, 850, 1150, 2000],
            [C['a36'], C['a37'], C['a38'], C['a39'], C['a40'], C['a41'],
             C['a42'], C['a42']], extrapolate=True)
        return f3 + C['a29'] * rrup
    return 0.


def _ask_deramp(C, sa1180, vs30):
    # abrahamson_2014.py:126-138
    with np.errstate(all='ignore'):
        d = (C['b'] * sa1180 * (-1. / (sa1180 + C['c']) + 1. / (
            sa1180 + C['c'] * (vs30 / C['vlin']) ** 1.5)))
    return np.where(vs30 < C['vlin'], d, 0.)


def _ask_std(region, C, ctx, sa1180):
    # abrahamson_2014.py:61-90, 180-199 and 306-316
    mag = ctx.mag
    meas = np.asarray(ctx.vs30measured, bool)
    if region == 'JPN':
        r = ctx.rrup
        phi_al = np.where(
            r < 30, C['s5'], np.where(
                r <= 80., C['s5'] + (C['s6'] - C['s5']) / 50. * (r - 30.),
                C['s6']))
        phi_al = 1. * phi_al
    else:
        s1 = np.where(meas, C['s1m'], C['s1e'])
        s2 = np.where(meas, C['s2m'], C['s2e'])
        phi_al = s1 + (s2 - s1) / 2. * (mag - 4.)
        phi_al = np.where(mag < 4, s1, phi_al)
        phi_al = np.where(mag >= 6, s2, phi_al)
    der = _ask_deramp(C, sa1180, ctx.vs30)
    phi_amp = np.where(phi_al < 0.4, 0.99 * phi_al, 0.4)
    phi_b = np.sqrt(phi_al ** 2 - phi_amp ** 2)
    phi = np.sqrt(phi_b ** 2 * (1 + der) ** 2 + phi_amp ** 2)
    tau_al = C['s3'] + (C['s4'] - C['s3']) / 2. * (mag - 5.)
    tau_al = np.where(mag < 5., C['s3'], tau_al)
    tau_al = np.where(mag >= 7., C['s4'], tau_al)
    tau = tau_al * (1 + der)
    return np.sqrt(phi ** 2 + tau ** 2), tau, phi


def ask14(ctx, imts, region=None, sigma_mu_epsilon=0.0):
    """AbrahamsonEtAl2014.compute, abrahamson_2014.py:517-560"""
    out = _out(len(imts), len(ctx.mag))
    for m, imt in enumerate(imts):
        C = coeffs('ASK14', imt)
        v1180 = np.ones_like(ctx.vs30) * 1180.
        # _get_sa_at_1180, :419-438
        sa1180 = np.exp(
            _ask_basic(C, ctx) + _ask_sof(C, ctx) +
            _ask_site(C, imt, v1180, np.zeros_like(ctx.vs30)) +
            _ask_hw(C, ctx) +
            C['a15'] * np.clip(ctx.ztor / 20.0, None, 1.) +
            _ask_basin(C, ctx, region, rock=True) +
            _ask_regional(region, C, imt, v1180, ctx.rrup))
        mean = (_ask_basic(C, ctx) + _ask_hw(C, ctx) +
                _ask_site(C, imt, ctx.vs30, sa1180) +
                C['a15'] * np.clip(ctx.ztor / 20.0, None, 1.) +
                _ask_sof(C, ctx) + _ask_basin(C, ctx, region, rock=False))
        mean = mean + _ask_regional(region, C, imt, ctx.vs30, ctx.rrup)
        mean = mean + sigma_mu_epsilon * epistemic_sigma(ctx)
        out[0, m] = mean
        out[1, m], out[2, m], out[3, m] = _ask_std(region, C, ctx, sa1180)
    return out


# --------------------------------------------------------------------- #
#                      Campbell & Bozorgnia (2014)                       #
# --------------------------------------------------------------------- #

_CB_C, _CB_N = 1.88, 1.18


def _cb_z25ref(SJ, vs30):
    # campbell_bozorgnia_2014.py:102-113
    if SJ:
        return np.exp(5.359 - 1.102 * np.log(vs30))
    return np.exp(7.089 - 1.144 * np.log(vs30))


def _cb_basin_fn(C, z25, SJ):
    # campbell_bozorgnia_2014.py:80-99 (no CyberShake)
    k = C['c16'] * C['k3']
    fb = np.zeros(len(z25))
    lo = z25 < 1.0
    fb[lo] = (C['c14'] + C['c15'] * SJ) * (z25[lo] - 1.0)
    hi = z25 > 3.0
    fb[hi] = k * math.exp(-0.75) * (1. - np.exp(-0.25 * (z25[hi] - 3.)))
    return fb


def _cb_mean(SJ, C, ctx, a1100):
    # get_mean_values campbell_bozorgnia_2014.py:349-370; a1100 None means
    # the rock pass (vs30 = 1100, reference basin depth)
    mag, rrup, rx = ctx.mag, ctx.rrup, ctx.rx
    rock = a1100 is None
    vs30 = 1100.0 * np.ones(len(mag)) if rock else ctx.vs30
    # magnitude term :268-283
    fmag = C['c0'] + C['c1'] * mag
    fmag = fmag + np.where(
        mag > 6.5, (C['c2'] * (mag - 4.5) + C['c3'] * (mag - 5.5) +
                    C['c4'] * (mag - 6.5)),
        np.where(mag > 5.5, C['c2'] * (mag - 4.5) + C['c3'] * (mag - 5.5),
                 np.where(mag > 4.5, C['c2'] * (mag - 4.5), 0.)))
    # geometric attenuation :171-176
    fdis = (C['c5'] + C['c6'] * mag) * np.log(np.sqrt(rrup ** 2 + C['c7'] ** 2))
    # style of faulting :326-338
    frv = np.where((ctx.rake > 30.) & (ctx.rake < 150.), 1., 0.)
    fnm = np.where((ctx.rake > -150.) & (ctx.rake < -30.), 1., 0.)
    fflt_m = np.where(mag <= 4.5, 0., np.where(mag > 5.5, 1., mag - 4.5))
    fflt = (C['c8'] * frv + C['c9'] * fnm) * fflt_m
    # hanging wall :179-253
    r1 = ctx.width * np.cos(np.radians(ctx.dip))
    r2 = 62.0 * mag - 350.0
    with np.errstate(all='ignore'):
        q = rx / r1
        f1 = C['h1'] + C['h2'] * q + C['h3'] * q ** 2
        d = (rx - r1) / (r2 - r1)
        f2 = np.maximum(1.0 + C['h5'] * d + C['h6'] * d ** 2, 0.0)
        # NB: numpy assigns f2rx[f2rx < 0] = 0, NaNs stay NaN; maximum keeps
        # NaN too
        fhng_rx = np.where(rx >= r1, f2, np.where(rx >= 0., f1, 0.))
        fhng_rrup = np.where(rrup > 0.0, (rrup - ctx.rjb) / rrup, 1.)
    fhng_m = (mag - 5.5) * (1.0 + C['a2'] * (mag - 6.5))
    fhng_m = np.where(mag < 5.5, 0.0, fhng_m)
    fhng_m = np.where(mag > 6.5, 1.0 + C['a2'] * (mag - 6.5), fhng_m)
    fhng_z = np.where(ctx.ztor > 16.66, 0., 1. - 0.06 * ctx.ztor)
    fhng_d = (90.0 - ctx.dip) / 45.0
    fhng = C['c10'] * fhng_rx * fhng_rrup * fhng_m * fhng_z * fhng_d
    # shallow site response :297-323
    vmod = vs30 / C['k1']
    fsite = C['c11'] * np.log(vmod)
    stiff = vs30 > C['k1']
    fsite = np.where(stiff, fsite + (C['k2'] * _CB_N * np.log(vmod)), fsite)
    if not rock:
        nl = C['k2'] * (np.log(a1100 + _CB_C * (vmod ** _CB_N)) -
                        np.log(a1100 + _CB_C))
        fsite = np.where(stiff, fsite, fsite + nl)
    if SJ:
        fj = (C['c13'] + C['k2'] * _CB_N) * np.log(vmod)
        soft = (C['c12'] + C['k2'] * _CB_N) * (
            np.log(vmod) - np.log(200.0 / C['k1']))
        fsite = fsite + np.where(vs30 <= 200.0, fj + soft, fj)
    # basin :116-150
    zref = _cb_z25ref(SJ, 1100.0) * np.ones_like(ctx.vs30)
    if rock:
        z25 = zref
    else:
        z25 = np.array(ctx.z2pt5, dtype=float)
        unknown = z25 == -999
        z25[unknown] = _cb_z25ref(SJ, ctx.vs30[unknown])
    fsed = _cb_basin_fn(C, z25, SJ)
    # hypocentral depth :256-265
    fh = np.clip(ctx.hypo_depth - 7.0, 0., 13.)
    fm = C['c17'] + (C['c18'] - C['c17']) * (mag - 5.5)
    fm = np.where(mag <= 5.5, C['c17'], fm)
    fm = np.where(mag > 6.5, C['c18'], fm)
    fhyp = fh * fm
    # dip :160-168
    fdip = C['c19'] * (5.5 - mag) * ctx.dip
    fdip = np.where(mag < 4.5, C['c19'] * ctx.dip, fdip)
    fdip = np.where(mag > 5.5, 0.0, fdip)
    # anelastic :69-77
    fatn = np.where(rrup >= 80.0, (C['c20'] + C['Dc20']) * (rrup - 80.0), 0.)
    return fmag + fdis + fflt + fhng + fsite + fsed + fhyp + fdip + fatn


def _cb_lin(C, mag, a, b):
    # _get_taulny / _get_philny, campbell_bozorgnia_2014.py:286-294, 341-349
    res = C[b] + (C[a] - C[b]) * (5.5 - mag)
    res = np.where(mag <= 4.5, C[a], res)
    return np.where(mag >= 5.5, C[b], res)


def _cb_sigma(C, ctx, pga1100, tau_pga, phi_pga):
    # compute_sigma campbell_bozorgnia_2014.py:432-463 (non CAV/IA branch)
    tau_b = _cb_lin(C, ctx.mag, 'tau1', 'tau2')
    phi_b = np.sqrt(_cb_lin(C, ctx.mag, 'phi1', 'phi2') ** 2. -
                    C['philnAF'] ** 2.)
    vsk1 = ctx.vs30 / C['k1']
    af1 = pga1100 + _CB_C * (vsk1 ** _CB_N)
    af2 = pga1100 + _CB_C
    alpha = np.where(ctx.vs30 < C['k1'],
                     C['k2'] * pga1100 * ((1.0 / af1) - (1.0 / af2)), 0.)
    t = np.sqrt(tau_b ** 2 + alpha ** 2 * tau_pga ** 2 +
                2.0 * alpha * C['rholny'] * tau_b * tau_pga)
    p = np.sqrt(phi_b ** 2 + C['philnAF'] ** 2 + alpha ** 2 * phi_pga ** 2 +
                2.0 * alpha * C['rholny'] * phi_b * phi_pga)
    return np.sqrt(t ** 2 + p ** 2), t, p


def cb14(ctx, imts, SJ=False, sigma_mu_epsilon=0.0):
    """CampbellBozorgnia2014.compute, campbell_bozorgnia_2014.py:520-567"""
    out = _out(len(imts), len(ctx.mag))
    Cp = coeffs('CB14', 'PGA')
    SJ = int(bool(SJ))
    pga1100 = np.exp(_cb_mean(SJ, Cp, ctx, None))
    tau_pga = _cb_lin(Cp, ctx.mag, 'tau1', 'tau2')
    phi_pga = np.sqrt(_cb_lin(Cp, ctx.mag, 'phi1', 'phi2') ** 2. -
                      Cp['philnAF'] ** 2.)
    for m, imt in enumerate(imts):
        C = coeffs('CB14', imt)
        kind, period = parse_imt(imt)
        mean = _cb_mean(SJ, C, ctx, pga1100)
        mean = mean + sigma_mu_epsilon * epistemic_sigma(ctx)
        if kind == 'SA' and period < 0.25:
            pga = _cb_mean(SJ, Cp, ctx, pga1100)
            mean = np.where(mean <= pga, pga, mean)
            # the epistemic term is added a second time in the reference
            mean = mean + sigma_mu_epsilon * epistemic_sigma(ctx)
        out[0, m] = mean
        out[1, m], out[2, m], out[3, m] = _cb_sigma(
            C, ctx, pga1100, tau_pga, phi_pga)
    return out


# --------------------------------------------------------------------- #
#        Boore & Atkinson (2008) and Atkinson & Boore (2006)             #
# --------------------------------------------------------------------- #

def _ba_fault_dummies(rake):
    # gsim/utils.py:60-78
    ss = np.where((np.abs(rake) <= 30.) | (180. - np.abs(rake) <= 30.), 1., 0.)
    rs = np.where((rake > 30.) & (rake < 150.), 1., 0.)
    ns = np.where((rake > -150.) & (rake < -30), 1., 0.)
    return ss, ns, rs


def _ba_fm(ctx, C):
    # _compute_ms atkinson_boore_2006.py:127-137
    ss, ns, rs = _ba_fault_dummies(ctx.rake)
    dm = ctx.mag - C['Mh']
    above = C['e2'] * ss + C['e3'] * ns + C['e4'] * rs + C['e7'] * dm
    below = (C['e2'] * ss + C['e3'] * ns + C['e4'] * rs + C['e5'] * dm +
             C['e6'] * dm ** 2)
    return np.where(ctx.mag <= C['Mh'], below, above)


def _ba_fd(ctx, C):
    # atkinson_boore_2006.py:63-71
    R = np.sqrt(ctx.rjb ** 2 + C['h'] ** 2)
    return ((C['c1'] + C['c2'] * (ctx.mag - 4.5)) * np.log(R / 1.0) +
            C['c3'] * (R - 1.0))


def _ba_bnl(vs30, S):
    # atkinson_boore_2006.py:140-163
    bnl = np.zeros(vs30.shape)
    bnl[vs30 <= 180.0] = S['b1']
    i = (vs30 > 180.0) & (vs30 <= 300.0)
    bnl[i] = ((S['b1'] - S['b2']) * np.log(vs30[i] / 300.0) /
              np.log(180.0 / 300.0) + S['b2'])
    i = (vs30 > 300.0) & (vs30 < 760.0)
    bnl[i] = S['b2'] * np.log(vs30[i] / 760.0) / np.log(300.0 / 760.0)
    return bnl


def _ba_fnl(pga4nl, bnl):
    # atkinson_boore_2006.py:166-195
    a1, a2, pga_low = 0.03, 0.09, 0.06
    dx = np.log(a2 / a1)
    dy = bnl * np.log(a2 / pga_low)
    c = (3 * dy - bnl * dx) / dx ** 2
    d = -(2 * dy - bnl * dx) / dx ** 3
    low = bnl * np.log(pga_low / 0.1)
    with np.errstate(all='ignore'):
        lp = np.log(pga4nl / a1)
        mid = bnl * np.log(pga_low / 0.1) + c * (lp ** 2) + d * (lp ** 3)
        top = bnl * np.log(pga4nl / 0.1)
    return np.where(pga4nl <= a1, low, np.where(pga4nl <= a2, mid, top))


def _ba_site(vs30, pga4nl, S):
    return S['blin'] * np.log(vs30 / 760.0), _ba_fnl(pga4nl, _ba_bnl(vs30, S))


def ba08(ctx, imts):
    """BooreAtkinson2008.compute (kind='base'), boore_atkinson_2008.py:100-151"""
    out = _out(len(imts), len(ctx.mag))
    Cp = coeffs('BA08', 'PGA')
    pga4nl = np.exp(_ba_fm(ctx, Cp) + _ba_fd(ctx, Cp))
    for m, imt in enumerate(imts):
        C = coeffs('BA08', imt)
        S = coeffs('AB06_SOIL_RESPONSE', imt)
        lin, nl = _ba_site(ctx.vs30, pga4nl, S)
        if imt in ('PGA', 'SA(0)', 'SA(0.0)'):
            out[0, m] = np.log(pga4nl) + lin + nl
        else:
            out[0, m] = _ba_fm(ctx, C) + _ba_fd(ctx, C) + lin + nl
        out[1, m], out[2, m], out[3, m] = C['std'], C['tau'], C['sigma']
    return out


_G = 9.80665  # scipy.constants.g


def _ab06_rock(C, f0, f1, f2, SC, mag, rrup, scale_fac):
    # _compute_mean + stress adjustment, atkinson_boore_2006.py:109-124,213-221
    cm = np.clip(mag - SC['M1'], 0, None)
    sda = scale_fac * np.clip(
        0.05 + SC['delta'] * cm / (SC['Mh'] - SC['M1']), None,
        SC['delta'] + 0.05)
    return (C['c1'] + C['c2'] * mag + C['c3'] * (mag ** 2) +
            (C['c4'] + C['c5'] * mag) * f1 + (C['c6'] + C['c7'] * mag) * f2 +
            (C['c8'] + C['c9'] * mag) * f0 + C['c10'] * rrup + sda)


def ab06(ctx, imts, mag_eq='NA', scale_fac=0):
    """AtkinsonBoore2006.compute, atkinson_boore_2006.py:421-448"""
    out = _out(len(imts), len(ctx.mag))
    mag = ctx.mag
    if mag_eq == 'Mblg87':
        mag = 2.715 - 0.277 * mag + 0.127 * mag * mag
    elif mag_eq == 'Mblg96':
        mag = 1.14 + 0.24 * mag + 0.0933 * mag * mag
    if mag_eq == 'NA':
        scale_fac = 0
    vs30 = ctx.vs30
    rrup = np.where(ctx.rrup < 1, 1., ctx.rrup)
    f0 = np.maximum(np.log10(10.0 / rrup), 0.0)
    f1 = np.minimum(np.log10(rrup), np.log10(70.0))
    f2 = np.maximum(np.log10(rrup / 140.0), 0.0)
    for m, imt in enumerate(imts):
        kind, period = parse_imt(imt)
        HR, BC = coeffs('AB06_HARD_ROCK', imt), coeffs('AB06_BC', imt)
        SR, SC = coeffs('AB06_SOIL_RESPONSE', imt), coeffs('AB06_STRESS', imt)
        soft = vs30 < 2000.0
        pga_bc = np.where(soft, _ab06_rock(
            coeffs('AB06_BC', 'PGA'), f0, f1, f2, SC, mag, rrup, scale_fac), 0.)
        pga_bc = (10 ** pga_bc) * 1e-2 / _G
        mean = np.where(soft,
                        _ab06_rock(BC, f0, f1, f2, SC, mag, rrup, scale_fac),
                        _ab06_rock(HR, f0, f1, f2, SC, mag, rrup, scale_fac))
        lin, nl = _ba_site(vs30, pga_bc, SR)
        sal = np.log10(np.exp(lin))
        sanl = np.log10(np.exp(nl))
        mean = np.where(soft, mean + sal + sanl, mean)
        if imt == 'PGV':
            mean = np.log(10 ** mean)
        else:
            mean = np.log((10 ** mean) * 1e-2 / _G)
        if mag_eq != 'NA':  # clip_mean, gsim/utils.py:45-57
            if period == 0:
                mean = np.minimum(mean, 0.405)
            if 0.02 < period < 0.55:
                mean = np.minimum(mean, 1.099)
        out[0, m] = mean
        out[1, m] = np.log(10 ** 0.30)
    return out


# --------------------------------------------------------------------- #
#                  Akkar, Sandikkaya & Bommer (2014)                     #
# --------------------------------------------------------------------- #

def _akkar_rock(C, ctx, dist):
    # akkar_2014.py:33-93, 121-126 (c1 = 6.75)
    mag, c1 = ctx.mag, 6.75
    fn = (ctx.rake > -135.0) & (ctx.rake < -45.0)
    fr = (ctx.rake > 45.0) & (ctx.rake < 135.0)
    return (C['a1'] +
            np.where(mag <= c1, C['a2'] * (mag - c1), C['a7'] * (mag - c1)) +
            C['a3'] * (8.5 - mag) ** 2 +
            (C['a4'] + C['a5'] * (mag - c1)) * np.log(
                np.sqrt(dist ** 2 + C['a6'] ** 2)) +
            (C['a8'] * fn + C['a9'] * fr))


def _akkar_site(C, pga, vs30):
    # akkar_2014.py:96-118
    vref, vcon = 750.0, 1000.0
    with np.errstate(all='ignore'):
        soft = (C['b1'] * np.log(vs30 / vref) + C['b2'] * np.log(
            (pga + C['c'] * (vs30 / vref) ** C['n']) /
            ((pga + C['c']) * (vs30 / vref) ** C['n'])))
    return np.where(vs30 < vref, soft, np.where(
        vs30 <= vcon, C['b1'] * np.log(vs30 / vref),
        C['b1'] * np.log(vcon / vref)))


def akkar14(ctx, imts, kind='rjb', adjustment_factor=1.0):
    """AkkarEtAl{Rjb,Repi,Rhyp}2014.compute, akkar_2014.py:177-198"""
    out = _out(len(imts), len(ctx.mag))
    table = {'rjb': 'AKKAR14_RJB', 'repi': 'AKKAR14_REPI',
             'rhypo': 'AKKAR14_RHYP'}[kind]
    dist = getattr(ctx, kind)
    pga = np.exp(_akkar_rock(coeffs(table, 'PGA'), ctx, dist))
    for m, imt in enumerate(imts):
        C = coeffs(table, imt)
        out[0, m] = (_akkar_rock(C, ctx, dist) + _akkar_site(C, pga, ctx.vs30)
                     + np.log(adjustment_factor))
        out[1, m] = np.sqrt(C['sigma'] ** 2 + C['tau'] ** 2)
        out[2, m] = C['tau']
        out[3, m] = C['sigma']
    return out


# --------------------------------------------------------------------- #
#        the models of the LogicTreeCase2ClassicalPSHA demo (cfg2)       #
# --------------------------------------------------------------------- #

def cy08(ctx, imts):
    """ChiouYoungs2008.compute, chiou_youngs_2008.py:204-228 with
    _get_ln_y_ref :56-99, _get_mean :102-124, _get_basin_term :37-53,
    _set_stddevs :127-164"""
    out = _out(len(imts), len(ctx.mag))
    mag, rake, rrup = ctx.mag, ctx.rake, ctx.rrup
    frv = ((30 <= rake) & (rake <= 150)).astype(float)
    fnm = ((-120 <= rake) & (rake <= -60)).astype(float)
    fhw = np.where(ctx.rx >= 0, 1, 0)
    for m, imt in enumerate(imts):
        C = coeffs('CY08', imt)
        ln_y_ref = (
            C['c1'] + (C['c1a'] * frv + C['c1b'] * fnm
                       + C['c7'] * (ctx.ztor - 4))
            + C['c2'] * (mag - 6)
            + ((C['c2'] - C['c3']) / C['cn'])
            * np.log(1 + np.exp(C['cn'] * (C['cm'] - mag)))
            + C['c4'] * np.log(rrup + C['c5'] * np.cosh(
                C['c6'] * np.fmax(mag - C['chm'], 0)))
            + (C['c4a'] - C['c4']) * np.log(np.sqrt(rrup ** 2 + C['crb'] ** 2))
            + (C['cg1'] + C['cg2'] / (np.cosh(np.fmax(mag - C['cg3'], 0))))
            * rrup
            + C['c9'] * fhw * np.tanh(
                ctx.rx * (np.cos(np.radians(ctx.dip)) ** 2) / C['c9a'])
            * (1 - np.sqrt(ctx.rjb ** 2 + ctx.ztor ** 2) / (rrup + 0.001)))
        exp1 = np.exp(C['phi3'] * (ctx.vs30.clip(-np.inf, 1130) - 360))
        exp2 = np.exp(C['phi3'] * (1130 - 360))
        z1pt0 = np.array(ctx.z1pt0, float)
        mask = z1pt0 == -999
        z1pt0[mask] = np.exp(28.5 - (3.82 / 8) * np.log(
            ctx.vs30[mask] ** 8 + 378.7 ** 8))
        deep = C['phi5'] * (1.0 - 1.0 / np.cosh(
            C['phi6'] * (z1pt0 - C['phi7']).clip(0, np.inf)))
        shallow = C['phi8'] / np.cosh(0.15 * (z1pt0 - 15).clip(0, np.inf))
        out[0, m] = (
            ln_y_ref + C['phi1'] * np.log(ctx.vs30 / 1130).clip(-np.inf, 0)
            + C['phi2'] * (exp1 - exp2)
            * np.log((np.exp(ln_y_ref) + C['phi4']) / C['phi4'])
            + (deep + shallow))
        fmeas = np.where(ctx.vs30measured, True, False)
        mag_test = np.clip(mag, 5.0, 7.0) - 5.0
        t = C['tau1'] + (C['tau2'] - C['tau1']) / 2 * mag_test
        b = C['phi2'] * (exp1 - exp2)
        y_ref = np.exp(ln_y_ref)
        NL = b * y_ref / (y_ref + C['phi4'])
        sigma = ((C['sig1'] + 0.5 * (C['sig2'] - C['sig1']) * mag_test)
                 * np.sqrt((C['sig3'] * ~fmeas + 0.7 * fmeas) + (1 + NL) ** 2))
        out[1, m] = np.sqrt(((1 + NL) ** 2) * (t ** 2) + sigma ** 2)
        out[2, m] = np.abs((1 + NL) * t)
        out[3, m] = sigma
    return out


def toro02(ctx, imts):
    """ToroEtAl2002.compute, toro_2002.py:103-128 with _compute_mean :31-57
    (total standard deviation only: tau and phi stay zero)"""
    out = _out(len(imts), len(ctx.mag))
    mag, rjb = ctx.mag, ctx.rjb
    for m, imt in enumerate(imts):
        C = coeffs('TORO02', imt)
        period = parse_imt(imt)[1]
        md = mag - 6
        RM = np.sqrt(rjb ** 2 + (C['c7'] ** 2) *
                     np.exp(-1.25 + 0.227 * mag) ** 2)
        term2 = (-C['c4'] * np.log(RM) - (C['c5'] - C['c4']) *
                 np.maximum(np.log(RM / 100), 0) - C['c6'] * RM)
        out[0, m] = C['c1'] + (C['c2'] * md + C['c3'] * md ** 2) + term2
        ale_m = np.interp(mag, [5.0, 5.5, 8.0],
                          [C['m50'], C['m55'], C['m80']])
        ale_r = np.interp(rjb, [5.0, 20.0], [C['r5'], C['r20']])
        ale = np.sqrt(ale_m ** 2 + ale_r ** 2)
        if period < 1:
            epi = 0.36 + 0.07 * (mag - 6)
        else:
            epi = 0.34 + 0.06 * (mag - 6)
        out[1, m] = np.sqrt(ale ** 2 + epi ** 2)
        if period == 3.0:
            out[0, m] /= 0.612
        if period == 4.0:
            out[0, m] /= 0.559
    return out


def campbell03(ctx, imts):
    """Campbell2003.compute (kind "base"), campbell_2003.py:178-189 with
    _compute_mean_1 :39-47 and the terms :71-106"""
    out = _out(len(imts), len(ctx.mag))
    mag, rrup = ctx.mag, ctx.rrup
    for m, imt in enumerate(imts):
        C = coeffs('C03', imt)
        term1 = (C['c2'] * mag) + C['c3'] * (8.5 - mag) ** 2
        R = np.sqrt(rrup ** 2 + (C['c7'] * np.exp(C['c8'] * mag)) ** 2)
        term2 = C['c4'] * np.log(R) + (C['c5'] + C['c6'] * mag) * rrup
        f3 = np.zeros_like(rrup)
        mid = (rrup > 70) & (rrup <= 130)
        far = rrup > 130
        f3[mid] = C['c9'] * (np.log(rrup[mid]) - np.log(70))
        f3[far] = (C['c9'] * (np.log(rrup[far]) - np.log(70)) +
                   C['c10'] * (np.log(rrup[far]) - np.log(130)))
        out[0, m] = C['c1'] + term1 + term2 + f3
        out[1, m] = np.where(mag < 7.16, C['c11'] + C['c12'] * mag, C['c13'])
    return out


def sadigh97(ctx, imts):
    """SadighEtAl1997.compute, sadigh_1997.py:134-161 with get_mean_rock
    :71-87, get_mean_deep_soil :47-68 and the two standard deviations
    :90-107 (total only)"""
    out = _out(len(imts), len(ctx.mag))
    mag, rrup = ctx.mag, ctx.rrup
    reverse = (45 <= ctx.rake) & (ctx.rake <= 135)
    rock = ctx.vs30 > 750
    low = mag <= 6.5
    mc = np.clip(mag, None, 8.5)
    for m, imt in enumerate(imts):
        lo = coeffs('SADIGH97_ROCK_LOWMAG', imt)
        hi = coeffs('SADIGH97_ROCK_HIMAG', imt)
        sd = coeffs('SADIGH97_ROCK_STDDERR', imt)
        so = coeffs('SADIGH97_SOIL', imt)
        C = {k: np.where(low, lo[k], hi[k]) for k in lo}
        mean_r = (C['c1'] + C['c2'] * mc + C['c3'] * ((8.5 - mc) ** 2.5)
                  + C['c4'] * np.log(rrup + np.exp(C['c5'] + C['c6'] * mc))
                  + C['c7'] * np.log(rrup + 2))
        mean_r = np.where(reverse, mean_r + 0.1823215567939546, mean_r)
        sig_r = np.where(mag > sd['maxmag'], sd['maxsigma'],
                         sd['sigma0'] + sd['magfactor'] * mag)
        c1 = np.where(reverse, -1.92, -2.17)
        c4 = np.where(low, 2.1863, 0.3825)
        c5 = np.where(low, 0.32, 0.5882)
        c6 = np.where(reverse, so['c6r'], so['c6ss'])
        mean_s = (c1 + 1.0 * mc + c6 + so['c7'] * ((8.5 - mc) ** 2.5)
                  - 1.7 * np.log(rrup + c4 * np.exp(c5 * mc)))
        sig_s = so['sigma0'] + so['magfactor'] * np.clip(mag, None, 7)
        out[0, m] = np.where(rock, mean_r, mean_s)
        out[1, m] = np.where(rock, sig_r, sig_s)
    return out


def _share_term(Frss, rake, pR=0.81, Fnss=0.95, pN=0.01):
    """_compute_faulting_style_term, campbell_2003.py:272-281 (CONSTS_FS of
    both SHARE classes)"""
    term = np.full_like(rake, np.power(Frss, -pR) * np.power(Fnss, -pN))
    term[(rake > 30.0) & (rake <= 150.0)] = (
        np.power(Frss, 1 - pR) * np.power(Fnss, -pN))
    term[(rake > -120.0) & (rake <= -60.0)] = (
        np.power(Frss, -pR) * np.power(Fnss, 1 - pN))
    return term


def toro02_share(ctx, imts):
    """ToroEtAl2002SHARE, toro_2002.py:134-142 on top of toro02"""
    out = toro02(ctx, imts)
    for m, imt in enumerate(imts):
        C = coeffs('TORO02_FS_ROCK', imt)
        out[0, m] = np.log(np.exp(out[0, m]) * (
            _share_term(C['Frss'], ctx.rake) * C['AFrock']))
    return out


def campbell03_share(ctx, imts):
    """Campbell2003SHARE, campbell_2003.py:236-245 on top of campbell03"""
    out = campbell03(ctx, imts)
    for m, imt in enumerate(imts):
        C = coeffs('C03_FS_ROCK', imt)
        out[0, m] = np.log(np.exp(out[0, m]) *
                           _share_term(C['Frss'], ctx.rake) * C['AFrock'])
    return out


# --------------------------------------------------------------------- #
#                              registry                                  #
# --------------------------------------------------------------------- #

MODELS = {
    'BooreEtAl2014': bssa14,
    'ChiouYoungs2014': cy14,
    'AbrahamsonEtAl2014': ask14,
    'CampbellBozorgnia2014': cb14,
    'BooreAtkinson2008': ba08,
    'AtkinsonBoore2006': ab06,
    'AkkarEtAlRjb2014': lambda ctx, imts, **kw: akkar14(ctx, imts, 'rjb', **kw),
    'AkkarEtAlRepi2014':
        lambda ctx, imts, **kw: akkar14(ctx, imts, 'repi', **kw),
    'AkkarEtAlRhyp2014':
        lambda ctx, imts, **kw: akkar14(ctx, imts, 'rhypo', **kw),
    'ChiouYoungs2008': cy08,
    'ToroEtAl2002': toro02,
    'Campbell2003': campbell03,
    'ToroEtAl2002SHARE': toro02_share,
    'Campbell2003SHARE': campbell03_share,
    'SadighEtAl1997': sadigh97,
}


def mean_stds(gsims, ctx, imts):
    """
    (4, G, M, N) float64 like ContextMaker.get_mean_stds
    (openquake/hazardlib/contexts.py:1250-1303); `gsims` is a list of
    (name, kwargs) pairs.
    """
    return np.stack([MODELS[name](ctx, imts, **kw) for name, kw in gsims],
                    axis=1)
