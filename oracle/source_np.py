"""
CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path)
of the rupture generation of point sources, SURVEY.md section 8f row 1:

    TruncatedGRMFD.get_annual_occurrence_rates  mfd/truncated_gr.py:103-164
    PointSource.get_planin / get_planar         source/point.py:192-289
    build_planar / _build_corners               geo/surface/planar.py:107-251
    WC1994.get_median_area                      scalerel/wc1994.py
    PointMSR.get_median_area                    scalerel/point.py:40-52

Pinned against the reference's own PointSource.get_planar by
oracle/make_golden_sources.py (tests/golden/point_sources.npz).
"""
import math
import numpy as np
import geo_np


def round_half_away(x):
    """baselib/general.py:566-571 (the Python 2 rounding the MFDs use)"""
    return float(math.floor(x + math.copysign(0.5, x)))


def truncated_gr_rates(min_mag, max_mag, bin_width, a_val, b_val):
    """(mags, rates) of mfd/truncated_gr.py:103-164"""
    lo = round_half_away(min_mag / bin_width) * bin_width
    hi = round_half_away(max_mag / bin_width) * bin_width
    if lo != hi:
        lo += bin_width / 2.0
        hi -= bin_width / 2.0
    nbins = int(round_half_away((hi - lo) / bin_width)) + 1
    mags, rates, mag = [], [], lo
    for _ in range(nbins):
        mags.append(mag)
        rates.append(10 ** (a_val - b_val * (mag - bin_width / 2.0)) -
                     10 ** (a_val - b_val * (mag + bin_width / 2.0)))
        mag += bin_width
    return np.array(mags), np.array(rates)


def median_area(msr, mag, rake):
    """scalerel/wc1994.py get_median_area, scalerel/point.py:40-52"""
    if msr == 'PointMSR':
        return 1e-4
    if msr == 'PeerMSR':         # scalerel/peer.py:32-36
        return 10 ** (mag - 4.0)
    assert msr == 'WC1994', msr
    if abs(rake) <= 45 or abs(rake) >= 135:
        return 10.0 ** (-3.42 + 0.90 * mag)
    elif rake > 0:
        return 10.0 ** (-3.99 + 0.98 * mag)
    return 10.0 ** (-2.87 + 0.82 * mag)


def point_source_ruptures(src, shift_hypo=False):
    """
    All the ruptures of one point source in the order of
    PointSource.get_planar (magnitude, nodal plane, hypocentral depth):
    returns (planar dict of arrays (U, ...), mag (U,), occurrence_rate (U,)).

    src: dict with lon, lat, usd, lsd, rar, msr, mags, mag_rates,
         npd = [(prob, strike, dip, rake)], hdd = [(prob, depth)],
         optional dip_fracs (one per depth)
    """
    fracs = src.get('dip_fracs')
    corners, sdr, hypo, wrate, mags, orate = [], [], [], [], [], []
    for mag, mrate in zip(src['mags'], src['mag_rates']):
        for nprob, strike, dip, rake in src['npd']:
            area = median_area(src['msr'], mag, rake)
            for d, (dprob, dep) in enumerate(src['hdd']):
                frac = 0.5 if fracs is None else fracs[d]
                c, cdep = geo_np.rupture_corners(
                    src['usd'], src['lsd'], src['rar'], area, strike, dip,
                    src['lon'], src['lat'], dep, frac)
                corners.append(c)
                sdr.append((strike, dip, rake))
                hypo.append((src['lon'], src['lat'],
                             cdep if shift_hypo else dep))
                wrate.append(nprob * dprob)        # planar.wlr[:, 2]
                mags.append(mag)
                orate.append(mrate * (nprob * dprob))  # contexts.py:416
    pla = geo_np.planar_from_corners(np.array(corners), np.array(sdr),
                                     np.array(hypo), np.array(wrate))
    return pla, np.array(mags), np.array(orate)
