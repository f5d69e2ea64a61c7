"""
CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path)
of the post-processing of the rates map, SURVEY.md section 8f row 3:

    to_rates / to_probs           hazardlib/calc/mean_rates.py:28-50
    MapGetter.get_hcurve          calculators/getters.py:351-364
    mean_curve / std_curve /
    quantile_curve                hazardlib/stats.py:156-214
    build_stat_curve              calculators/getters.py:47-76
    compute_hazard_maps           hazardlib/map_array.py:94-141
    postclassical                 calculators/classical.py:258-325

Pinned against those functions of the reference by
oracle/make_golden_post.py (tests/golden/post.npz).
"""
import numpy as np

CUTOFF = 1E-12
EPSILON = 1E-30


def to_rates(probs, itime=1., minrate=0.):
    pnes = 1. - np.asarray(probs, float)
    pnes[pnes == 0] = 1E-45
    rates = - np.log(pnes) / itime
    rates[rates < CUTOFF] = minrate
    rates[rates > 100.] = 100.
    return rates


def to_probs(rates, itime=1.):
    return 1. - np.exp(- np.asarray(rates, float) * itime)


def hcurves_from_rates(rates, trt_rlzs, R):
    """(N, L, Gt) rates -> (N, L, R) PoEs, getters.py:351-364"""
    N, L, Gt = rates.shape
    r0 = np.zeros((N, L, R))
    for g, rlzs in enumerate(trt_rlzs):
        for rlz in np.asarray(rlzs) % 2 ** 24:
            r0[:, :, rlz] += rates[:, :, g]
    return to_probs(r0)


def mean_curve(values, weights):
    values, weights = np.asarray(values, float), np.asarray(weights, float)
    return np.tensordot(weights, values, 1) / weights.sum()


def std_curve(values, weights):
    values, weights = np.asarray(values, float), np.asarray(weights, float)
    m = mean_curve(values, weights)
    return np.sqrt(np.tensordot(weights, (m - values) ** 2, 1))


def quantile_curve(q, curves, weights):
    curves, weights = np.asarray(curves, float), np.asarray(weights, float)
    out = np.zeros(curves.shape[1:])
    for idx in np.ndindex(*out.shape):
        c = curves[(slice(None),) + idx]
        # cw.sort(order='c'): numpy breaks ties with the remaining field w
        order = np.lexsort((weights, c))
        out[idx] = np.interp(q, np.cumsum(weights[order]), c[order])
    return out


def stat_curve(hcurve, kind, q, weights, M, use_rates=False):
    """
    build_stat_curve, getters.py:47-76.  hcurve (L, R); weights (R, W) with
    W = 1 (one weight per realization) or M + 1 (IMT-dependent weights: one
    column per IMT, the last column is the default weight)
    """
    L, R = hcurve.shape
    L1 = L // M
    poes = hcurve.T
    fn = {'mean': lambda v, w: mean_curve(v, w),
          'std': lambda v, w: std_curve(v, w),
          'max': lambda v, w: np.max(v, axis=0),   # stats.py:228-237
          'quantile': lambda v, w: quantile_curve(q, v, w)}[kind]
    out = np.zeros(L)
    for m in range(M):
        slc = slice(m * L1, (m + 1) * L1)
        ws = weights[:, m] if weights.shape[1] > 1 else weights[:, -1]
        if not ws.sum():
            continue
        if use_rates:
            out[slc] = to_probs(fn(to_rates(poes[:, slc]), ws))
        else:
            out[slc] = fn(poes[:, slc], ws)
    return out


def compute_hazard_maps(curves, imls, poes):
    """
    map_array.py:94-141: (N, L1) curves -> (N, P) intensity levels.

    The reference takes the logarithm of a Python list of the curve's
    elements: for a float32 curve (MapArray arrays are float32) that is a
    float32 logarithm, unless an element was replaced by the Python float
    EPSILON, which makes the whole list float64.  Restated as such.
    """
    curves = np.asarray(curves)
    N, L1 = curves.shape
    log_poes = np.log(poes)
    hmap = np.zeros((N, len(poes)))
    with np.errstate(divide='ignore'):
        limls = np.log(np.array(imls[::-1], float))
    for n, curve in enumerate(curves):
        rev = curve[::-1]
        if (rev < EPSILON).any():
            rev = np.maximum(rev.astype(float), EPSILON)
        log_curve = np.log(rev)
        for p, lp in enumerate(log_poes):
            if lp > log_curve[-1]:
                continue
            hmap[n, p] = np.exp(np.interp(lp, log_curve, limls))
    return hmap
