"""
CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path)
of the disaggregation hot function, SURVEY.md section 8f row 4:

    get_eps4               hazardlib/calc/disagg.py:176-186
    _disaggregate          hazardlib/calc/disagg.py:188-243 (poissonian lane)
    _disagg_eps            hazardlib/calc/disagg.py:246-256
    _build_disagg_matrix   hazardlib/calc/disagg.py:262-296
    _digitize_lons         hazardlib/calc/disagg.py:311-337

Pinned against the reference by oracle/make_golden_disagg.py.
"""
import numpy as np
from scipy.stats import truncnorm
import poe_np


def eps4(eps_edges, truncation_level):
    tn = truncnorm(-truncation_level, truncation_level)
    bands = tn.cdf(eps_edges[1:]) - tn.cdf(eps_edges[:-1])
    cum = np.array([bands[e:].sum() for e in range(len(bands))] + [0])
    return min(eps_edges), max(eps_edges), bands, cum


def lon_extent(lon1, lon2):
    # geo/utils.py get_longitudinal_extent
    return (lon2 - lon1 + 180) % 360 - 180


def digitize_lons(lons, lon_bins):
    if lon_bins[0] * lon_bins[-1] < 0 and abs(lon_bins[0] - lon_bins[-1]) > 180:
        idx = np.zeros_like(lons, dtype=int)
        for i in range(len(lon_bins) - 1):
            ok = lon_extent(lons, lon_bins[i + 1]) > 0
            if i != 0:
                ok &= lon_extent(lon_bins[i], lons) >= 0
            idx[ok] = i
        return np.array(idx)
    return np.digitize(lons, lon_bins) - 1


def disaggregate(rrup, clon, clat, rate, mea, std, iml2, dist_edges, lon_edges,
                 lat_edges, eps_edges, truncation_level, itime, epsstar, gp):
    """mea, std (M, U); iml2 (M, P) log intensities -> (D, Lo, La, E, M, P)"""
    U = len(rrup)
    min_eps, max_eps, bands, cum = eps4(eps_edges, truncation_level)
    E = len(bands)
    M, P = iml2.shape
    phi_b = poe_np.ndtr(truncation_level)
    poes = np.zeros((U, E, M, P))
    for m in range(M):
        for p in range(P):
            iml = iml2[m, p]
            if iml == -np.inf:
                continue
            lvls = (iml - mea[m]) / std[m]
            idxs = np.searchsorted(eps_edges, lvls)
            sf = poe_np.truncnorm_sf(phi_b, lvls)
            if epsstar:
                ok = (lvls >= min_eps) & (lvls < max_eps)
                poes[ok, idxs[ok] - 1, m, p] = gp * sf[ok]
            else:
                res = np.zeros((U, E))
                for e in range(E):
                    res[idxs <= e, e] = bands[e]
                    inside = idxs == e + 1
                    res[inside, e] = sf[inside] - cum[idxs[inside]]
                poes[:, :, m, p] = gp * res
    pnes = np.exp(-rate[:, None, None, None] * poes * itime)
    D, Lo, La = len(dist_edges) - 1, len(lon_edges) - 1, len(lat_edges) - 1
    di = np.digitize(rrup, dist_edges) - 1
    lo = digitize_lons(clon, lon_edges)
    la = np.digitize(clat, lat_edges) - 1
    di[di == D] = D - 1
    lo[lo == Lo] = Lo - 1
    la[la == La] = La - 1
    mat = np.ones((D, Lo, La, E, M, P))
    for u in range(U):
        mat[di[u], lo[u], la[u]] *= pnes[u]   # negative indices wrap
    return 1. - mat
