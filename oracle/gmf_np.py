"""
TEST INFRASTRUCTURE (never imported by the product): numpy / scipy restatement
of the sampling side of the reference's ground motion fields,
GmfComputer.compute_all (openquake/hazardlib/calc/gmf.py:394-474), pinned on
the reference's own run (oracle/make_golden_gmf_all.py ->
tests/golden/gmf_all.npz).
"""
import numpy as np
from scipy import stats, special

F32 = np.float32
MULT = 0x2360ED051FC65DA44385DF649FCCF645      # PCG_DEFAULT_MULTIPLIER_128
M128 = (1 << 128) - 1
M64 = (1 << 64) - 1


def pcg64_uniform(state, inc, n, offset=0):
    """The first n doubles (after `offset`) of numpy's Generator.uniform for a
    PCG64 in the given state: numpy/random/src/pcg64/pcg64.h (setseq 128,
    XSL-RR 128/64 output of the stepped state) and next_double =
    (next_uint64 >> 11) / 2^53"""
    out = np.zeros(n)
    s = state
    for i in range(offset + n):
        s = (s * MULT + inc) & M128
        if i >= offset:
            hi, lo = s >> 64, s & M64
            x, rot = hi ^ lo, s >> 122
            o = ((x >> rot) | (x << ((64 - rot) & 63))) & M64
            out[i - offset] = (o >> 11) * (1.0 / 9007199254740992.0)
    return out


def truncnorm_ppf(u, tl):
    """ndtri(Phi(-tl) + u (Phi(tl) - Phi(-tl))): what the device evaluates;
    equal to scipy.stats.truncnorm(-tl, tl).ppf(u) to 1e-12"""
    pa = special.ndtr(-tl)
    return special.ndtri(pa + u * (special.ndtr(tl) - pa))


def jb2009_matrix(lons, lats, period, vs30_clustering=False):
    """jayaram_baker_2009.py:33-44 on Mesh.get_distance_matrix
    (geo/mesh.py:401-429, geodetic.geodetic_distance)"""
    lo, la = np.radians(lons), np.radians(lats)
    d = 2 * 6371.0 * np.arcsin(np.sqrt(
        np.sin((la[:, None] - la) / 2.) ** 2 +
        np.cos(la[:, None]) * np.cos(la) * np.sin((lo[:, None] - lo) / 2.) ** 2))
    if period < 1:
        b = 40.7 - 15.0 * period if vs30_clustering else 8.5 + 17.2 * period
    else:
        b = 22.0 + 3.7 * period
    return np.exp((-3.0 / b) * d)


def compute_all(mean_stds_by_gsim, rlzs_by_gsim, imts, periods, sids, lons,
                lats, n_occ, seed, tl=3., spatial=None, min_iml=None,
                max_iml=None, scenario=True, total_only=()):
    """
    :param mean_stds_by_gsim: list of (4, M, N) float32
    :returns: dict of columns (as strip_zeros), sig, eps
    """
    G = len(mean_stds_by_gsim)
    M, N = mean_stds_by_gsim[0].shape[1:]
    rlzs_all = np.concatenate(rlzs_by_gsim)
    eid = np.arange(n_occ, dtype=np.uint32)
    assert scenario
    rlz = rlzs_all[np.arange(n_occ) // (n_occ // len(rlzs_all))]
    rng = np.random.default_rng(seed)
    dist = stats.truncnorm(-tl, tl)
    sig = np.zeros((n_occ, M), F32)
    eps = np.zeros((n_occ, M), F32)
    min_iml = np.zeros(M) if min_iml is None else min_iml
    max_iml = np.full(M, np.nan) if max_iml is None else max_iml
    gmvs = []
    for g in range(G):
        ms = mean_stds_by_gsim[g]
        idxs, = np.where(np.isin(rlz, rlzs_by_gsim[g]))
        E = len(idxs)
        if E == 0:
            continue
        within = np.array([dist.rvs((N, E), rng).astype(F32)
                           for _ in range(M)])
        eps[idxs] = np.array([dist.rvs(E, rng) for _ in range(M)]).T
        res = np.zeros((M, N, E), F32)
        for m in range(M):
            mean, sg, tau, phi = ms[:, m]
            if g in total_only:
                v = mean[:, None] + sg[:, None] * within[m]
                sig[idxs, m] = np.nan
            else:
                wres = phi[:, None] * within[m]
                if spatial is not None:
                    L = np.linalg.cholesky(jb2009_matrix(
                        lons, lats, periods[m], spatial))
                    wres = (L @ wres).astype(F32)
                v = mean[:, None] + wres + tau[:, None] * eps[idxs, m]
                sig[idxs, m] = tau.max()
            res[m] = v if imts[m] == 'MMI' else np.exp(v)
        # set_max_min
        for m in range(M):
            med = ms[0, m] if imts[m] == 'MMI' else np.exp(ms[0, m])
            over = res[m] > max_iml[m]
            res[m][over] = np.broadcast_to(med[:, None], res[m].shape)[over]
        low = (res < min_iml[:, None, None]).all(axis=0)
        res[:, low] = 0
        gmvs.append(res)
    gmv = np.concatenate(gmvs, axis=2)
    Et = gmv.shape[2]
    ok = gmv.sum(axis=0).T.reshape(-1) > 0
    data = {imt: gmv[m].T.reshape(-1)[ok] for m, imt in enumerate(imts)}
    eids = np.concatenate([eid[rlz == r] for r in rlzs_all])
    rl = np.concatenate([rlz[rlz == r] for r in rlzs_all])
    data['eid'] = np.repeat(eids, N)[:Et * N][ok]
    data['sid'] = np.tile(sids, Et)[ok]
    data['rlz'] = np.repeat(rl, N)[:Et * N][ok]
    return data, sig, eps
