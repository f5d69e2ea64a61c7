"""
ORACLE (test infrastructure, never imported by the product): the whole
classical-PSHA hot path on the CPU in numpy, composed from geo_np (distances),
gmpe_np (mean / stddevs) and poe_np (PoEs, map update):

    planar ruptures x sites -> context rows -> (4,G,M,P) mean/std
        -> (P, L, G) float32 PoEs -> (N, L, G) float32 pne / rate map

It restates ContextMaker.get_ctxs for planar ruptures
(openquake/hazardlib/contexts.py:374-474), get_mean_stds (:1250-1303),
gen_poes (:1155-1203) and update / get_pmap (:1206-1247).  Pinned by
tests/golden/pipeline_*.npz, which hold the outputs of the reference itself
on the same inputs (oracle/make_golden.py).

This module is also the "port" CPU baseline timed by bench.py.
"""
import numpy as np
import geo_np
import gmpe_np
import poe_np

# what each model reads from a context row (Appendix B of SURVEY.md; the
# REQUIRES_* class attributes of the reference GMPE classes)
REQUIRES = {
    'BooreEtAl2014': (('rjb',), ('mag', 'rake'), ('vs30',)),
    'ChiouYoungs2014': (('rrup', 'rjb', 'rx'), ('mag', 'rake', 'dip', 'ztor'),
                        ('vs30', 'vs30measured', 'z1pt0')),
    'AbrahamsonEtAl2014': (('rrup', 'rjb', 'rx', 'ry0'),
                           ('mag', 'rake', 'dip', 'ztor', 'width'),
                           ('vs30', 'vs30measured', 'z1pt0')),
    'CampbellBozorgnia2014': (('rrup', 'rjb', 'rx'),
                              ('mag', 'rake', 'dip', 'ztor', 'width',
                               'hypo_depth'), ('vs30', 'z2pt5')),
    'BooreAtkinson2008': (('rjb',), ('mag', 'rake'), ('vs30',)),
    'AtkinsonBoore2006': (('rrup',), ('mag',), ('vs30',)),
    'AkkarEtAlRjb2014': (('rjb',), ('mag', 'rake'), ('vs30',)),
    'AkkarEtAlRepi2014': (('repi',), ('mag', 'rake'), ('vs30',)),
    'AkkarEtAlRhyp2014': (('rhypo',), ('mag', 'rake'), ('vs30',)),
    'ChiouYoungs2008': (('rrup', 'rjb', 'rx'), ('mag', 'rake', 'dip', 'ztor'),
                        ('vs30', 'vs30measured', 'z1pt0')),
    'ToroEtAl2002': (('rjb',), ('mag',), ()),
    'Campbell2003': (('rrup',), ('mag',), ()),
}


class Ctx(object):
    """Bag of equally long arrays with attribute access (recarray-like)"""
    def __init__(self, **arrays):
        vars(self).update(arrays)

    def __len__(self):
        return len(self.sids)

    def __getitem__(self, idx):
        return Ctx(**{k: v[idx] for k, v in vars(self).items()})


def required(gsims):
    """(distances, rupture params, site params) the models ask for; rrup is
    always computed on top of these (contexts.py:659) but is clamped by
    minimum_distance only if a model requires it (contexts.py:391-394)"""
    dists, rups, sites = set(), set(), set()
    for name, kw in gsims:
        d, r, s = REQUIRES[name]
        dists.update(d)
        rups.update(r)
        sites.update(s)
        if name == 'BooreEtAl2014' and kw.get('region', 'nobasin') != 'nobasin':
            sites.add('z1pt0')
    return dists, rups, sites


def site_xyz(sites):
    return geo_np.spherical_to_cartesian(
        sites['lon'], sites['lat'], sites['depth'])


def get_ctx(gsims, planar, mag, sites, magdist=None, minimum_distance=0,
            probs_occur=None):
    """
    Context rows for U planar ruptures x N sites, flattened rupture-major
    (row = u * N + n) and filtered by rrup < magdist[u] like
    _get_ctx_planar + genctxs_Pp (contexts.py:374-474).

    :param planar: dict from geo_np.planar_from_corners (U ruptures)
    :param mag: (U,) magnitudes; occurrence rates are planar['wlr'][:, 2]
    :param sites: dict of (N,) arrays sids, lon, lat, depth, vs30, ...
    :param magdist: None or (U,) maximum distances
    :param probs_occur: None or (U, K) occurrence pmf (NegativeBinomialTOM)
    """
    dists, rups, sitepars = required(gsims)
    U, N = len(mag), len(sites['sids'])
    xyz = site_xyz(sites)
    d = {}
    for name in sorted(dists | {'rrup'}):
        arr = geo_np.DIST[name](planar, xyz)
        if minimum_distance and name in dists:  # contexts.py:391-394
            arr = np.where(arr < minimum_distance, minimum_distance, arr)
        d[name] = arr

    def rup(x):
        return np.repeat(np.asarray(x)[:, None], N, axis=1)

    def sit(x):
        return np.repeat(np.asarray(x)[None, :], U, axis=0)
    cols = dict(d)
    cols['mag'] = rup(mag)
    cols['occurrence_rate'] = rup(planar['wlr'][:, 2])
    cols['rup_id'] = rup(np.arange(U, dtype=np.uint32))
    par = dict(width=planar['wlr'][:, 0], dip=planar['sdr'][:, 1],
               rake=planar['sdr'][:, 2], strike=planar['sdr'][:, 0],
               ztor=planar['corners'][:, 2, 0],
               hypo_depth=planar['hypo'][:, 2])
    for name in rups - {'mag'}:
        cols[name] = rup(par[name])
    for name in sitepars:
        cols[name] = sit(sites[name])
    cols['sids'] = sit(sites['sids'])
    if probs_occur is not None:
        cols['probs_occur'] = np.repeat(
            np.asarray(probs_occur)[:, None, :], N, axis=1)
    else:
        cols['probs_occur'] = np.zeros((U, N, 0))
    flat = {k: v.reshape((U * N,) + v.shape[2:]) for k, v in cols.items()}
    ctx = Ctx(**flat)
    if magdist is not None:
        ctx = ctx[ctx.rrup < np.repeat(np.asarray(magdist, float), N)]
    return ctx


def get_mean_stds(gsims, ctx, imts):
    """(4, G, M, P) float64, contexts.py:1250-1303"""
    out = gmpe_np.mean_stds(gsims, ctx, imts)
    return out


def get_poes(gsims, ctx, imtls, truncation_level):
    """
    (P, L, G) float32 as built by ContextMaker._gen_poes
    (contexts.py:1155-1186); also returns the (4, G, M, P) mean/stds
    """
    imts = list(imtls)
    loglevels = np.log(np.array([imtls[imt] for imt in imts], float))
    ms = get_mean_stds(gsims, ctx, imts)
    P, G = len(ctx), len(gsims)
    poes = np.zeros((P, loglevels.size, G), np.float32)
    for g in range(G):
        poes[:, :, g] = poe_np.set_poes(ms[:2, g], loglevels, truncation_level)
    return poes, ms


def update(arr, sid_to_row, gsims, ctx, imtls, truncation_level, itime,
           rates_mode=False, chunk=200_000):
    """
    ContextMaker.update (contexts.py:1232-1247) followed by
    MapArray.update_indep (map_array.py:443-455) on `arr` (N, L, G) float32.
    Rows are processed in the reference's order: by unique magnitude rounded
    to 3 digits (contexts.py:1196-1198), then in row order.
    """
    mags = np.round(ctx.mag, 3)
    for mag in np.unique(mags):
        sub = ctx[mags == mag]
        for start in range(0, len(sub), chunk):
            c = sub[start:start + chunk]
            poes, _ = get_poes(gsims, c, imtls, truncation_level)
            poe_np.update_pmap(arr, poes.astype(np.float64),
                               c.occurrence_rate, c.probs_occur,
                               sid_to_row[c.sids], itime, rates_mode)
    return arr


def get_pmap(gsims, ctx, sids, imtls, truncation_level, itime):
    """ContextMaker.get_pmap (contexts.py:1206-1218): PoE map (N, L, G) f32"""
    L = sum(len(v) for v in imtls.values())
    arr = np.ones((len(sids), L, len(gsims)), np.float32)
    sid_to_row = np.zeros(int(np.max(sids)) + 1, np.uint32)
    sid_to_row[sids] = np.arange(len(sids))
    update(arr, sid_to_row, gsims, ctx, imtls, truncation_level, itime)
    return 1. - arr   # MapArray.__invert__, map_array.py


def exact_rate_sum(gsims, ctx, sids, imtls, truncation_level, itime):
    """
    float64 re-accumulation of sum(rate * poe * itime) per (site, level,
    gsim) from the float32 PoEs: the order-independent quantity the GPU map
    is compared with (the reference's float32 read-modify-write is order
    dependent, SURVEY.md section 7 "hard parts").
    """
    L = sum(len(v) for v in imtls.values())
    acc = np.zeros((len(sids), L, len(gsims)))
    sid_to_row = np.zeros(int(np.max(sids)) + 1, np.int64)
    sid_to_row[sids] = np.arange(len(sids))
    poes, _ = get_poes(gsims, ctx, imtls, truncation_level)
    x = ctx.occurrence_rate[:, None, None] * poes.astype(np.float64) * itime
    np.add.at(acc, sid_to_row[ctx.sids], x)
    return acc
