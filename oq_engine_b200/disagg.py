"""Disaggregation: the hot function of openquake/hazardlib/calc/disagg.py
(`_disaggregate`, :188-243 -- "the crucial bit for performance") on the
device.

`disaggregate(ctx, mea, std, cmaker, g, iml2, bin_edges, epsstar, gp)` keeps
the argument meaning of the reference (minus the eps4 tuple, which is
derived from the epsilon edges and the truncation level as get_eps4 does,
and the monitors): ctx are the context rows of ONE site and ONE magnitude
bin with columns rrup, clon, clat, occurrence_rate; mea, std the (G, M, U)
arrays of ContextMaker.get_mean_stds; iml2 the (M, P) logarithmic
intensities; bin_edges = [mag, dist, lon, lat, eps] edges.  Returns the
6D matrix (D, Lo, La, E, M, P) of probabilities of exceedance.
Non-Poissonian ruptures (probs_occur) are an error: no device path.
"""
import ctypes
import numpy

from . import _lib


def disaggregate(ctx, mea, std, cmaker, g, iml2, bin_edges, epsstar=False,
                 gp=1., device=None):
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    po = getattr(ctx, 'probs_occur', None)
    if po is not None and numpy.size(po):
        raise NotImplementedError(
            'disaggregation of non-Poissonian ruptures has no device '
            'implementation')
    h = get_handle(_device_index(device if device is not None
                                 else getattr(cmaker, 'device', None)))
    dev = h.device
    iml2 = numpy.ascontiguousarray(iml2, float)
    M, P = iml2.shape
    dist_edges, lon_edges, lat_edges, eps_edges = (
        numpy.ascontiguousarray(e, float) for e in bin_edges[1:5])
    D, Lo, La, E = (len(e) - 1 for e in (dist_edges, lon_edges, lat_edges,
                                         eps_edges))
    U = len(ctx)
    mea_g = numpy.ascontiguousarray(numpy.asarray(mea)[g], float)
    std_g = numpy.ascontiguousarray(numpy.asarray(std)[g], float)
    assert mea_g.shape == (M, U), (mea_g.shape, M, U)
    t = {n: _dev(numpy.asarray(ctx[n] if not hasattr(ctx, n)
                               else getattr(ctx, n), float), dev)
         for n in ('rrup', 'clon', 'clat', 'occurrence_rate')} if U else {}
    t_mea, t_std = (_dev(mea_g, dev), _dev(std_g, dev)) if U else (None, None)
    out = torch.empty((D, Lo, La, E, M, P), dtype=torch.float64,
                      device='cuda:%d' % dev)
    dp = _lib.c_double_p
    h.call('oqg_disaggregate', _ptr(t.get('rrup')), _ptr(t.get('clon')),
           _ptr(t.get('clat')), _ptr(t.get('occurrence_rate')), U,
           _ptr(t_mea), _ptr(t_std), M, P, iml2.ctypes.data_as(dp),
           dist_edges.ctypes.data_as(dp), D, lon_edges.ctypes.data_as(dp), Lo,
           lat_edges.ctypes.data_as(dp), La, eps_edges.ctypes.data_as(dp), E,
           float(cmaker.truncation_level), float(cmaker.investigation_time),
           int(bool(epsstar)), float(gp), _ptr(out))
    return out.cpu().numpy()


# ------------------------------------------------------------ Disaggregator
class FarAwayRupture(Exception):
    """Raised if no rupture affects the site (contexts.py:277-278)"""


def to_rates(probs, itime=1., minrate=0.):
    # hazardlib/calc/mean_rates.py:28-41
    pnes = 1. - numpy.asarray(probs, float)
    pnes[pnes == 0] = 1E-45
    rates = - numpy.log(pnes) / itime
    rates[rates < 1E-12] = minrate
    rates[rates > 100.] = 100.
    return rates


def split_by_magbin(ctxt, mag_edges):
    """disagg.py:339-350: magnitude bin -> context rows"""
    ctx = ctxt[numpy.argsort(ctxt.mag)]
    fullmagi = numpy.searchsorted(mag_edges, ctx.mag) - 1
    fullmagi[fullmagi == -1] = 0  # magnitude on the edge
    return {magi: ctx[fullmagi == magi] for magi in numpy.unique(fullmagi)}


class Disaggregator(object):
    """
    Single-site disaggregation, mirroring the reference's driver
    (openquake/hazardlib/calc/disagg.py:353-538) for independent sources and
    for a group of mutually exclusive sources (`src_mutex`):
    `.init(magi)` selects the rows of a magnitude bin and evaluates mean and
    standard deviations with `cmaker.get_mean_stds`, `._disagg6D(imldic, g)`
    returns the (D, Lo, La, E, M, P) matrix of rates of gsim `g` through
    `disaggregate` (oqg_disaggregate), `.disagg_mag_dist_eps(imldic,
    rlz_weights)` the (Ma, D, E, M) matrix of rates weighted over the
    realizations.

    :param ctxs: context rows (a recarray or a list of them) with the clon,
        clat columns (few sites: ContextMaker.get_ctxs adds them)
    :param sitecol: a site collection of length 1
    :param bin_edges: [mag, dist, lon, lat, eps] edges; lon and lat may be
        dictionaries by site ID as in the reference
    :param rlzs_by_g: for each gsim of the cmaker the realizations using it
        (the reference keeps them as the values of cmaker.gsims)
    """
    def __init__(self, ctxs, sitecol, cmaker, bin_edges, rlzs_by_g,
                 epsstar=False):
        assert len(sitecol) == 1, sitecol
        self.sitecol = sitecol
        self.sid = sid = int(sitecol.sids[0])
        self.cmaker = cmaker
        self.epsstar = epsstar

        def of_site(e):
            return e[sid] if isinstance(e, dict) else e
        self.bin_edges = (bin_edges[0], bin_edges[1], of_site(bin_edges[2]),
                          of_site(bin_edges[3]), bin_edges[4])
        for i, name in enumerate(['Ma', 'D', 'Lo', 'La', 'E']):
            setattr(self, name, len(self.bin_edges[i]) - 1)
        self.mea, self.std = {}, {}   # magi -> array[G, M, U]
        self._key = {}                # magi -> (magi, mutex) of mea / std
        self.g_by_rlz = {}
        for g, rlzs in enumerate(rlzs_by_g):
            for rlz in rlzs:
                self.g_by_rlz[int(rlz)] = g
        ctxt = cmaker.recarray(ctxs) if isinstance(ctxs, (list, tuple)) \
            else ctxs
        ctx = ctxt[ctxt.sids == sid]
        if len(ctx) == 0:
            raise FarAwayRupture('No ruptures affecting site #%d' % sid)
        if 'clon' not in ctx.dtype.names:
            raise ValueError('the context rows need the clon, clat columns')
        self.ctx_by_magi = split_by_magbin(ctx, self.bin_edges[0])

    def init(self, magi, src_mutex=None):
        """
        :param src_mutex: empty, or a dictionary with keys src_id, weight and
            grp_probability for a group of mutually exclusive sources
            (disagg.py:736-742)
        :returns: the sum of the weights of the mutually exclusive sources
            reaching the magnitude bin, or 1
        """
        try:
            ctx = self.ctx_by_magi[magi]
        except KeyError:
            raise FarAwayRupture
        self.magi = magi
        self.src_mutex = dict(src_mutex) if src_mutex else {}
        key = (magi, bool(self.src_mutex))
        if self.src_mutex:
            # rows ordered by source, one slice per source (disagg.py:421-441)
            ctx = ctx[numpy.argsort(ctx.src_id, kind='stable')]
        self.ctx = ctx
        if self._key.get(magi) != key:
            self.mea[magi], self.std[magi] = self.cmaker.get_mean_stds(
                [ctx])[:2]
            self._key[magi] = key
        if not self.src_mutex:
            return 1.
        src_ids, start = numpy.unique(ctx.src_id, return_index=True)
        self.src_mutex['start'] = start
        self.src_mutex['stop'] = numpy.append(start[1:], len(ctx))
        present = set(int(s) for s in src_ids)
        pairs = [(int(s), w) for s, w in zip(self.src_mutex['src_id'],
                                             self.src_mutex['weight'])
                 if int(s) in present]
        if sorted(s for s, _ in pairs) != sorted(present):
            raise ValueError('a source of the context has no mutex weight')
        # the slices are in increasing src_id order; the weights follow the
        # order of src_mutex['src_id'] like in the reference (disagg.py:438)
        self.weights = [w for _, w in pairs]
        return sum(self.weights)

    def _disagg6D(self, imldic, g):
        # imldic: imt -> list of P intensities; returns rates
        iml2 = numpy.array([numpy.asarray(v, float) for v in imldic.values()])
        imlog2 = numpy.zeros_like(iml2)
        for m, imt in enumerate(imldic):
            with numpy.errstate(divide='ignore'):
                imlog2[m] = iml2[m] if str(imt) == 'MMI' else numpy.log(iml2[m])
        mea, std = self.mea[self.magi], self.std[self.magi]
        if not self.src_mutex:
            poes = disaggregate(self.ctx, mea, std, self.cmaker, g, imlog2,
                                self.bin_edges, self.epsstar, 1.)
            return to_rates(poes)
        # mutually exclusive sources: the weighted sum of the matrices of
        # probabilities of the sources (disagg.py:463-475); poes, not rates
        gp = self.src_mutex.get('grp_probability', 1.)
        poes = 0.
        for w, s1, s2 in zip(self.weights, self.src_mutex['start'],
                             self.src_mutex['stop']):
            poes = poes + w * disaggregate(
                self.ctx[s1:s2], mea[:, :, s1:s2], std[:, :, s1:s2],
                self.cmaker, g, imlog2, self.bin_edges, self.epsstar, gp)
        return poes

    def disagg_by_magi(self, imtls, rlzs, rwdic=None, src_mutex=None):
        """
        disagg.py:477-513: for each magnitude bin with ruptures a dictionary
        with keys trti, magi, sid, one 6D matrix of rates per realization and
        'mean' (the mean over the realizations weighted by `rwdic`, rlz ->
        weight) if `rwdic` is given
        """
        for magi in range(self.Ma):
            try:
                mw = self.init(magi, src_mutex)
            except FarAwayRupture:
                continue
            res = {'trti': getattr(self.cmaker, 'trti', 0), 'magi': magi,
                   'sid': self.sid}
            for rlz in rlzs:
                try:
                    g = self.g_by_rlz[int(rlz)]
                except KeyError:   # non-contributing realization
                    continue
                arr6D = self._disagg6D(imtls, g)
                res[rlz] = to_rates(arr6D) if src_mutex else arr6D
                if rwdic:
                    res['mean'] = res.get('mean', 0.) + arr6D * rwdic[rlz] * mw
            if rwdic and src_mutex:
                res['mean'] = to_rates(res['mean'])
            yield res

    def disagg_mag_dist_eps(self, imldic, rlz_weights, src_mutex=None):
        """
        :param imldic: a dictionary imt -> iml
        :param rlz_weights: the weights of the realizations
        :returns: a 4D matrix of rates of shape (Ma, D, E, M)
        """
        M = len(imldic)
        imtls = {imt: [iml] for imt, iml in imldic.items()}
        out = numpy.zeros((self.Ma, self.D, self.E, M))
        for magi in range(self.Ma):
            try:
                mw = self.init(magi, src_mutex)
            except FarAwayRupture:
                continue
            for rlz, g in self.g_by_rlz.items():
                mat5 = self._disagg6D(imtls, g)[..., 0]  # p = 0
                out[magi] += mat5.sum(axis=(1, 2)) * rlz_weights[rlz] * mw
        return to_rates(out) if src_mutex else out


# ------------------------------------------------------------ bin edges
KM_TO_DEGREES = 0.0089932          # geo/utils.py:47
DEGREES_TO_RAD = 0.01745329252     # geo/utils.py:48


def uniform_bins(min_value, max_value, bin_width):
    """disagg.py:299-310: bins of the given width including all values"""
    return bin_width * numpy.arange(
        int(numpy.floor(min_value / bin_width)),
        int(numpy.ceil(max_value / bin_width) + 1))


def lon_lat_bins(lon, lat, size_km, coord_bin_width):
    """disagg.py:63-83: lon, lat bin edges around a site"""
    import math
    nbins = numpy.ceil(size_km * KM_TO_DEGREES / coord_bin_width)
    delta_lon = min(size_km * KM_TO_DEGREES / math.cos(lat * DEGREES_TO_RAD),
                    180)
    delta_lat = min(size_km * KM_TO_DEGREES, 90)
    EPS = .001  # avoid discarding the last edge
    lon_bins = lon + numpy.arange(-delta_lon, delta_lon + EPS,
                                  2 * delta_lon / nbins)
    lat_bins = lat + numpy.arange(-delta_lat, delta_lat + EPS,
                                  2 * delta_lat / nbins)
    l1, l2 = lon_bins.min(), lon_bins.max()
    if l1 * l2 < 0 and abs(l1 - l2) > 180:    # geo/utils.py cross_idl
        lon_bins %= 360
    return lon_bins, lat_bins


def build_bin_edges(mags, maxdist, site_lon, site_lat, truncation_level,
                    n_epsilons, mag_bin_width, dist_bin_width,
                    coord_bin_width, bin_edges={}):
    """
    _build_bin_edges, disagg.py:86-142, for one site
    :returns: [mag, dist, lon, lat, eps] edges
    """
    if 'mag' in bin_edges:
        mag_edges = bin_edges['mag']
    else:
        mags = sorted(set(float(m) for m in mags))
        n1 = int(numpy.floor(mags[0] / mag_bin_width))
        n2 = int(numpy.ceil(mags[-1] / mag_bin_width))
        if n2 == n1 or mags[-1] >= round((mag_bin_width * n2), 3):
            n2 += 1
        mag_edges = mag_bin_width * numpy.arange(n1, n2 + 1)
    if 'dist' in bin_edges:
        dist_edges = bin_edges['dist']
    elif dist_bin_width:
        dist_edges = uniform_bins(0, maxdist, dist_bin_width)
    else:
        dist_edges = [0, maxdist]
    if 'lon' in bin_edges or 'lat' in bin_edges:
        lon_edges, lat_edges = bin_edges['lon'], bin_edges['lat']
    else:
        lon_edges, lat_edges = lon_lat_bins(site_lon, site_lat, maxdist,
                                            coord_bin_width)
    if 'eps' in bin_edges:
        eps_edges = bin_edges['eps']
    else:
        eps_edges = numpy.linspace(-truncation_level, truncation_level,
                                   n_epsilons + 1)
    return [numpy.asarray(e, float) for e in (
        mag_edges, dist_edges, lon_edges, lat_edges, eps_edges)]


def disaggregation(sources, site, imt, iml, gsim_by_trt, truncation_level,
                   n_epsilons=None, mag_bin_width=None, dist_bin_width=None,
                   coord_bin_width=None, maximum_distance=None, epsstar=False,
                   bin_edges={}, investigation_time=None):
    """
    The hazardlib function calc.disagg.disaggregation (disagg.py:544-669):
    the disaggregation matrix of `imt` exceeding `iml` at one site by
    magnitude, distance, longitude and latitude of the closest point,
    epsilon and tectonic region type.

    :param sources: point-like sources with a `tectonic_region_type`
    :param site: a SiteCollection with one site
    :param gsim_by_trt: tectonic region type -> gsim
    :param maximum_distance: an IntegrationDistance (the reference's
        `source_filter.integration_distance`); None = no filter: 1000 km for
        the contexts, the largest rupture distance for the bins
    :param investigation_time: the time span of the temporal occurrence
        model (the reference reads it from the first source)
    :returns: (mag, dist, lon, lat, eps edges, trts), 6D matrix of
        probabilities
    """
    from .contexts import ContextMaker, IntegrationDistance
    from .tom import PoissonTOM
    sources = list(sources)
    assert len(site) == 1, site
    trts = sorted(set(src.tectonic_region_type for src in sources))
    tom = getattr(sources[0], 'temporal_occurrence_model', None)
    if tom is None:
        if investigation_time is None:
            raise ValueError('the sources have no temporal occurrence model: '
                             'pass the investigation_time')
        tom = PoissonTOM(investigation_time)
    param = dict(imtls={str(imt): [iml]}, truncation_level=truncation_level,
                 investigation_time=tom.time_span)
    if maximum_distance is not None:
        param['maximum_distance'] = maximum_distance
    ctxs, cmaker, mags, dists = {}, {}, set(), []
    for trt in trts:
        # NB: like the reference (disagg.py:621, 645, 667) every ContextMaker
        # gets the models of ALL the tectonic region types and the matrix is
        # computed with g = 0, i.e. with the model of the first tectonic
        # region type in sorted order
        cmaker[trt] = cm = ContextMaker(
            trt, [gsim_by_trt[t] for t in trts], param)
        cm.tom = tom
        ctxs[trt] = cm.from_srcs(
            [s for s in sources if s.tectonic_region_type == trt], site)
        mags |= set(ctxs[trt].mag)
        dists.extend(ctxs[trt].rrup)
    if not dists:
        raise FarAwayRupture('No ruptures affecting the site')
    if maximum_distance is None:
        maxdist = max(dists)
    else:   # filters.upper_maxdist
        maxdist = max(maximum_distance[t][-1][1] for t in maximum_distance)
    edges = build_bin_edges(
        mags, maxdist, float(site.array['lon'][0]),
        float(site.array['lat'][0]), truncation_level, n_epsilons,
        mag_bin_width, dist_bin_width, coord_bin_width, bin_edges)
    shape = [len(e) - 1 for e in edges] + [len(trts)]
    matrix = numpy.zeros(shape)
    sid = int(site.sids[0])
    for t, trt in enumerate(trts):
        if len(ctxs[trt]) == 0:
            continue
        dis = Disaggregator(
            ctxs[trt], site, cmaker[trt],
            [edges[0], edges[1], {sid: edges[2]}, {sid: edges[3]}, edges[4]],
            rlzs_by_g=[[0]] * len(trts), epsstar=epsstar)
        for magi in range(dis.Ma):
            try:
                dis.init(magi)
            except FarAwayRupture:
                continue
            matrix[magi, ..., t] = dis._disagg6D({imt: [iml]}, 0)[..., 0, 0]
    return edges + [trts], 1. - numpy.exp(-matrix)     # to_probs
