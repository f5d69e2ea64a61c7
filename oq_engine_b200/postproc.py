"""Post-processing of the rates map on the device (SURVEY.md section 8f row
3): hazard curves per realization, statistics over the realizations and
hazard maps.

Mirrors `postclassical` (openquake/calculators/classical.py:258-325), which
for every site calls MapGetter.get_hcurve (calculators/getters.py:351-364)
and build_stat_curve (getters.py:47-76) with the statistics of
hazardlib/stats.py:156-214, then calc.make_hmaps -> compute_hazard_maps
(hazardlib/map_array.py:94-141).  Here all the sites go through one kernel
(`oqg_postclassical`) and the maps through another (`oqg_hazard_maps`).
"""
import ctypes
import numpy

from .map_array import MapArray

TWO24 = 2 ** 24
KINDS = {'mean': 0, 'std': 1, 'quantile': 2, 'max': 3}


def _stat_spec(name):
    """'mean' | 'std' | 'max' | 'quantile-0.05' -> (kind, q); like the keys of
    OqParam.hazard_stats()"""
    if name in ('mean', 'std', 'max'):
        return KINDS[name], 0.
    if name.startswith('quantile-'):
        q = float(name.split('-', 1)[1])
        if not 0. <= q <= 1.:
            raise ValueError('quantile %s out of range' % q)
        return KINDS['quantile'], q
    raise NotImplementedError('statistic %r has no device implementation'
                              % name)


def _handle(device):
    from .contexts import get_handle, _device_index
    return get_handle(_device_index(device))


def compute_hazard_maps(curves, imls, poes, device=None):
    """
    hazardlib/map_array.py:94-141: curves (N, L1) PoEs -> (N, P) intensity
    levels by log-log interpolation; 0 where the PoE is above the curve.
    """
    curves = numpy.asarray(curves, numpy.float32)
    N, L1 = curves.shape
    if L1 != len(imls):
        raise ValueError('The curves have %d levels, %d were passed' %
                         (L1, len(imls)))
    return compute_hmaps(curves[:, None, :], {'_': imls}, poes,
                         device)[:, 0]


def compute_hmaps(curvesNML, imtls, poes, device=None):
    """hazardlib/map_array.py:144-158: (N, M, L1) -> (N, M, P)"""
    import torch
    from .contexts import _dev, _ptr
    h = _handle(device)
    curves = numpy.ascontiguousarray(curvesNML, numpy.float32) \
        if not isinstance(curvesNML, torch.Tensor) else curvesNML
    N, M, L1 = curves.shape
    assert M == len(imtls), (M, len(imtls))
    imls = numpy.array([numpy.asarray(v, float) for v in imtls.values()])
    assert imls.shape == (M, L1), imls.shape
    P = len(poes)
    t_c = _dev(curves, h.device, torch.float32)
    t_i = _dev(imls, h.device)
    t_p = _dev(numpy.asarray(poes, float), h.device)
    t_o = torch.empty((N, M, P), dtype=torch.float32,
                      device='cuda:%d' % h.device)
    h.call('oqg_hazard_maps', _ptr(t_c), N, M, L1, _ptr(t_i), P, _ptr(t_p),
           _ptr(t_o))
    return t_o.cpu().numpy()


def calc_mean_rates(rmap, gweights, wget, imtls, device=None,
                    as_tensor=False):
    """
    hazardlib/calc/mean_rates.py:62-74: the mean hazard rates, shape
    (N, M, L1) float64.

    :param rmap: a rates MapArray (N, L, Gt) (or its array, or a CUDA tensor)
    :param gweights: (Gt,) weights of the gsim columns, or a 2D array of
        IMT-dependent weights from which `wget(gweights, imt)` selects the
        (Gt,) weights of each IMT
    :param wget: None for 1D weights
    """
    import torch
    from .contexts import _dev, _ptr
    h = _handle(device)
    arr = rmap if isinstance(rmap, (numpy.ndarray, torch.Tensor)) \
        else rmap.array
    N, L, Gt = arr.shape
    M = len(imtls)
    if L % M:
        raise ValueError('%d levels for %d IMTs' % (L, M))
    gweights = numpy.asarray(gweights, float)
    if gweights.ndim == 1:
        w = gweights[None]
    else:
        w = numpy.array([numpy.asarray(wget(gweights, imt), float)
                         for imt in imtls])
    if w.shape[1] != Gt:
        raise ValueError('%d weights for %d gsim columns' % (w.shape[1], Gt))
    t_r = _dev(arr, h.device, torch.float32)
    t_w = _dev(numpy.ascontiguousarray(w), h.device)
    out = torch.empty((N, M, L // M), dtype=torch.float64,
                      device='cuda:%d' % h.device)
    h.call('oqg_mean_rates', _ptr(t_r), N, L, Gt, M, _ptr(t_w), len(w),
           _ptr(out))
    return out if as_tensor else out.cpu().numpy()


def postclassical(rmap, trt_rlzs, weights, imtls, hstats=('mean',),
                  individual_rlzs=False, poes=(), use_rates=False,
                  device=None, grouping=True):
    """
    :param rmap: a rates MapArray (N, L, Gt) (or its array)
    :param trt_rlzs: for each gsim column g the realizations it belongs to
        (possibly with the TRT index in the high bits, as in the reference)
    :param weights: (R,) realization weights, or (R, M + 1) IMT-dependent
        weights (one column per IMT, the last one the default weight)
    :param imtls: dict imt -> levels (M IMTs of L1 levels, L = M * L1)
    :param hstats: names of the statistics: 'mean', 'std', 'quantile-X'
    :param grouping: False = do not group the realizations with the same
        gsim columns (tests and measurements; the results are the same)
    :returns: a dictionary kind -> list of MapArray (N, M, L1) [or (N, M, P)
        for the hazard maps], kinds 'hcurves-rlzs', 'hcurves-stats',
        'hmaps-rlzs', 'hmaps-stats' as in the reference
    """
    import torch
    from .contexts import _dev, _ptr
    h = _handle(device)
    dev = h.device
    cuda = 'cuda:%d' % dev
    if isinstance(rmap, MapArray):
        assert rmap.rates, 'postclassical needs a map of rates'
        sids, arr = rmap.sids, rmap.array
    else:
        arr = rmap
        sids = numpy.arange(len(arr), dtype=numpy.uint32)
    N, L, Gt = arr.shape
    M = len(imtls)
    L1 = L // M
    assert M * L1 == L, (M, L1, L)
    rlzs = [numpy.asarray(t, numpy.int64) % TWO24 for t in trt_rlzs]
    assert len(rlzs) == Gt, (len(rlzs), Gt)
    weights = numpy.asarray(weights, float)
    R = len(weights)
    if max(int(r.max()) for r in rlzs if len(r)) >= R:
        raise ValueError('realization index out of range')
    if weights.ndim == 1 or weights.shape[1] == 1:
        wdev = weights.reshape(R, -1)[:, -1][None]         # (1, R)
    else:
        if weights.shape[1] != M + 1:
            raise ValueError('IMT-dependent weights must have M + 1 columns')
        wdev = weights[:, :M].T                            # (M, R)
    wdev = numpy.ascontiguousarray(wdev)
    wsum = wdev.sum(axis=1)
    specs = [_stat_spec(name) for name in hstats]
    S = len(specs)
    if R > 4096 and any(k == KINDS['quantile'] for k, _ in specs):
        # realizations with the same gsim columns share their curve: the
        # library works on the distinct groups
        member = numpy.zeros((R, Gt), bool)
        for g, r in enumerate(rlzs):
            member[r, g] = True
        ngroups = len(numpy.unique(member, axis=0)) if grouping else R
        if ngroups > 256:
            raise ValueError(
                'quantiles need at most 4096 realizations or 256 distinct '
                'sets of gsim columns, got %d / %d (mean, std and max have '
                'no limit)' % (R, ngroups))
    do_rlzs = R == 1 or individual_rlzs
    t_rates = _dev(arr, dev, torch.float32)
    g_off = numpy.concatenate([[0], numpy.cumsum([len(r) for r in rlzs])]
                              ).astype(numpy.int32)
    g_rlz = numpy.concatenate(rlzs).astype(numpy.int32)
    t_off, t_rlz, t_w = _dev(g_off, dev), _dev(g_rlz, dev), _dev(wdev, dev)
    t_r = torch.empty((R, N, L), dtype=torch.float32, device=cuda) \
        if do_rlzs else None
    t_s = torch.empty((S, N, L), dtype=torch.float32, device=cuda) \
        if S else None
    kinds = (ctypes.c_int32 * max(S, 1))(*[k for k, _ in specs])
    qs = (ctypes.c_double * max(S, 1))(*[q for _, q in specs])
    ws = (ctypes.c_double * len(wsum))(*wsum)
    h.call('oqg_set_post_grouping', int(bool(grouping)))
    h.call('oqg_postclassical', _ptr(t_rates), N, L, Gt, M, _ptr(t_off),
           _ptr(t_rlz), R, _ptr(t_w), len(wsum), ws, S, kinds, qs,
           int(use_rates), _ptr(t_r), _ptr(t_s))
    out = {}

    def maps(t, P_or_L1):
        res = []
        for a in t.cpu().numpy():
            m = MapArray(sids, M, P_or_L1)
            m.array = a.reshape(N, M, P_or_L1)
            res.append(m)
        return res
    P = len(poes)
    if P:
        t_i = _dev(numpy.array([numpy.asarray(v, float)
                                for v in imtls.values()]), dev)
        t_p = _dev(numpy.asarray(poes, float), dev)
    for kind, t in (('rlzs', t_r), ('stats', t_s)):
        if t is None:
            continue
        out['hcurves-' + kind] = maps(t, L1)
        if P:
            K = t.shape[0] * N
            t_o = torch.empty((t.shape[0], N, M * P), dtype=torch.float32,
                              device=cuda)
            h.call('oqg_hazard_maps', _ptr(t), K, M, L1, _ptr(t_i), P,
                   _ptr(t_p), _ptr(t_o))
            out['hmaps-' + kind] = maps(t_o, P)
    return out


def last_post_path(device=None):
    """(path, groups) of the last postclassical call on the device: path 0 =
    groups in registers / shared memory, 1 = streaming, 2 = one CTA per
    element, 3 = generic kernel"""
    h = _handle(device)
    n = ctypes.c_int(0)
    from ._lib import lib
    path = lib().oqg_last_post_path(h._h, ctypes.byref(n))
    return path, n.value


rates_dt = numpy.dtype([('sid', numpy.uint32), ('lid', numpy.uint16),
                        ('gid', numpy.uint16), ('rate', numpy.float32)])
slice_dt = numpy.dtype([('idx', numpy.uint32), ('start', int), ('stop', int)])


def rates_to_records(rmap, gid=None, num_chunks=1, level0=0, device=None):
    """
    The compaction of a rates map into the `_rates` records the reference
    stores: MapArray.to_array (hazardlib/map_array.py:396-405, to_rates
    :520-540) followed by the chunk grouping of `_store`
    (calculators/classical.py:58-101, gen_chunks map_array.py:49-64), as one
    device stream compaction (oqg_rates_to_records).

    :param rmap: a rates MapArray (N, L, G) (host or device resident)
    :param gid: the G gsim IDs of the columns (default: rmap.gid or 0..G-1)
    :param num_chunks: the records are grouped by chunk = sid % num_chunks
    :returns: (records, slice_by_idx): a structured array of rates_dt with
        the elements of rate > 0 ordered by (chunk, gsim column, site,
        level), and an array of (idx, start, stop) with the slice of every
        chunk that has records of sites in it
    """
    import torch
    from .contexts import _dev, _ptr
    assert rmap.rates, 'a map of rates is needed'
    h = _handle(device if device is not None else rmap.device)
    dev = h.device
    N, L, G = rmap.shape
    if gid is None:
        gid = getattr(rmap, 'gid', None)
    gid = numpy.arange(G) if gid is None else numpy.asarray(gid)
    assert len(gid) == G, (len(gid), G)
    sids = numpy.asarray(rmap.sids, numpy.uint32)
    chunk = sids % num_chunks if num_chunks > 1 else numpy.zeros(N, int)
    # rows g * N + s in output order: by chunk, then column, then site
    ch = numpy.repeat(chunk[None, :], G, axis=0).ravel()      # (G * N,)
    rows = numpy.argsort(ch, kind='stable').astype(numpy.int64)
    t_rates = _dev(rmap.array, dev, torch.float32)
    t_sids, t_gid = _dev(sids, dev), _dev(gid.astype(numpy.int32), dev)
    t_rows = _dev(rows, dev)
    total = ctypes.c_int64(0)
    R = len(rows)
    if R == 0:
        return numpy.zeros(0, rates_dt), numpy.zeros(0, slice_dt)
    # counting pass, then the records in a buffer of the exact size
    h.call('oqg_rates_to_records', _ptr(t_rates), N, L, G, _ptr(t_sids),
           _ptr(t_gid), int(level0), _ptr(t_rows), R, None, 0,
           ctypes.byref(total))
    n = total.value
    out = torch.empty((max(n, 1), 12), dtype=torch.uint8,
                      device='cuda:%d' % dev)
    h.call('oqg_rates_to_records', _ptr(t_rates), N, L, G, _ptr(t_sids),
           _ptr(t_gid), int(level0), _ptr(t_rows), R, _ptr(out), n,
           ctypes.byref(total))
    torch.cuda.current_stream(dev).synchronize()
    recs = out[:n].cpu().numpy().reshape(-1).view(rates_dt)
    # slices of the chunks: the records are sorted by chunk
    rch = recs['sid'] % num_chunks if num_chunks > 1 else numpy.zeros(
        n, numpy.uint32)
    idx = numpy.unique(chunk)
    start = numpy.searchsorted(rch, idx, 'left')
    stop = numpy.searchsorted(rch, idx, 'right')
    slices = numpy.zeros(len(idx), slice_dt)
    slices['idx'], slices['start'], slices['stop'] = idx, start, stop
    return recs, slices
