"""Ground motion fields given the epsilons: the arithmetic of
GmfComputer._compute (openquake/hazardlib/calc/gmf.py:476-525) on the device
(SURVEY.md section 8f row 4).  The random numbers stay with the caller, as in
the reference (numpy generators on the host): this module evaluates
``exp(mean + phi * within_eps + tau * between_eps)`` for all the IMTs, sites
and events at once, in float32 like the reference.  Spatial / cross-IMT
correlation of the residuals is applied by the caller to `within_eps` before
the call (the reference's non-spatial models do the same, gmf.py:441-457).

`GmfComputer` is the sampling side (compute_all, gmf.py:394-474): the
epsilons are numpy's own stream (PCG64 of default_rng(seed), reproduced bit
for bit on the device with jump-ahead) through the inverse CDF of the
truncated normal, the within-event residuals are correlated with the Cholesky
factor of the JB2009 matrix (cuSOLVER / cuBLAS through torch: plain library
calls, FP64), and set_max_min runs on the device.  Cross-IMT models other
than NoCrossCorrelation, amplification and secondary perils are not built.
"""
import ctypes
import numpy

from .imt import from_string

F32 = numpy.float32
TRUNCATION_THRESHOLD = 1e-9


def compute_gmf(mean_stds, imts, within_eps, between_eps=None,
                total_stddev_only=False, truncation_within=3.,
                truncation_between=3., device=None):
    """
    :param mean_stds: (4, M, N) mean, sig, tau, phi of one gsim
        (ContextMaker.get_mean_stds(...)[:, g])
    :param imts: the M intensity measure types (strings); MMI is not
        exponentiated
    :param within_eps: (M, N, E) float32
    :param between_eps: (E, M) float32 (not needed without between-event
        variability)
    :param total_stddev_only: the gsim defines only the total standard
        deviation (gmf.py:497-506)
    :returns: (M, N, E) float32 ground motion values
    """
    import torch
    from .contexts import get_handle, _device_index, _dev, _ptr
    h = get_handle(_device_index(device))
    dev = h.device
    ms = numpy.ascontiguousarray(mean_stds, F32)
    _4, M, N = ms.shape
    assert len(imts) == M
    within = numpy.ascontiguousarray(within_eps, F32)
    E = within.shape[2]
    assert within.shape == (M, N, E), within.shape
    if truncation_within <= TRUNCATION_THRESHOLD and \
            truncation_between <= TRUNCATION_THRESHOLD:
        mode = 0
    elif total_stddev_only:
        mode = 1
    else:
        mode = 2
        if between_eps is None:
            raise ValueError('between_eps is needed')
    t_ms, t_w = _dev(ms, dev), _dev(within, dev)
    t_b = _dev(numpy.ascontiguousarray(between_eps, F32), dev) \
        if mode == 2 else None
    if mode == 2:
        assert t_b.shape == (E, M), t_b.shape
    not_mmi = (ctypes.c_ubyte * M)(*[int(str(imt) != 'MMI') for imt in imts])
    out = torch.empty((M, N, E), dtype=torch.float32, device='cuda:%d' % dev)
    h.call('oqg_gmf', _ptr(t_ms), _ptr(t_w), _ptr(t_b), not_mmi, M, N, E, mode,
           _ptr(out))
    return out.cpu().numpy()


# --------------------------------------------------------------------------
# The sampling side: GmfComputer.compute_all (gmf.py:394-474)
# --------------------------------------------------------------------------
class JB2009CorrelationModel(object):
    """Jayaram and Baker (2009) spatial correlation of the within-event
    residuals (correlation_models/spatial/jayaram_baker_2009.py:33-69):
    exp(-3 d / b) with b from the period and `vs30_clustering`.  The matrix
    and its Cholesky factor live on the device, one per IMT (the `cache` of
    SpatialCorrelationModel.apply_correlation, correlation_models/base.py
    :322-336)."""
    def __init__(self, vs30_clustering):
        self.vs30_clustering = vs30_clustering
        self.cache = {}

    def decay_coeff(self, imt):
        imt = from_string(str(imt))
        if imt.string == 'PGV':
            period = 1.0
        elif imt.string == 'PGA':
            period = 0.
        elif imt.string.startswith('SA('):
            period = imt.period
            if not 0.01 <= period <= 10.:
                raise ValueError('JayaramBaker2009 supports SA periods from '
                                 '0.01 to 10 s, not %g s' % period)
        else:
            raise ValueError('JayaramBaker2009 does not support %s'
                             % imt.string)
        if period < 1:
            b = 40.7 - 15.0 * period if self.vs30_clustering \
                else 8.5 + 17.2 * period
        else:
            b = 22.0 + 3.7 * period
        return -3.0 / b

    def lower_triangle(self, lons, lats, imt, h):
        """Cholesky factor (N, N) float64 on the device of handle `h` for the
        complete site collection"""
        import torch
        from .contexts import _dev, _ptr
        key = str(imt)
        if key not in self.cache:
            N = len(lons)
            t_lon, t_lat = _dev(numpy.asarray(lons, float), h.device), _dev(
                numpy.asarray(lats, float), h.device)
            corr = torch.empty((N, N), dtype=torch.float64,
                               device='cuda:%d' % h.device)
            h.call('oqg_corr_matrix', _ptr(t_lon), _ptr(t_lat), N,
                   self.decay_coeff(imt), _ptr(corr))
            self.cache[key] = torch.linalg.cholesky(corr)   # cuSOLVER potrf
        return self.cache[key]


def get_eid_rlz(e0, n_occ, rlzs, scenario, seed):
    """source/rupture.py:785-800: the event IDs of a rupture and the
    realization of each event"""
    eid = numpy.arange(e0, e0 + n_occ, dtype=numpy.uint32)
    rlzs = numpy.asarray(rlzs)
    if scenario:   # the rlzs are distributed evenly
        rlz = rlzs[numpy.arange(n_occ) // (n_occ // len(rlzs))]
    else:          # event based: random_choice(rlzs, n, 0, seed)
        rng = numpy.random.default_rng(seed)
        cumsum = numpy.repeat(1. / len(rlzs), len(rlzs)).cumsum()
        rlz = rlzs[numpy.searchsorted(cumsum, rng.random(n_occ))]
    return eid, rlz


class GmfComputer(object):
    """
    The reference's GmfComputer (gmf.py:168-474) for one rupture, with the
    random numbers generated, correlated and turned into ground motion values
    on the device.

    :param ctx: the context rows of the rupture on its N sites (a recarray
        from ContextMaker.get_ctxs, field `sids` included)
    :param lons, lats: coordinates of the COMPLETE site collection (used by
        the correlation model; `ctx.sids` index them)
    :param cmaker: a ContextMaker
    :param rlzs_by_gsim: for each model of cmaker its realizations
    :param e0, n_occ, seed, rup_id: of the EBRupture
    :param scenario: how the realizations are spread over the events
    :param correlation_model: None or a JB2009CorrelationModel
    :param min_iml, max_iml: (M,) thresholds of set_max_min
    """
    def __init__(self, ctx, lons, lats, cmaker, rlzs_by_gsim, n_occ, seed,
                 e0=0, rup_id=0, scenario=False, correlation_model=None,
                 min_iml=None, max_iml=None, truncation_level_within=None,
                 truncation_level_between=None):
        if len(ctx) == 0:
            raise ValueError('No sites')
        self.ctx, self.cmaker = ctx, cmaker
        self.lons, self.lats = numpy.asarray(lons), numpy.asarray(lats)
        self.rlzs_by_gsim = [numpy.asarray(r) for r in rlzs_by_gsim]
        assert len(self.rlzs_by_gsim) == len(cmaker.gsims)
        self.n_occ, self.seed, self.e0, self.rup_id = n_occ, seed, e0, rup_id
        self.scenario = scenario
        self.correlation_model = correlation_model
        self.imts = [str(im) for im in cmaker.imtls]
        self.N, self.M = len(ctx), len(self.imts)
        self.tlw = cmaker.truncation_level if truncation_level_within is None \
            else truncation_level_within
        self.tlb = cmaker.truncation_level if truncation_level_between is None \
            else truncation_level_between
        self.min_iml = numpy.zeros(self.M) if min_iml is None \
            else numpy.asarray(min_iml, float)
        self.max_iml = numpy.full(self.M, numpy.nan) if max_iml is None \
            else numpy.asarray(max_iml, float)

    def init_eid_rlz_sig_eps(self):
        self.rlzs = numpy.concatenate(self.rlzs_by_gsim)
        self.eid, self.rlz = get_eid_rlz(self.e0, self.n_occ, self.rlzs,
                                         self.scenario, self.seed)
        self.E = len(self.eid)
        self.sig = numpy.zeros((self.E, self.M), F32)
        self.between_eps = numpy.zeros((self.E, self.M), F32)
        st = numpy.random.PCG64(self.seed).state['state']
        mask = (1 << 64) - 1
        self._state = (ctypes.c_uint64 * 2)(st['state'] >> 64,
                                            st['state'] & mask)
        self._inc = (ctypes.c_uint64 * 2)(st['inc'] >> 64, st['inc'] & mask)
        self._offset = 0     # uniforms of the generator consumed so far

    def _rvs(self, h, shape, tl):
        """stats.truncnorm(-tl, tl).rvs(shape, rng).astype(F32) on the device,
        continuing the stream of default_rng(seed)"""
        import torch
        from .contexts import _ptr
        n = int(numpy.prod(shape))
        out = torch.empty(shape, dtype=torch.float32,
                          device='cuda:%d' % h.device)
        h.call('oqg_truncnorm_rvs', self._state, self._inc, self._offset, n,
               float(max(tl, 1e-9)), _ptr(out))
        self._offset += n
        return out

    def compute_all(self):
        """
        :returns: a dictionary of arrays with the columns of the reference's
            data frame (one per IMT, then eid, sid, rlz), the rows with all
            zero values removed (strip_zeros, gmf.py:341-375)
        """
        import torch
        from .contexts import _dev, _ptr
        cm = self.cmaker
        h = cm.handle()
        dev = 'cuda:%d' % h.device
        self.init_eid_rlz_sig_eps()
        M, N = self.M, self.N
        not_mmi = (ctypes.c_ubyte * M)(*[int(im != 'MMI') for im in self.imts])
        c_max = self.max_iml.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        c_min = self.min_iml.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        zero_trunc = self.tlw <= TRUNCATION_THRESHOLD and \
            self.tlb <= TRUNCATION_THRESHOLD
        spatial = self.correlation_model is not None
        sids = numpy.asarray(self.ctx.sids)
        gmvs = []
        for g, (gs, rlzs) in enumerate(zip(cm.gsims, self.rlzs_by_gsim)):
            with cm.gmf_mon:
                ms = cm.get_4MN([self.ctx], gs).astype(F32)      # (4, M, N)
            idxs, = numpy.where(numpy.isin(self.rlz, rlzs))
            E = len(idxs)
            if E == 0:
                continue
            t_ms = _dev(numpy.ascontiguousarray(ms), h.device)
            # arrays of random numbers of shape (M, N, E) and (M, E), in the
            # order the reference draws them (gmf.py:441-446, 414-419)
            if self.tlw <= TRUNCATION_THRESHOLD:
                within = torch.zeros((M, N, E), dtype=torch.float32,
                                     device=dev)
            else:
                within = self._rvs(h, (M, N, E), self.tlw)
            if self.tlb <= TRUNCATION_THRESHOLD:
                between = torch.zeros((E, M), dtype=torch.float32, device=dev)
            else:
                between = self._rvs(h, (M, E), self.tlb).T.contiguous()
            self.between_eps[idxs] = between.cpu().numpy()
            total_only = gs.DEFINED_FOR_STANDARD_DEVIATION_TYPES == {'Total'}
            if zero_trunc:
                if spatial:
                    raise ValueError('truncation_level_within=0 requires '
                                     'no correlation model')
                mode = 0
            elif total_only:
                if spatial:
                    raise ValueError(
                        '%s requires a GSIM with inter/intra standard '
                        'deviations, %s has only the total one' % (
                            self.correlation_model.__class__.__name__, gs))
                mode = 1
                self.sig[idxs] = numpy.nan
            else:
                mode = 2
                self.sig[idxs] = ms[2].max(axis=1)         # tau.max() per IMT
                if spatial:
                    # apply_correlation: L @ (phi * eps), float64 product
                    # rounded to float32 (correlation_models/base.py:322-336)
                    mode = 3
                    res = t_ms[3][:, :, None] * within     # float32
                    for m, imt in enumerate(self.imts):
                        L = self.correlation_model.lower_triangle(
                            self.lons, self.lats, imt, h)
                        if len(sids) < len(self.lons):
                            tsid = torch.as_tensor(sids.astype(numpy.int64),
                                                   device=dev)
                            L = L[tsid][:, tsid]
                        within[m] = (L @ res[m].double()).float()
            out = torch.empty((M, N, E), dtype=torch.float32, device=dev)
            h.call('oqg_gmf', _ptr(t_ms), _ptr(within), _ptr(between),
                   not_mmi, M, N, E, mode, _ptr(out))
            h.call('oqg_gmf_max_min', _ptr(out), _ptr(t_ms[0]), c_max, c_min,
                   not_mmi, M, N, E)
            gmvs.append(out)
        return self.strip_zeros(gmvs)

    def strip_zeros(self, gmvs):
        import torch
        E = sum(t.shape[2] for t in gmvs)
        sids = numpy.asarray(self.ctx.sids, numpy.uint32)
        # build_eid_sid_rlz, gmf.py:95-105
        eids = numpy.concatenate([self.eid[self.rlz == r] for r in self.rlzs])
        rlzs = numpy.concatenate([self.rlz[self.rlz == r] for r in self.rlzs])
        data = {}
        if not gmvs:
            gmv = numpy.zeros((self.M, self.N, 0), F32)
        else:
            gmv = torch.cat(gmvs, dim=2).cpu().numpy()        # (M, N, E)
        ok = gmv.sum(axis=0).T.reshape(-1) > 0
        for m, imt in enumerate(self.imts):
            data[imt] = gmv[m].T.reshape(-1)[ok]
        data['eid'] = numpy.repeat(eids, self.N)[:E * self.N][ok]
        data['sid'] = numpy.tile(sids, E)[ok]
        data['rlz'] = numpy.repeat(rlzs, self.N)[:E * self.N][ok]
        return data
