"""ContextMaker: the reference's operator surface for the classical-PSHA hot
path (openquake/hazardlib/contexts.py:507-1303) over the CUDA kernels.

    get_ctxs        :1051-1099 (planar ruptures, :374-474)
    get_mean_stds   :1250-1303
    gen_poes        :1155-1203
    update/get_pmap :1206-1247
    hazard_planar   the fused equivalent of RmapMaker._make_src_indep
                    (:1548-1579): ruptures x sites -> map with no context rows
                    materialised

The host side is Python as in the reference; PyTorch tensors are used only as
device buffers.  There is no CPU fallback: an unsupported GMPE, a missing
library or a missing CUDA device raises.
"""
import os
import ast
import ctypes
import numpy
from scipy.special import ndtr

from . import _lib
from .imt import from_string
from .gsim import GMPE, get_gsim
from .map_array import MapArray
from .tom import PoissonTOM
from .performance import Monitor
from .planar import planar_array_dt, as_flat

U32 = numpy.uint32
F32 = numpy.float32
MINMAG, MAXMAG = 2.5, 10.2   # calc/filters.py

_handles = {}


def default_device():
    return int(os.environ.get('OQ_GPU_DEVICE',
                              os.environ.get('LOCAL_RANK', 0)))


def get_handle(device=None):
    """One oqg handle per process and device, bound to torch's current
    stream on that device so that torch copies and kernels are ordered"""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('no CUDA device: oq-engine_b200 has no CPU path')
    device = default_device() if device is None else device
    if device not in _handles:
        _handles[device] = _lib.Handle(device)
    h = _handles[device]
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.lib().oqg_set_stream(h._h, ctypes.c_void_p(stream))
    return h


def _device_index(device):
    """None | int | 'cuda:1' -> None | int"""
    if device is None or isinstance(device, int):
        return device
    s = str(device)
    return int(s.split(':')[1]) if ':' in s else 0


def _dev(arr, device, dtype=None):
    """numpy array (or torch tensor) -> contiguous torch CUDA tensor"""
    import torch
    if isinstance(arr, torch.Tensor):
        t = arr.to('cuda:%d' % device)
        if dtype is not None:
            t = t.to(dtype)
        return t.contiguous()
    arr = numpy.array(arr, order='C')   # copy: fields of records are strided
    if arr.dtype == bool:
        arr = arr.view(numpy.uint8)
    if arr.dtype == numpy.uint32:
        arr = arr.view(numpy.int32)
    t = torch.from_numpy(arr).to('cuda:%d' % device, non_blocking=False)
    if dtype is not None:
        t = t.to(dtype)
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


class IntegrationDistance(dict):
    """trt -> [(mag, dist), ...] (calc/filters.py:196-225); calling it with a
    TRT gives a function mag -> maximum distance"""
    @classmethod
    def new(cls, value):
        self = cls()
        if value == 'magdist':
            self['default'] = [(3, 0), (6, 150), (10, 600)]
            return self
        if isinstance(value, str):
            # '300' or '[(4., 50), (8., 200.)]' or a dict literal, parsed
            # like the reference's floatdict (literal_eval, never eval)
            value = ast.literal_eval(value)
        if isinstance(value, dict):
            items = value
        else:
            items = {'default': value}
        for trt, val in items.items():
            if isinstance(val, (list, tuple)):
                pairs = [tuple(it) for it in val]
                if len(set(pairs)) < len(pairs):
                    raise ValueError('Found duplicates in %s' % pairs)
                if pairs != sorted(pairs):
                    raise ValueError('%s is not ordered' % pairs)
                for mag, dist in pairs:
                    if mag < MINMAG or mag > MAXMAG:
                        raise ValueError('Invalid magnitude %s' % mag)
                self[trt] = pairs
            else:
                assert val >= 0, val
                self[trt] = [(MINMAG, val), (MAXMAG, val)]
        return self

    def __call__(self, trt):
        pairs = self.get(trt, self.get('default'))
        return magdepdist(pairs)

    def cut(self, min_mag_by_trt):
        """
        The engine's `minimum_magnitude` (calc/filters.py:228-259, applied by
        OqParam to maximum_distance): the (mag, dist) pairs of every tectonic
        region type start at its minimum magnitude, with the distance
        interpolated there, so that smaller magnitudes get distance zero and
        their ruptures are dropped by the distance filter.  In place.
        """
        lower = dict(min_mag_by_trt)
        if 'default' not in self:
            self['default'] = max(
                self.values(), key=lambda prs: max(d for _m, d in prs))
        if 'default' not in lower:
            lower['default'] = min(lower.values())
        for trt in set(self) | set(lower):
            mmin = lower.get(trt, lower['default'])
            if not mmin:
                continue
            pairs = self.get(trt, self['default'])
            at_min = float(magdepdist(pairs)(mmin))
            kept = [(m, d) for m, d in pairs if m >= mmin]
            self[trt] = kept if kept and kept[0][0] <= mmin else [
                (mmin, at_min)] + kept


class magdepdist(object):
    """Linear interpolation mag -> distance (calc/filters.py:183-190:
    interp1d(bounds_error=False, fill_value=0.)): the distance is ZERO for a
    magnitude outside the pairs, so such ruptures are dropped by the
    `rrup < magdist` filter, as in the reference"""
    def __init__(self, pairs):
        self.mags = numpy.array([m for m, d in pairs], float)
        self.dists = numpy.array([d for m, d in pairs], float)

    def __call__(self, mag):
        return numpy.interp(mag, self.mags, self.dists, left=0., right=0.)


def _interp(param, name, trt):
    mdd = param.get(name)
    if mdd is None:
        return magdepdist([(MINMAG, 1000), (MAXMAG, 1000)])
    if isinstance(mdd, IntegrationDistance):
        return mdd(trt)
    if isinstance(mdd, dict):
        return IntegrationDistance.new(mdd)(trt)
    if callable(mdd):
        return mdd
    return magdepdist([(MINMAG, float(mdd)), (MAXMAG, float(mdd))])


class PlanarRuptures(object):
    """
    The rupture side of the path: U planar ruptures with their magnitudes
    (what `src.get_planar()` feeds to genctxs_Pp, contexts.py:443-474).

    :param planar: (U, 3) record array of planar_array_dt (wlr[:, 2] is the
        occurrence rate)
    :param mag: (U,) magnitudes
    """
    def __init__(self, planar, mag, src_id=0):
        assert planar.dtype == planar_array_dt, planar.dtype
        self.planar = planar
        self.mag = numpy.asarray(mag, float)
        self.src_id = src_id
        assert len(self.mag) == len(planar)

    def __len__(self):
        return len(self.mag)

    @property
    def rate(self):
        return self.planar['wlr'][:, 2]


def _gsim_structs(gsims, imts):
    """ctypes arrays (oqg_gsim[G], coeffs[G][M+1][KMAX]) for the models"""
    G, M = len(gsims), len(imts)
    arr = (_lib.OqgGsim * G)()
    coeffs = numpy.zeros((G, M + 1, _lib.KMAX))
    for g, gsim in enumerate(gsims):
        fam, flags, par = gsim.device_spec()
        arr[g].family, arr[g].flags = fam, flags
        for k in range(4):
            arr[g].par[k] = par[k]
        for m, imt in enumerate(imts):
            coeffs[g, m] = gsim.coeff_row(imt)   # KeyError if undefined
        try:
            coeffs[g, M] = gsim.pga_row()
        except KeyError:
            pass
    return arr, coeffs


CTX_COLUMNS = ('mag', 'rake', 'dip', 'ztor', 'width', 'hypo_depth', 'rrup',
               'rjb', 'rx', 'ry0', 'repi', 'rhypo', 'vs30', 'z1pt0', 'z2pt5',
               'vs30measured')


def _ctx_struct(ctx, device):
    """OqgCtx of device pointers for a recarray-like ctx; returns the
    struct and the tensors to keep alive"""
    names = (ctx.dtype.names if hasattr(ctx, 'dtype') and ctx.dtype.names
             else list(vars(ctx)))
    st, keep = _lib.OqgCtx(), []
    for col in CTX_COLUMNS:
        if col in names:
            a = ctx[col] if hasattr(ctx, 'dtype') else getattr(ctx, col)
            if col == 'vs30measured':
                t = _dev(numpy.asarray(a, bool), device)
            else:
                t = _dev(numpy.asarray(a, float), device)
            keep.append(t)
            setattr(st, col, t.data_ptr())
    return st, keep


def mean_stds_device(gsims, ctx, imts, device=None, handle=None,
                     configure=True):
    """(4, G, M, N) float64 numpy array computed on the GPU"""
    import torch
    h = handle or get_handle(device)
    if configure:
        arr, coeffs = _gsim_structs(gsims, imts)
        h.call('oqg_set_gsims', len(gsims), arr, len(imts),
               coeffs.ctypes.data_as(_lib.c_double_p))
        h.owner = h.cfg = None
    N = len(ctx)
    out = torch.empty((4, len(gsims), len(imts), N), dtype=torch.float64,
                      device='cuda:%d' % h.device)
    if N:
        st, keep = _ctx_struct(ctx, h.device)
        h.call('oqg_mean_stds', ctypes.byref(st), N, _ptr(out))
    return out.cpu().numpy()


def mean_stds_staged_device(gsims, ctx, imts, device=None, handle=None):
    """(2, G, M, N) float64 = mean, sig as evaluated by the staged model
    code of the fused kernel (verification of csrc/oqg_fast.cuh)"""
    import torch
    h = handle or get_handle(device)
    arr, coeffs = _gsim_structs(gsims, imts)
    h.call('oqg_set_gsims', len(gsims), arr, len(imts),
           coeffs.ctypes.data_as(_lib.c_double_p))
    h.owner = h.cfg = None
    N = len(ctx)
    out = torch.empty((2, len(gsims), len(imts), N), dtype=torch.float64,
                      device='cuda:%d' % h.device)
    if N:
        st, keep = _ctx_struct(ctx, h.device)
        h.call('oqg_mean_stds_staged', ctypes.byref(st), N, _ptr(out))
    return out.cpu().numpy()


def _group_by_site(sidxs, N):
    order = numpy.argsort(sidxs, kind='stable').astype(numpy.int64)
    seg = numpy.searchsorted(sidxs[order], numpy.arange(N + 1)).astype(
        numpy.int64)
    return order, seg


def update_device(pmap, poes, rates, probs_occur, sidxs, itime, handle=None,
                  weights=None):
    """MapArray.update_indep (or, with `weights`, update_mutex) on the GPU:
    poes (n, L, G) host or device"""
    import torch
    h = handle or get_handle(_device_index(pmap.device))
    dev = h.device
    n = len(sidxs)
    if n == 0:
        return
    N, L, G = pmap.shape
    # the kernel reads its shapes from the handle configuration
    if getattr(h, 'cfg', None) != (L, G):
        _configure_shapes(h, L, G)
    tpoes = _dev(poes, dev, torch.float32)
    order, seg = _group_by_site(numpy.asarray(sidxs, numpy.int64), N)
    probs = numpy.asarray(probs_occur, float)
    K = probs.shape[1] if probs.ndim == 2 else 0
    t_order, t_seg = _dev(order, dev), _dev(seg, dev)
    t_rate = _dev(numpy.asarray(rates, float), dev)
    t_probs = _dev(probs, dev) if K else None
    on_host = pmap.device is None
    t_map = _dev(pmap.array, dev) if on_host else pmap.array
    if weights is not None:
        t_w = _dev(numpy.asarray(weights, float), dev)
        h.call('oqg_update_mutex', _ptr(t_map), N, _ptr(tpoes), n,
               _ptr(t_order), _ptr(t_seg), _ptr(t_rate), _ptr(t_probs), K,
               float(itime), _ptr(t_w))
    else:
        h.call('oqg_update', _ptr(t_map), N, _ptr(tpoes), n, _ptr(t_order),
               _ptr(t_seg), _ptr(t_rate), _ptr(t_probs), K, float(itime),
               int(pmap.rates))
    if on_host:
        pmap.array[:] = t_map.cpu().numpy()


def _configure_shapes(h, L, G):
    """Minimal configuration for oqg_update when no ContextMaker set the
    handle up: G dummy models, one IMT with L levels"""
    arr = (_lib.OqgGsim * G)()
    for g in range(G):
        arr[g].family = _lib.FAMILIES['BA08']
    coeffs = numpy.zeros((G, 2, _lib.KMAX))
    h.call('oqg_set_gsims', G, arr, 1, coeffs.ctypes.data_as(_lib.c_double_p))
    lev = numpy.zeros(L)
    h.call('oqg_set_levels', 1, L, lev.ctypes.data_as(_lib.c_double_p), 99.)
    h.cfg = (L, G)
    h.owner = None


class ContextMaker(object):
    """
    :param trt: tectonic region type string
    :param gsims: list of GMPE instances (or class names)
    :param oq: dictionary with imtls, truncation_level, investigation_time,
        maximum_distance, minimum_distance (same keys as the reference)
    """
    tom = None
    cluster = 0

    def __init__(self, trt, gsims, oq, device=None, monitor=None):
        param = dict(oq)
        self.trt = trt
        self.init_monitoring(monitor or Monitor())
        self.gsims = [get_gsim(g) for g in gsims]
        for g in self.gsims:
            if not isinstance(g, GMPE):
                raise NotImplementedError('%r has no device implementation'
                                          % g)
        self.gid = numpy.arange(len(self.gsims), dtype=numpy.uint16)
        if 'imtls' not in param:
            raise KeyError('Missing imtls in ContextMaker!')
        for imt in param['imtls']:
            if not isinstance(imt, str):
                raise TypeError('Expected string, got %s' % type(imt))
        self.imtls = {imt: numpy.asarray(imls, float)
                      for imt, imls in param['imtls'].items()}
        self.maximum_distance = _interp(param, 'maximum_distance', trt)
        self.minimum_distance = param.get('minimum_distance', 0)
        self.investigation_time = param.get('investigation_time')
        self.shift_hypo = param.get('shift_hypo', False)
        self.max_sites_disagg = param.get('max_sites_disagg', 10)
        psd = param.get('pointsource_distance') or 0
        if isinstance(psd, dict):   # floatdict of the job.ini, by TRT
            psd = psd.get(trt, psd.get('default', 0))
        self.pointsource_distance = float(psd)
        self.truncation_level = param.get('truncation_level', 99.)
        self.phi_b = float(ndtr(self.truncation_level))
        # horizontal component conversion (set_imts_conv, contexts.py:726-746):
        # `conv` = {gsim index: {imt string: (conv_median, conv_sigma, rstd)}}
        # as the reference builds it from IMC.apply_conversion; the name of a
        # component (`horiz_comp`) alone is an error here: the conversion
        # tables are not part of this package
        self.conv = param.get('conv') or {}
        if param.get('horiz_comp') and not self.conv:
            raise NotImplementedError(
                'horiz_comp needs the conversion coefficients: pass '
                'conv={g: {imt: (conv_median, conv_sigma, rstd)}}')
        self.device = device
        self.cfactor = numpy.zeros(2)
        self.set_imts_conv()
        for req in ('DISTANCES', 'SITES_PARAMETERS', 'RUPTURE_PARAMETERS'):
            reqset = set()
            for gsim in self.gsims:
                reqset.update(getattr(gsim, 'REQUIRES_' + req))
            setattr(self, 'REQUIRES_' + req, reqset)
        # validate models and IMTs now: unsupported -> error at construction
        self._structs = _gsim_structs(self.gsims, self.imts)

    def init_monitoring(self, monitor):
        """Child monitors with the reference's names (contexts.py:681-690);
        `fused_mon` is the one-kernel path (distances + mean_std + poes +
        map), which the reference has no name for.  With a monitor made with
        device_sync=True the durations are device times."""
        self.gmf_mon = monitor('computing mean_std', measuremem=False)
        self.poe_mon = monitor('get_poes', measuremem=False)
        self.ir_mon = monitor('iter_ruptures', measuremem=False)
        self.sec_mon = monitor('building dparam', measuremem=False)
        self.delta_mon = monitor('getting delta_rates', measuremem=False)
        self.fused_mon = monitor('fused distances+mean_std+poes+map',
                                 measuremem=False)
        self.task_no = getattr(monitor, 'task_no', 0)
        self.out_no = getattr(monitor, 'out_no', self.task_no)
        self.cfactor = numpy.zeros(2)

    def set_imts_conv(self):
        self.imts = tuple(from_string(im) for im in self.imtls)
        lens = {len(v) for v in self.imtls.values()}
        if len(lens) != 1:
            raise ValueError('all IMTs must have the same number of levels')
        self.L1 = lens.pop()
        with numpy.errstate(divide='ignore'):
            self.loglevels = {
                imt: (numpy.log(imls) if imt != 'MMI' else imls)
                for imt, imls in self.imtls.items()}

    @property
    def M(self):
        return len(self.imts)

    @property
    def G(self):
        return len(self.gsims)

    @property
    def L(self):
        return self.M * self.L1

    # ------------------------------------------------------------- device
    def handle(self):
        """The configured device handle (created lazily in the worker
        process, like init_monitoring in the reference)"""
        return self.configure(get_handle(_device_index(self.device)))

    def configure(self, h):
        """Load the models, levels and IMT mask of this ContextMaker into the
        oqg handle `h` (a _lib.Handle; any number of handles may exist on a
        device, each with its own stream)"""
        if getattr(h, 'owner', None) is not self:
            arr, coeffs = self._structs
            h.call('oqg_set_gsims', self.G, arr, self.M,
                   coeffs.ctypes.data_as(_lib.c_double_p))
            lev = numpy.ascontiguousarray(
                [self.loglevels[im] for im in self.imtls], float)
            h.call('oqg_set_levels', self.M, self.L1,
                   lev.ctypes.data_as(_lib.c_double_p),
                   float(self.truncation_level))
            # set_poes, contexts.py:1469-1475: an ImtWeight of 0 (set by the
            # engine when parsing the gsim logic tree) zeroes the PoEs of
            # that IMT for that gsim
            mask = numpy.zeros((self.G, self.M), numpy.uint8)
            for g, gsim in enumerate(self.gsims):
                weight = getattr(gsim, 'weight', None)
                dic = getattr(weight, 'dic', None)
                if dic:
                    for m, imt in enumerate(self.imtls):
                        if dic.get(imt) == 0:
                            mask[g, m] = 1
            if mask.any():
                h.call('oqg_set_imt_mask', mask.ctypes.data_as(
                    ctypes.POINTER(ctypes.c_ubyte)))
            if self.conv:
                cv = numpy.zeros((self.G, self.M, 3))
                for g, dic in self.conv.items():
                    for m, imt in enumerate(self.imtls):
                        if imt in dic:
                            cv[g, m] = dic[imt]
                h.call('oqg_set_conversion',
                       cv.ctypes.data_as(_lib.c_double_p))
            h.owner = self
            h.cfg = (self.L, self.G)
        return h

    def __getstate__(self):
        # ctypes arrays do not pickle: rebuild them in the worker
        dic = self.__dict__.copy()
        dic.pop('_structs', None)
        return dic

    def __setstate__(self, dic):
        self.__dict__.update(dic)
        self._structs = _gsim_structs(self.gsims, self.imts)

    # --------------------------------------------------------------- ctxs
    fewsites = False

    def ctx_dtype(self, K=0):
        """The dtype of the context rows (contexts.py:660-678); with few
        sites the closest points clon, clat are added (contexts.py:385-386,
        425-427)"""
        dists = self.REQUIRES_DISTANCES | {'rrup'}
        if self.fewsites:
            dists = dists | {'clon', 'clat'}
        fields = []
        for name in sorted(self.REQUIRES_RUPTURE_PARAMETERS):
            fields.append((name, float))
        for name in sorted(self.REQUIRES_SITES_PARAMETERS):
            fields.append((name, bool if name == 'vs30measured' else float))
        for name in sorted(dists):
            fields.append((name, float))
        fields += [('src_id', numpy.int32), ('rup_id', U32), ('sids', U32),
                   ('occurrence_rate', float), ('probs_occur', float, (K,))]
        return numpy.dtype(fields)

    def new_ctx(self, size, K=0):
        return numpy.zeros(size, self.ctx_dtype(K)).view(numpy.recarray)

    def get_ctxs(self, src, sitecol, src_id=0, step=1):
        """
        Context rows of the ruptures of `src` (a PlanarRuptures or a
        source.PointSource) on the sites, one recarray per unique magnitude
        like genctxs_Pp (contexts.py:443-474): rows with rrup >=
        maximum_distance(mag) are dropped, required distances are clamped by
        minimum_distance.  For a PointSource the ruptures are built on the
        device and, with a pointsource_distance below 1000 km, the near / far
        split of _quintets (contexts.py:321-370) selects the rows.
        """
        import torch
        from .source import PointSource
        from . import fault
        if isinstance(src, fault.SimpleFaultSource):
            src_id = src.id if src_id == 0 else src_id
            src = src.ruptures()
        from .multifault import MultiFaultRuptures
        if isinstance(src, (fault.MeshRuptures, MultiFaultRuptures)):
            # non-planar ruptures: genctxs (contexts.py:959-1029), one
            # context per rupture
            with self.ir_mon:
                return fault.get_ctxs_mesh(self, src, sitecol, src_id, step)
        vis = None
        if isinstance(src, PointSource):
            src, vis = self._planar_of_point_source(src, sitecol, src_id)
        if not isinstance(src, PlanarRuptures):
            raise NotImplementedError(
                'only planar ruptures (PlanarRuptures, PointSource) and mesh '
                'ruptures (fault.MeshRuptures, fault.SimpleFaultSource) have '
                'a device path')
        with self.ir_mon:
            return list(self._gen_ctxs_planar(src, sitecol, src_id, step, vis))

    # distance planes pulled to the host per block of ruptures of get_ctxs
    GET_CTXS_BYTES = 256 << 20

    def _gen_ctxs_planar(self, src, sitecol, src_id, step, vis):
        """Generator behind get_ctxs / get_ctx_iter for planar ruptures: one
        magnitude at a time -- the distances of the ruptures of that
        magnitude only are computed on the device and copied to the host,
        in blocks of at most GET_CTXS_BYTES -- so that the memory does not
        grow with the size of the source (the reference generates per
        magnitude too, genctxs_Pp contexts.py:443-474)"""
        import torch
        h = self.handle()
        U, N = len(src), len(sitecol)
        if U == 0 or N == 0:
            return
        # contexts.py:1064: few sites -> closest points in the rows
        self.fewsites = N <= self.max_sites_disagg
        dists = sorted(self.REQUIRES_DISTANCES | {'rrup'} |
                       ({'clon', 'clat'} if self.fewsites else set()))
        mask = 0
        for d in dists:
            mask |= 1 << _lib.DIST.index(d)
        order = [d for d in _lib.DIST if d in dists]
        dev = h.device
        st, keep = self._sites_struct(sitecol, dev)
        tom = self.tom
        K = 0
        probs = None
        if tom is not None and hasattr(tom, 'get_pmf'):
            probs = tom.get_pmf(src.rate)
            K = probs.shape[1]
        mags = src.mag
        flat = as_flat(src.planar)
        per = max(1, self.GET_CTXS_BYTES // (len(order) * N * 8))
        for mag in numpy.unique(mags)[::step]:
            allidx = numpy.flatnonzero(mags == mag)
            parts = []
            for b0 in range(0, len(allidx), per):
                idx = allidx[b0:b0 + per]
                t_pla = _dev(flat[idx], dev)
                out = torch.empty((len(order), len(idx), N),
                                  dtype=torch.float64, device='cuda:%d' % dev)
                h.call('oqg_distances', _ptr(t_pla), len(idx),
                       ctypes.byref(st), N, mask, _ptr(out))
                dd = dict(zip(order, out.cpu().numpy()))
                ctx = self.new_ctx(len(idx) * N, K).reshape(len(idx), N)
                for name in dists:
                    arr = dd[name]
                    if self.minimum_distance and \
                            name in self.REQUIRES_DISTANCES:
                        arr = numpy.where(arr < self.minimum_distance,
                                          self.minimum_distance, arr)
                    ctx[name] = arr
                pla = src.planar[idx]
                par = dict(mag=mags[idx], width=pla['wlr'][:, 0],
                           dip=pla['sdr'][:, 1], rake=pla['sdr'][:, 2],
                           strike=pla['sdr'][:, 0],
                           ztor=pla['corners'][:, 2, 0],
                           hypo_depth=pla['hypo'][:, 2])
                for name in self.REQUIRES_RUPTURE_PARAMETERS:
                    ctx[name] = par[name][:, None]
                for name in self.REQUIRES_SITES_PARAMETERS:
                    ctx[name] = sitecol.array[name][None, :]
                ctx['occurrence_rate'] = pla['wlr'][:, 2][:, None]
                ctx['rup_id'] = idx[:, None]
                ctx['src_id'] = src_id
                ctx['sids'] = sitecol.sids[None, :]
                if K:
                    ctx['probs_occur'] = probs[idx][:, None, :]
                if vis is not None and (vis[idx, 4] != 0).any():
                    # near sites see all the planes and depths (mode 1), far
                    # sites the average rupture (mode 2), contexts.py:347-370
                    aux = vis[idx]
                    d2 = ((sitecol.xyz[None, :, :] - aux[:, None, :3]) ** 2
                          ).sum(axis=2)
                    close = d2 <= (aux[:, 3] ** 2)[:, None]
                    mode = aux[:, 4][:, None]
                    keepm = numpy.where(mode == 1, close, numpy.where(
                        mode == 2, ~close, mode == 0))
                    ctx = ctx[keepm]
                else:
                    ctx = ctx.flatten()
                ctx = ctx[ctx['rrup'] < self.maximum_distance(mag)]
                if len(ctx):
                    parts.append(ctx)
            if parts:
                yield (parts[0] if len(parts) == 1 else
                       numpy.concatenate(parts)).view(numpy.recarray)
        del keep

    def get_ctx_iter(self, src, sitecol, src_id=0, step=1):
        """The name BASELINE.json's north_star uses for get_ctxs: the same
        contexts, generated lazily (one magnitude of a planar source in
        memory at a time)"""
        from .source import PointSource
        from . import fault
        if isinstance(src, (fault.SimpleFaultSource, fault.MeshRuptures)):
            return iter(self.get_ctxs(src, sitecol, src_id, step))
        vis = None
        if isinstance(src, PointSource):
            src, vis = self._planar_of_point_source(src, sitecol, src_id)
        if not isinstance(src, PlanarRuptures):
            raise NotImplementedError(type(src).__name__)
        return self._gen_ctxs_planar(src, sitecol, src_id, step, vis)

    def _planar_of_point_source(self, src, sitecol, src_id=0):
        """PlanarRuptures of a PointSource built on the device and the
        per-rupture visibility (U, 5) of the pointsource_distance
        approximation (None when it does not apply)"""
        from .source import PointSourceBatch, point_ruptures_device
        batch = PointSourceBatch([src], self.pointsource_distance)
        h = self.handle()
        mag_dist = self.magdist(batch.mag)
        mode = None
        if batch.pointsource_distance > 0 and \
                len(sitecol) <= self.max_sites_disagg:
            mode = batch.fewsites_mode(mag_dist, sitecol)
        t_pla, t_mag, _md, t_aux = point_ruptures_device(
            batch, h, mag_dist, self.shift_hypo, mode, with_aux=True)
        pla = t_pla.cpu().numpy().reshape(-1, 3, 14).view(
            planar_array_dt).reshape(-1, 3)
        vis = t_aux.cpu().numpy() if batch.pointsource_distance > 0 else None
        return PlanarRuptures(pla, t_mag.cpu().numpy(), src_id), vis

    def _sites_struct(self, sitecol, dev):
        st, keep = _lib.OqgSites(), []
        arr = sitecol.array
        for name in ('lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5',
                     'vs30measured'):
            if name in arr.dtype.names:
                a = arr[name]
                t = _dev(numpy.asarray(a, bool if name == 'vs30measured'
                                       else float), dev)
                keep.append(t)
                setattr(st, name, t.data_ptr())
        return st, keep

    def recarray(self, ctxs):
        """Concatenate context arrays.  An empty list (no rupture within
        the integration distance) gives an EMPTY context array with the
        fields of this ContextMaker, not numpy's bare ValueError: callers
        test ``len(ctx)`` (the reference asserts a non-empty list,
        contexts.py:798-803, and its callers filter the sources first)"""
        if len(ctxs) == 0:
            return self.new_ctx(0)
        if len(ctxs) == 1:
            return ctxs[0]
        return numpy.concatenate(ctxs).view(numpy.recarray)

    # --------------------------------------------------------- mean, stds
    def get_mean_stds(self, ctxs, split_by_mag=True):
        """
        :returns: (4, G, M, N) float64: mean, sig, tau, phi
        (contexts.py:1250-1279), in the order of the rows.  `split_by_mag`
        only makes the reference evaluate runs of equal magnitude one at a
        time (split_array on ORDERED indices, baselib/performance.py:579-592):
        it does not reorder anything, and has no effect here.
        """
        ctx = self.recarray(list(ctxs))
        h = self.handle()
        with self.gmf_mon:
            return mean_stds_device(self.gsims, ctx, self.imts, handle=h,
                                    configure=False)

    # --------------------------------------------------------------- poes
    def _device_poes(self, h, ctx):
        """device tensors (poes (n, L, G) f32, mean_stds (4, G, M, n))"""
        import torch
        n = len(ctx)
        dev = 'cuda:%d' % h.device
        ms = torch.empty((4, self.G, self.M, n), dtype=torch.float64,
                         device=dev)
        with self.gmf_mon:     # contexts.py:1161
            st, keep = _ctx_struct(ctx, h.device)
            h.call('oqg_mean_stds', ctypes.byref(st), n, _ptr(ms))
        with self.poe_mon:     # contexts.py:1171
            poes = torch.empty((n, self.L, self.G), dtype=torch.float32,
                               device=dev)
            h.call('oqg_poes', _ptr(ms), n, _ptr(poes))
        return poes, ms

    # rows per block of gen_poes: the (n, L, G) PoEs of a block are at most
    # this many bytes as float32 on the device (and twice that as the float64
    # array the caller gets)
    GEN_POES_BYTES = 32 << 20

    def gen_poes(self, ctx):
        """
        :yields: poes (n, L, G) float64, mea, sig, tau (G, M, n) and the
            context rows: one or more blocks per unique magnitude (rounded to
            3 digits, in increasing order).  Like the reference, which cuts
            every magnitude into slabs so that `poes` stays small
            (contexts.py:1155-1203, split_in_slices), a block never holds
            more than GEN_POES_BYTES of PoEs, whatever the size of `ctx`:
            mean / stds and PoEs are computed block by block on the device.
        """
        h = self.handle()
        ctx.mag = numpy.round(ctx.mag, 3)
        order = numpy.argsort(ctx.mag, kind='stable')
        sctx = ctx[order]
        if len(sctx) == 0:
            return
        mags, starts = numpy.unique(sctx.mag, return_index=True)
        stops = list(starts[1:]) + [len(sctx)]
        self.cfactor += [len(sctx), len(mags)]
        rows = max(1, self.GEN_POES_BYTES // (self.L * self.G * 4))
        for a, b in zip(starts, stops):
            for c in range(a, b, rows):
                d = min(b, c + rows)
                blk = sctx[c:d]
                poes, ms = self._device_poes(h, blk)
                ms = ms.cpu().numpy()
                yield (poes.cpu().numpy().astype(numpy.float64), ms[0], ms[1],
                       ms[2], blk)

    # ---------------------------------------------------------------- map
    def update(self, pmap, ctx, rup_mutex=None):
        """
        Update the map in place (contexts.py:1232-1247): mean/stds, PoEs and
        the fold into the map all stay on the device.
        """
        if len(ctx) == 0:
            return
        h = self.handle()
        itime = self.tom.time_span if self.tom is not None else (
            self.investigation_time)
        poes, _ms = self._device_poes(h, ctx)
        probs_occur = ctx.probs_occur
        if rup_mutex:      # contexts.py:1241-1242 -> MapArray.update_mutex
            weights = numpy.array([
                rup_mutex[src_id, rup_id]
                for src_id, rup_id in zip(ctx.src_id, ctx.rup_id)])
            update_device(pmap, poes, ctx.occurrence_rate, probs_occur,
                          pmap.sidx[ctx.sids], itime, handle=h,
                          weights=weights)
            self.cfactor += [len(ctx),
                             len(numpy.unique(numpy.round(ctx.mag, 3)))]
            return
        if self.cluster:   # contexts.py:1243-1246: arr[sidx] *= 1 - poe
            if pmap.rates:
                raise ValueError('a cluster needs a map of probabilities')
            probs_occur, itime = numpy.ones((len(ctx), 1)), 0.
        update_device(pmap, poes, ctx.occurrence_rate, probs_occur,
                      pmap.sidx[ctx.sids], itime, handle=h)
        self.cfactor += [len(ctx), len(numpy.unique(numpy.round(ctx.mag, 3)))]

    def get_pmap(self, ctx, tom=None, rup_mutex={}):
        """PoE map (contexts.py:1205-1218); with `rup_mutex` (a dictionary
        (src_id, rup_id) -> weight) the ruptures are mutually exclusive"""
        rup_indep = not rup_mutex
        self.tom = tom or self.tom or PoissonTOM(self.investigation_time)
        sids = numpy.unique(ctx.sids)
        pmap = MapArray(sids, self.L, self.G).fill(rup_indep)
        self.update(pmap, ctx, rup_mutex)
        return ~pmap if rup_indep else pmap

    # ------------------------------------------- notebook-level operators
    def from_srcs(self, srcs, sitecol):
        """
        contexts.py:866-881: the context rows of a list of sources on the
        sites within the integration distance, one array; a source without
        an `id` gets its position in the list as src_id
        """
        ctxs = []
        for i, src in enumerate(srcs):
            sid = getattr(src, 'id', -1)
            if sid == -1:
                sid = getattr(src, 'src_id', i) \
                    if isinstance(src, PlanarRuptures) else i
            ctxs.extend(self.get_ctxs(src, sitecol, src_id=sid))
        return self.recarray(ctxs)

    def get_poes(self, srcs, sitecol, tom=None, rup_mutex={}):
        """
        contexts.py:1145-1153
        :returns: an array of PoEs of shape (N, L, G), N the sites reached
        """
        ctx = self.from_srcs(srcs, sitecol)
        return self.get_pmap(ctx, tom, rup_mutex).array

    def get_rmap(self, ctx):
        """
        contexts.py:1220-1230
        :returns: a MapArray of rates of shape (N, L, G)
        """
        rmap = (~self.get_pmap(ctx)).to_rates()
        rmap.gid = self.gid
        rmap.wei = getattr(self, 'wei', None)
        return rmap

    def max_intensity(self, sitecol1, mags, dists):
        """
        contexts.py:1101-1130: the largest median intensity over the models
        and IMTs for every (magnitude, distance), all the required distances
        set to the distance, the other rupture parameters to zero and a
        width of 10 m; one batched evaluation instead of a loop of contexts

        :param sitecol1: a SiteCollection with a single site
        :returns: an array of GMVs of shape (#mags, #dists)
        """
        assert len(sitecol1) == 1, sitecol1
        nmags, ndists = len(mags), len(dists)
        ctx = self.new_ctx(nmags * ndists)
        for dst in self.REQUIRES_DISTANCES | {'rrup'}:
            ctx[dst] = numpy.tile(numpy.asarray(dists, float), nmags)
        for par in self.REQUIRES_SITES_PARAMETERS:
            ctx[par] = sitecol1.array[par][0]
        ctx['sids'] = sitecol1.sids[0]
        ctx['mag'] = numpy.repeat(numpy.asarray(mags, float), ndists)
        if 'width' in ctx.dtype.names:
            ctx['width'] = .01
        mean = self.get_mean_stds([ctx])[0]          # (G, M, n)
        return numpy.exp(mean.max(axis=(0, 1))).reshape(nmags, ndists)

    def get_occ_rates(self, ctxt):
        """
        contexts.py:1132-1142: the occurrence rates, from probs_occur[0] for
        non-parametric ruptures
        """
        if numpy.isfinite(ctxt[0].occurrence_rate):
            return ctxt.occurrence_rate
        probs = numpy.array([rec.probs_occur[0] for rec in ctxt])
        return -numpy.log(probs) / self.investigation_time

    def get_4MN(self, ctxs, gsim):
        """
        contexts.py:1281-1303 (called by the GmfComputer): mean, sig, tau, phi
        of one model of this ContextMaker, shape (4, M, N)
        """
        for g, gs in enumerate(self.gsims):
            if gs is gsim or gs == gsim:
                break
        else:
            raise ValueError('%s is not a model of this ContextMaker' % gsim)
        out = self.get_mean_stds(ctxs, split_by_mag=False)[:, g]
        if self.truncation_level not in (0, 1E-9, 99.) and \
                (out[1] == 0.).any():
            raise ValueError('Total StdDev is zero for %s' % gsim)
        return out

    # -------------------------------------------------------------- fused
    def magdist(self, mags):
        """maximum distance for each magnitude"""
        return numpy.array([float(self.maximum_distance(m)) for m in mags])

    def hazard_planar(self, pmap, src, sitecol, count=False, aux=None):
        """
        Fold all the ruptures of `src` (PlanarRuptures) on the sites into
        `pmap` in one fused pass: the equivalent of
        ``for ctx in get_ctxs(src, sitecol): update(pmap, ctx)`` without
        materialising the context rows.  Returns the number of rupture-site
        pairs kept by the distance filter if `count`.  `aux` (U, 5) float64,
        optional: the per-rupture records [x, y, z of the source location in
        km, psdist, mode] of the pointsource_distance approximation
        (oqg_hazard_planar_aux; mode 0 all sites, 1 the sites within psdist
        of the location, 2 the others).
        """
        import torch
        h = self.handle()
        dev = h.device
        tom = self.tom or PoissonTOM(self.investigation_time)
        self.tom = tom
        U, N = len(src), len(sitecol)
        if U == 0 or N == 0:
            return 0
        K, t_probs = 0, None
        if self.cluster:
            if pmap.rates:
                raise ValueError('a cluster needs a map of probabilities')
            t_probs, K = self._fated(U, dev)
        elif hasattr(tom, 'get_pmf'):
            probs = numpy.ascontiguousarray(tom.get_pmf(src.rate))
            K = probs.shape[1]
            t_probs = _dev(probs, dev)
        t_pla = _dev(as_flat(src.planar), dev)
        t_mag = _dev(src.mag, dev)
        t_md = _dev(self.magdist(src.mag), dev)
        t_aux = None if aux is None else _dev(
            numpy.ascontiguousarray(aux, float), dev)
        return self._hazard_planar_dev(pmap, t_pla, t_mag, t_md, t_probs, K,
                                       U, sitecol, count, t_aux)

    def hazard_mesh(self, pmap, src, sitecol, count=False):
        """
        Fold the ruptures of a fault source (fault.MeshRuptures or
        fault.SimpleFaultSource) into `pmap`: distances of the rupture meshes
        on the device, then the fused GMPE -> PoE -> map pass
        (oqg_hazard_dist); the equivalent of genctxs + update of the
        reference (contexts.py:959-1029, 1232-1247)
        """
        from . import fault
        if isinstance(src, fault.SimpleFaultSource):
            src = src.ruptures()
        return fault.hazard_mesh(self, pmap, src, sitecol, count)

    def _fated(self, U, dev):
        # cluster groups (contexts.py:1243-1246): every rupture happens once
        # when the cluster does, pne = prod (1 - poe); the kernels take that
        # as a one-element probs_occur with investigation time 0
        import torch
        return torch.ones((U, 1), dtype=torch.float64,
                          device='cuda:%d' % dev), 1

    def _hazard_planar_dev(self, pmap, t_pla, t_mag, t_md, t_probs, K, U,
                           sitecol, count, t_aux=None):
        # ruptures already on the device -> oqg_hazard_planar_aux
        import torch
        h = self.handle()
        dev = h.device
        N = len(sitecol)
        st, keep = self._sites_struct(sitecol, dev)
        t_sidx = _dev(pmap.sidx[sitecol.sids].astype(numpy.uint32), dev)
        on_host = pmap.device is None
        t_map = _dev(pmap.array, dev) if on_host else pmap.array
        t_cnt = torch.zeros(2, dtype=torch.int64, device='cuda:%d' % dev) \
            if count else None
        with self.fused_mon:
            h.call('oqg_hazard_planar_aux', _ptr(t_pla), _ptr(t_mag),
                   _ptr(t_md), _ptr(t_aux), _ptr(t_probs), K, U,
                   ctypes.byref(st), _ptr(t_sidx), N,
                   float(self.minimum_distance),
                   0. if self.cluster else float(self.tom.time_span),
                   int(pmap.rates), _ptr(t_map), _ptr(t_cnt))
        if on_host:
            pmap.array[:] = t_map.cpu().numpy()
        if count:
            return int(t_cnt[0].item())

    def hazard_point_sources(self, pmap, sources, sitecol, count=False):
        """
        Fold all the ruptures of a list of point sources (source.PointSource,
        or a source.PointSourceBatch) into `pmap`: the equivalent of
        ``for src in sources: for ctx in get_ctxs(src, sitecol):
        update(pmap, ctx)`` (genctxs_Pp, contexts.py:443-474) with the
        ruptures built on the device (oqg_point_ruptures) and no context
        rows materialised.  With a `pointsource_distance` below 1000 km in
        the parameters (the engine default is 100) the approximation of
        _quintets (contexts.py:321-370) applies: sites farther than
        psdist(mag) from a source with several nodal planes / depths see one
        average rupture per magnitude; with at most `max_sites_disagg`
        sites the choice is all-or-nothing per magnitude.
        """
        from .source import PointSourceBatch, point_ruptures_device
        # the reference's default (1000 km) keeps every site "close", i.e.
        # no approximation; a smaller value replaces, for far sites, the
        # nodal planes / depths by one average rupture (_quintets)
        batch = sources if isinstance(sources, PointSourceBatch) \
            else PointSourceBatch(sources, self.pointsource_distance)
        h = self.handle()
        tom = self.tom or PoissonTOM(self.investigation_time)
        self.tom = tom
        if batch.U == 0 or len(sitecol) == 0:
            return 0
        K, t_probs = 0, None
        if self.cluster:
            if pmap.rates:
                raise ValueError('a cluster needs a map of probabilities')
            t_probs, K = self._fated(batch.U, h.device)
        elif hasattr(tom, 'get_pmf'):
            _mag, rate = batch.expand()
            probs = numpy.ascontiguousarray(tom.get_pmf(rate))
            K = probs.shape[1]
            t_probs = _dev(probs, h.device)
        mag_dist = self.magdist(batch.mag)
        mode = None
        if batch.pointsource_distance > 0 and \
                len(sitecol) <= self.max_sites_disagg:
            # few sites: all-or-nothing per magnitude, contexts.py:358-362
            mode = batch.fewsites_mode(mag_dist, sitecol)
        t_pla, t_mag, t_md, t_aux = point_ruptures_device(
            batch, h, mag_dist, self.shift_hypo, mode, with_aux=True)
        return self._hazard_planar_dev(pmap, t_pla, t_mag, t_md, t_probs, K,
                                       batch.U, sitecol, count, t_aux)

    def get_pmap_planar(self, src, sitecol):
        """PoE map of planar ruptures on the sites through the fused path"""
        pmap = MapArray(sitecol.sids, self.L, self.G).fill(1)
        self.hazard_planar(pmap, src, sitecol)
        return ~pmap


def get_mean_stds(gsim, ctx, imts, **kw):
    """
    Module-level helper of the reference (contexts.py:1723): (4, M, N) array
    for a single model
    """
    imtls = {from_string(i).string: [0] for i in imts}
    cm = ContextMaker('*', [gsim], dict(imtls=imtls, **kw))
    return cm.get_mean_stds([ctx], split_by_mag=False)[:, 0]
