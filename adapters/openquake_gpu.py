"""Reference-side binding of liboqg: the module a maintainer of gem/oq-engine
adds (e.g. as ``openquake/hazardlib/gpu.py``) to run the classical-PSHA hot
path of a stock engine on a B200.

It imports **openquake** and feeds the C ABI (include/oqg.h) from the
reference's OWN objects:

* the models come from the reference's GSIM instances: family, flags and
  scalar parameters from their attributes, coefficient rows from their
  ``COEFFS`` tables evaluated with the reference's ``CoeffsTable.__getitem__``
  (log-period interpolation and KeyError included), column order from
  ``oqg_family_fields`` (:func:`pack_gsims`);
* the sites from ``SiteCollection.array`` (lon, lat, depth, vs30,
  vs30measured, z1pt0, z2pt5 columns, :func:`site_columns`);
* the ruptures of point-like sources from ``PointSource.get_planar()``: its
  ``planar_array_dt`` records upload unchanged (:func:`planar_of_source`);
* ``GpuContextMaker(ContextMaker)`` overrides ``get_mean_stds``, ``gen_poes``
  and ``update`` (``get_pmap`` / ``get_rmap`` / ``get_poes`` of the reference
  then run on the device through them) and adds the fused pass
  ``update_from_sources``; ``GpuRmapMaker(RmapMaker)`` overrides
  ``_make_src_indep`` with it.

Unsupported models or model options raise ``NotImplementedError`` when the
GpuContextMaker is built (there is no silent CPU fallback inside the GPU
classes; a caller that wants the stock path keeps the stock ContextMaker).

The device side is reached through ``oq_engine_b200`` (ctypes binding
``_lib`` and the device plumbing of its ContextMaker); nothing here
duplicates arithmetic.  Followed reference interfaces: contexts.py:542-735
(ContextMaker.__init__, _init1/_init2, set_imts_conv), :681-690
(init_monitoring), :1155-1303 (gen_poes, update, get_mean_stds, get_4MN),
:1548-1579 (RmapMaker._make_src_indep), gsim/base.py:389-401 (compute),
gsim/coeffs_table.py:231-286, source/point.py:258-289 (get_planar),
site.py:249-324.
"""
import copy
import numpy

from openquake.baselib.performance import Monitor
from openquake.hazardlib import imt as imt_module
from openquake.hazardlib.contexts import ContextMaker, RmapMaker
from openquake.hazardlib.map_array import MapArray

import oq_engine_b200 as _gpu
from oq_engine_b200 import _lib
from oq_engine_b200.gsim import GMPE as _DeviceGMPE

KMAX = _lib.KMAX


# --------------------------------------------------------------- the models
def _no(gsim, **flags):
    """Options without a device implementation are a hard error"""
    bad = sorted(k for k, v in flags.items() if v)
    if bad:
        raise NotImplementedError(
            '%s with %s has no device implementation'
            % (gsim.__class__.__name__, ', '.join(bad)))


def _spec_bssa14(g):
    _no(g, usgs_basin_scaling=g.usgs_basin_scaling,
        cybershake_basin_adj=g.cybershake_basin_adj)
    try:
        region = {'nobasin': 0, 'CAL': 1, 'JPN': 2}[g.region]
    except KeyError:
        raise NotImplementedError(
            'BooreEtAl2014(region=%r) has no device implementation'
            % g.region) from None
    return 'BSSA14', region | (int(bool(g.sof)) << 4), [
        g.sigma_mu_epsilon, 0., 0., 0.]


def _spec_cy14(g):
    c = g.conf
    _no(g, add_delta_c1=c.get('add_delta_c1'),
        stress_par_host=c.get('stress_par_host'),
        stress_par_target=c.get('stress_par_target'),
        source_function_table=c.get('source_function_table') is not None
        or getattr(g, 'source_function_tab', None),
        delta_gamma_tab=getattr(g, 'delta_gamma_tab', None),
        usgs_basin_scaling=getattr(g, 'usgs_basin_scaling', False),
        cybershake_basin_adj=getattr(g, 'cybershake_basin_adj', False))
    try:
        region = {'CAL': 0, 'JPN': 1, 'ITA': 2, 'WEN': 3}[g.region]
    except KeyError:
        raise NotImplementedError(
            'ChiouYoungs2014(region=%r) has no device implementation'
            % g.region) from None
    flags = region | (int(bool(c['use_hw'])) << 4) | (
        int(bool(c['peer'])) << 5)
    return 'CY14', flags, [g.sigma_mu_epsilon, c['alpha_nm'], 0., 0.]


def _spec_ask14(g):
    _no(g, usgs_basin_scaling=g.usgs_basin_scaling,
        cybershake_basin_adj=g.cybershake_basin_adj)
    region = {None: 0, 'CAL': 0, 'JPN': 1, 'TWN': 2, 'CHN': 3}[g.region]
    return 'ASK14', region, [g.sigma_mu_epsilon, 0., 0., 0.]


def _spec_cb14(g):
    _no(g, estimate_ztor=g.estimate_ztor, estimate_width=g.estimate_width,
        estimate_hypo_depth=g.estimate_hypo_depth,
        coeffs=bool(g.kwargs.get('coeffs')),
        usgs_basin_scaling=getattr(g, 'usgs_basin_scaling', False),
        cybershake_basin_adj=getattr(g, 'cybershake_basin_adj', False))
    return 'CB14', int(bool(g.SJ)), [g.sigma_mu_epsilon, 0., 0., 0.]


def _spec_ba08(g):
    return 'BA08', 0, [0., 0., 0., 0.]


def _spec_ab06(g):
    try:
        mag_eq = {'NA': 0, 'Mblg87': 1, 'Mblg96': 2}[g.mag_eq]
    except KeyError:
        raise NotImplementedError(
            'AtkinsonBoore2006(mag_eq=%r) has no device implementation'
            % g.mag_eq) from None
    return 'AB06', mag_eq, [float(g.scale_fac), 0., 0., 0.]


def _spec_akkar(kind):
    def spec(g):
        # the reference stores log(adjustment_factor) (akkar_2014.py:175)
        return 'AKKAR14', kind, [float(g.adjustment_factor), 0., 0., 0.]
    return spec


def _spec_plain(family, flags=0):
    def spec(g):
        return family, flags, [0., 0., 0., 0.]
    return spec


# class name of the reference -> (spec function, {prefix: table attribute})
SUPPORTED = {
    'BooreEtAl2014': (_spec_bssa14, {'': 'COEFFS'}),
    'ChiouYoungs2014': (_spec_cy14, {'': 'COEFFS'}),
    'ChiouYoungs2014PEER': (_spec_cy14, {'': 'COEFFS'}),
    'AbrahamsonEtAl2014': (_spec_ask14, {'': 'COEFFS'}),
    'CampbellBozorgnia2014': (_spec_cb14, {'': 'COEFFS'}),
    'BooreAtkinson2008': (_spec_ba08, {
        '': 'COEFFS', 'so_': 'COEFFS_SOIL_RESPONSE'}),
    'AtkinsonBoore2006': (_spec_ab06, {
        'bc_': 'COEFFS_BC', 'hr_': 'COEFFS_HARD_ROCK',
        'st_': 'COEFFS_STRESS', 'so_': 'COEFFS_SOIL_RESPONSE'}),
    'AkkarEtAlRjb2014': (_spec_akkar(0), {'': 'COEFFS'}),
    'AkkarEtAlRepi2014': (_spec_akkar(1), {'': 'COEFFS'}),
    'AkkarEtAlRhyp2014': (_spec_akkar(2), {'': 'COEFFS'}),
    'ChiouYoungs2008': (_spec_plain('CY08'), {'': 'COEFFS'}),
    'ToroEtAl2002': (_spec_plain('TORO02'), {'': 'COEFFS'}),
    'Campbell2003': (_spec_plain('C03'), {'': 'COEFFS'}),
    'ToroEtAl2002SHARE': (_spec_plain('TORO02', 1), {'': 'COEFFS'}),
    'Campbell2003SHARE': (_spec_plain('C03', 1), {'': 'COEFFS'}),
    'SadighEtAl1997': (_spec_plain('SADIGH97'), {
        'lo_': 'COEFFS_ROCK_LOWMAG', 'hi_': 'COEFFS_ROCK_HIMAG',
        'sd_': 'COEFFS_ROCK_STDDERR', 'so_': 'COEFFS_SOIL'}),
}


def device_spec(gsim):
    """(family name, flags, [par0..par3]) of a reference GSIM instance;
    NotImplementedError if the class or its options have no device
    implementation"""
    name = gsim.__class__.__name__
    if name not in SUPPORTED:
        raise NotImplementedError(
            'GMPE %s has no device implementation (supported: %s)'
            % (name, ', '.join(sorted(SUPPORTED))))
    return SUPPORTED[name][0](gsim)


def _table(gsim, attr):
    if attr == 'COEFFS_SOIL_RESPONSE' and not hasattr(gsim, attr):
        # BooreAtkinson2008 takes the soil response of AtkinsonBoore2006
        from openquake.hazardlib.gsim.atkinson_boore_2006 import (
            AtkinsonBoore2006)
        return AtkinsonBoore2006.COEFFS_SOIL_RESPONSE
    return getattr(gsim, attr)


def _share_fields(gsim, imt):
    """The SHARE adjustment of ToroEtAl2002SHARE / Campbell2003SHARE
    (toro_2002.py:134-142, campbell_2003.py:236-245): the faulting style term
    of the three rake classes from the class's own COEFFS_FS_ROCK / CONSTS_FS,
    zeros for the base classes"""
    family = SUPPORTED[gsim.__class__.__name__][0](gsim)[0]
    names = (('adj_ss', 'adj_rv', 'adj_nm') if family == 'TORO02'
             else ('fs_ss', 'fs_rv', 'fs_nm', 'afrock'))
    if not getattr(gsim, 'CONSTS_FS', None):
        return dict.fromkeys(names, 0.)
    from openquake.hazardlib.gsim.campbell_2003 import (
        _compute_faulting_style_term)
    C, K = gsim.COEFFS_FS_ROCK[imt], gsim.CONSTS_FS
    # one rake per class: strike slip / other, reverse, normal
    terms = _compute_faulting_style_term(
        C['Frss'], K['pR'], K['Fnss'], K['pN'], numpy.array([0., 90., -90.]))
    if family == 'TORO02':
        return dict(zip(names, terms * C['AFrock']))
    return dict(zip(names, list(terms) + [C['AFrock']]))


def _extra_fields(family, imt, gsim=None):
    """The non-tabulated fields of a device coefficient row"""
    t = imt.period
    out = {'period': t}
    if family in ('TORO02', 'C03') and gsim is not None:
        out.update(_share_fields(gsim, imt))
    if family == 'ASK14':
        # _get_v1, abrahamson_2014.py:345-365
        if imt.string.startswith('SA') and t > 0.5:
            out['v1'] = 800. if t >= 3.0 else float(
                numpy.exp(-0.35 * numpy.log(t / 0.5) + numpy.log(1500.)))
        else:
            out['v1'] = 1500.
    elif family == 'BA08':
        out['is_pga'] = float(imt.string in ('PGA', 'SA(0.0)', 'SA(0)'))
    elif family == 'AB06':
        out['is_pgv'] = float(imt.string == 'PGV')
    return out


def coeff_row(gsim, imt):
    """The KMAX doubles of the device coefficient row of a reference GSIM
    for an IMT, read from the reference's own CoeffsTable objects; KeyError
    if the model is not defined for the IMT (as in the reference)"""
    if isinstance(imt, str):
        imt = imt_module.from_string(imt)
    family, _flags, _par = device_spec(gsim)
    if family == 'CB14' and imt.string in ('CAV', 'IA'):
        raise KeyError('%s has no device implementation' % imt.string)
    tables = SUPPORTED[gsim.__class__.__name__][1]
    extra = _extra_fields(family, imt, gsim)
    recs = {}
    out = []
    for name in _lib.family_fields(family):
        if name in extra:
            out.append(float(extra[name]))
            continue
        prefix, _, col = name.rpartition('_')
        if prefix + '_' in tables:
            key = prefix + '_'
        else:
            key, col = '', name
        if key not in recs:
            recs[key] = _table(gsim, tables[key])[imt]   # KeyError if undefined
        out.append(float(recs[key][col]))
    assert len(out) <= KMAX
    return out + [0.] * (KMAX - len(out))


def pack_gsims(gsims, imts):
    """
    The arguments of ``oqg_set_gsims`` from reference objects:
    ctypes array oqg_gsim[G] and coefficient rows float64 (G, M + 1, KMAX)
    (row M of a model is its PGA row, zeros if it has none)
    """
    G, M = len(gsims), len(imts)
    arr = (_lib.OqgGsim * G)()
    rows = numpy.zeros((G, M + 1, KMAX))
    for g, gsim in enumerate(gsims):
        family, flags, par = device_spec(gsim)
        arr[g].family, arr[g].flags = _lib.FAMILIES[family], flags
        for k in range(4):
            arr[g].par[k] = par[k]
        for m, imt in enumerate(imts):
            rows[g, m] = coeff_row(gsim, imt)
        try:
            rows[g, M] = coeff_row(gsim, 'PGA')
        except KeyError:
            pass
    return arr, rows


class PackedGsim(_DeviceGMPE, register=False):
    """A model described by what :func:`pack_gsims` made of a reference GSIM
    instance: the device plumbing of oq_engine_b200 takes it like one of its
    own GMPE objects"""
    def __init__(self, ref_gsim):
        self.ref = ref_gsim
        self.kwargs = dict(getattr(ref_gsim, 'kwargs', {}))
        self.family, self._flags, self._par = device_spec(ref_gsim)
        for req in ('DISTANCES', 'RUPTURE_PARAMETERS', 'SITES_PARAMETERS'):
            setattr(self, 'REQUIRES_' + req,
                    frozenset(getattr(ref_gsim, 'REQUIRES_' + req)))
        if hasattr(ref_gsim, 'weight'):     # ImtWeight of the logic tree
            self.weight = ref_gsim.weight

    def __repr__(self):
        return 'PackedGsim(%r)' % self.ref

    def device_spec(self):
        return _lib.FAMILIES[self.family], self._flags, list(self._par)

    def coeff_row(self, imt):
        return coeff_row(self.ref, imt if isinstance(imt, str) else imt.string)

    def pga_row(self):
        return coeff_row(self.ref, 'PGA')


# ---------------------------------------------------------------- the sites
SITE_COLUMNS = ('lon', 'lat', 'depth', 'vs30', 'vs30measured', 'z1pt0',
                'z2pt5')


def site_columns(sitecol):
    """
    The columns of the reference's ``SiteCollection.array`` that
    ``oqg_sites`` takes, as contiguous host arrays (float64; vs30measured
    uint8); a missing basin column is the reference's -999 sentinel
    """
    arr = sitecol.array
    out = {}
    for name in SITE_COLUMNS:
        if name in arr.dtype.names:
            col = arr[name]
            out[name] = numpy.ascontiguousarray(
                col, numpy.uint8 if name == 'vs30measured' else float)
    return out


def _device_sitecol(sitecol):
    cols = site_columns(sitecol)
    sc = _gpu.SiteCollection.from_points(
        cols['lon'], cols['lat'], cols.get('depth'), {
            k: (v.astype(bool) if k == 'vs30measured' else v)
            for k, v in cols.items() if k not in ('lon', 'lat', 'depth')})
    sc.array['sids'] = sitecol.sids
    return sc


# ------------------------------------------------------------- the ruptures
def _records(arrays, rate):
    """planar records (U, 3) of the arrays of one magnitude with the
    occurrence rate in wlr[:, 2] (_get_ctx_planar, contexts.py:432: rate of
    the magnitude times the weight of the plane; several arrays only with
    ps_grid_spacing, averaged like genctxs_Pp, contexts.py:460-463)"""
    pla = numpy.concatenate(arrays).copy().view(numpy.recarray)
    pla.wlr[:, 2] *= rate / len(arrays)
    return pla


def planar_of_source(src, shift_hypo=False, cmaker=None, sitecol=None):
    """
    The ruptures of a point-like reference source from its own
    ``get_planar`` (source/point.py:258-289), ordered by magnitude, nodal
    plane, depth as genctxs_Pp visits them: returns (planar records (U, 3) of
    planar_array_dt with the occurrence rate in wlr[:, 2], magnitudes (U,),
    aux (U, 5) or None).

    With `cmaker` and `sitecol` the ``pointsource_distance`` approximation of
    ``_quintets`` (contexts.py:321-370) is reproduced: a source with several
    nodal planes / depths contributes, per magnitude, its ruptures for the
    sites within psdist of its location (aux mode 1) and the average rupture
    of ``iruptures`` for the others (mode 2); with at most max_sites_disagg
    sites the choice is all-or-nothing (mode 0).  aux = [x, y, z of the
    location in km, psdist, mode] per rupture is what oqg_hazard_planar_aux
    takes.
    """
    mrate = dict(src.get_annual_occurrence_rates())
    planardict = src.get_planar(shift_hypo)
    approx = cmaker is not None and src.count_nphc() > 1
    planars, mags, aux = [], [], []
    if not approx:
        for mag, arrays in planardict.items():
            if cmaker is not None and not cmaker.maximum_distance(mag) > 0:
                continue
            pla = _records(arrays, mrate[mag])
            planars.append(pla)
            mags.append(numpy.full(len(pla), mag))
            aux.append(numpy.zeros((len(pla), 5)))
    else:
        from openquake.hazardlib.geo.utils import spherical_to_cartesian
        magdist = {mag: cmaker.maximum_distance(mag) for mag in planardict}
        maxmag = max(mag for mag, dist in magdist.items() if dist > 0)
        src.max_radius(magdist[maxmag])           # sets src.radius
        cdist = sitecol.get_cdist(src.location)
        loc = spherical_to_cartesian(src.location.x, src.location.y, 0.)
        minmag = cmaker.maximum_distance.x[0]
        maxmag = cmaker.maximum_distance.x[-1]
        fewsites = len(sitecol) <= cmaker.max_sites_disagg
        for m, rup in enumerate(src.iruptures()):
            mag = rup.mag
            if mag > maxmag or mag < minmag:
                continue
            psdist = src.get_psdist(m, mag, cmaker.pointsource_distance,
                                    magdist)
            avg = _records([rup.surface.array.reshape(-1, 3)], mrate[mag])
            pla = _records(planardict[mag], mrate[mag])
            if fewsites:
                parts = [(avg if not (cdist <= psdist).any() else pla, 0.)]
            else:
                parts = [(pla, 1.), (avg, 2.)]
            for recs, mode in parts:
                planars.append(recs)
                mags.append(numpy.full(len(recs), mag))
                ax = numpy.zeros((len(recs), 5))
                ax[:, :3], ax[:, 3], ax[:, 4] = loc, psdist, mode
                aux.append(ax)
    planar = numpy.concatenate(planars)
    assert planar.dtype.itemsize == _gpu.planar_array_dt.itemsize
    aux = numpy.concatenate(aux)
    return (planar.view(_gpu.planar_array_dt).reshape(-1, 3),
            numpy.concatenate(mags), aux if aux[:, 4].any() else None)


# ------------------------------------------------------------ ContextMaker
def _msr_name(msr):
    """The name of a reference magnitude-scaling relationship as the device
    sources take it (WC1994, PointMSR, PeerMSR, 'CScalingMSR.C = <C>'), None
    if it has no device implementation"""
    name = msr.__class__.__name__
    if name == 'CScalingMSR':
        return 'CScalingMSR.C = %r' % float(msr.C)
    return name if name in ('WC1994', 'PointMSR', 'PeerMSR') else None


def fault_of_source(src):
    """
    The device-side description of a reference SimpleFaultSource
    (source/simple_fault.py): its fault mesh from the reference's own
    SimpleFaultSurface.from_fault_data, its magnitudes and rates from
    get_annual_occurrence_rates; the floating ruptures are then windows of
    that mesh (oq_engine_b200.fault.SimpleFaultSource).
    """
    from openquake.hazardlib.geo.surface.simple_fault import (
        SimpleFaultSurface)
    msr = _msr_name(src.magnitude_scaling_relationship)
    _no(src, hypo_list=len(src.hypo_list), slip_list=len(src.slip_list),
        hypo_depth_list=len(src.hypo_depth_list), msr=msr is None,
        magnitude_dependent_aspect_ratio=not isinstance(
            src.rupture_aspect_ratio, (int, float)))
    mesh = SimpleFaultSurface.from_fault_data(
        src.fault_trace, src.upper_seismogenic_depth,
        src.lower_seismogenic_depth, src.dip, src.rupture_mesh_spacing).mesh
    mags, rates = zip(*src.get_annual_occurrence_rates())
    out = _gpu.SimpleFaultSource(
        src.source_id, (mesh.lons, mesh.lats, mesh.depths),
        _gpu.ArbitraryMFD(mags, rates), src.rupture_mesh_spacing, msr,
        float(src.rupture_aspect_ratio), float(src.rake),
        src.tectonic_region_type)
    out.id = getattr(src, 'id', 0)
    out.offset = getattr(src, 'offset', 0)
    return out


class GpuContextMaker(ContextMaker):
    """
    The reference's ContextMaker with the hot operators on the device.
    Same constructor; building it with a model (or model option) without a
    device implementation raises NotImplementedError.
    """
    def __init__(self, trt, gsims, oq, monitor=Monitor(), extraparams=()):
        super().__init__(trt, gsims, oq, monitor, extraparams)
        pack_gsims(list(self.gsims), [i.string for i in self.imts])  # validate
        self._engine = None

    # the device handle lives in the worker process: never pickle it
    def __getstate__(self):
        dic = self.__dict__.copy()
        dic['_engine'] = None
        return dic

    def init_monitoring(self, monitor):
        # contexts.py:681-690: the per-worker hook; the engine (and through
        # it the device handle) is created lazily on first use after it
        super().init_monitoring(monitor)
        self._engine = None

    @property
    def engine(self):
        """The device ContextMaker configured from this one's parameters"""
        if self._engine is None:
            param = dict(
                imtls={imt: numpy.asarray(imls, float)
                       for imt, imls in self.imtls.items()},
                truncation_level=self.truncation_level,
                investigation_time=self.investigation_time,
                maximum_distance=self.maximum_distance,
                minimum_distance=self.minimum_distance,
                pointsource_distance=getattr(self, 'pointsource_distance', 0),
                max_sites_disagg=self.max_sites_disagg,
                shift_hypo=self.shift_hypo,
                # horiz_comp_to_geom_mean: the reference's own coefficients
                conv={g: {imt.string: tuple(map(float, cv))
                          for imt, cv in self.conv[gsim].items()}
                      for g, gsim in enumerate(self.gsims)
                      if self.conv.get(gsim)})
            self._engine = _gpu.ContextMaker(
                self.trt, [PackedGsim(g) for g in self.gsims], param)
        eng = self._engine
        eng.tom = getattr(self, 'tom', None)
        eng.cluster = getattr(self, 'cluster', 0)
        return eng

    # ---- operators of the hot path --------------------------------------
    def get_mean_stds(self, ctxs, split_by_mag=True):
        """(4, G, M, N) float64 on the device (contexts.py:1250-1279)"""
        with self.gmf_mon:
            return self.engine.get_mean_stds(ctxs, split_by_mag)

    def gen_poes(self, ctx):
        """Yields (poes (n, L, G), mea, sig, tau, ctxt) per magnitude like
        contexts.py:1155-1203, PoEs computed on the device"""
        eng = self.engine
        with self.poe_mon:
            for tup in eng.gen_poes(ctx):
                yield tup
        self.cfactor += eng.cfactor
        eng.cfactor[:] = 0

    def update(self, pmap, ctx, rup_mutex=None):
        """contexts.py:1232-1247 with the PoEs and the read-modify-write of
        the map on the device; `pmap` is the reference's MapArray and is
        updated in place"""
        dmap = _device_map(pmap)
        with self.poe_mon:
            self.engine.update(dmap, ctx, rup_mutex)
        if dmap.array is not pmap.array:      # pragma: no cover
            pmap.array[:] = dmap.array

    # ---- the fused pass ---------------------------------------------------
    def update_from_sources(self, pmap, sources, sitecol):
        """
        Fold every rupture of point-like `sources` on `sitecol` into `pmap`
        in fused passes (distances, GMPE, PoEs and accumulation without
        materialising context rows): the equivalent of
        ``for src in sources: for ctx in get_ctxs(src, sitecol):
        update(pmap, ctx)``.  Returns the number of rupture-site pairs kept.
        """
        eng = self.engine
        if eng.conv:
            raise NotImplementedError(
                'the fused pass has no horizontal component conversion: use '
                'update()')
        dsites = _device_sitecol(sitecol)
        dmap = _device_map(pmap)
        kept = 0
        with self.poe_mon:
            for src in sources:
                if src.__class__.__name__ == 'SimpleFaultSource':
                    # genctxs of a fault source: distance planes of the
                    # rupture meshes + the fused pass (oqg_hazard_dist)
                    kept += eng.hazard_mesh(dmap, fault_of_source(src),
                                            dsites, count=True)
                    continue
                if not hasattr(src, 'get_planar'):
                    raise NotImplementedError(
                        '%s has no device implementation of its ruptures '
                        '(point-like and simple fault sources only)'
                        % src.__class__.__name__)
                planar, mags, aux = planar_of_source(
                    src, self.shift_hypo, self, sitecol)
                kept += eng.hazard_planar(
                    dmap, _gpu.PlanarRuptures(planar, mags, src.id), dsites,
                    count=True, aux=aux)
        return kept


def _device_map(pmap):
    """An oq_engine_b200 MapArray sharing the numpy array of a reference
    MapArray: the device operators update the reference's array in place"""
    dmap = _gpu.MapArray(pmap.sids, pmap.shape[1], pmap.shape[2],
                         bool(pmap.rates))
    dmap.array = pmap.array
    return dmap


class GpuRmapMaker(RmapMaker):
    """RmapMaker whose independent-sources driver runs the fused pass
    (contexts.py:1548-1579)"""
    def _make_src_indep(self):
        cm = self.cmaker
        sitecol = self.srcfilter.sitecol
        pnemap = MapArray(sitecol.sids, cm.imtls.size, len(cm.gsims),
                          not self.cluster).fill(self.cluster)
        kept = cm.update_from_sources(pnemap, self.sources, sitecol)
        for src in self.sources:
            self.update_source_data(src, cm.task_no, 0., kept / max(
                len(self.sources), 1))
        return pnemap


def use_gpu(cmaker):
    """A GpuContextMaker with the parameters of a stock ContextMaker"""
    new = copy.copy(cmaker)
    new.__class__ = GpuContextMaker
    new._engine = None
    pack_gsims(list(new.gsims), [i.string for i in new.imts])
    return new
