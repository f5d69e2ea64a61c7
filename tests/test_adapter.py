"""Tests of the reference-side binding adapters/openquake_gpu.py.

* With the reference present (the build container; needs oracle/refshim for
  its absent h5py / shapely): the adapter packs, from the reference's OWN
  GSIM objects and CoeffsTables, structs and coefficient rows that are
  byte-identical to those of the stand-alone mirror (oq_engine_b200.gsim) for
  all the supported classes and flags; unsupported models and options are
  NotImplementedError; an IMT outside a table is KeyError; GpuContextMaker
  pickles; sites and planar records come out in the layouts the C ABI takes.
* Everywhere (no reference needed): the packed bytes of the committed fixture
  tests/golden/adapter_packed.npz equal the mirror's packing, and - on the GPU
  box - feeding them to the C ABI reproduces what the STOCK reference computed
  from the same objects on the CPU (test_replay_packed_fixture).
"""
import os
import sys
import ast
import ctypes
import pickle
import numpy as np
import pytest

from helpers import ROOT, load

REFERENCE = os.environ.get('OQ_REFERENCE', '/root/reference')
needs_ref = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REFERENCE, 'openquake')),
    reason='the reference is not on this machine')

CASES = [
    ('BooreEtAl2014', {}), ('BooreEtAl2014', dict(region='CAL')),
    ('BooreEtAl2014', dict(region='JPN', sof=False)),
    ('BooreEtAl2014', dict(sigma_mu_epsilon=1.)),
    ('ChiouYoungs2014', {}), ('ChiouYoungs2014', dict(region='JPN')),
    ('ChiouYoungs2014', dict(region='ITA', alpha_nm=0.8)),
    ('ChiouYoungs2014', dict(region='WEN', use_hw=False)),
    ('AbrahamsonEtAl2014', {}), ('AbrahamsonEtAl2014', dict(region='JPN')),
    ('AbrahamsonEtAl2014', dict(region='TWN')),
    ('AbrahamsonEtAl2014', dict(region='CHN', sigma_mu_epsilon=-1.)),
    ('CampbellBozorgnia2014', {}),
    ('CampbellBozorgnia2014', dict(SJ=True, sigma_mu_epsilon=1.)),
    ('BooreAtkinson2008', {}), ('AtkinsonBoore2006', {}),
    ('AtkinsonBoore2006', dict(mag_eq='Mblg87')),
    ('AtkinsonBoore2006', dict(mag_eq='Mblg96', scale_fac=0.5)),
    ('AkkarEtAlRjb2014', {}), ('AkkarEtAlRepi2014', {}),
    ('AkkarEtAlRhyp2014', dict(adjustment_factor=1.2)),
    ('ChiouYoungs2008', {}), ('ToroEtAl2002', {}), ('Campbell2003', {}),
    ('ToroEtAl2002SHARE', {}), ('Campbell2003SHARE', {}),
]
IMTS = ['PGA', 'SA(0.1)', 'SA(0.27)', 'SA(1.0)', 'SA(3.0)']


@pytest.fixture(scope='module')
def ref():
    """(adapter module, {name: reference GSIM class})"""
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
    for p in (os.path.join(ROOT, 'oracle', 'refshim'),
              os.path.join(ROOT, 'adapters')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oq_bootstrap
    oq_bootstrap.install()
    import openquake_gpu
    from openquake.hazardlib.gsim import get_available_gsims
    return openquake_gpu, get_available_gsims()


@needs_ref
@pytest.mark.parametrize('name,kw', CASES)
def test_packing_is_byte_identical_to_the_mirror(ref, name, kw):
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import _gsim_structs
    og, avail = ref
    arr, rows = og.pack_gsims([avail[name](**kw)], IMTS)
    arr2, rows2 = _gsim_structs([oq.get_gsim(name, **kw)],
                                [oq.from_string(i) for i in IMTS])
    assert bytes(arr) == bytes(arr2)
    assert rows.tobytes() == rows2.tobytes()
    # the REQUIRES_* sets seen by the device plumbing are the reference's
    pg = og.PackedGsim(avail[name](**kw))
    mir = oq.get_gsim(name, **kw)
    for req in ('DISTANCES', 'RUPTURE_PARAMETERS', 'SITES_PARAMETERS'):
        assert getattr(pg, 'REQUIRES_' + req) == getattr(
            mir, 'REQUIRES_' + req), (name, req)


@needs_ref
def test_error_mapping(ref):
    og, avail = ref
    with pytest.raises(NotImplementedError, match='no device implementation'):
        og.pack_gsims([avail['ZhaoEtAl2006Asc']()], ['PGA'])
    with pytest.raises(NotImplementedError, match='usgs_basin_scaling'):
        og.pack_gsims([avail['BooreEtAl2014'](
            region='CAL', usgs_basin_scaling=True)], ['PGA'])
    with pytest.raises(NotImplementedError, match='estimate_width'):
        og.pack_gsims([avail['CampbellBozorgnia2014'](estimate_width=True)],
                      ['PGA'])
    with pytest.raises(NotImplementedError, match='add_delta_c1'):
        og.pack_gsims([avail['ChiouYoungs2014'](add_delta_c1=True)], ['PGA'])
    with pytest.raises(KeyError):      # outside the table, as the reference
        og.pack_gsims([avail['BooreEtAl2014']()], ['SA(20.0)'])
    with pytest.raises(KeyError):
        og.pack_gsims([avail['CampbellBozorgnia2014']()], ['IA'])
    # a GpuContextMaker refuses at construction, not at first use
    with pytest.raises(NotImplementedError):
        og.GpuContextMaker('*', [avail['ZhaoEtAl2006Asc']()],
                           dict(imtls={'PGA': [0.1]}))


@needs_ref
def test_gpu_context_maker_pickles_and_keeps_the_reference_api(ref):
    og, avail = ref
    from openquake.hazardlib.contexts import ContextMaker
    gsims = [avail['BooreEtAl2014'](), avail['ChiouYoungs2014']()]
    param = dict(imtls={'PGA': [0.01, 0.1], 'SA(1.0)': [0.01, 0.1]},
                 truncation_level=3., investigation_time=50.,
                 maximum_distance={'default': [(2.5, 200.), (10.2, 200.)]})
    cm = og.GpuContextMaker('Active Shallow Crust', gsims, param)
    assert isinstance(cm, ContextMaker)
    for name in ('get_mean_stds', 'gen_poes', 'update', 'get_pmap',
                 'get_rmap', 'get_ctxs', 'from_srcs', 'init_monitoring'):
        assert callable(getattr(cm, name))
    for name in ('poe_mon', 'gmf_mon', 'ir_mon'):
        assert hasattr(cm, name), name       # contexts.py:681-690
    cm2 = pickle.loads(pickle.dumps(cm))
    assert isinstance(cm2, og.GpuContextMaker) and cm2._engine is None
    assert [repr(g) for g in cm2.gsims] == [repr(g) for g in cm.gsims]
    # without a device the operators fail loudly (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        ctx = cm.new_ctx(1)
        with pytest.raises(RuntimeError, match='no CUDA device'):
            cm.get_mean_stds([ctx])
    stock = ContextMaker('Active Shallow Crust', gsims, param)
    assert isinstance(og.use_gpu(stock), og.GpuContextMaker)


@needs_ref
def test_sites_and_planar_records_come_from_reference_objects(ref):
    og, avail = ref
    import oq_engine_b200 as oq
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import make_golden_adapter as mk
    sc = mk.sitecol(17)
    cols = og.site_columns(sc)
    assert set(cols) == set(og.SITE_COLUMNS)
    for name, col in cols.items():
        assert col.flags['C_CONTIGUOUS'] and len(col) == 17
        assert col.dtype == (np.uint8 if name == 'vs30measured' else float)
        np.testing.assert_array_equal(col, sc.array[name])
    dsc = og._device_sitecol(sc)
    np.testing.assert_array_equal(dsc.sids, sc.sids)
    np.testing.assert_array_equal(dsc.array['vs30'], sc.array['vs30'])
    # PointSource.get_planar(): the records upload unchanged, ordered by
    # magnitude, nodal plane, depth, with the rate in wlr[:, 2]
    src = mk.sources()[0]
    planar, mags, aux = og.planar_of_source(src)
    assert aux is None
    assert planar.dtype == oq.planar_array_dt and planar.shape == (16, 3)
    rates = dict(src.get_annual_occurrence_rates())
    assert sorted(set(mags)) == sorted(rates)
    for mag in rates:                 # plane weights 0.6 / 0.4 x depths 0.5
        np.testing.assert_allclose(planar['wlr'][mags == mag][:, 2].sum(),
                                   rates[mag], rtol=1e-12)
    # the geometry is the reference's, untouched
    raw = np.concatenate([p.reshape(-1, 3) for [p] in
                          src.get_planar().values()])
    for name in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'sdr', 'hypo'):
        np.testing.assert_array_equal(planar[name], raw[name])
    np.testing.assert_array_equal(planar['wlr'][:, :2], raw['wlr'][:, :2])


# ----------------------------------------------- fixture: runs everywhere
def fixture_gsims(oq):
    d = load('adapter_packed')
    return d, [oq.get_gsim(str(n), **ast.literal_eval(str(k)))
               for n, k in zip(d['gsim_names'], d['gsim_kwargs'])]


def test_fixture_bytes_equal_the_mirror_packing():
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import _gsim_structs
    d, gsims = fixture_gsims(oq)
    arr, rows = _gsim_structs(gsims, [oq.from_string(str(i))
                                      for i in d['imts']])
    assert bytes(arr) == d['gsim_bytes'].tobytes()
    assert rows.tobytes() == d['rows'].tobytes()


@pytest.mark.gpu
def test_replay_packed_fixture():
    """What the adapter packed from the reference's objects, fed to the C ABI
    (no mirror gsim, site or source objects involved), against the stock
    reference's CPU results from the same objects"""
    import torch
    import oq_engine_b200 as oq
    from oq_engine_b200 import _lib
    from oq_engine_b200.contexts import _dev, _ptr
    d = load('adapter_packed')
    G, M = len(d['gsim_names']), len(d['imts'])
    L1 = d['loglevels'].shape[1]
    h = _lib.Handle(0)
    arr = (_lib.OqgGsim * G).from_buffer_copy(d['gsim_bytes'].tobytes())
    rows = np.ascontiguousarray(d['rows'])
    h.call('oqg_set_gsims', G, arr, M, rows.ctypes.data_as(_lib.c_double_p))
    lev = np.ascontiguousarray(d['loglevels'])
    h.call('oqg_set_levels', M, L1, lev.ctypes.data_as(_lib.c_double_p),
           float(d['truncation_level']))
    keep, st = [], _lib.OqgSites()
    for name in ('lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5',
                 'vs30measured'):
        t = _dev(d['site_' + name], 0)
        keep.append(t)
        setattr(st, name, t.data_ptr())
    N, U = len(d['sids']), len(d['mag'])
    t_pla, t_mag = _dev(d['planar'], 0), _dev(d['mag'], 0)
    t_md = _dev(d['magdist'], 0)
    t_map = torch.ones((N, M * L1, G), dtype=torch.float32, device='cuda:0')
    t_cnt = torch.zeros(2, dtype=torch.int64, device='cuda:0')
    t_aux = _dev(d['aux'], 0)     # pointsource_distance of _quintets
    h.call('oqg_hazard_planar_aux', _ptr(t_pla), _ptr(t_mag), _ptr(t_md),
           _ptr(t_aux), None, 0, U, ctypes.byref(st), None, N, 0.,
           float(d['investigation_time']), 0, _ptr(t_map), _ptr(t_cnt))
    h.call('oqg_synchronize')
    assert int(t_cnt[0]) == int(d['nrows'])      # same rows as from_srcs
    pne = t_map.cpu().numpy()
    # the reference's get_pmap returns PoEs built by float32 read-modify-write
    np.testing.assert_allclose(pne, 1. - d['poes'], rtol=2e-5, atol=1e-7)
    # mean and sigma of the rows the reference built
    ctx = oq._lib.OqgCtx()
    for name in ('mag', 'rake', 'dip', 'ztor', 'width', 'hypo_depth', 'rrup',
                 'rjb', 'rx', 'ry0', 'vs30', 'z1pt0', 'z2pt5',
                 'vs30measured'):
        t = _dev(d['ctx_' + name], 0)
        keep.append(t)
        setattr(ctx, name, t.data_ptr())
    P = int(d['nrows'])
    out = torch.empty((4, G, M, P), dtype=torch.float64, device='cuda:0')
    h.call('oqg_mean_stds', ctypes.byref(ctx), P, _ptr(out))
    np.testing.assert_allclose(out.cpu().numpy(), d['mean_stds'], rtol=1e-10,
                               atol=1e-12)
    h.close()


@pytest.mark.gpu
def test_horizontal_component_conversion_replay():
    """ContextMaker.horiz_comp_to_geom_mean (contexts.py:726-776): the
    reference's conversion coefficients for RotD50 (BSSA14) and GMRotI50
    (BA08) fed to oqg_set_conversion; oqg_mean_stds then reproduces the
    reference's converted mean and sigma"""
    import torch
    from oq_engine_b200 import _lib
    from oq_engine_b200.contexts import _dev, _ptr
    d = load('adapter_packed')
    G, M = len(d['conv_names']), len(d['imts'])
    h = _lib.Handle(0)
    arr = (_lib.OqgGsim * G).from_buffer_copy(d['conv_gsim_bytes'].tobytes())
    rows = np.ascontiguousarray(d['conv_rows'])
    h.call('oqg_set_gsims', G, arr, M, rows.ctypes.data_as(_lib.c_double_p))
    conv = np.ascontiguousarray(d['conv'])
    h.call('oqg_set_conversion', conv.ctypes.data_as(_lib.c_double_p))
    keep, ctx = [], _lib.OqgCtx()
    for name in ('mag', 'rake', 'rjb', 'vs30'):
        t = _dev(d['conv_' + name], 0)
        keep.append(t)
        setattr(ctx, name, t.data_ptr())
    P = len(d['conv_mag'])
    out = torch.empty((4, G, M, P), dtype=torch.float64, device='cuda:0')
    h.call('oqg_mean_stds', ctypes.byref(ctx), P, _ptr(out))
    got = out.cpu().numpy()
    np.testing.assert_allclose(got, d['conv_mean_stds'], rtol=1e-10,
                               atol=1e-12)
    # and it is a real change for the converted models only
    h.call('oqg_set_conversion', None)
    h.call('oqg_mean_stds', ctypes.byref(ctx), P, _ptr(out))
    raw = out.cpu().numpy()
    assert np.abs(raw[0, 0] - got[0, 0]).max() > 1e-3      # RotD50
    assert np.abs(raw[1, 1] - got[1, 1]).max() > 1e-3      # GMRotI50 sigma
    np.testing.assert_array_equal(raw[:, 2], got[:, 2])    # geometric mean
    # the fused pass refuses to run with a conversion set
    h.call('oqg_set_conversion', conv.ctypes.data_as(_lib.c_double_p))
    lev = np.zeros((M, 2))
    h.call('oqg_set_levels', M, 2, lev.ctypes.data_as(_lib.c_double_p), 3.)
    with pytest.raises(RuntimeError, match='horizontal-component'):
        h.call('oqg_hazard_planar', None, None, None, None, 0, 1,
               ctypes.byref(_lib.OqgSites()), None, 1, 0., 1., 0, None, None)
    h.close()


def test_reference_arm_runs_the_installed_reference():
    """bench.py --impl reference: the unmodified reference from baseline/_ref
    (oracle/ref_runner.py) gives the map of the oracle on the same inputs"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_runner
    if not ref_runner.available():
        pytest.skip('the reference is not installed under baseline/_ref')
    import hazard_np
    import geo_np
    from helpers import synth_ruptures, synth_sites, IMTS
    rng = np.random.default_rng(3)
    r, sites = synth_ruptures(24, rng), synth_sites(40, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    names = ['BooreEtAl2014', 'ChiouYoungs2014']
    pla = ref_runner.build_planar(r)
    rows, dt, arr = ref_runner.hazard_pass(names, imtls, pla, r['mag'], sites)
    assert rows == 24 * 40 and arr.shape == (40, 140, 2)
    pairs = [(n, {}) for n in names]
    args = (r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
            r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    octx = hazard_np.get_ctx(pairs, geo_np.build_planar(*args), r['mag'],
                             sites, None, 0)
    acc = hazard_np.exact_rate_sum(pairs, octx, sites['sids'], imtls, 3., 50.)
    np.testing.assert_allclose(arr, np.exp(-acc), rtol=2e-5, atol=1e-7)


@needs_ref
def test_fault_source_of_the_reference(ref):
    """a reference SimpleFaultSource becomes the device-side description
    whose floating ruptures are the reference's own (the fault of the
    LogicTreeCase2ClassicalPSHA demo, fixture of oracle/make_golden_fault.py)"""
    og, _avail = ref
    from openquake.hazardlib.geo import Point, Line
    from openquake.hazardlib.mfd import TruncatedGRMFD
    from openquake.hazardlib.scalerel.wc1994 import WC1994
    from openquake.hazardlib.source import SimpleFaultSource
    from openquake.hazardlib.tom import PoissonTOM
    d = load('fault_ctxs')
    src = SimpleFaultSource(
        'second', 'Simple Fault Source', 'Active Shallow Crust',
        TruncatedGRMFD(*d['mfd']), float(d['spacing']), WC1994(), 2.0,
        PoissonTOM(50.), 8., 20., Line([Point(1.0, -0.5), Point(1.4, 0.0),
                                        Point(1.4, 0.3)]), 30., 90.)
    dev = og.fault_of_source(src)
    np.testing.assert_array_equal(dev.mesh[0], d['mesh_lons'])
    np.testing.assert_array_equal(dev.mesh[2], d['mesh_deps'])
    mag, rate, win = dev.windows()
    np.testing.assert_array_equal(win, d['windows'])
    np.testing.assert_allclose(mag, d['mag'], rtol=1e-15)
    np.testing.assert_allclose(rate, d['rate'], rtol=1e-14)
    assert dev.count_ruptures() == src.count_ruptures() == 400
    import copy
    bad = copy.copy(src)
    bad.hypo_list = ((0.5, 0.5, 1.),)
    with pytest.raises(NotImplementedError, match='hypo_list'):
        og.fault_of_source(bad)
