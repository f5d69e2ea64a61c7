"""CPU tests of the host side: the C-ABI library loads and exports every
symbol include/oqg.h declares, the coefficient packing, the planar rupture
builder, the TOMs, MapArray and the loud failure without a GPU."""
import os
import re
import json
import ctypes
import numpy as np
import pytest

import geo_np
import gmpe_np
from helpers import ROOT, load, rup_of, sites_of, pipeline_cases, ctx_of

import oq_engine_b200 as oq


def header_functions():
    with open(os.path.join(ROOT, 'include', 'oqg.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(oqg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(oq._lib.LIBPATH)
    names = header_functions()
    assert len(names) >= 18, names
    for name in names:
        assert hasattr(lib, name), 'missing symbol %s' % name
    # and the Python binding knows all of them
    bound = oq._lib.lib()._signatures
    assert set(names) == set(bound), set(names) ^ set(bound)


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        oq.contexts.get_handle()
    h = ctypes.c_void_p()
    L = oq._lib.lib()
    assert L.oqg_create(ctypes.byref(h), 0) != 0
    assert b'no CPU fallback' in L.oqg_last_error(None)
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()],
                         dict(imtls={'PGA': [0.1]}))
    with pytest.raises(RuntimeError):
        cm.handle()


def test_family_fields_resolve_for_all_models():
    imts = ['PGA', 'PGV', 'SA(0.05)', 'SA(0.27)', 'SA(1.0)', 'SA(3.0)']
    for name, cls in oq.registry.items():
        gsim = cls()
        fields = oq._lib.family_fields(gsim.family)
        assert 'period' in fields and len(fields) <= oq._lib.KMAX
        for imt in imts:
            if imt == 'PGV' and name in ('ToroEtAl2002', 'Campbell2003',
                                         'ToroEtAl2002SHARE',
                                         'Campbell2003SHARE',
                                         'SadighEtAl1997'):
                with pytest.raises(KeyError):     # not in their tables
                    gsim.coeff_row(imt)
                continue
            if imt == 'SA(0.05)' and name == 'SadighEtAl1997':
                with pytest.raises(KeyError):     # below its shortest period
                    gsim.coeff_row(imt)
                continue
            row = gsim.coeff_row(imt)
            assert len(row) == oq._lib.KMAX and np.isfinite(row).all()
        fam, flags, par = gsim.device_spec()
        assert 0 <= fam < 11 and len(par) == 4


def test_coeffs_table_interpolation_matches_oracle():
    # log-period interpolation, coeffs_table.py:263-277
    for table in ('BSSA14', 'CY14', 'ASK14', 'CB14', 'BA08', 'AKKAR14_RJB'):
        for imt in ('PGA', 'SA(0.27)', 'SA(0.7)', 'SA(2.0)'):
            mine = oq.CoeffsTable.get(table)[imt]
            ref = gmpe_np.coeffs(table, imt)
            assert mine.keys() == ref.keys()
            for k in ref:
                assert mine[k] == ref[k], (table, imt, k)
    with pytest.raises(KeyError):
        oq.CoeffsTable.get('AB06_BC')['SA(9.0)']
    with pytest.raises(KeyError):
        oq.CoeffsTable.get('BSSA14')['MMI']


def test_unsupported_models_and_options():
    with pytest.raises(NotImplementedError):
        oq.get_gsim('ZhaoEtAl2006Asc')
    with pytest.raises(NotImplementedError):
        oq.ChiouYoungs2014(usgs_basin_scaling=True)
    with pytest.raises(ValueError):
        oq.BooreEtAl2014(region='XYZ')
    with pytest.raises(KeyError):
        oq.ContextMaker('*', [oq.CampbellBozorgnia2014()],
                        dict(imtls={'CAV': [0.1]}))
    with pytest.raises(ValueError):
        oq.ContextMaker('*', [oq.BooreEtAl2014()],
                        dict(imtls={'PGA': [0.1, 0.2], 'SA(1.0)': [0.1]}))


def test_build_planar_matches_reference_fixture():
    d = load('planar')
    r = rup_of(d)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], float(d['usd']), float(d['lsd']),
                          float(d['rar']))
    assert pla.dtype == oq.planar_array_dt and pla.shape == (96, 3)
    for f in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr', 'sdr', 'hypo'):
        np.testing.assert_allclose(pla[f], d['planar_' + f], rtol=0,
                                   atol=1e-10, err_msg=f)
    flat = oq.planar.as_flat(pla)
    assert flat.shape == (96, 42)
    # memory layout [k][14]: corners xyz normal uv1 uv2 wlr sdr hypo
    np.testing.assert_array_equal(flat[:, 12], pla['sdr'][:, 0])
    np.testing.assert_array_equal(flat[:, 14 + 4], pla['xyz'][:, 1, 0])
    np.testing.assert_allclose(oq.wc1994_median_area(r['mag'], r['rake']),
                               r['area'], rtol=1e-14)


def test_build_planar_clamps_to_seismogenic_layer():
    # shallow and deep hypocentres: the rupture is shifted to fit [usd, lsd]
    n = 6
    mag = np.full(n, 7.0)
    dep = np.array([1., 3., 10., 18., 19.5, 12.])
    dip = np.array([90., 45., 30., 60., 90., 45.])
    strike = np.linspace(0, 300, n)
    rake = np.zeros(n)
    area = oq.wc1994_median_area(mag, rake)
    args = (mag, strike, dip, rake, np.ones(n), area, np.zeros(n),
            np.zeros(n), dep, 0., 20., 2.)
    mine = oq.build_planar(*args)
    ref = geo_np.build_planar(*args)
    for f in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr', 'hypo'):
        np.testing.assert_allclose(mine[f], ref[f], rtol=0, atol=1e-9,
                                   err_msg=f)
    assert mine['corners'][:, 2].min() >= -1e-9
    assert mine['corners'][:, 2].max() <= 20 + 1e-9


def test_negbin_tom_matches_reference_fixture():
    d = load('pipeline_negbin')
    nb = json.loads(str(d['negbin']))
    tom = oq.NegativeBinomialTOM(float(d['investigation_time']), **nb)
    ctx = ctx_of(d)
    pmf = tom.get_pmf(ctx.occurrence_rate)
    np.testing.assert_allclose(pmf, ctx.probs_occur, rtol=1e-12, atol=1e-300)
    assert tom.get_pmf(0.01).shape == (10,)
    with pytest.raises(ValueError):
        oq.NegativeBinomialTOM(50., mu=1., alpha=0.)
    with pytest.raises(ValueError):
        oq.PoissonTOM(0)


def test_integration_distance():
    md = oq.IntegrationDistance.new('[(4., 50), (8., 200.)]')
    f = md('Active Shallow Crust')
    assert f(6.0) == 125. and f(4.0) == 50. and f(8.0) == 200.
    # outside the pairs the distance is ZERO like the reference's
    # interp1d(bounds_error=False, fill_value=0.) (calc/filters.py:183-190):
    # such ruptures are dropped by the rrup < magdist filter
    assert f(3.0) == 0. and f(9.0) == 0.
    g = oq.IntegrationDistance.new('300')('x')
    assert g(7.3) == 300. and g(2.4) == 0. and g(10.3) == 0.
    # parsed like the reference's floatdict (ast.literal_eval, never eval)
    with pytest.raises(ValueError):
        oq.IntegrationDistance.new('__import__("os").getcwd()')
    with pytest.raises(ValueError, match='Invalid magnitude'):
        oq.IntegrationDistance.new('[(1., 50), (8., 200.)]')
    with pytest.raises(ValueError, match='not ordered'):
        oq.IntegrationDistance.new('[(8., 200.), (4., 50)]')
    md2 = oq.IntegrationDistance.new(
        "{'default': 200, 'Stable Shallow Crust': [(5, 100), (7, 300)]}")
    assert md2('Stable Shallow Crust')(6.) == 200. and md2('x')(6.) == 200.
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()],
                         dict(imtls={'PGA': [0.1]}, maximum_distance=md))
    np.testing.assert_allclose(cm.magdist([4., 6., 8.]), [50., 125., 200.])
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()], dict(imtls={'PGA': [.1]}))
    assert cm.magdist([5.])[0] == 1000.   # default, contexts.py:292-296


def test_map_array_host_semantics():
    sids = np.array([5, 2, 9], np.uint32)
    pmap = oq.MapArray(sids, 4, 2).fill(1)
    assert pmap.array.dtype == np.float32 and pmap.array.shape == (3, 4, 2)
    np.testing.assert_array_equal(pmap.sidx[sids], [0, 1, 2])
    pmap.array[:] = np.float32(0.9)
    poes = ~pmap
    np.testing.assert_allclose(poes.array, 0.1, rtol=1e-6)
    rates = pmap.to_rates()
    assert rates.rates
    np.testing.assert_allclose(rates.array, -np.log(np.float32(0.9)),
                               rtol=1e-6)
    rates += rates.new(rates.array.copy())
    np.testing.assert_allclose(rates.array, -2 * np.log(np.float32(0.9)),
                               rtol=1e-6)
    with pytest.raises(AssertionError):
        oq.MapArray(sids, 4, 2).fill(2)


def test_context_maker_layout():
    gsims = [oq.BooreEtAl2014(), oq.ChiouYoungs2014(),
             oq.AbrahamsonEtAl2014(), oq.CampbellBozorgnia2014(),
             oq.BooreAtkinson2008()]
    cm = oq.ContextMaker('*', gsims, dict(
        imtls={'PGA': [.1, .2], 'SA(1.0)': [.1, .2]}, truncation_level=3.))
    # field order observed in the reference
    names = list(cm.ctx_dtype(0).names)
    assert names == ['dip', 'hypo_depth', 'mag', 'rake', 'width', 'ztor',
                     'vs30', 'vs30measured', 'z1pt0', 'z2pt5', 'rjb', 'rrup',
                     'rx', 'ry0', 'src_id', 'rup_id', 'sids',
                     'occurrence_rate', 'probs_occur']
    assert (cm.G, cm.M, cm.L1, cm.L) == (5, 2, 2, 4)
    np.testing.assert_allclose(cm.phi_b, 0.9986501019683699, rtol=1e-15)
    import pickle
    cm2 = pickle.loads(pickle.dumps(cm))   # crosses process boundaries
    assert cm2.G == 5 and cm2._structs[1].shape == (5, 3, oq._lib.KMAX)


def test_site_tiles():
    sc = oq.SiteCollection.from_points(np.linspace(0, 1, 10), np.zeros(10),
                                       None, dict(vs30=760.))
    tiles = [oq.sharding.site_tile(sc, k, 3) for k in range(3)]
    assert sorted(np.concatenate([t.sids for t in tiles])) == list(range(10))
    np.testing.assert_array_equal(tiles[1].sids, [1, 4, 7])
    blocks = [oq.sharding.site_tile(sc, k, 3, 'block') for k in range(3)]
    np.testing.assert_array_equal(blocks[0].sids, [0, 1, 2])
    assert oq.sharding.groups_of(1, 4, 8) == [1, 5]


def test_monitors_have_the_reference_names():
    """init_monitoring (contexts.py:681-690): the child monitors the
    reference's calculators read, created under the monitor passed in"""
    mon = oq.Monitor('classical')
    mon.task_no = 7
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()], dict(
        imtls={'PGA': [.1, .2]}, truncation_level=3.), monitor=mon)
    ops = [c.operation for c in mon.children]
    assert ops[:5] == ['computing mean_std', 'get_poes', 'iter_ruptures',
                       'building dparam', 'getting delta_rates']
    assert cm.gmf_mon is mon.children[0] and cm.poe_mon is mon.children[1]
    assert cm.task_no == 7 and cm.out_no == 7
    with cm.poe_mon:
        pass
    with cm.poe_mon:
        pass
    assert cm.poe_mon.counts == 2 and cm.poe_mon.duration >= 0
    assert mon.get_data() == [('get_poes', cm.poe_mon.duration, 2)]
    # a ContextMaker made without a monitor has its own
    cm2 = oq.ContextMaker('*', [oq.BooreEtAl2014()], dict(
        imtls={'PGA': [.1, .2]}, truncation_level=3.))
    assert cm2.gmf_mon.operation == 'computing mean_std'


def test_multi_point_source_splits_into_point_sources():
    """source/multi_point.py: one point source per location with its own MFD"""
    mfds = [oq.TruncatedGRMFD(5.5, 7.0, 0.1, a, 0.9) for a in (3.0, 3.2)]
    mps = oq.MultiPointSource('3', [0.2, 0.8], [-0.3, -0.3], mfds, 'WC1994',
                              1.5, 0., 20., [(1., 0., 60., 90.)],
                              [(.25, 5.), (.5, 10.), (.25, 15.)],
                              [0.5, 0.6667, 0.75])
    pts = list(mps)
    assert [p.source_id for p in pts] == ['3:0', '3:1'] and len(mps) == 2
    assert (pts[1].lon, pts[1].lat) == (0.8, -0.3)
    assert pts[0].mfd is mfds[0] and pts[1].mfd is mfds[1]
    assert pts[0].hypo_dip_fracs == [0.5, 0.6667, 0.75]
    assert mps.count_ruptures() == 2 * 15 * 3
    with pytest.raises(ValueError):
        oq.MultiPointSource('x', [0.], [0., 1.], mfds, 'WC1994', 1., 0., 10.,
                            [(1., 0., 90., 0.)], [(1., 5.)])
    with pytest.raises(NotImplementedError):
        oq.MultiPointSource('x', [0.], [0.], mfds[:1], 'Leonard2014_SCR', 1.,
                            0., 10., [(1., 0., 90., 0.)], [(1., 5.)])


def test_integration_distance_cut():
    """IntegrationDistance.cut, the doctest values of calc/filters.py:232-240"""
    md = oq.IntegrationDistance.new('[(4., 50), (8., 200.)]')
    md.cut({'default': 5.})
    assert md == {'default': [(5.0, 87.5), (8.0, 200.0)]}
    md = oq.IntegrationDistance.new('200')
    md.cut({"Active Shallow Crust": 5.2, "default": 4.})
    assert md == {'default': [(4.0, 200.0), (10.2, 200)],
                  'Active Shallow Crust': [(5.2, 200.0), (10.2, 200)]}
    # smaller magnitudes have no integration distance: their ruptures go
    f = md('Active Shallow Crust')
    assert f(5.1) == 0. and f(5.2) == 200. and f(7.) == 200.
