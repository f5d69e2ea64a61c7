"""RmapMaker.make (openquake/hazardlib/contexts.py:1478-1662): independent
sources, mutually exclusive sources and a cluster group against the rates
maps produced by the reference's own RmapMaker (oracle/make_golden_rmap.py).

Tolerance: the reference accumulates in float32 per context row (rate maps:
`arr += rate * poe * itime`; mutex / cluster: pne maps), the device once per
map element: rtol 2e-5 on rates >= 1e-7 (the float32 noise of ~30 updates),
plus an absolute 2e-7 floor for the maps that pass through 1 - pne.
"""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'rmap_maker.npz'))
SRC = [  # as in oracle/make_golden_rmap.py
    (0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
    (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
    (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90., 10.),
]


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def setup(oq):
    srcs = []
    for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC):
        src = oq.PointSource(
            str(i), lon, lat, oq.EvenlyDiscretizedMFD(5.05, 0.3, rates),
            'WC1994', 1.5, 0., 20., [(1., strike, dip, rake)], [(1., dep)])
        src.mutex_weight = float(D['mutex_weights'][i])
        srcs.append(src)
    N = len(D['site_lon'])
    sc = oq.SiteCollection.from_points(
        D['site_lon'], D['site_lat'], np.zeros(N), dict(vs30=D['vs30']))
    imtls = {str(imt): D['levels'] for imt in D['imts']}
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('100')))
    return cm, sc, srcs


def close(got, want, atol):
    assert got.shape == want.shape and got.dtype == np.float32
    big = want >= 1e-7
    np.testing.assert_allclose(got[big], want[big], rtol=2e-5, atol=atol)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=max(atol, 1e-9))


def test_host_validation():
    import oq_engine_b200 as oq
    with pytest.raises(ValueError):
        oq.SourceGroup('*', [], src_interdep='sometimes')
    g = oq.SourceGroup('*', [1, 2], grp_probability=0.5)
    assert list(g) == [1, 2] and g.grp_probability == 0.5


@pytest.mark.gpu
def test_independent_sources(oq):
    cm, sc, srcs = setup(oq)
    dic = oq.RmapMaker(cm, sc, srcs).make()
    assert set(dic) >= {'rmap', 'cfactor', 'rup_data', 'source_data',
                        'task_no'}
    rmap = dic['rmap']
    assert rmap.rates and (rmap.gid == np.arange(2)).all()
    close(rmap.array, D['indep_rates'], 0.)
    sd = dic['source_data']
    assert sd['nrupts'] == [4, 3, 5] and len(sd['ctimes']) == 3
    # the reference reports the context rows of each source
    np.testing.assert_allclose(sum(sd['nctxs']), D['indep_nctxs'].sum(),
                               rtol=0, atol=0.5)


@pytest.mark.gpu
def test_mutex_sources(oq):
    cm, sc, srcs = setup(oq)
    grp = oq.SourceGroup('Active Shallow Crust', srcs, 'mutex group',
                         src_interdep='mutex', grp_probability=0.9)
    dic = oq.RmapMaker(cm, sc, grp).make()
    close(dic['rmap'].array, D['mutex_rates'], 2e-7)
    with pytest.raises(NotImplementedError):
        oq.RmapMaker(cm, sc, oq.SourceGroup('*', srcs, rup_interdep='mutex'))


@pytest.mark.gpu
def test_cluster_group(oq):
    cm, sc, srcs = setup(oq)
    tom = oq.PoissonTOM(50.)
    tom.occurrence_rate = float(D['cluster_rate'])
    grp = oq.SourceGroup('Active Shallow Crust', srcs, 'cluster group',
                         cluster=True, temporal_occurrence_model=tom)
    dic = oq.RmapMaker(cm, sc, grp).make()
    close(dic['rmap'].array, D['cluster_rates'], 2e-7)
    assert cm.cluster == 1
    # the unfused update gives the same pne map as the fused pass
    cm2, _, _ = setup(oq)
    cm2.cluster, cm2.tom = 1, tom
    pla = oq.build_planar([6.], [10.], [80.], [0.], [1e-3], [80.], [0.],
                          [0.], [7.], 0., 20., 1.5)
    src = oq.PlanarRuptures(pla, np.array([6.]))
    a = oq.MapArray(sc.sids, cm2.L, cm2.G).fill(1)
    cm2.hazard_planar(a, src, sc)
    b = oq.MapArray(sc.sids, cm2.L, cm2.G).fill(1)
    ctx = cm2.recarray(cm2.get_ctxs(src, sc))
    cm2.update(b, ctx)
    rows = np.unique(b.sidx[ctx.sids])
    np.testing.assert_allclose(a.array[rows], b.array[rows], rtol=2e-6,
                               atol=1e-7)   # float32 PoEs of the unfused path
    assert (a.array[rows] < 1).any()


@pytest.mark.gpu
def test_mutex_ruptures(oq):
    """ContextMaker.get_pmap(ctx, tom, rup_mutex): update_pmap_m
    (map_array.py:262-269) against the reference on the same ruptures"""
    M = np.load(os.path.join(GOLDEN, 'rup_mutex.npz'))
    cm, _, srcs = setup(oq)
    N = len(M['site_lon'])
    sc = oq.SiteCollection.from_points(
        M['site_lon'], M['site_lat'], np.zeros(N), dict(vs30=np.full(N, 500.)))
    h = cm.handle()
    ctxs = []
    for i, src in enumerate(srcs):
        b = oq.PointSourceBatch([src])
        t_pla, t_mag, _ = oq.point_ruptures_device(b, h)
        pla = t_pla.cpu().numpy().reshape(-1, 3, 14).view(
            oq.planar_array_dt).reshape(-1, 3)
        ctxs.extend(cm.get_ctxs(oq.PlanarRuptures(pla, t_mag.cpu().numpy()),
                                sc, src_id=i))
    ctx = cm.recarray(ctxs)
    assert len(ctx) == int(M['nrows'])
    rup_mutex = {(int(s), int(r)): w
                 for (s, r), w in zip(M['keys'], M['weights'])}
    assert set(zip(ctx.src_id.tolist(), ctx.rup_id.tolist())) == set(rup_mutex)
    pmap = cm.get_pmap(ctx, oq.PoissonTOM(50.), rup_mutex)
    assert (pmap.sids == M['sids']).all()
    np.testing.assert_allclose(pmap.array, M['poes'], rtol=2e-5, atol=1e-9)
    # independent ruptures through the same call: the usual PoE map
    indep = cm.get_pmap(ctx, oq.PoissonTOM(50.))
    assert (indep.array <= 1).all() and (indep.array != pmap.array).any()


@pytest.mark.gpu
def test_imt_weight_zero(oq):
    """set_poes, contexts.py:1469-1475: an ImtWeight of zero zeroes the PoEs
    of that IMT for that gsim -- fused pass and unfused operators against
    the reference's get_pmap"""
    import types
    W = np.load(os.path.join(GOLDEN, 'imt_weight.npz'))
    cm, _, srcs = setup(oq)
    cm.gsims[1].weight = types.SimpleNamespace(dic={'SA(0.2)': 0, 'PGA': 0.5})
    N = len(W['site_lon'])
    sc = oq.SiteCollection.from_points(
        W['site_lon'], W['site_lat'], np.zeros(N), dict(vs30=np.full(N, 500.)))
    L1 = len(D['levels'])
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    kept = cm.hazard_point_sources(pmap, srcs, sc, count=True)
    assert kept == int(W['nrows'])
    poes = (~pmap).array
    assert (poes[:, L1:2 * L1, 1] == 0).all()
    np.testing.assert_allclose(poes, W['poes'], rtol=2e-5, atol=2e-7)
    # unfused: get_ctxs -> get_pmap
    h = cm.handle()
    ctxs = []
    for i, src in enumerate(srcs):
        t_pla, t_mag, _ = oq.point_ruptures_device(oq.PointSourceBatch([src]), h)
        pla = t_pla.cpu().numpy().reshape(-1, 3, 14).view(
            oq.planar_array_dt).reshape(-1, 3)
        ctxs.extend(cm.get_ctxs(oq.PlanarRuptures(pla, t_mag.cpu().numpy()),
                                sc, src_id=i))
    unf = cm.get_pmap(cm.recarray(ctxs), oq.PoissonTOM(50.))
    assert (unf.array[:, L1:2 * L1, 1] == 0).all()
    np.testing.assert_allclose(unf.array, W['poes'], rtol=2e-5, atol=2e-7)
    # a new ContextMaker without weights on the same handle resets the mask
    cm2, _, _ = setup(oq)
    pm2 = oq.MapArray(sc.sids, cm2.L, cm2.G).fill(1)
    cm2.hazard_point_sources(pm2, srcs, sc)
    assert ((~pm2).array[:, L1:2 * L1, 1] > 0).any()


@pytest.mark.gpu
def test_calc_hazard_curves(oq):
    """hazardlib's calc_hazard_curves (calc/hazard_curve.py:76-164): two
    tectonic region types with their own GSIM, composed as
    1 - (1 - pmap) exp(-rmap), against the reference"""
    H = np.load(os.path.join(GOLDEN, 'hazard_curves.npz'))
    _, _, srcs = setup(oq)
    for src in srcs[:2]:
        src.tectonic_region_type = 'Active Shallow Crust'
    srcs[2].tectonic_region_type = 'Stable Shallow Crust'
    N = len(H['site_lon'])
    sc = oq.SiteCollection.from_points(
        H['site_lon'], H['site_lat'], np.zeros(N), dict(vs30=np.full(N, 600.)))
    imtls = {str(imt): H['levels'] for imt in H['imts']}
    gsim_by_trt = {'Active Shallow Crust': oq.get_gsim('BooreEtAl2014'),
                   'Stable Shallow Crust': oq.get_gsim('AtkinsonBoore2006')}
    curves = oq.calc_hazard_curves(srcs, sc, imtls, gsim_by_trt,
                                   truncation_level=3.,
                                   investigation_time=50.)
    assert curves.dtype.names == tuple(imtls) and len(curves) == N
    for imt in imtls:
        np.testing.assert_allclose(curves[imt], H['curve_' + imt], rtol=2e-5,
                                   atol=2e-7, err_msg=imt)
    with pytest.raises(KeyError):
        oq.calc_hazard_curves(srcs, sc, imtls, {'Active Shallow Crust':
                                                gsim_by_trt['Active Shallow Crust']},
                              truncation_level=3., investigation_time=50.)


@pytest.mark.gpu
def test_notebook_operators(oq):
    """from_srcs / get_poes / get_rmap / get_occ_rates / get_4MN
    (contexts.py:866-881, 1132-1153, 1220-1230, 1281-1303) against the
    reference on the three point sources (oracle/make_golden_rmap.py)"""
    A = np.load(os.path.join(GOLDEN, 'cm_api.npz'))
    cm, _, srcs = setup(oq)
    N = len(A['site_lon'])
    sc = oq.SiteCollection.from_points(
        A['site_lon'], A['site_lat'], np.zeros(N), dict(vs30=np.full(N, 500.)))
    ctx = cm.from_srcs(srcs, sc)
    assert len(ctx) == int(A['nrows'])

    def order(src_id, sids, mag, rrup):
        return np.lexsort((rrup, np.round(mag, 3), sids, src_id))
    mine = order(ctx.src_id, ctx.sids, ctx.mag, ctx.rrup)
    ref = order(A['ctx_src_id'], A['ctx_sids'], A['ctx_mag'], A['ctx_rrup'])
    assert (ctx.sids[mine] == A['ctx_sids'][ref]).all()
    assert (ctx.src_id[mine] == A['ctx_src_id'][ref]).all()
    np.testing.assert_allclose(ctx.rrup[mine], A['ctx_rrup'][ref], rtol=1e-9)
    np.testing.assert_allclose(cm.get_occ_rates(ctx)[mine],
                               A['occ_rates'][ref], rtol=1e-12)
    poes = cm.get_poes(srcs, sc)
    # the reference folds pne in float32, chunk after chunk: one ulp of 1
    close(poes, A['poes'], 3e-7)
    rmap = cm.get_rmap(ctx)
    assert (rmap.sids == A['rmap_sids']).all() and (rmap.gid == [0, 1]).all()
    np.testing.assert_allclose(rmap.array, A['rates'], rtol=3e-5, atol=3e-7)
    out4 = cm.get_4MN([ctx], cm.gsims[1])
    assert out4.shape == A['out4'].shape
    np.testing.assert_allclose(out4[:, :, mine], A['out4'][:, :, ref],
                               rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        cm.get_4MN([ctx], oq.get_gsim('ChiouYoungs2014'))
    # max_intensity with the four NGA-West2 models
    cm4 = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim(n) for n in ('AbrahamsonEtAl2014', 'ChiouYoungs2014',
                                  'CampbellBozorgnia2014', 'BooreEtAl2014')],
        dict(imtls=cm.imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('100')))
    sc1 = oq.SiteCollection.from_points(
        A['site_lon'][:1], A['site_lat'][:1], np.zeros(1), dict(
            vs30=np.full(1, 500.), vs30measured=np.ones(1, bool),
            z1pt0=np.full(1, 100.), z2pt5=np.full(1, 5.)))
    got = cm4.max_intensity(sc1, A['mi_mags'], A['mi_dists'])
    np.testing.assert_allclose(got, A['maxint'], rtol=1e-9)
