"""Disaggregation hot function: the oracle
(oracle/disagg_np.py) and, on the GPU, oqg_disaggregate against the
reference's own _disaggregate (oracle/make_golden_disagg.py): both epsilon
modes, a group probability, lon bins across the date line, rows on and
beyond the bin edges, a zero-hazard level.  Tolerance rtol 1e-10, atol 1e-14
(float64 throughout; the product of the reference is a sum of logarithms on
the device, in a non-deterministic order)."""
import os
import types
import numpy as np
import pytest

import disagg_np
from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'disagg.npz'))


def edges(lon='lon_edges'):
    return [np.array([5., 6.]), D['dist_edges'], D[lon], D['lat_edges'],
            D['eps_edges']]


@pytest.mark.parametrize('epsstar,g,gp', [(0, 0, 1.), (0, 1, .8), (1, 0, 1.),
                                          (1, 1, .8)])
def test_oracle_golden(epsstar, g, gp):
    mine = disagg_np.disaggregate(
        D['rrup'], D['clon'], D['clat'], D['rate'], D['mea'][g], D['std'][g],
        D['iml2'], D['dist_edges'], D['lon_edges'], D['lat_edges'],
        D['eps_edges'], float(D['tl']), float(D['itime']), bool(epsstar), gp)
    np.testing.assert_allclose(mine, D['mat_%d_%d' % (epsstar, g)],
                               rtol=1e-12, atol=1e-15)


def test_oracle_eps4():
    _, _, bands, cum = disagg_np.eps4(D['eps_edges'], float(D['tl']))
    np.testing.assert_allclose(bands, D['eps_bands'], rtol=1e-13)
    np.testing.assert_allclose(cum, D['eps_cum'], rtol=1e-13)
    assert abs(bands.sum() - 1) < 1e-12


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def ctx_of(clon='clon'):
    return types.SimpleNamespace(
        rrup=D['rrup'], clon=D[clon], clat=D['clat'],
        occurrence_rate=D['rate'], probs_occur=np.zeros((len(D['rrup']), 0)),
        __len__=None)


class Ctx(object):
    def __init__(self, clon='clon', n=None):
        sl = slice(0, n)
        self.rrup, self.clon, self.clat = (D['rrup'][sl], D[clon][sl],
                                           D['clat'][sl])
        self.occurrence_rate = D['rate'][sl]
        self.probs_occur = np.zeros((len(self.rrup), 0))

    def __len__(self):
        return len(self.rrup)


CM = types.SimpleNamespace(truncation_level=3., investigation_time=50.,
                           device=None)


@pytest.mark.gpu
@pytest.mark.parametrize('epsstar,g,gp', [(0, 0, 1.), (0, 1, .8), (1, 0, 1.),
                                          (1, 1, .8)])
def test_disaggregate_golden(oq, epsstar, g, gp):
    got = oq.disaggregate(Ctx(), D['mea'], D['std'], CM, g, D['iml2'],
                          edges(), bool(epsstar), gp)
    want = D['mat_%d_%d' % (epsstar, g)]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-14)
    assert (got[..., 2, 1] == 0).all()        # the zero-hazard level


@pytest.mark.gpu
def test_disaggregate_no_date_line_and_edge_cases(oq):
    got = oq.disaggregate(Ctx('clon2'), D['mea'], D['std'], CM, 0, D['iml2'],
                          edges('lon_edges2'))
    np.testing.assert_allclose(got, D['mat_nodl'], rtol=1e-10, atol=1e-14)
    # no rows: a matrix of zeros
    empty = oq.disaggregate(Ctx(n=0), D['mea'][:, :, :0], D['std'][:, :, :0],
                            CM, 0, D['iml2'], edges())
    assert empty.shape == D['mat_0_0'].shape and (empty == 0).all()
    # a subset of the rows against the oracle
    n = 37
    sub = oq.disaggregate(Ctx(n=n), D['mea'][:, :, :n], D['std'][:, :, :n],
                          CM, 1, D['iml2'], edges(), False, 0.5)
    mine = disagg_np.disaggregate(
        D['rrup'][:n], D['clon'][:n], D['clat'][:n], D['rate'][:n],
        D['mea'][1][:, :n], D['std'][1][:, :n], D['iml2'], D['dist_edges'],
        D['lon_edges'], D['lat_edges'], D['eps_edges'], 3., 50., False, 0.5)
    np.testing.assert_allclose(sub, mine, rtol=1e-10, atol=1e-14)
    # non-Poissonian rows: no device path
    bad = Ctx()
    bad.probs_occur = np.ones((len(bad), 3))
    with pytest.raises(NotImplementedError):
        oq.disaggregate(bad, D['mea'], D['std'], CM, 0, D['iml2'], edges())


# ----------------------------------------------------------- Disaggregator
DG = np.load(os.path.join(GOLDEN, 'disaggregator.npz'))
SRC = [(0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
       (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
       (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90., 10.)]


def test_host_split_by_magbin_and_to_rates():
    import oq_engine_b200.disagg as dg
    ctx = np.zeros(5, [('mag', float)]).view(np.recarray)
    ctx.mag = [5.0, 6.4, 5.5, 5.49, 6.0]
    got = dg.split_by_magbin(ctx, np.array([5., 5.5, 6., 6.5]))
    assert {k: list(v.mag) for k, v in got.items()} == {
        0: [5.0, 5.49, 5.5], 1: [6.0], 2: [6.4]}
    np.testing.assert_allclose(dg.to_rates(np.array([0., 0.5, 1.])),
                               [0., np.log(2.), 100.])


@pytest.mark.gpu
def test_disaggregator_golden(oq):
    """the driver on the rows of three point sources at one site against the
    reference's Disaggregator (disagg_mag_dist_eps and the 6D matrices)"""
    srcs = [oq.PointSource(
        str(i), lon, lat, oq.EvenlyDiscretizedMFD(5.05, 0.3, rates), 'WC1994',
        1.5, 0., 20., [(1., strike, dip, rake)], [(1., dep)])
        for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC)]
    sc = oq.SiteCollection.from_points([0.05], [0.0], [0.],
                                       dict(vs30=500.))
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls={'PGA': [0.05, 0.2], 'SA(0.2)': [0.05, 0.2]},
             truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('100')))
    ctxs = []
    for i, src in enumerate(srcs):
        ctxs.extend(cm.get_ctxs(src, sc, src_id=i))
    ctx = cm.recarray(ctxs)
    assert len(ctx) == int(DG['nrows'])
    np.testing.assert_allclose(np.sort(ctx.rrup), np.sort(DG['ctx_rrup']),
                               rtol=1e-9)
    np.testing.assert_allclose(np.sort(ctx.clon), np.sort(DG['ctx_clon']),
                               atol=1e-9)
    edges = [DG['mag_edges'], DG['dist_edges'], {0: DG['lon_edges']},
             {0: DG['lat_edges']}, DG['eps_edges']]
    dis = oq.Disaggregator(ctx, sc, cm, edges, rlzs_by_g=[[0, 2], [1]])
    assert (dis.Ma, dis.D, dis.Lo, dis.La, dis.E) == (3, 5, 4, 3, 4)
    mde = dis.disagg_mag_dist_eps({'PGA': 0.1, 'SA(0.2)': 0.25},
                                  DG['rlz_weights'])
    assert mde.shape == DG['mde'].shape
    np.testing.assert_allclose(mde, DG['mde'], rtol=1e-8, atol=1e-13)
    imtls2 = {'PGA': [0.1, 0.3], 'SA(0.2)': [0.25, 0.6]}
    for magi in range(dis.Ma):
        dis.init(magi)
        for g in range(2):
            np.testing.assert_allclose(
                dis._disagg6D(imtls2, g), DG['mat6_%d_%d' % (magi, g)],
                rtol=1e-8, atol=1e-13)
    # disagg_by_magi: the same matrices per realization and their mean
    rw = {0: .5, 1: .3, 2: .2}
    for res in dis.disagg_by_magi(imtls2, [0, 1, 2], rw):
        magi = res['magi']
        m0, m1 = (DG['mat6_%d_%d' % (magi, g)] for g in range(2))
        np.testing.assert_allclose(res[1], m1, rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(res['mean'], .7 * m0 + .3 * m1, rtol=1e-8,
                                   atol=1e-13)
    # a group of mutually exclusive sources (disagg.py:421-441, 463-475)
    src_mutex = {'grp_probability': float(DG['mutex_gp']),
                 'src_id': np.array([0, 1, 2]), 'weight': DG['mutex_weight']}
    dis2 = oq.Disaggregator(ctx, sc, cm, edges, rlzs_by_g=[[0, 2], [1]])
    mde = dis2.disagg_mag_dist_eps({'PGA': 0.1, 'SA(0.2)': 0.25},
                                   DG['rlz_weights'], src_mutex)
    np.testing.assert_allclose(mde, DG['mde_mutex'], rtol=1e-8, atol=1e-13)
    for magi in range(dis2.Ma):
        mw = dis2.init(magi, src_mutex)
        np.testing.assert_allclose(mw, DG['mutex_mw_%d' % magi], rtol=1e-15)
        for g in range(2):
            np.testing.assert_allclose(
                dis2._disagg6D(imtls2, g), DG['mat6m_%d_%d' % (magi, g)],
                rtol=1e-8, atol=1e-13)
    # back to independent sources on the same object: mea / std follow the
    # row order
    dis2.init(0)
    np.testing.assert_allclose(dis2._disagg6D(imtls2, 0), DG['mat6_0_0'],
                               rtol=1e-8, atol=1e-13)
    with pytest.raises(ValueError):
        dis2.init(0, {'grp_probability': 1., 'src_id': [0], 'weight': [1.]})
    with pytest.raises(oq.FarAwayRupture):
        oq.Disaggregator(ctx[:0], sc, cm, edges, rlzs_by_g=[[0], [1]])


# ------------------------------------------------- disaggregation (top level)
def test_host_bin_edges():
    """uniform_bins / lon_lat_bins / the magnitude edges of _build_bin_edges
    (disagg.py:63-142) against the reference's edges of the fixture"""
    from oq_engine_b200 import disagg as dg
    T = np.load(os.path.join(GOLDEN, 'disaggregation.npz'))
    np.testing.assert_allclose(dg.uniform_bins(1, 10, 1.1),
                               [0., 1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8,
                                9.9, 11.])
    maxdist = T['dist'][-1] - 1e-3      # any value in the last 15 km bin
    edges = dg.build_bin_edges([5.05, 5.35, 6.25], maxdist, 0.05, 0.0, 3., 6,
                               0.5, 15., 0.2)
    np.testing.assert_allclose(edges[0], T['mag'], rtol=1e-15)
    np.testing.assert_allclose(edges[1], T['dist'], rtol=1e-15)
    np.testing.assert_allclose(edges[4], T['eps'], rtol=1e-15)
    lon, lat = dg.lon_lat_bins(179.9, 10., 100., 0.4)
    assert (lon >= 0).all() and lon.max() > 180      # across the date line


@pytest.mark.gpu
@pytest.mark.parametrize('epsstar', [False, True])
def test_disaggregation_function_golden(oq, epsstar):
    """calc.disagg.disaggregation (disagg.py:544-669) on three point sources
    of two tectonic region types against the reference's edges and matrix"""
    T = np.load(os.path.join(GOLDEN, 'disaggregation.npz'))
    tag = '_star' if epsstar else ''
    TRT = ['Active Shallow Crust', 'Stable Shallow Crust',
           'Active Shallow Crust']
    srcs = [oq.PointSource(
        str(i), lon, lat, oq.EvenlyDiscretizedMFD(5.05, 0.3, rates), 'WC1994',
        1.5, 0., 20., [(1., strike, dip, rake)], [(1., dep)],
        tectonic_region_type=TRT[i])
        for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC)]
    sc = oq.SiteCollection.from_points([0.05], [0.0], [0.], dict(vs30=500.))
    gsim_by_trt = {'Active Shallow Crust': oq.get_gsim('BooreEtAl2014'),
                   'Stable Shallow Crust': oq.get_gsim('BooreAtkinson2008')}
    edges, matrix = oq.disaggregation(
        srcs, sc, 'SA(0.2)', 0.15, gsim_by_trt, 3., n_epsilons=6,
        mag_bin_width=0.5, dist_bin_width=15., coord_bin_width=0.2,
        epsstar=epsstar, investigation_time=50.)
    for got, name in zip(edges, ['mag', 'dist', 'lon', 'lat', 'eps']):
        np.testing.assert_allclose(got, T[name + tag], rtol=1e-9, atol=1e-12,
                                   err_msg=name)
    assert edges[5] == ['Active Shallow Crust', 'Stable Shallow Crust']
    assert matrix.shape == T['matrix' + tag].shape
    np.testing.assert_allclose(matrix, T['matrix' + tag], rtol=1e-8,
                               atol=1e-13)
    assert (matrix[..., 1] > 0).any()
    with pytest.raises(ValueError):
        oq.disaggregation(srcs, sc, 'SA(0.2)', 0.15, gsim_by_trt, 3., 6, 0.5,
                          15., 0.2)
