"""
Golden fixture for the disaggregation kernel:
the reference's `_disaggregate` ("the crucial bit for performance",
openquake/hazardlib/calc/disagg.py:188-243) with `_disagg_eps` (:246-256),
`get_eps4` (:176-186) and `_build_disagg_matrix` (:262-296), run in this
container on synthetic rows of one site and one magnitude bin; the numpy
oracle (oracle/disagg_np.py) is pinned against it in the same run.

    python oracle/make_golden_disagg.py   -> tests/golden/disagg.npz
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.baselib.performance import Monitor  # noqa
from openquake.hazardlib.calc import disagg  # noqa
from openquake.hazardlib.stats import ndtr  # noqa
import disagg_np  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


def main():
    rng = np.random.default_rng(21)
    U, G, M, P = 500, 2, 3, 2
    tl = 3.
    dt = [('rrup', float), ('clon', float), ('clat', float),
          ('occurrence_rate', float), ('probs_occur', float, (0,))]
    ctx = np.zeros(U, dt).view(np.recarray)
    ctx.rrup = rng.uniform(0, 210, U)       # some beyond the last edge
    ctx.rrup[:3] = [0., 200., 40.]          # on the edges
    ctx.clon = 179.6 + rng.uniform(-1.5, 1.5, U)   # across the date line
    ctx.clon = np.where(ctx.clon > 180, ctx.clon - 360, ctx.clon)
    ctx.clat = 10. + rng.uniform(-1.2, 1.2, U)
    ctx.occurrence_rate = 10 ** rng.uniform(-5, -2, U)
    mea = rng.uniform(-5, -1, (G, M, U))
    std = rng.uniform(0.4, 0.9, (G, M, U))
    iml2 = np.log(np.array([[0.05, 0.3], [0.1, 0.5], [0.02, 0.2]]))
    iml2[2, 1] = -np.inf                     # zero hazard: skipped
    dist_edges = np.arange(0, 201, 20.)
    lon_edges = np.array([178.0, 178.8, 179.6, -179.6, -178.8, -178.0])
    lat_edges = np.arange(8.5, 11.6, 0.5)
    eps_edges = np.linspace(-tl, tl, 7)
    bin_edges = [np.array([5., 6.]), dist_edges, lon_edges, lat_edges,
                 eps_edges]
    cm = types.SimpleNamespace(phi_b=ndtr(tl), investigation_time=50.)
    eps4 = disagg.get_eps4(tuple(eps_edges), tl)
    mon = Monitor('x')
    out = dict(rrup=ctx.rrup, clon=ctx.clon, clat=ctx.clat,
               rate=ctx.occurrence_rate, mea=mea, std=std, iml2=iml2,
               dist_edges=dist_edges, lon_edges=lon_edges,
               lat_edges=lat_edges, eps_edges=eps_edges, tl=tl, itime=50.,
               eps_bands=eps4[2], eps_cum=eps4[3])
    for epsstar in (False, True):
        for g, gp in ((0, 1.), (1, 0.8)):
            mat = disagg._disaggregate(ctx, mea, std, cm, g, iml2, bin_edges,
                                       eps4, epsstar, gp, False, mon, mon, mon)
            mine = disagg_np.disaggregate(
                ctx.rrup, ctx.clon, ctx.clat, ctx.occurrence_rate, mea[g],
                std[g], iml2, dist_edges, lon_edges, lat_edges, eps_edges, tl,
                50., epsstar, gp)
            np.testing.assert_allclose(mine, mat, rtol=1e-12, atol=1e-15)
            out['mat_%d_%d' % (epsstar, g)] = mat
            print('epsstar=%s g=%d' % (epsstar, g), mat.shape, mat.max(),
                  (mat > 0).mean())
    # a lon grid that does not cross the date line (numpy.digitize branch)
    lon2 = np.arange(-181.2, -177.9, 0.8)
    ctx2 = ctx.copy()
    ctx2.clon = np.where(ctx.clon > 0, ctx.clon - 360, ctx.clon)
    be2 = [bin_edges[0], dist_edges, lon2, lat_edges, eps_edges]
    mat = disagg._disaggregate(ctx2, mea, std, cm, 0, iml2, be2, eps4, False,
                               1., False, mon, mon, mon)
    mine = disagg_np.disaggregate(
        ctx2.rrup, ctx2.clon, ctx2.clat, ctx2.occurrence_rate, mea[0], std[0],
        iml2, dist_edges, lon2, lat_edges, eps_edges, tl, 50., False, 1.)
    np.testing.assert_allclose(mine, mat, rtol=1e-12, atol=1e-15)
    out['lon_edges2'], out['clon2'], out['mat_nodl'] = lon2, ctx2.clon, mat
    np.savez_compressed(os.path.join(GOLDEN, 'disagg.npz'), **out)
    print('disagg.npz written; oracle == reference')


if __name__ == '__main__':
    main()


def make_disaggregator():
    """
    The Disaggregator driver (disagg.py:353-538) on the context rows of three
    point sources at one site: disagg_mag_dist_eps -> (Ma, D, E, M) rates and
    the 6D matrices per magnitude bin and gsim -> tests/golden/disaggregator.npz
    """
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, NodalPlane
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import EvenlyDiscretizedMFD
    from openquake.hazardlib.scalerel.wc1994 import WC1994
    from openquake.hazardlib.source import PointSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.baselib.general import DictArray
    tom = PoissonTOM(50.)
    SRC = [(0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
           (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
           (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90.,
            10.)]
    srcs = []
    for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC):
        src = PointSource(
            str(i), 'src%d' % i, 'Active Shallow Crust',
            EvenlyDiscretizedMFD(5.05, 0.3, rates), 2., WC1994(), 1.5, tom,
            0., 20., Point(lon, lat),
            PMF([(1., NodalPlane(strike, dip, rake))]), PMF([(1., dep)]))
        src.id, src.grp_id, src.weight, src.offset = i, 0, 1., 0
        srcs.append(src)
    sitecol = SiteCollection.from_points(
        [0.05], [0.0], sitemodel=types.SimpleNamespace(
            reference_vs30_value=500., reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=100.,
            reference_depth_to_2pt5km_per_sec=5.))
    imts = ['PGA', 'SA(0.2)']
    imtls = DictArray({imt: [0.05, 0.2] for imt in imts})
    gsims = {valid.gsim('[BooreEtAl2014]'): [0, 2],
             valid.gsim('[BooreAtkinson2008]'): [1]}
    oq = types.SimpleNamespace(
        imtls=imtls, truncation_level=3., investigation_time=50.,
        maximum_distance=IntegrationDistance.new('100'), epsilon_star=False,
        infer_occur_rates=False, split_sources=True, min_iml=np.zeros(2),
        af=None, calculation_mode='disaggregation', keep_rupdata=False,
        disagg_by_src=False, mea_tau_phi=False, cross_correl=None,
        max_sites_disagg=10, use_rates=False, with_betw_ratio=None,
        disagg_bin_edges={}, mag_bin_width=0.5, collect_rlzs=False,
        pointsource_distance={'default': 1000}, ps_grid_spacing=0,
        maximum_distance_stations=None)
    cm = ContextMaker('Active Shallow Crust', gsims, dict(
        imtls=imtls, truncation_level=3., investigation_time=50.,
        maximum_distance=IntegrationDistance.new('100')))
    cm.oq = oq
    cm.tom = tom
    ctxs = []
    for src in srcs:
        ctxs.extend(cm.get_ctxs(src, sitecol, src_id=src.id))
    ctx = cm.recarray(ctxs)
    mag_edges = np.array([5., 5.5, 6., 6.5])
    dist_edges = np.arange(0, 101, 20.)
    lon_edges = np.linspace(-0.6, 0.7, 5)
    lat_edges = np.linspace(-0.5, 0.6, 4)
    eps_edges = np.linspace(-3, 3, 5)
    bin_edges = [mag_edges, dist_edges, {0: lon_edges}, {0: lat_edges},
                 eps_edges]
    dis = disagg.Disaggregator([ctx], sitecol, cm, bin_edges)
    imldic = {'PGA': 0.1, 'SA(0.2)': 0.25}
    rlz_weights = np.array([0.5, 0.3, 0.2])
    mde = dis.disagg_mag_dist_eps(imldic, rlz_weights)
    out = dict(mag_edges=mag_edges, dist_edges=dist_edges, lon_edges=lon_edges,
               lat_edges=lat_edges, eps_edges=eps_edges, mde=mde,
               rlz_weights=rlz_weights, iml=np.array([0.1, 0.25]),
               nrows=len(ctx), ctx_mag=ctx.mag, ctx_rrup=ctx.rrup,
               ctx_clon=ctx.clon, ctx_clat=ctx.clat)
    imtls2 = {'PGA': [0.1, 0.3], 'SA(0.2)': [0.25, 0.6]}
    for magi in range(dis.Ma):
        try:
            dis.init(magi, {})
        except disagg.FarAwayRupture:
            continue
        for g in range(2):
            out['mat6_%d_%d' % (magi, g)] = dis._disagg6D(imtls2, g)
    # mutually exclusive sources (disagg.py:421-441, :463-475, :534): a
    # magnitude bin reached by a subset of the sources uses the sum of the
    # weights of that subset (the last bin here: the third source only)
    src_mutex = {'grp_probability': 0.8, 'src_id': np.array([0, 1, 2]),
                 'weight': [0.5, 0.3, 0.2]}
    dis2 = disagg.Disaggregator([ctx], sitecol, cm, bin_edges)
    out['mde_mutex'] = dis2.disagg_mag_dist_eps(imldic, rlz_weights,
                                                dict(src_mutex))
    out['mutex_weight'] = np.array(src_mutex['weight'])
    out['mutex_gp'] = 0.8
    for magi in range(dis2.Ma):
        try:
            out['mutex_mw_%d' % magi] = dis2.init(magi, dict(src_mutex))
        except disagg.FarAwayRupture:
            continue
        for g in range(2):
            out['mat6m_%d_%d' % (magi, g)] = dis2._disagg6D(imtls2, g)
    np.savez_compressed(os.path.join(GOLDEN, 'disaggregator.npz'), **out)
    print('disaggregator.npz', mde.shape, mde.sum(), len(ctx),
          [k for k in out if k.startswith('mat6')])


if __name__ == '__main__':
    make_disaggregator()


def make_disaggregation():
    """
    The top-level hazardlib function calc.disagg.disaggregation
    (disagg.py:544-669) on three point sources of two tectonic region types
    at one site, bin edges built from the bin widths
    -> tests/golden/disaggregation.npz
    """
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, NodalPlane
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import EvenlyDiscretizedMFD
    from openquake.hazardlib.scalerel.wc1994 import WC1994
    from openquake.hazardlib.source import PointSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import Site
    from openquake.hazardlib.imt import from_string
    tom = PoissonTOM(50.)
    SRC = [(0.10, 0.10, [0.02, 0.008, 0.003, 0.001], 30., 60., 90., 6.),
           (0.45, -0.20, [0.05, 0.01, 0.002], 120., 90., 0., 8.),
           (-0.30, 0.35, [0.01, 0.006, 0.004, 0.002, 0.001], 250., 45., -90.,
            10.)]
    TRT = ['Active Shallow Crust', 'Stable Shallow Crust',
           'Active Shallow Crust']
    srcs = []
    for i, (lon, lat, rates, strike, dip, rake, dep) in enumerate(SRC):
        src = PointSource(
            str(i), 'src%d' % i, TRT[i],
            EvenlyDiscretizedMFD(5.05, 0.3, rates), 2., WC1994(), 1.5, tom,
            0., 20., Point(lon, lat),
            PMF([(1., NodalPlane(strike, dip, rake))]), PMF([(1., dep)]))
        src.id, src.grp_id, src.weight, src.offset = -1, 0, 1., 0
        srcs.append(src)
    site = Site(Point(0.05, 0.0), vs30=500., vs30measured=True, z1pt0=100.,
                z2pt5=5.)
    gsim_by_trt = {'Active Shallow Crust': valid.gsim('[BooreEtAl2014]'),
                   'Stable Shallow Crust': valid.gsim('[BooreAtkinson2008]')}
    out = {}
    for tag, epsstar in (('', False), ('_star', True)):
        edges, matrix = disagg.disaggregation(
            srcs, site, from_string('SA(0.2)'), 0.15, gsim_by_trt, 3.,
            n_epsilons=6, mag_bin_width=0.5, dist_bin_width=15.,
            coord_bin_width=0.2, epsstar=epsstar)
        mag, dist, lon, lat, eps, trts = edges
        out.update({'mag' + tag: mag, 'dist' + tag: dist, 'lon' + tag: lon[0],
                    'lat' + tag: lat[0], 'eps' + tag: eps,
                    'matrix' + tag: matrix})
        print('disaggregation%s' % tag, matrix.shape, matrix.max(), trts,
              len(mag), len(dist), len(lon[0]), len(lat[0]))
    np.savez_compressed(os.path.join(GOLDEN, 'disaggregation.npz'), **out)


if __name__ == '__main__':
    make_disaggregation()
