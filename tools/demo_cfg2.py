"""BASELINE.json configs[1], demos/hazard/LogicTreeCase2ClassicalPSHA, end to
end on the device path (hazardlib level): a simple fault source and an area
source, 9 magnitude-frequency variants each (absolute uncertainties on the
Gutenberg-Richter a / b values and on the maximum magnitude), two GMPEs per
tectonic region type -> 324 realizations; mean and quantile hazard curves and
hazard maps.

    python tools/demo_cfg2.py            (on a machine with a GPU)

`run(oq, fixture)` is what tests/test_cfg2.py calls; the fixture
(tests/golden/cfg2_logic_tree.npz, oracle/make_golden_cfg2.py) carries the
fault mesh, since building it from the trace is source-model geometry outside
this package, and the reference's results."""
import os
import sys
import time
import itertools
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

AB = {'area': [(4.6, 1.1), (4.5, 1.0), (4.4, 0.9)],
      'fault': [(3.3, 1.0), (3.2, 0.9), (3.1, 0.8)]}
MMAX = {'area': [7.0, 7.3, 7.6], 'fault': [7.5, 7.8, 8.0]}
W3 = [0.333, 0.333, 0.334]
GSIMS = {'fault': ['BooreAtkinson2008', 'ChiouYoungs2008'],
         'area': ['ToroEtAl2002', 'Campbell2003']}
TRT = {'fault': 'Active Shallow Crust', 'area': 'Stable Continental Crust'}


def realizations():
    """full enumeration of the two logic trees of the demo: (area variant,
    fault variant, area gsim, fault gsim, weight) per realization"""
    out = []
    for (ia, jf, ka, lf), (ga, gs) in itertools.product(
            itertools.product(range(3), repeat=4),
            itertools.product(range(2), repeat=2)):
        w = W3[ia] * W3[jf] * W3[ka] * W3[lf] * 0.5 * 0.5
        out.append((ia * 3 + ka, jf * 3 + lf, gs, ga, w))
    return out


def run(oq, d):
    """:returns: (rates (1, L, 36), columns, trt_rlzs, weights, result of
    postclassical)"""
    levels = d['levels']
    imtls = {'PGA': levels}
    sc = oq.SiteCollection.from_points(
        [0.5], [-0.5], [0.], dict(vs30=[600.], vs30measured=[True],
                                  z1pt0=[100.], z2pt5=[5.]))
    mesh = (d['mesh_lons'], d['mesh_lats'], d['mesh_deps'])
    cols, rates, kept = [], [], {}
    for kind in ('fault', 'area'):
        cm = oq.ContextMaker(TRT[kind], [oq.get_gsim(n) for n in GSIMS[kind]],
                             dict(imtls=imtls, truncation_level=3.,
                                  investigation_time=50.,
                                  maximum_distance=oq.IntegrationDistance.new(
                                      '200'),
                                  pointsource_distance=50.))
        for v, ((a, b), mmax) in enumerate(itertools.product(AB[kind],
                                                             MMAX[kind])):
            if kind == 'fault':
                srcs = [oq.SimpleFaultSource(
                    'second', mesh, oq.TruncatedGRMFD(6.5, mmax, 0.1, a, b),
                    4.0, 'WC1994', 2.0, 90., TRT[kind])]
            else:
                srcs = list(oq.AreaSource(
                    'first', [-0.5, -0.3, 0.1, 0.3], [-0.5, -0.1, 0.2, -0.8],
                    10., oq.TruncatedGRMFD(5.0, mmax, 0.1, a, b), 'WC1994',
                    1.0, 0., 15., [(1., 0., 90., 0.)],
                    [(.2, 5.), (.6, 10.), (.2, 15.)]))
            res = oq.RmapMaker(cm, sc, srcs).make()
            kept[kind, v] = int(round(sum(res['source_data']['nctxs'])))
            for g in range(2):
                cols.append((kind, v, g))
                rates.append(np.asarray(res['rmap'].array)[0, :, g])
    rates = np.array(rates).T[None].astype(np.float32)
    rlzs = realizations()
    weights = np.array([r[4] for r in rlzs])
    trt_rlzs = [[] for _ in cols]
    for r, (va, vf, ga, gf, _w) in enumerate(rlzs):
        trt_rlzs[cols.index(('area', va, ga))].append(r)
        trt_rlzs[cols.index(('fault', vf, gf))].append(r)
    post = oq.postclassical(
        rates, trt_rlzs, weights, imtls,
        hstats=('mean', 'quantile-0.05', 'quantile-0.5', 'quantile-0.95'),
        individual_rlzs=True, poes=(0.1, 0.02))
    return rates, cols, trt_rlzs, weights, post, kept


if __name__ == '__main__':
    import oq_engine_b200 as oq
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'cfg2_logic_tree.npz'))
    run(oq, d)      # warm up (library load, constant tables)
    t0 = time.time()
    rates, cols, trt_rlzs, weights, post, kept = run(oq, d)
    dt = time.time() - t0
    mean = post['hcurves-stats'][0].array[0, 0]
    print('cfg2: %d gsim columns, %d realizations, %.1f ms' % (
        len(cols), len(weights), dt * 1e3))
    print('mean curve     ', np.round(mean, 5))
    print('reference      ', np.round(d['stat_curves'][0], 5))
    print('hazard maps    ', [m.array[0, 0] for m in post['hmaps-stats']])
    print('reference      ', d['hmaps'])
