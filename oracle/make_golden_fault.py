"""
Golden fixture for the ruptures of fault sources (SURVEY.md section 8f row 2:
genctxs for non-planar ruptures, contexts.py:959-1029, 1071-1099), generated
by running the REFERENCE in this container:

    python oracle/make_golden_fault.py   ->  tests/golden/fault_ctxs.npz

The source is the simple fault of demos/hazard/LogicTreeCase2ClassicalPSHA
(trace (1.0, -0.5) (1.4, 0.0) (1.4, 0.3), dip 30, 8-20 km, WC1994, aspect
ratio 2, TruncatedGRMFD(6.5, 7.5, 0.1, 3.2, 0.9), rake 90, mesh spacing 4 km).
Stored: the whole fault mesh (SimpleFaultSurface.from_fault_data), every
rupture the reference floats over it (window, magnitude, rate, hypocentre),
and for two sets of models -- the demo's BooreAtkinson2008 + ChiouYoungs2008,
and the four NGA-West2 models (which read every rupture parameter and
distance: dip, ztor, width, hypo_depth, rrup, rjb, rx, ry0) -- the context
rows of ContextMaker.get_ctxs and the float32 PoE map of get_pmap.

All the sites are farther than 40 km from the fault: below that distance the
reference's Joyner-Boore distance needs shapely (geo/mesh.py:507-538), which
this image does not have, so the near field of rjb is pinned separately on the
reference's own test values (oracle/make_golden_mesh_rjb.py).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.contexts import ContextMaker  # noqa
from openquake.hazardlib import valid  # noqa
from openquake.hazardlib.geo import Point, Line  # noqa
from openquake.hazardlib.geo.surface.simple_fault import (  # noqa
    SimpleFaultSurface)
from openquake.hazardlib.mfd import TruncatedGRMFD  # noqa
from openquake.hazardlib.scalerel.wc1994 import WC1994  # noqa
from openquake.hazardlib.source import SimpleFaultSource  # noqa
from openquake.hazardlib.tom import PoissonTOM  # noqa
from openquake.hazardlib.site import SiteCollection  # noqa
from openquake.hazardlib.calc.filters import IntegrationDistance  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
TRACE = [(1.0, -0.5), (1.4, 0.0), (1.4, 0.3)]
SPACING = 4.0
LEVELS = [0.005, 0.007, 0.0098, 0.0137, 0.0192, 0.0269, 0.0376, 0.0527, 0.0738,
          0.103, 0.145, 0.203, 0.284, 0.397, 0.556, 0.778, 1.09, 1.52, 2.13]
MODELS = {
    'demo': ['BooreAtkinson2008', 'ChiouYoungs2008'],
    'nga': ['BooreEtAl2014', 'ChiouYoungs2014', 'AbrahamsonEtAl2014',
            'CampbellBozorgnia2014'],
}
CTX_FIELDS = ('mag rake dip ztor width hypo_depth rrup rjb rx ry0 vs30 '
              'vs30measured z1pt0 z2pt5 sids rup_id occurrence_rate').split()


def fault_source(mfd=None):
    src = SimpleFaultSource(
        'second', 'Simple Fault Source', 'Active Shallow Crust',
        mfd or TruncatedGRMFD(6.5, 7.5, 0.1, 3.2, 0.9), SPACING, WC1994(),
        2.0, PoissonTOM(50.), 8., 20., Line([Point(*p) for p in TRACE]), 30.,
        90.)
    src.id, src.offset = 0, 0
    return src


def sites():
    rng = np.random.default_rng(11)
    # the demo's site, then sites on both sides of the fault, 45-260 km away
    lon = np.concatenate([[0.5], rng.uniform(-0.6, 0.55, 14),
                          rng.uniform(2.0, 3.3, 9)])
    lat = np.concatenate([[-0.5], rng.uniform(-1.5, 1.5, 23)])
    N = len(lon)

    class SM:
        pass
    sm = SM()
    sm.vs30 = np.concatenate([[600.], rng.uniform(200, 1200, N - 1)])
    sm.vs30measured = np.concatenate([[True], rng.random(N - 1) < .5])
    sm.z1pt0 = np.concatenate([[100.], rng.uniform(5, 600, N - 1)])
    sm.z2pt5 = np.concatenate([[5.], rng.uniform(0.3, 6, N - 1)])
    sc = SiteCollection.from_points(
        lon, lat, np.zeros(N), sm,
        req_site_params=('vs30', 'vs30measured', 'z1pt0', 'z2pt5'))
    return sc


if __name__ == '__main__':
    src = fault_source()
    surf = SimpleFaultSurface.from_fault_data(
        src.fault_trace, 8., 20., 30., SPACING)
    mesh = surf.mesh
    rups = list(src.iter_ruptures())
    # the window of every rupture inside the fault mesh
    wins = []
    for rup in rups:
        m = rup.surface.mesh
        r0c0 = np.argwhere((mesh.lons == m.lons[0, 0]) &
                           (mesh.lats == m.lats[0, 0]) &
                           (mesh.depths == m.depths[0, 0]))
        assert len(r0c0) == 1
        r0, c0 = r0c0[0]
        nr, nc = m.lons.shape
        np.testing.assert_array_equal(mesh.lons[r0:r0 + nr, c0:c0 + nc],
                                      m.lons)
        wins.append((r0, c0, nr, nc))
    sc = sites()
    # every site beyond the 40 km threshold of the Joyner-Boore distance
    dmin = surf.get_min_distance(sc.mesh).min()
    assert dmin > 41., dmin
    dic = dict(
        mesh_lons=mesh.lons, mesh_lats=mesh.lats, mesh_deps=mesh.depths,
        spacing=SPACING, rar=2.0, rake=90., mfd=[6.5, 7.5, 0.1, 3.2, 0.9],
        windows=np.array(wins), mag=np.array([r.mag for r in rups]),
        rate=np.array([r.occurrence_rate for r in rups]),
        hypo=np.array([[r.hypocenter.x, r.hypocenter.y, r.hypocenter.z]
                       for r in rups]),
        dip=np.array([r.surface.get_dip() for r in rups]),
        strike=np.array([r.surface.get_strike() for r in rups]),
        width=np.array([r.surface.get_width() for r in rups]),
        ztor=np.array([r.surface.get_top_edge_depth() for r in rups]),
        levels=np.array(LEVELS), maxdist=200.,
        **{'site_' + n: np.array(sc.array[n]) for n in (
            'lon', 'lat', 'depth', 'vs30', 'vs30measured', 'z1pt0', 'z2pt5')})
    for tag, names in MODELS.items():
        gsims = [valid.gsim('[%s]' % n) for n in names]
        imts = ['PGA'] if tag == 'demo' else ['PGA', 'SA(0.2)', 'SA(1.0)']
        cm = ContextMaker('Active Shallow Crust', gsims, dict(
            imtls={i: LEVELS for i in imts}, truncation_level=3.,
            investigation_time=50.,
            maximum_distance=IntegrationDistance.new('200')))
        ctxs = list(cm.get_ctxs(src, sc))
        ctx = cm.recarray(ctxs)
        pmap = cm.get_pmap(ctx, PoissonTOM(50.))
        ms = cm.get_mean_stds([ctx])
        dic[tag + '_gsims'] = np.array(names)
        dic[tag + '_imts'] = np.array(imts)
        dic[tag + '_nctxs'] = len(ctxs)
        dic[tag + '_ctx_lens'] = np.array([len(c) for c in ctxs])
        for f in CTX_FIELDS:
            if f in ctx.dtype.names:
                dic['%s_ctx_%s' % (tag, f)] = np.array(ctx[f])
        dic[tag + '_poes'] = pmap.array
        dic[tag + '_mean_stds'] = ms
        print(tag, len(rups), 'ruptures', len(ctxs), 'contexts', len(ctx),
              'rows; max poe', pmap.array.max())
    np.savez_compressed(os.path.join(GOLDEN, 'fault_ctxs.npz'), **dic)
    print('fault mesh', mesh.lons.shape, 'closest site', dmin, 'km')
