"""
The PEER verification tests the reference runs as acceptance tests of this
path (openquake/hazardlib/tests/acceptance/peer_test.py, data
_peer_test_data.py: "PEER 2010/106 - Verification of Probabilistic Seismic
Hazard Analysis Computer Programs", Thomas, Wong, Abrahamson): set 1 cases
10 and 11 (an area source of 100 km radius, SadighEtAl1997 on rock, no
truncation) and cases 2 and 5 (a vertical strike-slip fault with PeerMSR).

The fixture holds the INPUTS of the reference's tests (polygon, fault trace,
MFDs, sites, levels) and the PUBLISHED probabilities of exceedance with the
tolerances the reference's tests use; for the fault cases, which need no
polygon code, also the reference's own curves (calc_hazard_curves run here).

    python oracle/make_golden_peer.py   -> tests/golden/peer_set1.npz
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib import const  # noqa
from openquake.hazardlib.site import SiteCollection  # noqa
from openquake.hazardlib.source import SimpleFaultSource  # noqa
from openquake.hazardlib.scalerel import PeerMSR  # noqa
from openquake.hazardlib.geo.surface import SimpleFaultSurface  # noqa
from openquake.hazardlib.gsim.sadigh_1997 import SadighEtAl1997  # noqa
from openquake.hazardlib.contexts import ContextMaker  # noqa
from openquake.hazardlib.map_array import MapArray  # noqa
from openquake.hazardlib.tom import PoissonTOM  # noqa
from openquake.hazardlib.tests.acceptance import _peer_test_data as T  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


def sites_of(sites):
    return np.array([[s.location.longitude, s.location.latitude, s.vs30]
                     for s in sites])


def mfd_of(mfd):
    return np.array(mfd.get_annual_occurrence_rates())       # (mag, rate)


def main():
    out = {}
    poly = T.SET1_CASE10_SOURCE_POLYGON
    out['poly_lons'], out['poly_lats'] = poly.lons, poly.lats
    out['rar'] = T.SET1_RUPTURE_ASPECT_RATIO
    out['area_sites'] = sites_of([T.SET1_CASE10_SITE1, T.SET1_CASE10_SITE2,
                                  T.SET1_CASE10_SITE3, T.SET1_CASE10_SITE4])
    m = T.SET1_CASE10_MFD
    out['area_gr'] = np.array([m.min_mag, m.max_mag, m.bin_width, m.a_val,
                               m.b_val])
    out['area_mfd'] = mfd_of(m)
    out['c10_depth'] = T.SET1_CASE10_HYPOCENTER_DEPTH
    out['c10_imls'] = np.array(T.SET1_CASE10_IMLS)
    out['c10_poes'] = np.array([T.SET1_CASE10_SITE1_POES,
                                T.SET1_CASE10_SITE2_POES,
                                T.SET1_CASE10_SITE3_POES,
                                T.SET1_CASE10_SITE4_POES])
    out['c11_depths'] = np.array(T.SET1_CASE11_HYPOCENTERS)
    out['c11_imls'] = np.array(T.SET1_CASE11_IMLS)
    out['c11_poes'] = np.array([T.SET1_CASE11_SITE1_POES,
                                T.SET1_CASE11_SITE2_POES,
                                T.SET1_CASE11_SITE3_POES,
                                T.SET1_CASE11_SITE4_POES])
    # the fault of cases 1 to 9
    tr = T.SET1_CASE1TO9_FAULT_TRACE
    out['trace'] = np.array([[p.longitude, p.latitude] for p in tr.points])
    out['fault'] = np.array([T.SET1_CASE1TO9_UPPER_SEISMOGENIC_DEPTH,
                             T.SET1_CASE1TO9_LOWER_SEISMOGENIC_DEPTH,
                             T.SET1_CASE1TO9_DIP, T.SET1_CASE1TO9_RAKE])
    # the fault mesh as the reference builds it (rupture_mesh_spacing 1 km)
    surf = SimpleFaultSurface.from_fault_data(
        tr, T.SET1_CASE1TO9_UPPER_SEISMOGENIC_DEPTH,
        T.SET1_CASE1TO9_LOWER_SEISMOGENIC_DEPTH, T.SET1_CASE1TO9_DIP, 1.0)
    out['mesh_lons'], out['mesh_lats'] = surf.mesh.lons, surf.mesh.lats
    out['mesh_deps'] = surf.mesh.depths
    fsites = [getattr(T, 'SET1_CASE1TO9_SITE%d' % i) for i in range(1, 8)]
    out['fault_sites'] = sites_of(fsites)
    for case, mfd in (('c2', T.SET1_CASE2_MFD), ('c5', T.SET1_CASE5_MFD)):
        C = case.upper().replace('C', 'CASE')
        imls = getattr(T, 'SET1_%s_IMLS' % C)
        out[case + '_mfd'] = mfd_of(mfd)
        out[case + '_imls'] = np.array(imls)
        out[case + '_poes'] = np.array([
            getattr(T, 'SET1_%s_SITE%d_POES' % (C, i)) for i in range(1, 8)])
        src = SimpleFaultSource(
            source_id='fault1', name='fault1',
            tectonic_region_type=const.TRT.ACTIVE_SHALLOW_CRUST, mfd=mfd,
            rupture_mesh_spacing=1.0,
            magnitude_scaling_relationship=PeerMSR(),
            rupture_aspect_ratio=T.SET1_RUPTURE_ASPECT_RATIO,
            temporal_occurrence_model=PoissonTOM(1.),
            upper_seismogenic_depth=T.SET1_CASE1TO9_UPPER_SEISMOGENIC_DEPTH,
            lower_seismogenic_depth=T.SET1_CASE1TO9_LOWER_SEISMOGENIC_DEPTH,
            fault_trace=tr, dip=T.SET1_CASE1TO9_DIP,
            rake=T.SET1_CASE1TO9_RAKE)
        out[case + '_nrup'] = src.count_ruptures()
        try:
            # calc_hazard_curves goes through the task machinery (h5py);
            # its per-group body is ContextMaker.get_ctxs -> get_pmap
            # (calc/hazard_curve.py:40-164, contexts.py:1205-1218)
            cm = ContextMaker(const.TRT.ACTIVE_SHALLOW_CRUST,
                              [SadighEtAl1997()],
                              dict(imtls={'PGA': imls}, truncation_level=0,
                                   investigation_time=1.))
            cm.tom = PoissonTOM(1.)
            ctx = cm.recarray(list(cm.get_ctxs(src, SiteCollection(fsites))))
            out[case + '_nrows'] = len(ctx)
            out[case + '_ref'] = cm.get_pmap(ctx).array[:, :, 0]
            # the same through a map of rates, as RmapMaker._make_src_indep
            # does (contexts.py:1548-1579): float32 additions, no rounding
            # of factors 1 - 1e-5 to float32 as in the map of probabilities
            rmap = MapArray(np.arange(len(fsites), dtype=np.uint32),
                            len(imls), 1, rates=True).fill(0)
            cm.update(rmap, ctx)
            out[case + '_ref_rates'] = rmap.array[:, :, 0]
            print(case, 'reference run: max |ref - PEER| =',
                  np.abs(out[case + '_ref'] - out[case + '_poes']).max())
        except Exception as exc:      # rjb of a mesh needs shapely here
            print(case, 'reference run not possible here:', repr(exc)[:200])
    # the non-parametric source equivalent to case 2 (peer_test.py:216-266):
    # 6 x 12 rupture meshes of 8 x 15 nodes, each with the probabilities of
    # 0 and 1 occurrences
    nd = T.SET1_CASE2_SOURCE_DATA
    lons, lats, deps, hypo = [], [], [], []
    for i in range(nd['num_rups_dip']):
        for j in range(nd['num_rups_strike']):
            lons.append(nd['lons'])
            lats.append(nd['lats'][j])
            deps.append(nd['depths'][i])
            hypo.append([nd['hypo_lons'][i, j], nd['hypo_lats'][i, j],
                         nd['hypo_depths'][i, j]])
    out['np_lons'], out['np_lats'] = np.array(lons), np.array(lats)
    out['np_deps'], out['np_hypo'] = np.array(deps), np.array(hypo)
    out['np_pmf'] = np.array([p for p, _n in nd['pmf'].data])
    assert [n for _p, n in nd['pmf'].data] == [0, 1]
    out['np_mag'], out['np_rake'] = nd['mag'], nd['rake']
    np.savez_compressed(os.path.join(GOLDEN, 'peer_set1.npz'), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
