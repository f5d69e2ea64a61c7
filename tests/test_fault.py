"""Ruptures of fault sources against the
reference's own run of the simple fault of the LogicTreeCase2ClassicalPSHA
demo (oracle/make_golden_fault.py -> tests/golden/fault_ctxs.npz):

* the floating ruptures (windows, magnitudes, rates, hypocentres) of the
  host mirror are those of SimpleFaultSource.iter_ruptures;
* dip / strike / width / ztor of every rupture, the context rows of
  ContextMaker.get_ctxs (order, sids, every distance and parameter) and the
  float32 PoE map of get_pmap, through the mesh kernels and the fused pass
  with external distance planes (oqg_hazard_dist).

Tolerances: distances rtol 1e-9 (rx of the reference goes through azimuths:
1e-7 km absolute), PoEs rtol 2e-5 on pne (float32 map) as for the planar
path."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'fault_ctxs.npz'))


def source(oq):
    return oq.SimpleFaultSource(
        'second', (D['mesh_lons'], D['mesh_lats'], D['mesh_deps']),
        oq.TruncatedGRMFD(*D['mfd']), float(D['spacing']), 'WC1994',
        float(D['rar']), float(D['rake']))


def sitecol(oq):
    return oq.SiteCollection.from_points(
        D['site_lon'], D['site_lat'], D['site_depth'],
        {k: D['site_' + k] for k in ('vs30', 'vs30measured', 'z1pt0',
                                     'z2pt5')})


def test_floating_ruptures_match_the_reference():
    import oq_engine_b200 as oq
    src = source(oq)
    mag, rate, win = src.windows()
    np.testing.assert_array_equal(win, D['windows'])
    np.testing.assert_allclose(mag, D['mag'], rtol=1e-15)
    np.testing.assert_allclose(rate, D['rate'], rtol=1e-14)
    rups = src.ruptures()
    assert len(rups) == src.count_ruptures() == 400
    np.testing.assert_allclose(rups.hypo, D['hypo'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(rups.ztor(), D['ztor'], rtol=1e-15)
    # the meshes of the ruptures are the windows of the fault mesh
    for u in (0, 57, 399):
        r0, c0, nr, nc = D['windows'][u]
        lo, la, de = rups.mesh(u)
        np.testing.assert_array_equal(
            lo, D['mesh_lons'][r0:r0 + nr, c0:c0 + nc])
        np.testing.assert_array_equal(
            de, D['mesh_deps'][r0:r0 + nr, c0:c0 + nc])
    # the same ruptures given as a list of meshes
    few = oq.MeshRuptures.from_meshes(
        [rups.mesh(u) for u in (3, 200, 399)], rups.mag[[3, 200, 399]], 90.,
        rups.rate[[3, 200, 399]])
    np.testing.assert_allclose(few.hypo, D['hypo'][[3, 200, 399]], atol=1e-12)
    with pytest.raises(ValueError):
        oq.MeshRuptures.from_windows(src.mesh, [[0, 0, 8, 3]], [6.], 0., [1.])
    with pytest.raises(ValueError):     # a mesh of one row has no dip
        oq.MeshRuptures.from_windows(src.mesh, [[0, 0, 1, 3]], [6.], 0., [1.])
    with pytest.raises(ValueError):     # too coarse a mesh for the magnitude
        oq.SimpleFaultSource('x', src.mesh, oq.TruncatedGRMFD(
            4.0, 5.0, 0.1, 3., 1.), 40.)


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def maker(oq, tag):
    imtls = {str(i): D['levels'] for i in D[tag + '_imts']}
    return oq.ContextMaker(
        'Active Shallow Crust', [oq.get_gsim(str(n))
                                 for n in D[tag + '_gsims']],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new(
                 str(float(D['maxdist'])))))


@pytest.mark.gpu
def test_rupture_parameters(oq):
    from oq_engine_b200 import fault
    cm = maker(oq, 'nga')
    rups = source(oq).ruptures()
    planes, par = fault.rupture_planes(cm, rups, sitecol(oq),
                                       ['rrup', 'rjb', 'rx', 'ry0'])
    np.testing.assert_allclose(par['dip'], D['dip'], rtol=1e-10)
    np.testing.assert_allclose(par['strike'], D['strike'], rtol=1e-10,
                               atol=1e-9)   # strike ~ 0 on the N-S segment
    np.testing.assert_allclose(par['width'], D['width'], rtol=1e-10)
    np.testing.assert_allclose(par['ztor'], D['ztor'], rtol=1e-15)
    assert set(planes) == {'rrup', 'rjb', 'rx', 'ry0'}
    assert planes['rrup'].shape == (400, len(D['site_lon']))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['demo', 'nga'])
def test_get_ctxs_of_a_fault_source(oq, tag):
    cm = maker(oq, tag)
    sc = sitecol(oq)
    ctxs = cm.get_ctxs(source(oq), sc)
    assert len(ctxs) == int(D[tag + '_nctxs'])
    np.testing.assert_array_equal([len(c) for c in ctxs],
                                  D[tag + '_ctx_lens'])
    ctx = cm.recarray(ctxs)
    for f in ('sids', 'rup_id'):
        np.testing.assert_array_equal(ctx[f], D['%s_ctx_%s' % (tag, f)])
    for f in ('mag', 'rake', 'dip', 'ztor', 'width', 'hypo_depth', 'rrup',
              'rjb', 'rx', 'ry0', 'vs30', 'z1pt0', 'z2pt5',
              'occurrence_rate'):
        key = '%s_ctx_%s' % (tag, f)
        if key in D.files:
            np.testing.assert_allclose(ctx[f], D[key], rtol=1e-9, atol=1e-7,
                                       err_msg=f)
    # the unfused operators on those rows: mean and standard deviations,
    # then the map
    ms = cm.get_mean_stds([ctx])
    np.testing.assert_allclose(ms, D[tag + '_mean_stds'], rtol=1e-8,
                               atol=1e-10)
    pmap = cm.get_pmap(ctx, oq.PoissonTOM(50.))
    np.testing.assert_allclose(1. - pmap.array, 1. - D[tag + '_poes'],
                               rtol=2e-5, atol=1.5e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['demo', 'nga'])
def test_hazard_of_a_fault_source_fused(oq, tag):
    cm = maker(oq, tag)
    sc = sitecol(oq)
    src = source(oq)
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
    kept = cm.hazard_mesh(pmap, src, sc, count=True)
    assert kept == int(D[tag + '_ctx_lens'].sum())
    # the reference's map has a row for each site with contexts
    # (get_pmap, contexts.py:1206-1218); the others stay at pne = 1
    sids = np.unique(D[tag + '_ctx_sids'])
    assert len(sids) == len(D[tag + '_poes']) < len(sc)
    far = np.setdiff1d(sc.sids, sids)
    assert (pmap.array[far] == 1.).all()
    np.testing.assert_allclose(pmap.array[sids], 1. - D[tag + '_poes'],
                               rtol=2e-5, atol=1.5e-7)
    poes = D[tag + '_poes'].astype(np.float64)
    big = poes > 1e-2       # float32 pnes: 6e-8 absolute each
    np.testing.assert_allclose(
        1. - pmap.array[sids][big].astype(np.float64), poes[big], rtol=5e-4)
    # rates map, device resident, through RmapMaker
    res = oq.RmapMaker(cm, sc, [src]).make()
    rates = res['rmap'].array
    np.testing.assert_allclose(np.exp(-rates[sids].astype(np.float64)),
                               1. - D[tag + '_poes'], rtol=2e-5, atol=1.5e-7)
    assert res['source_data']['nrupts'] == [400]
    # in chunks of ruptures (bounded scratch): same map
    from oq_engine_b200 import fault
    old = fault.PLANE_BYTES
    fault.PLANE_BYTES = 64 * len(sc) * 8 * 4
    try:
        pm2 = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
        assert cm.hazard_mesh(pm2, src, sc, count=True) == kept
    finally:
        fault.PLANE_BYTES = old
    np.testing.assert_allclose(pm2.array, pmap.array, rtol=3e-6)
