"""Multi-fault ruptures (the `dparam` cache)
against the reference's own run (oracle/make_golden_multifault.py ->
tests/golden/multifault.npz): the host bookkeeping of the lines (projection
centre, length, azimuth; flipped / order / shift / u_max of the MultiLine of
every rupture; the area-weighted parameters), the GC2 coordinates of the
sites for both orientations of every section (oqg_line_tuw), the composed
rrup / rjb / rx / ry0 planes (oqg_multi_compose), the context rows of genctxs
and the PoE map through the fused pass.

The reference keeps tuw, the shifts and the multi-surface parameters in
float32 and accumulates t and u in float32: rx / ry0 and dip / width / ztor
are compared at rtol 2e-6 (+ 2e-4 km), everything else as the other tests."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'multifault.npz'))
S = int(D['nsec'])
RUP_SECS = [D['rup_sec'][D['rup_off'][u]:D['rup_off'][u + 1]]
            for u in range(len(D['mags']))]


def ruptures(oq):
    from oq_engine_b200.multifault import MultiFaultRuptures
    secs = [(D['sec%d_lons' % k], D['sec%d_lats' % k], D['sec%d_deps' % k])
            for k in range(S)]
    tors = [D['sec%d_tor' % k] for k in range(S)]
    hypo = np.array([D['rup%d_hypo' % u] for u in range(len(RUP_SECS))])
    return MultiFaultRuptures(secs, tors, D['secparam_area'], RUP_SECS,
                              D['mags'], D['rakes'], D['rate'], hypo)


def test_lines_and_multilines_match_the_reference():
    import oq_engine_b200 as oq
    from oq_engine_b200 import multifault as mf
    for k in range(S):
        ln = mf.TorLine(D['sec%d_tor' % k])
        np.testing.assert_allclose(
            [ln.lam0, ln.phi0, ln.length, ln.azimuth], D['sec%d_tor_par' % k],
            rtol=1e-12, atol=1e-14)
        fl = ln.flip()
        np.testing.assert_allclose(
            [fl.lam0, fl.phi0, fl.length, fl.azimuth],
            D['sec%d_tor_par_flip' % k], rtol=1e-12, atol=1e-14)
        # GC2 coordinates on the host (used for u_max) = get_dparam 'tuw'
        for tag in ('far', 'near'):
            lon, lat = D[tag + '_site_lon'], D[tag + '_site_lat']
            ref = D[tag + '_sec_tuw'][k]               # (N, 2, 3) float32
            np.testing.assert_allclose(ln.get_tuw(lon, lat), ref[:, 0],
                                       rtol=2e-6, atol=2e-5)
            np.testing.assert_allclose(fl.get_tuw(lon, lat), ref[:, 1],
                                       rtol=2e-6, atol=2e-5)
    rups = ruptures(oq)
    for u, secs in enumerate(RUP_SECS):
        flipped, soidx, shift = mf.multiline([rups.lines[i] for i in secs])
        np.testing.assert_array_equal(flipped, D['rup%d_flipped' % u])
        np.testing.assert_array_equal(soidx, D['rup%d_soidx' % u])
        np.testing.assert_allclose(shift, D['rup%d_shift' % u], rtol=1e-6,
                                   atol=1e-5)
    np.testing.assert_allclose(rups.umax, D['msparam_u_max'], rtol=1e-6)
    ms = rups.msparams({k: D['secparam_' + k] for k in (
        'dip', 'strike', 'width', 'ztor')})
    for k in ('dip', 'strike', 'width', 'ztor'):
        np.testing.assert_allclose(ms[k], D['msparam_' + k], rtol=1e-6)
    with pytest.raises(ValueError):
        mf.MultiFaultRuptures(rups.sections, [D['sec%d_tor' % k]
                                              for k in range(S)],
                              D['secparam_area'], [[0, 99]], [6.], 0., [1.],
                              [[0, 0, 5.]])


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def sitecol(oq, tag):
    return oq.SiteCollection.from_points(
        D[tag + '_site_lon'], D[tag + '_site_lat'], D[tag + '_site_depth'],
        {k: D['%s_site_%s' % (tag, k)] for k in ('vs30', 'vs30measured',
                                                 'z1pt0', 'z2pt5')})


def maker(oq):
    return oq.ContextMaker(
        'Active Shallow Crust', [oq.get_gsim(str(n)) for n in D['gsims']],
        dict(imtls={'PGA': D['levels'], 'SA(1.0)': D['levels']},
             truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('300')))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['far', 'near'])
def test_section_cache(oq, tag):
    """the `dparam` cache: rrup, rjb and tuw of every section"""
    from oq_engine_b200 import multifault as mf
    cm, rups, sc = maker(oq), ruptures(oq), sitecol(oq, tag)
    names = ['rrup', 'rx', 'ry0'] + (['rjb'] if tag == 'far' else [])
    cache = mf.section_planes(cm, rups, sc, names)
    np.testing.assert_allclose(cache['rrup'].cpu().numpy(),
                               D[tag + '_sec_rrup'], rtol=1e-10, atol=1e-8)
    if tag == 'far':
        np.testing.assert_allclose(cache['rjb'].cpu().numpy(),
                                   D['far_sec_rjb'], rtol=1e-9, atol=1e-7)
    tuw = cache['tuw'].cpu().numpy()                  # (S, 2, N, 3)
    ref = D[tag + '_sec_tuw'].transpose(0, 2, 1, 3)   # (S, N, 2, 3) ->
    np.testing.assert_allclose(tuw, ref, rtol=2e-6, atol=2e-5)
    for k in ('dip', 'strike', 'width', 'ztor'):
        np.testing.assert_allclose(cache['par'][k], D['secparam_' + k],
                                   rtol=1e-6)


@pytest.mark.gpu
def test_rx_ry0_near_the_faults(oq):
    """T and U of every rupture on 40 sites on top of and around the faults
    against the reference's _get_tu on its own cache"""
    from oq_engine_b200 import multifault as mf
    cm, rups, sc = maker(oq), ruptures(oq), sitecol(oq, 'near')
    planes, _par = mf.rupture_planes(cm, rups, sc, ['rrup', 'rx', 'ry0'])
    rx, ry0 = planes['rx'].cpu().numpy(), planes['ry0'].cpu().numpy()
    rrup = planes['rrup'].cpu().numpy()
    for u, secs in enumerate(RUP_SECS):
        t, uu = D['rup%d_near_t' % u], D['rup%d_near_u' % u]
        um = D['msparam_u_max'][u]
        want = np.where(uu < 0, np.abs(uu), np.where(uu > um, uu - um, 0.))
        np.testing.assert_allclose(rx[u], t, rtol=5e-6, atol=5e-4)
        np.testing.assert_allclose(ry0[u], want, rtol=5e-6, atol=5e-4)
        np.testing.assert_allclose(rrup[u], D['near_sec_rrup'][secs].min(0),
                                   rtol=1e-10, atol=1e-8)


@pytest.mark.gpu
def test_contexts_and_map_of_multifault_ruptures(oq):
    cm, rups, sc = maker(oq), ruptures(oq), sitecol(oq, 'far')
    ctxs = cm.get_ctxs(rups, sc)
    np.testing.assert_array_equal([len(c) for c in ctxs], D['ctx_lens'])
    ctx = cm.recarray(ctxs)
    for f in ('sids', 'rup_id'):
        np.testing.assert_array_equal(ctx[f], D['ctx_' + f])
    for f in ('mag', 'rake', 'rrup', 'rjb', 'occurrence_rate', 'hypo_depth'):
        np.testing.assert_allclose(ctx[f], D['ctx_' + f], rtol=1e-9,
                                   atol=1e-7, err_msg=f)
    for f in ('dip', 'ztor', 'width', 'rx', 'ry0'):   # float32 in the reference
        np.testing.assert_allclose(ctx[f], D['ctx_' + f], rtol=5e-6,
                                   atol=5e-4, err_msg=f)
    ms = cm.get_mean_stds([ctx])
    np.testing.assert_allclose(ms, D['mean_stds'], rtol=1e-5, atol=1e-6)
    # fused: section cache -> composed planes -> oqg_hazard_dist
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
    kept = cm.hazard_mesh(pmap, rups, sc, count=True)
    assert kept == int(D['ctx_lens'].sum())
    sids = np.unique(D['ctx_sids'])
    np.testing.assert_allclose(pmap.array[sids], 1. - D['poes'], rtol=3e-5,
                               atol=2e-7)
    res = oq.RmapMaker(cm, sc, [rups]).make()
    np.testing.assert_allclose(
        np.exp(-res['rmap'].array[sids].astype(np.float64)), 1. - D['poes'],
        rtol=3e-5, atol=2e-7)
