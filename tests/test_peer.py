"""The PEER verification tests (PEER 2010/106, Thomas, Wong, Abrahamson) that
the reference runs as acceptance tests of this path
(openquake/hazardlib/tests/acceptance/peer_test.py): set 1 cases 10 and 11
(area source, SadighEtAl1997 on rock, truncation level 0) and cases 2 and 5
(vertical strike-slip fault, PeerMSR), on the device path, against the
PUBLISHED probabilities with the tolerances of the reference's own tests --
and, for the fault cases, against the reference's own curves
(oracle/make_golden_peer.py)."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'peer_set1.npz'))
TRT = 'Active Shallow Crust'


def test_fixture_is_the_reference_test_data():
    # page 14 of the report: truncated Gutenberg-Richter, 15 magnitude bins
    assert D['area_mfd'].shape == (15, 2) and len(D['poly_lons']) == 90
    assert D['c10_poes'].shape == (4, 10) and D['c11_poes'].shape == (4, 11)
    assert D['c2_ref'].shape == D['c2_poes'].shape == (7, 15)
    # the reference itself meets the published values within its tolerances
    for case, atol in (('c2', 3e-3), ('c5', 1e-3)):
        assert np.abs(D[case + '_ref'] - D[case + '_poes']).max() < atol


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def sites(oq, arr):
    return oq.SiteCollection.from_points(
        arr[:, 0], arr[:, 1], np.zeros(len(arr)),
        dict(vs30=arr[:, 2], vs30measured=True, z1pt0=1., z2pt5=2.))


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['c10', 'c11'])
def test_peer_set1_area_source(oq, case):
    """peer_test.py:48-126: an area source of 100 km radius discretised at
    10 km (polygon discretisation on the host, area.py; ruptures, distances,
    GMPE, PoEs and map on the device), PointMSR, aspect ratio 2"""
    hdd = [(1., float(D['c10_depth']))] if case == 'c10' else [
        (1. / len(D['c11_depths']), float(d)) for d in D['c11_depths']]
    src = oq.AreaSource(
        'area', D['poly_lons'], D['poly_lats'], 10.,
        oq.TruncatedGRMFD(*D['area_gr']), 'PointMSR', float(D['rar']), 0.,
        10., [(1., 0., 90., 0.)], hdd, tectonic_region_type=TRT)
    mags = [m for m, _ in src.mfd.get_annual_occurrence_rates()]
    np.testing.assert_allclose(mags, D['area_mfd'][:, 0], rtol=1e-14)
    curves = oq.calc_hazard_curves(
        list(src), sites(oq, D['area_sites']), {'PGA': D[case + '_imls']},
        {TRT: oq.SadighEtAl1997()}, 0, investigation_time=1.)
    got, exp = curves['PGA'], D[case + '_poes']
    assert got.shape == exp.shape
    # assert_hazard_curve_is(..., atol=1e-4, rtol=1e-1), peer_test.py:77-84
    np.testing.assert_allclose(got, exp, atol=1e-4, rtol=1e-1)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['c2', 'c5'])
def test_peer_set1_fault(oq, case):
    """peer_test.py:128-214: the 25 km vertical strike-slip fault, floating
    ruptures with PeerMSR (area 10^(M - 4)) and aspect ratio 2 over the 1 km
    mesh, through the fused pass with the mesh distance planes"""
    usd, lsd, dip, rake = D['fault']
    mfd = oq.ArbitraryMFD(D[case + '_mfd'][:, 0], D[case + '_mfd'][:, 1])
    src = oq.SimpleFaultSource(
        'fault1', (D['mesh_lons'], D['mesh_lats'], D['mesh_deps']), mfd, 1.0,
        'PeerMSR', float(D['rar']), float(rake), tectonic_region_type=TRT)
    assert len(src.ruptures()) == int(D[case + '_nrup'])
    curves = oq.calc_hazard_curves(
        [src], sites(oq, D['fault_sites']), {'PGA': D[case + '_imls']},
        {TRT: oq.SadighEtAl1997()}, 0, investigation_time=1.)
    got = curves['PGA']
    # the reference's own run.  (1) Its map of rates (RmapMaker's mode):
    # float32 additions, tight.  (2) Its map of probabilities multiplies
    # float32 factors 1 - 1e-5 .. 1 - 1e-7, each rounded to 6e-8: with
    # thousands of ruptures (case 5) that alone is 1e-4 relative, so there
    # the tolerance is the reference's own float32 error, not the device's
    np.testing.assert_allclose(got, 1. - np.exp(-D[case + '_ref_rates'].astype(
        float)), rtol=2e-5, atol=1.5e-7)     # the curves are float32
    np.testing.assert_allclose(got, D[case + '_ref'], rtol=3e-4, atol=2.5e-7)
    # the published values with the reference's tolerances per site
    atols = ([3e-3, 2e-5, 2e-5, 1e-3, 1e-3, 1e-3, 2e-5] if case == 'c2' else
             [1e-3] * 7)
    for s in range(7):
        np.testing.assert_allclose(got[s], D[case + '_poes'][s],
                                   atol=atols[s], rtol=1e-5)


@pytest.mark.gpu
def test_peer_set1_non_parametric_source(oq):
    """peer_test.py:216-266: the non-parametric source equivalent to case 2
    -- 72 rupture meshes of 8 x 15 nodes, each with the probabilities of 0
    and 1 occurrences (probs_occur) instead of a rate -- against the
    published values of case 2 with the tolerances of the reference's test"""
    U = len(D['np_lons'])
    assert U == 72 and D['np_lons'].shape[1:] == (8, 15)
    rups = oq.MeshRuptures.from_meshes(
        [(D['np_lons'][u], D['np_lats'][u], D['np_deps'][u])
         for u in range(U)],
        np.full(U, float(D['np_mag'])), float(D['np_rake']),
        np.full(U, np.nan), hypo=D['np_hypo'],
        probs_occur=np.tile(D['np_pmf'], (U, 1)))
    group = oq.SourceGroup(TRT, [rups], 'non parametric', 'indep', 'indep')
    curves = oq.calc_hazard_curves(
        [group], sites(oq, D['fault_sites']), {'PGA': D['c2_imls']},
        {TRT: oq.SadighEtAl1997()}, 0, investigation_time=1.)
    got = curves['PGA']
    atols = [3e-3, 2e-5, 2e-5, 1e-3, 1e-3, 1e-3, 2e-5]   # peer_test.py:253-266
    for s in range(7):
        np.testing.assert_allclose(got[s], D['c2_poes'][s], atol=atols[s],
                                   rtol=1e-5)
