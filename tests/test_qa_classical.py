"""Engine-level known answers: the expected hazard
curves of the reference's QA test classical/case_01
(openquake/qa_tests_data/classical/case_01/expected/*.csv, checked by
calculators/tests/classical_test.py:73-76 at delta 1e-5) reproduced by the
device path from the QA test's own inputs (oracle/make_golden_qa.py reads its
job.ini and source model)."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'qa_classical_case_01.npz'))


def test_fixture():
    assert str(D['msr']) == 'CScalingMSR.C = 3.90'
    assert D['sites'].shape == (4, 3) and D['npd'].shape == (3, 4)
    assert D['expected_PGA'].shape == D['expected_SA(0.1)'].shape == (4, 3)


def test_msr_names():
    import oq_engine_b200 as oq
    from oq_engine_b200.source import parse_msr, median_area
    assert parse_msr('CScalingMSR.C = 3.90') == ('CScalingMSR', 3.9)
    assert parse_msr('PeerMSR') == ('PeerMSR', 4.0)
    np.testing.assert_allclose(median_area('CScalingMSR.C = 4.0', [6.], 0.),
                               median_area('PeerMSR', [6.], 0.), rtol=0)
    np.testing.assert_allclose(median_area('PeerMSR', [6.5], 0.),
                               10 ** 2.5, rtol=1e-15)
    for bad in ('Leonard2014_SCR', 'PeerMSR.C = 1', 'CScalingMSR'):
        with pytest.raises((NotImplementedError, ValueError)):
            oq.PointSource('x', 0., 0., oq.ArbitraryMFD([5.], [1.]), bad,
                           1., 0., 10., [(1., 0., 90., 0.)], [(1., 5.)])


@pytest.mark.gpu
def test_qa_classical_case_01():
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(D['trt'])
    src = oq.PointSource(
        '1', float(D['lon']), float(D['lat']),
        oq.EvenlyDiscretizedMFD(float(D['min_mag']), float(D['bin_width']),
                                list(D['occur_rates'])),
        str(D['msr']), float(D['rar']), float(D['usd']), float(D['lsd']),
        [tuple(r) for r in D['npd']], [tuple(r) for r in D['hdd']],
        tectonic_region_type=trt)
    assert src.count_ruptures() == 3      # the engine collapses the equal planes
    S = D['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], S[:, 2],
        dict(vs30=float(D['vs30']), vs30measured=True,
             z1pt0=float(D['z1pt0']), z2pt5=float(D['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(D['imts'], D['levels'])}
    curves = oq.calc_hazard_curves(
        [src], sc, imtls, {trt: oq.SadighEtAl1997()},
        float(D['truncation_level']),
        investigation_time=float(D['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(D['maximum_distance']))),
        pointsource_distance=100.)        # the engine's default
    for imt in imtls:
        exp = D['expected_' + imt]
        # the files hold 7 significant digits; delta 1e-5 of the QA test
        np.testing.assert_allclose(curves[imt], exp, rtol=1e-5, atol=1e-7,
                                   err_msg=imt)
        assert (curves[imt][exp == 0] == 0).all()


E = np.load(os.path.join(GOLDEN, 'qa_classical_case_24.npz'))


@pytest.mark.gpu
def test_qa_classical_case_24():
    """classical/case_24 (the area source of BASELINE configs[0] with three
    hypocentral depths): 52 point sources x 5 magnitudes x 3 depths = 780
    ruptures, BooreAtkinson2008 (rjb), PGV + PGA + 7 SA x 45 levels, two
    sites, against the engine's expected curves.  The engine bins the
    truncated Gutenberg-Richter MFD with width_of_mfd_bin.  Checked against
    the expected files at the QA test's delta 1e-3 (classical_test.py:355-363)
    and against this version's hazardlib run on the same inputs at 2e-5."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(E['trt'])
    mn, mx, a, b = E['gr']
    src = oq.AreaSource(
        '1', E['poly_lons'], E['poly_lats'], float(E['area_discretization']),
        oq.TruncatedGRMFD(mn, mx, float(E['width_of_mfd_bin']), a, b),
        str(E['msr']), float(E['rar']), float(E['usd']), float(E['lsd']),
        [tuple(r) for r in E['npd']], [tuple(r) for r in E['hdd']],
        tectonic_region_type=trt)
    srcs = list(src)
    assert sum(s.count_ruptures() for s in srcs) == int(E['num_ruptures'])
    S = E['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(E['vs30']), vs30measured=True,
             z1pt0=float(E['z1pt0']), z2pt5=float(E['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(E['imts'], E['levels'])}
    curves = oq.calc_hazard_curves(
        srcs, sc, imtls, {trt: oq.BooreAtkinson2008()},
        float(E['truncation_level']),
        investigation_time=float(E['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(E['maximum_distance']))),
        pointsource_distance=float(E['pointsource_distance']))
    for m, imt in enumerate(imtls):
        # the QA test's own comparison: assert_allclose(atol=delta,
        # rtol=delta) with delta = 1e-3 (the expected files were written by
        # engine 3.20; this version's hazardlib differs from them by up to
        # 2.4e-4 absolute, oracle/make_golden_qa.py prints it)
        np.testing.assert_allclose(curves[imt], E['expected_' + imt],
                                   rtol=1e-3, atol=1e-3, err_msg=imt)
        # this version of the reference, run on the same inputs
        ref = E['ref_poes'][:, m]
        np.testing.assert_allclose(curves[imt], ref, rtol=2e-5, atol=1.5e-7,
                                   err_msg=imt)


F = np.load(os.path.join(GOLDEN, 'qa_classical_case_26.npz'))


@pytest.mark.gpu
def test_qa_classical_case_26():
    """classical/case_26 (classical_test.py:373-375, delta 2e-5): an area
    source over a 1 x 1 degree square discretised at 5 km (484 point
    sources), YoungsCoppersmith1985MFD with characteristic magnitude 7.0
    (23 magnitudes, 11 132 ruptures, the large ones clipped to the
    seismogenic layer), BooreAtkinson2008, SA(1.0) x 19 levels, one site:
    the engine's expected curve at the QA test's delta, and this version's
    hazardlib on the same inputs."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(F['trt'])
    src = oq.AreaSource(
        '1', F['poly_lons'], F['poly_lats'], float(F['area_discretization']),
        oq.ArbitraryMFD(F['mfd'][:, 0], F['mfd'][:, 1]), str(F['msr']),
        float(F['rar']), float(F['usd']), float(F['lsd']),
        [tuple(r) for r in F['npd']], [tuple(r) for r in F['hdd']],
        tectonic_region_type=trt)
    srcs = list(src)
    assert len(srcs) == int(F['num_points'])
    assert sum(s.count_ruptures() for s in srcs) == int(F['num_ruptures'])
    S = F['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(F['vs30']), vs30measured=True,
             z1pt0=float(F['z1pt0']), z2pt5=float(F['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(F['imts'], F['levels'])}
    [imt] = imtls
    curves = oq.calc_hazard_curves(
        srcs, sc, imtls, {trt: oq.BooreAtkinson2008()},
        float(F['truncation_level']),
        investigation_time=float(F['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(F['maximum_distance']))))
    np.testing.assert_allclose(curves[imt], F['expected'], rtol=2e-5,
                               atol=2e-5)           # the QA test's delta
    np.testing.assert_allclose(curves[imt], F['ref_poes'][:, 0], rtol=2e-5,
                               atol=1.5e-7)


H = np.load(os.path.join(GOLDEN, 'qa_classical_case_31.npz'))


@pytest.mark.gpu
def test_qa_classical_case_31():
    """classical/case_31 (classical_test.py:472-474, delta 1e-5): an area
    source, a point source and a multi-point source with dipping reverse
    ruptures, three hypocentral depths each with its own fixedDipFrac
    (hypo_dip_fracs), ChiouYoungs2014 with its hanging-wall terms, the
    engine's default pointsource_distance: 6 075 ruptures on one site, the
    engine's expected mean curve and this version's hazardlib."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(H['trt'])
    mn, mx, a, b = H['gr']
    binw = float(H['width_of_mfd_bin'])
    npd = [tuple(r) for r in H['npd']]
    hdd = [(p, dep) for p, dep, _f in H['hdd']]
    fracs = [f for _p, _dep, f in H['hdd']]
    args = (str(H['msr']), float(H['rar']), float(H['usd']), float(H['lsd']),
            npd, hdd, fracs)
    # the three sources as objects: the area and the multi-point source are
    # split into their point sources by the RmapMaker, as the engine does
    srcs = [
        oq.AreaSource('1', H['poly_lons'], H['poly_lats'],
                      float(H['area_discretization']),
                      oq.TruncatedGRMFD(mn, mx, binw, a, b), *args,
                      tectonic_region_type=trt),
        oq.PointSource('2', H['point'][0], H['point'][1],
                       oq.TruncatedGRMFD(mn, mx, binw, a, b), *args,
                       tectonic_region_type=trt),
        oq.MultiPointSource(
            '3', H['multi_points'][:, 0], H['multi_points'][:, 1],
            [oq.TruncatedGRMFD(mn, mx, binw, av, b)
             for av in H['multi_a_vals']], *args, tectonic_region_type=trt)]
    assert sum(s.count_ruptures() for s in srcs) == int(H['num_ruptures'])
    S = H['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(H['vs30']), vs30measured=True,
             z1pt0=float(H['z1pt0']), z2pt5=float(H['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(H['imts'], H['levels'])}
    curves = oq.calc_hazard_curves(
        srcs, sc, imtls, {trt: oq.ChiouYoungs2014()},
        float(H['truncation_level']),
        investigation_time=float(H['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(H['maximum_distance']))),
        pointsource_distance=float(H['pointsource_distance']))
    np.testing.assert_allclose(curves['PGA'], H['expected'], rtol=1e-5,
                               atol=1e-5)           # the QA test's delta
    np.testing.assert_allclose(curves['PGA'], H['ref_poes'], rtol=2e-5,
                               atol=1.5e-7)


K = np.load(os.path.join(GOLDEN, 'qa_classical_case_28.npz'))


def test_case_28_hypo_depth_dist_does_not_change_the_curve():
    # the reference with and without the hypoDepthDist (renormalised weights,
    # a model that reads nothing of the hypocentre): the same curve
    np.testing.assert_allclose(K['ref_poes'], K['ref_poes_plain'], rtol=1e-6,
                               atol=1e-12)
    np.testing.assert_allclose(K['ref_poes'], K['expected'], rtol=1e-5,
                               atol=1e-5)


@pytest.mark.gpu
def test_qa_classical_case_28():
    """classical/case_28 (classical_test.py:406-408, delta 1e-5): a dipping
    simple fault (dip 60, reverse), floating ruptures over the 5 km mesh,
    ChiouYoungs2014 (rrup, rjb, rx, ztor, dip of mesh ruptures: the mesh
    kernels + oqg_hazard_dist), one site at 57 km: the engine's expected
    mean curve and this version's hazardlib.  The hypoDepthDist of the source
    does not change the curve (see the test above), the device path runs
    the plain fault."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(K['trt'])
    src = oq.SimpleFaultSource(
        '1', (K['mesh_lons'], K['mesh_lats'], K['mesh_deps']),
        oq.ArbitraryMFD(K['mfd'][:, 0], K['mfd'][:, 1]), float(K['spacing']),
        str(K['msr']), float(K['rar']), float(K['rake']),
        tectonic_region_type=trt)
    assert len(src.ruptures()) == int(K['num_ruptures_plain'])
    S = K['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(K['vs30']), vs30measured=True,
             z1pt0=float(K['z1pt0']), z2pt5=float(K['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(K['imts'], K['levels'])}
    curves = oq.calc_hazard_curves(
        [src], sc, imtls, {trt: oq.ChiouYoungs2014()},
        float(K['truncation_level']),
        investigation_time=float(K['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(K['maximum_distance']))))
    np.testing.assert_allclose(curves['PGA'], K['expected'], rtol=1e-5,
                               atol=1e-5)           # the QA test's delta
    np.testing.assert_allclose(curves['PGA'], K['ref_poes_plain'], rtol=2e-5,
                               atol=1.5e-7)


Q = np.load(os.path.join(GOLDEN, 'qa_classical_case_78.npz'))


@pytest.mark.gpu
def test_qa_classical_case_78():
    """classical/case_78 (classical_test.py:1014-1020, delta 1e-5): a point
    source and a dipping simple fault, each with its own NegativeBinomialTOM
    (non-Poissonian probs_occur per rupture), AbrahamsonEtAl2014 with an
    inferred vs30 and a basin depth, PGA x 50 levels: the engine's expected
    mean curve and this version's hazardlib.  The temporal occurrence model
    belongs to the source, not to the group."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(Q['trt'])
    T = float(Q['investigation_time'])
    point = oq.PointSource(
        '00000', Q['ps_loc'][0], Q['ps_loc'][1],
        oq.EvenlyDiscretizedMFD(Q['ps_mfd'][0], Q['ps_mfd'][1],
                                list(Q['ps_mfd'][2:])),
        str(Q['ps_msr']), float(Q['ps_rar']), float(Q['ps_usd']),
        float(Q['ps_lsd']), [tuple(r) for r in Q['ps_npd']],
        [tuple(r) for r in Q['ps_hdd']], tectonic_region_type=trt)
    point.temporal_occurrence_model = oq.NegativeBinomialTOM(T, *Q['ps_tom'])
    fault = oq.SimpleFaultSource(
        '00001', (Q['mesh_lons'], Q['mesh_lats'], Q['mesh_deps']),
        oq.ArbitraryMFD(Q['fs_mfd'][:, 0], Q['fs_mfd'][:, 1]),
        float(Q['spacing']), str(Q['fs_msr']), float(Q['fs_rar']),
        float(Q['fs_rake']), tectonic_region_type=trt)
    fault.temporal_occurrence_model = oq.NegativeBinomialTOM(T, *Q['fs_tom'])
    assert len(fault.ruptures()) == int(Q['fs_num_ruptures'])
    S = Q['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(Q['vs30']), vs30measured=bool(Q['vs30measured']),
             z1pt0=float(Q['z1pt0']), z2pt5=float(Q['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(Q['imts'], Q['levels'])}
    # one group: the RmapMaker folds the sources one temporal occurrence
    # model at a time
    curves = oq.calc_hazard_curves(
        [point, fault], sc, imtls, {trt: oq.AbrahamsonEtAl2014()},
        float(Q['truncation_level']), investigation_time=T,
        maximum_distance=oq.IntegrationDistance.new(
            str(float(Q['maximum_distance']))))
    np.testing.assert_allclose(curves['PGA'], Q['expected'], rtol=1e-5,
                               atol=1e-5)           # the QA test's delta
    # this version's hazardlib (float32 map of probabilities: its PoEs carry
    # an absolute 6e-8 per multiplication)
    np.testing.assert_allclose(curves['PGA'], Q['ref_poes'], rtol=2e-5,
                               atol=1e-6)


C3 = np.load(os.path.join(GOLDEN, 'qa_classical_case_03.npz'))


@pytest.mark.gpu
def test_qa_classical_case_03():
    """classical/case_03 (classical_test.py:118-120, delta 1e-5): two area
    sources over one polygon with minimum_magnitude = 5.7 -- the integration
    distance is cut at it (IntegrationDistance.cut), the smaller magnitudes
    have distance zero and their ruptures are dropped by the distance filter
    on the device, the second source (maximum magnitude 5.5) contributes
    nothing; MFD binned with width_of_mfd_bin = 0.2; BooreAtkinson2008."""
    import oq_engine_b200 as oq
    oq._lib.lib()
    trt = str(C3['trt'])
    binw = float(C3['width_of_mfd_bin'])
    srcs = [oq.AreaSource(
        str(k + 1), C3['poly_lons'], C3['poly_lats'],
        float(C3['area_discretization']),
        oq.TruncatedGRMFD(mn, mx, binw, a, b), str(C3['msr']),
        float(C3['rar']), float(C3['usd']), float(C3['lsd']),
        [tuple(r) for r in C3['npd']], [tuple(r) for r in C3['hdd']],
        tectonic_region_type=trt)
        for k, (mn, mx, a, b) in enumerate(C3['grs'])]
    maxdist = oq.IntegrationDistance.new(str(float(C3['maximum_distance'])))
    maxdist.cut({'default': float(C3['minimum_magnitude'])})
    S = C3['sites']
    sc = oq.SiteCollection.from_points(
        S[:, 0], S[:, 1], np.zeros(len(S)),
        dict(vs30=float(C3['vs30']), vs30measured=True,
             z1pt0=float(C3['z1pt0']), z2pt5=float(C3['z2pt5'])))
    imtls = {str(imt): lev for imt, lev in zip(C3['imts'], C3['levels'])}
    curves = oq.calc_hazard_curves(
        srcs, sc, imtls, {trt: oq.BooreAtkinson2008()},
        float(C3['truncation_level']),
        investigation_time=float(C3['investigation_time']),
        maximum_distance=maxdist)
    np.testing.assert_allclose(curves['PGA'], C3['expected'], rtol=1e-5,
                               atol=1e-5)           # the QA test's delta
    np.testing.assert_allclose(curves['PGA'], C3['ref_poes'], rtol=2e-5,
                               atol=1.5e-7)


@pytest.mark.parametrize('case, delta', [('03', 1e-5), ('26', 2e-5),
                                         ('28', 1e-5), ('31', 1e-5),
                                         ('78', 1e-5)])
def test_fixtures_reference_run_meets_the_expected_files(case, delta):
    """the pin of the fixtures: this version's hazardlib, run by
    oracle/make_golden_qa.py on the inputs of each QA test, reproduces the
    engine's expected file within the QA test's own delta (case_24 is the
    exception: its files date from engine 3.20 and are compared at 1e-3)"""
    d = np.load(os.path.join(GOLDEN, 'qa_classical_case_%s.npz' % case))
    ref = d['ref_poes'].reshape(d['expected'].shape)
    np.testing.assert_allclose(ref, d['expected'], rtol=delta, atol=delta)


def test_fixture_case_24_drift():
    for imt in E['imts']:
        m = list(E['imts']).index(imt)
        np.testing.assert_allclose(E['ref_poes'][:, m],
                                   E['expected_' + str(imt)], rtol=1e-3,
                                   atol=1e-3)
