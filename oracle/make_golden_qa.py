"""
Engine-level known answers: the expected hazard curves of the reference's QA
test classical/case_01 (openquake/qa_tests_data/classical/case_01, checked
by calculators/tests/classical_test.py:73-76 at delta 1e-5): one point source
(incremental MFD, CScalingMSR with C = 3.9, three nodal planes of which two
are equal, one hypocentral depth), SadighEtAl1997, 4 sites 100 m above sea
level, PGA and SA(0.1) x 3 levels, truncation level 2, 200 km.

The fixture is the INPUT of the QA test read from its own files (job.ini,
source_model.xml) and its expected CSV files; nothing is computed here.

    python oracle/make_golden_qa.py   -> tests/golden/qa_classical_case_01.npz
"""
import os
import ast
import csv
import configparser
import xml.etree.ElementTree as ET
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('OQ_REFERENCE', '/root/reference')
CASE = os.path.join(REF, 'openquake', 'qa_tests_data', 'classical', 'case_01')
GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
# SiteCollection.from_points sets z1pt0 / z2pt5 only when they are required
REQ_SITE_PARAMS = ('vs30', 'vs30measured', 'z1pt0', 'z2pt5')
NS = {'n': 'http://openquake.org/xmlns/nrml/0.4',
      'gml': 'http://www.opengis.net/gml'}


def main():
    ini = configparser.ConfigParser()
    ini.read(os.path.join(CASE, 'job.ini'))
    get = lambda sec, key: ini[sec][key]  # noqa
    sites = np.array([[float(x) for x in s.split()]
                      for s in get('geometry', 'sites').split(',')])
    imtls = ast.literal_eval(get('calculation',
                                 'intensity_measure_types_and_levels'))
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([imtls[i] for i in imtls]),
        vs30=float(get('site_params', 'reference_vs30_value')),
        z1pt0=float(get('site_params', 'reference_depth_to_1pt0km_per_sec')),
        z2pt5=float(get('site_params', 'reference_depth_to_2pt5km_per_sec')),
        investigation_time=float(get('calculation', 'investigation_time')),
        truncation_level=float(get('calculation', 'truncation_level')),
        maximum_distance=float(get('calculation', 'maximum_distance')))
    root = ET.parse(os.path.join(CASE, 'source_model.xml')).getroot()
    [src] = root.findall('.//n:pointSource', NS)
    lon, lat = [float(x) for x in src.find('.//gml:pos', NS).text.split()]
    mfd = src.find('n:incrementalMFD', NS)
    out.update(
        lon=lon, lat=lat, trt=src.get('tectonicRegion'),
        usd=float(src.find('.//n:upperSeismoDepth', NS).text),
        lsd=float(src.find('.//n:lowerSeismoDepth', NS).text),
        msr=src.find('n:magScaleRel', NS).text.strip(),
        rar=float(src.find('n:ruptAspectRatio', NS).text),
        min_mag=float(mfd.get('minMag')), bin_width=float(mfd.get('binWidth')),
        occur_rates=np.array([float(x) for x in mfd.find(
            'n:occurRates', NS).text.split()]),
        npd=np.array([[float(n.get(k)) for k in (
            'probability', 'strike', 'dip', 'rake')]
            for n in src.findall('.//n:nodalPlane', NS)]),
        hdd=np.array([[float(n.get(k)) for k in ('probability', 'depth')]
                      for n in src.findall('.//n:hypoDepth', NS)]))
    for imt in imtls:
        with open(os.path.join(CASE, 'expected',
                               'hazard_curve-%s.csv' % imt)) as f:
            rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
        assert rows[0][:3] == ['lon', 'lat', 'depth'], rows[0]
        body = np.array(rows[1:], float)
        np.testing.assert_allclose(body[:, :3], sites)
        out['expected_' + imt] = body[:, 3:]
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_01.npz'),
                        **out)
    print({k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})


def logscale(x_min, x_max, n):
    """valid.logscale of the reference's job.ini parser"""
    return np.logspace(np.log10(x_min), np.log10(x_max), int(n))


def reference_case_24(d, imtls, mfd=None, gsim='[BooreAtkinson2008]'):
    """The same calculation with the reference's hazardlib of THIS version
    (the expected files were written by engine 3.20 and are compared at delta
    1e-3): AreaSource -> point sources -> ContextMaker.get_ctxs -> update of
    a map of rates (RmapMaker's mode) -> 1 - exp(-rates); (N, M, L1)"""
    import sys
    sys.path.insert(0, HERE)
    # bootstraps the reference and installs the ray-casting Polygon.contains
    # in place of shapely's (absent here), as for BASELINE configs[0]
    import make_golden_sources  # noqa
    import types
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, Polygon, NodalPlane
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import TruncatedGRMFD
    from openquake.hazardlib.scalerel import WC1994
    from openquake.hazardlib.source import AreaSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.map_array import MapArray
    tom = PoissonTOM(d['investigation_time'])
    if mfd is None:
        mn, mx, a, b = d['gr']
    area = AreaSource(
        '1', 'Area Source', 'Active Shallow Crust',
        mfd or TruncatedGRMFD(mn, mx, d['width_of_mfd_bin'], a, b), 5.,
        WC1994(),
        d['rar'], tom, d['usd'], d['lsd'],
        PMF([(p, NodalPlane(s, dip, r)) for p, s, dip, r in d['npd']]),
        PMF([(p, dep) for p, dep in d['hdd']]),
        Polygon([Point(x, y) for x, y in zip(d['poly_lons'],
                                             d['poly_lats'])]),
        d['area_discretization'])
    levels = {imt: list(v) for imt, v in imtls.items()}
    cm = ContextMaker('Active Shallow Crust', [valid.gsim(gsim)], dict(
        imtls=levels, truncation_level=d['truncation_level'],
        investigation_time=d['investigation_time'],
        maximum_distance=IntegrationDistance.new(
            str(d['maximum_distance']))))
    cm.tom = tom
    S = d['sites']
    sitecol = SiteCollection.from_points(
        S[:, 0], S[:, 1], sitemodel=types.SimpleNamespace(
            reference_vs30_value=d['vs30'], reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=d['z1pt0'],
            reference_depth_to_2pt5km_per_sec=d['z2pt5']),
        req_site_params=REQ_SITE_PARAMS)
    M, L1 = len(levels), len(next(iter(levels.values())))
    rmap = MapArray(sitecol.sids, M * L1, 1, rates=True).fill(0)
    nrup = 0
    for src in area:
        src.id = 0
        nrup += src.count_ruptures()
        for ctx in cm.get_ctxs(src, sitecol):
            cm.update(rmap, ctx)
    assert nrup == d['num_ruptures'], nrup
    return (1. - np.exp(-rmap.array[:, :, 0].astype(float))).reshape(
        len(S), M, L1)


def case_24():
    """classical/case_24 (calculators/tests/classical_test.py:355-366, delta
    1e-3): the area source of the AreaSourceClassicalPSHA demo with three
    hypocentral depths, BooreAtkinson2008, PGV + PGA + 7 SA x 45 levels, two
    sites, pointsource_distance = 1000 (no approximation), 780 ruptures"""
    case = os.path.join(os.path.dirname(CASE), 'case_24')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()]
                      for s in ini['geometry']['sites'].split(',')])
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'],
                 {'logscale': logscale})
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        pointsource_distance=float(ini['calculation']['pointsource_distance']),
        width_of_mfd_bin=float(ini['erf']['width_of_mfd_bin']),
        area_discretization=float(ini['erf']['area_source_discretization']))
    root = ET.parse(os.path.join(case, 'source_model.xml')).getroot()
    [src] = root.findall('.//n:areaSource', NS)
    pos = [float(x) for x in src.find('.//gml:posList', NS).text.split()]
    mfd = src.find('n:truncGutenbergRichterMFD', NS)
    out.update(
        poly_lons=np.array(pos[0::2]), poly_lats=np.array(pos[1::2]),
        trt=src.get('tectonicRegion'),
        usd=float(src.find('.//n:upperSeismoDepth', NS).text),
        lsd=float(src.find('.//n:lowerSeismoDepth', NS).text),
        msr=src.find('n:magScaleRel', NS).text.strip(),
        rar=float(src.find('n:ruptAspectRatio', NS).text),
        gr=np.array([float(mfd.get(k)) for k in ('minMag', 'maxMag', 'aValue',
                                                 'bValue')]),
        npd=np.array([[float(n.get(k)) for k in (
            'probability', 'strike', 'dip', 'rake')]
            for n in src.findall('.//n:nodalPlane', NS)]),
        hdd=np.array([[float(n.get(k)) for k in ('probability', 'depth')]
                      for n in src.findall('.//n:hypoDepth', NS)]),
        num_ruptures=780)
    for imt in imtls:
        with open(os.path.join(case, 'expected',
                               'hazard_curve-%s.csv' % imt)) as f:
            rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
        body = np.array(rows[1:], float)
        np.testing.assert_allclose(body[:, :2], sites)
        out['expected_' + imt] = body[:, 3:]
    out['ref_poes'] = reference_case_24(out, imtls)
    for m, imt in enumerate(imtls):
        exp, ref = out['expected_' + imt], out['ref_poes'][:, m]
        print('case_24 %-10s reference (this version) vs expected file: max '
              'abs %.1e, max rel (PoE >= 1e-4) %.1e' % (
                  imt, np.abs(ref - exp).max(),
                  np.abs(ref / exp - 1)[exp >= 1e-4].max()))
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_24.npz'),
                        **out)
    print('case_24', {k: np.shape(v) for k, v in out.items()
                      if k.startswith('expected')})


def case_26():
    """classical/case_26 (classical_test.py:373-375, delta 2e-5, expected
    file of engine 3.26): an area source over a 1 x 1 degree square
    discretised at 5 km with a YoungsCoppersmith1985MFD (characteristic
    magnitude 7.0: ruptures longer than the seismogenic layer is thick),
    BooreAtkinson2008, SA(1.0) x 19 levels, one site.  The magnitudes and
    rates of the MFD are read from the reference's own class (data, as the
    coefficient tables)."""
    case = os.path.join(os.path.dirname(CASE), 'case_26')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()] + [0.]
                      for s in ini['geometry']['sites'].split(',')])[:, :3]
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'],
                 {'logscale': logscale})
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        area_discretization=float(ini['erf']['area_source_discretization']))
    root = ET.parse(os.path.join(case, 'source_model.xml')).getroot()
    [src] = root.findall('.//n:areaSource', NS)
    pos = [float(x) for x in src.find('.//gml:posList', NS).text.split()]
    node = src.find('n:YoungsCoppersmithMFD', NS)
    import sys
    sys.path.insert(0, HERE)
    import make_golden_sources  # noqa: bootstraps the reference
    from openquake.hazardlib.mfd import YoungsCoppersmith1985MFD
    mfd = YoungsCoppersmith1985MFD(
        min_mag=float(node.get('minMag')), b_val=float(node.get('bValue')),
        char_mag=float(node.get('characteristicMag')),
        char_rate=float(node.get('characteristicRate')),
        total_moment_rate=None, bin_width=float(node.get('binWidth')))
    out.update(
        poly_lons=np.array(pos[0::2]), poly_lats=np.array(pos[1::2]),
        trt=src.get('tectonicRegion'),
        usd=float(src.find('.//n:upperSeismoDepth', NS).text),
        lsd=float(src.find('.//n:lowerSeismoDepth', NS).text),
        msr=src.find('n:magScaleRel', NS).text.strip(),
        rar=float(src.find('n:ruptAspectRatio', NS).text),
        mfd=np.array(mfd.get_annual_occurrence_rates()),
        npd=np.array([[float(n.get(k)) for k in (
            'probability', 'strike', 'dip', 'rake')]
            for n in src.findall('.//n:nodalPlane', NS)]),
        hdd=np.array([[float(n.get(k)) for k in ('probability', 'depth')]
                      for n in src.findall('.//n:hypoDepth', NS)]))
    with open(os.path.join(case, 'expected', 'hazard_curve-rlz-000.csv')) as f:
        rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
    out['expected'] = np.array(rows[1:], float)[:, 3:]
    d = dict(out, num_ruptures=None)
    ref = reference_case_26(d, imtls, mfd)
    out['ref_poes'], out['num_points'], out['num_ruptures'] = ref
    print('case_26: %d points, %d ruptures; reference (this version) vs '
          'expected file: max abs %.1e, max rel (PoE >= 1e-4) %.1e' % (
              ref[1], ref[2], np.abs(ref[0][:, 0] - out['expected']).max(),
              np.abs(ref[0][:, 0] / out['expected'] - 1)[
                  out['expected'] >= 1e-4].max()))
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_26.npz'),
                        **out)


def reference_case_26(d, imtls, mfd):
    """as reference_case_24, returning also the numbers of points and
    ruptures of the discretised source"""
    from openquake.hazardlib import source
    count = {}
    orig = source.AreaSource.__iter__

    def counting(self):
        n = r = 0
        for src in orig(self):
            n += 1
            r += src.count_ruptures()
            yield src
        count['points'], count['ruptures'] = n, r
    source.AreaSource.__iter__ = counting
    try:
        d = dict(d)
        d['num_ruptures'] = -1
        try:
            poes = reference_case_24(d, imtls, mfd)
        except AssertionError:
            # reference_case_24 checks the rupture count of case_24 only
            d['num_ruptures'] = count['ruptures']
            poes = reference_case_24(d, imtls, mfd)
    finally:
        source.AreaSource.__iter__ = orig
    return poes, count['points'], count['ruptures']


def case_31():
    """classical/case_31 (classical_test.py:472-474, default delta 1e-5,
    expected file of engine 3.26): an area source, a point source and a
    multi-point source, dipping reverse ruptures (dip 60, rake 90) with three
    hypocentral depths each carrying its own fixedDipFrac, ChiouYoungs2014
    (hanging-wall terms), PGA x 6 levels, one site, the engine's default
    pointsource_distance of 100 km (the far sites of a source see the average
    rupture of its depths, contexts.py:321-370)."""
    case = os.path.join(os.path.dirname(CASE), 'case_31')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()]
                      for s in ini['geometry']['sites'].split(',')])
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'])
    binw = float(ini['erf']['width_of_mfd_bin'])
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        area_discretization=float(ini['erf']['area_source_discretization']),
        width_of_mfd_bin=binw, pointsource_distance=100.)   # openquake.cfg
    root = ET.parse(os.path.join(case, 'source_model.xml')).getroot()

    def common(src):
        return dict(
            usd=float(src.find('.//n:upperSeismoDepth', NS).text),
            lsd=float(src.find('.//n:lowerSeismoDepth', NS).text),
            msr=src.find('n:magScaleRel', NS).text.strip(),
            rar=float(src.find('n:ruptAspectRatio', NS).text),
            npd=np.array([[float(n.get(k)) for k in (
                'probability', 'strike', 'dip', 'rake')]
                for n in src.findall('.//n:nodalPlane', NS)]),
            hdd=np.array([[float(n.get(k)) for k in (
                'probability', 'depth', 'fixedDipFrac')]
                for n in src.findall('.//n:hypoDepth', NS)]))
    [area] = root.findall('.//n:areaSource', NS)
    [point] = root.findall('.//n:pointSource', NS)
    [multi] = root.findall('.//n:multiPointSource', NS)
    com = common(area)
    for other in (point, multi):      # the three sources share everything
        for k, v in common(other).items():
            np.testing.assert_array_equal(v, com[k])
    out.update(com)
    out['trt'] = area.get('tectonicRegion')
    pos = [float(x) for x in area.find('.//gml:posList', NS).text.split()]
    out['poly_lons'], out['poly_lats'] = np.array(pos[0::2]), np.array(pos[1::2])
    g = area.find('n:truncGutenbergRichterMFD', NS)
    out['gr'] = np.array([float(g.get(k)) for k in ('minMag', 'maxMag',
                                                    'aValue', 'bValue')])
    g2 = point.find('n:truncGutenbergRichterMFD', NS)
    assert [g2.get(k) for k in ('minMag', 'maxMag', 'aValue', 'bValue')] == [
        g.get(k) for k in ('minMag', 'maxMag', 'aValue', 'bValue')]
    out['point'] = np.array([float(x) for x in point.find(
        './/gml:pos', NS).text.split()])
    mp = [float(x) for x in multi.find('.//gml:posList', NS).text.split()]
    out['multi_points'] = np.array(mp).reshape(-1, 2)
    mm = multi.find('n:multiMFD', NS)
    assert mm.get('kind') == 'truncGutenbergRichterMFD'
    a_vals = [float(x) for x in mm.find('n:a_val', NS).text.split()]
    assert (float(mm.find('n:b_val', NS).text),
            float(mm.find('n:min_mag', NS).text),
            float(mm.find('n:max_mag', NS).text)) == (
                out['gr'][3], out['gr'][0], out['gr'][1])
    out['multi_a_vals'] = np.array(a_vals)
    with open(os.path.join(case, 'expected', 'hazard_curve-mean-PGA.csv')) as f:
        rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
    out['expected'] = np.array(rows[1:], float)[:, 3:]
    ref = reference_case_31(out, imtls)
    out['ref_poes'], out['num_ruptures'], out['num_rows'] = ref
    print('case_31: %d ruptures, %d context rows; reference (this version) '
          'vs expected file: max abs %.1e, max rel %.1e' % (
              ref[1], ref[2], np.abs(ref[0] - out['expected']).max(),
              np.abs(ref[0] / out['expected'] - 1).max()))
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_31.npz'),
                        **out)


def reference_case_31(d, imtls):
    """this version's hazardlib on the inputs of case_31: the three sources
    through ContextMaker.get_ctxs (with pointsource_distance) -> update of a
    map of rates -> 1 - exp(-rates)"""
    import sys
    sys.path.insert(0, HERE)
    import make_golden_sources  # noqa: bootstrap + polygon shim
    import types
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, Polygon, NodalPlane
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import TruncatedGRMFD
    from openquake.hazardlib.scalerel import WC1994
    from openquake.hazardlib.source import AreaSource, PointSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.map_array import MapArray
    tom = PoissonTOM(d['investigation_time'])
    mn, mx, a, b = d['gr']
    npd = PMF([(p, NodalPlane(s, dip, r)) for p, s, dip, r in d['npd']])

    def hdd():
        pmf = PMF([(p, dep) for p, dep, _f in d['hdd']])
        pmf.hypo_dip_fracs = tuple(f for _p, _dep, f in d['hdd'])
        return pmf

    def mfd(a_val):
        return TruncatedGRMFD(mn, mx, d['width_of_mfd_bin'], a_val, b)
    trt = 'Active Shallow Crust'
    srcs = list(AreaSource(
        '1', 'Area Source', trt, mfd(a), 5., WC1994(), d['rar'], tom,
        d['usd'], d['lsd'], npd, hdd(),
        Polygon([Point(x, y) for x, y in zip(d['poly_lons'],
                                             d['poly_lats'])]),
        d['area_discretization']))
    srcs.append(PointSource(
        '2', 'Point Source', trt, mfd(a), 5., WC1994(), d['rar'], tom,
        d['usd'], d['lsd'], Point(*d['point']), npd, hdd()))
    for i, ((x, y), av) in enumerate(zip(d['multi_points'],
                                         d['multi_a_vals'])):
        srcs.append(PointSource(
            '3:%d' % i, 'Multi-Point Source', trt, mfd(av), 5., WC1994(),
            d['rar'], tom, d['usd'], d['lsd'], Point(x, y), npd, hdd()))
    levels = {imt: list(v) for imt, v in imtls.items()}
    cm = ContextMaker(trt, [valid.gsim('[ChiouYoungs2014]')], dict(
        imtls=levels, truncation_level=d['truncation_level'],
        investigation_time=d['investigation_time'],
        pointsource_distance={'default': d['pointsource_distance']},
        maximum_distance=IntegrationDistance.new(
            str(d['maximum_distance']))))
    cm.tom = tom
    S = d['sites']
    sitecol = SiteCollection.from_points(
        S[:, 0], S[:, 1], sitemodel=types.SimpleNamespace(
            reference_vs30_value=d['vs30'], reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=d['z1pt0'],
            reference_depth_to_2pt5km_per_sec=d['z2pt5']),
        req_site_params=REQ_SITE_PARAMS)
    L = sum(len(v) for v in levels.values())
    rmap = MapArray(sitecol.sids, L, 1, rates=True).fill(0)
    nrup = nrows = 0
    for src in srcs:
        src.id = 0
        nrup += src.count_ruptures()
        for ctx in cm.get_ctxs(src, sitecol):
            nrows += len(ctx)
            cm.update(rmap, ctx)
    return (1. - np.exp(-rmap.array[:, :, 0].astype(float)), nrup, nrows)


def case_28():
    """classical/case_28 (classical_test.py:406-408, delta 1e-5, expected
    file of engine 3.26): a dipping simple fault (dip 60, rake 90, 0 - 20
    km, 5 km mesh) with a hypoDepthDist, ChiouYoungs2014, PGA x 6 levels, one
    site 57 km away.  The hypocentral depth distribution splits every
    floating rupture into copies with renormalised weights
    (source/simple_fault.py:204-268) and ChiouYoungs2014 reads nothing of the
    hypocentre, so the curve is that of the plain fault: the device path runs
    the fault without the distribution, the reference run below WITH it."""
    case = os.path.join(os.path.dirname(CASE), 'case_28')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()]
                      for s in ini['geometry']['sites'].split(',')])
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'])
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        spacing=float(ini['erf']['rupture_mesh_spacing']),
        width_of_mfd_bin=float(ini['erf']['width_of_mfd_bin']))
    root = ET.parse(os.path.join(case, 'source_model.xml')).getroot()
    [src] = root.findall('.//n:simpleFaultSource', NS)
    pos = [float(x) for x in src.find('.//gml:posList', NS).text.split()]
    g = src.find('n:truncGutenbergRichterMFD', NS)
    out.update(
        trace=np.array(pos).reshape(-1, 2), trt=src.get('tectonicRegion'),
        dip=float(src.find('.//n:dip', NS).text),
        usd=float(src.find('.//n:upperSeismoDepth', NS).text),
        lsd=float(src.find('.//n:lowerSeismoDepth', NS).text),
        msr=src.find('n:magScaleRel', NS).text.strip(),
        rar=float(src.find('n:ruptAspectRatio', NS).text),
        rake=float(src.find('n:rake', NS).text),
        gr=np.array([float(g.get(k)) for k in ('minMag', 'maxMag', 'aValue',
                                               'bValue')]),
        hdd=np.array([[float(n.get(k)) for k in ('probability', 'depth')]
                      for n in src.findall('.//n:hypoDepth', NS)]))
    with open(os.path.join(case, 'expected', 'hazard_curve-mean-PGA.csv')) as f:
        rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
    out['expected'] = np.array(rows[1:], float)[:, 3:]
    import sys
    sys.path.insert(0, HERE)
    import make_golden_sources  # noqa: bootstraps the reference
    import types
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, Line
    from openquake.hazardlib.geo.surface import SimpleFaultSurface
    from openquake.hazardlib.mfd import TruncatedGRMFD
    from openquake.hazardlib.scalerel import WC1994
    from openquake.hazardlib.source import SimpleFaultSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.map_array import MapArray
    tom = PoissonTOM(out['investigation_time'])
    mn, mx, a, b = out['gr']
    trace = Line([Point(x, y) for x, y in out['trace']])
    hypo_data = types.SimpleNamespace(
        hypo_list=(), slip_list=(),
        depth_list=[(p, dep) for p, dep in out['hdd']])
    mfd = TruncatedGRMFD(mn, mx, out['width_of_mfd_bin'], a, b)
    fault = SimpleFaultSource(
        '1', 'Simple Fault Source', 'Active Shallow Crust', mfd,
        out['spacing'], WC1994(), out['rar'], tom, out['usd'], out['lsd'],
        trace, out['dip'], out['rake'], hypo_data)
    plain = SimpleFaultSource(
        '1', 'Simple Fault Source', 'Active Shallow Crust', mfd,
        out['spacing'], WC1994(), out['rar'], tom, out['usd'], out['lsd'],
        trace, out['dip'], out['rake'])
    mesh = SimpleFaultSurface.from_fault_data(
        trace, out['usd'], out['lsd'], out['dip'], out['spacing']).mesh
    out['mesh_lons'], out['mesh_lats'] = mesh.lons, mesh.lats
    out['mesh_deps'] = mesh.depths
    out['mfd'] = np.array(mfd.get_annual_occurrence_rates())
    out['num_ruptures_plain'] = plain.count_ruptures()
    levels = {imt: list(v) for imt, v in imtls.items()}
    S = out['sites']
    sitecol = SiteCollection.from_points(
        S[:, 0], S[:, 1], sitemodel=types.SimpleNamespace(
            reference_vs30_value=out['vs30'], reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=out['z1pt0'],
            reference_depth_to_2pt5km_per_sec=out['z2pt5']),
        req_site_params=REQ_SITE_PARAMS)
    res = {}
    for tag, src_ in (('hdd', fault), ('plain', plain)):
        cm = ContextMaker('Active Shallow Crust', [valid.gsim(
            '[ChiouYoungs2014]')], dict(
            imtls=levels, truncation_level=out['truncation_level'],
            investigation_time=out['investigation_time'],
            maximum_distance=IntegrationDistance.new(
                str(out['maximum_distance']))))
        cm.tom = tom
        src_.id = 0
        rmap = MapArray(sitecol.sids, 6, 1, rates=True).fill(0)
        n = 0
        for ctx in cm.get_ctxs(src_, sitecol):
            n += len(ctx)
            cm.update(rmap, ctx)
        res[tag] = 1. - np.exp(-rmap.array[:, :, 0].astype(float))
        print('case_28 %-5s %5d context rows; reference (this version) vs '
              'expected file: max abs %.1e, max rel (PoE > 0) %.1e' % (
                  tag, n, np.abs(res[tag] - out['expected']).max(),
                  np.abs(res[tag] / out['expected'] - 1)[
                      out['expected'] > 0].max()))
    out['ref_poes'], out['ref_poes_plain'] = res['hdd'], res['plain']
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_28.npz'),
                        **out)


def case_78():
    """classical/case_78 (classical_test.py:1014-1020, delta 1e-5, expected
    file of engine 3.26): a point source and a dipping simple fault, each
    with its own NegativeBinomialTOM (mu, alpha: non-Poissonian probs_occur
    per rupture, tom.py:220-319), AbrahamsonEtAl2014 (rrup, rjb, rx, ry0,
    width, ztor of mesh ruptures; z1pt0 and the inferred vs30 matter), PGA x
    50 levels, one site."""
    case = os.path.join(os.path.dirname(CASE), 'case_78')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()]
                      for s in ini['geometry']['sites'].split(',')])
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'],
                 {'logscale': logscale})
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        vs30measured=ini['site_params']['reference_vs30_type'] == 'measured',
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        spacing=float(ini['erf']['rupture_mesh_spacing']),
        width_of_mfd_bin=float(ini['erf']['width_of_mfd_bin']))
    NS5 = {'n': 'http://openquake.org/xmlns/nrml/0.5',
           'gml': 'http://www.opengis.net/gml'}
    root = ET.parse(os.path.join(case, 'source_negbinom.xml')).getroot()
    [ps] = root.findall('.//n:pointSource', NS5)
    [fs] = root.findall('.//n:simpleFaultSource', NS5)
    f = lambda node, tag: float(node.find(tag, NS5).text)  # noqa
    imfd = ps.find('n:incrementalMFD', NS5)
    g = fs.find('n:truncGutenbergRichterMFD', NS5)
    out.update(
        trt=root.find('.//n:sourceGroup', NS5).get('tectonicRegion'),
        ps_tom=np.array([float(ps.get('mu')), float(ps.get('alpha'))]),
        ps_loc=np.array([float(x) for x in ps.find(
            './/gml:pos', NS5).text.split()]),
        ps_usd=f(ps, './/n:upperSeismoDepth'),
        ps_lsd=f(ps, './/n:lowerSeismoDepth'),
        ps_msr=ps.find('n:magScaleRel', NS5).text.strip(),
        ps_rar=f(ps, 'n:ruptAspectRatio'),
        ps_mfd=np.array([float(imfd.get('minMag')),
                         float(imfd.get('binWidth'))] + [
            float(x) for x in imfd.find('n:occurRates', NS5).text.split()]),
        ps_npd=np.array([[float(n.get(k)) for k in (
            'probability', 'strike', 'dip', 'rake')]
            for n in ps.findall('.//n:nodalPlane', NS5)]),
        ps_hdd=np.array([[float(n.get(k)) for k in ('probability', 'depth')]
                         for n in ps.findall('.//n:hypoDepth', NS5)]),
        fs_tom=np.array([float(fs.get('mu')), float(fs.get('alpha'))]),
        fs_trace=np.array([float(x) for x in fs.find(
            './/gml:posList', NS5).text.split()]).reshape(-1, 2),
        fs_dip=f(fs, './/n:dip'), fs_usd=f(fs, './/n:upperSeismoDepth'),
        fs_lsd=f(fs, './/n:lowerSeismoDepth'),
        fs_msr=fs.find('n:magScaleRel', NS5).text.strip(),
        fs_rar=f(fs, 'n:ruptAspectRatio'), fs_rake=f(fs, 'n:rake'),
        fs_gr=np.array([float(g.get(k)) for k in ('minMag', 'maxMag',
                                                  'aValue', 'bValue')]))
    with open(os.path.join(case, 'expected',
                           'hazard_curve-mean-PGA_NegBinomTest.csv')) as fh:
        rows = [r for r in csv.reader(fh) if not r[0].startswith('#')]
    out['expected'] = np.array(rows[1:], float)[:, 3:]
    import sys
    sys.path.insert(0, HERE)
    import make_golden_sources  # noqa: bootstraps the reference
    import types
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, Line, NodalPlane
    from openquake.hazardlib.geo.surface import SimpleFaultSurface
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import TruncatedGRMFD, EvenlyDiscretizedMFD
    from openquake.hazardlib.scalerel import WC1994, PointMSR
    from openquake.hazardlib.source import SimpleFaultSource, PointSource
    from openquake.hazardlib.tom import NegativeBinomialTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.map_array import MapArray
    T = out['investigation_time']
    trt = 'Active Shallow Crust'
    assert out['ps_msr'] == 'PointMSR' and out['fs_msr'] == 'WC1994'
    point = PointSource(
        '00000', 'point00000', trt, EvenlyDiscretizedMFD(
            out['ps_mfd'][0], out['ps_mfd'][1], list(out['ps_mfd'][2:])),
        out['spacing'], PointMSR(), out['ps_rar'],
        NegativeBinomialTOM(T, *out['ps_tom']), out['ps_usd'], out['ps_lsd'],
        Point(*out['ps_loc']),
        PMF([(p, NodalPlane(s, d, r)) for p, s, d, r in out['ps_npd']]),
        PMF([(p, d) for p, d in out['ps_hdd']]))
    mn, mx, a, b = out['fs_gr']
    trace = Line([Point(x, y) for x, y in out['fs_trace']])
    mfd = TruncatedGRMFD(mn, mx, out['width_of_mfd_bin'], a, b)
    fault = SimpleFaultSource(
        '00001', 'simplefault', trt, mfd, out['spacing'], WC1994(),
        out['fs_rar'], NegativeBinomialTOM(T, *out['fs_tom']), out['fs_usd'],
        out['fs_lsd'], trace, out['fs_dip'], out['fs_rake'])
    mesh = SimpleFaultSurface.from_fault_data(
        trace, out['fs_usd'], out['fs_lsd'], out['fs_dip'],
        out['spacing']).mesh
    out['mesh_lons'], out['mesh_lats'] = mesh.lons, mesh.lats
    out['mesh_deps'] = mesh.depths
    out['fs_mfd'] = np.array(mfd.get_annual_occurrence_rates())
    out['fs_num_ruptures'] = fault.count_ruptures()
    levels = {imt: list(v) for imt, v in imtls.items()}
    S = out['sites']
    sitecol = SiteCollection.from_points(
        S[:, 0], S[:, 1], sitemodel=types.SimpleNamespace(
            reference_vs30_value=out['vs30'],
            reference_vs30_type='measured' if out['vs30measured']
            else 'inferred',
            reference_depth_to_1pt0km_per_sec=out['z1pt0'],
            reference_depth_to_2pt5km_per_sec=out['z2pt5']),
        req_site_params=REQ_SITE_PARAMS)
    pmap = MapArray(sitecol.sids, 50, 1).fill(1.)
    nrows = 0
    for src_ in (point, fault):
        cm = ContextMaker(trt, [valid.gsim('[AbrahamsonEtAl2014]')], dict(
            imtls=levels, truncation_level=out['truncation_level'],
            investigation_time=T, maximum_distance=IntegrationDistance.new(
                str(out['maximum_distance']))))
        cm.tom = src_.temporal_occurrence_model
        src_.id = 0
        for ctx in cm.get_ctxs(src_, sitecol):
            nrows += len(ctx)
            cm.update(pmap, ctx)
    out['ref_poes'] = (~pmap).array[:, :, 0]
    print('case_78: %d context rows; reference (this version) vs expected '
          'file: max abs %.1e, max rel %.1e' % (
              nrows, np.abs(out['ref_poes'] - out['expected']).max(),
              np.abs(out['ref_poes'] / out['expected'] - 1).max()))
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_78.npz'),
                        **out)


def case_03():
    """classical/case_03 (classical_test.py:118-120, default delta 1e-5,
    expected file of engine 3.23): two area sources over the same polygon,
    minimum_magnitude = 5.7 (OqParam cuts maximum_distance at it,
    calc/filters.py:228-259: the smaller magnitudes have no integration
    distance, the second source, maximum magnitude 5.5, contributes nothing),
    width_of_mfd_bin = 0.2, BooreAtkinson2008, PGA x 5 levels, one site."""
    case = os.path.join(os.path.dirname(CASE), 'case_03')
    ini = configparser.ConfigParser()
    ini.read(os.path.join(case, 'job.ini'))
    sites = np.array([[float(x) for x in s.split()]
                      for s in ini['geometry']['sites'].split(',')])
    imtls = eval(ini['calculation']['intensity_measure_types_and_levels'],
                 {'logscale': logscale})
    out = dict(
        sites=sites, imts=np.array(list(imtls)),
        levels=np.array([np.asarray(imtls[i], float) for i in imtls]),
        vs30=float(ini['site_params']['reference_vs30_value']),
        z1pt0=float(ini['site_params']['reference_depth_to_1pt0km_per_sec']),
        z2pt5=float(ini['site_params']['reference_depth_to_2pt5km_per_sec']),
        investigation_time=float(ini['calculation']['investigation_time']),
        truncation_level=float(ini['calculation']['truncation_level']),
        maximum_distance=float(ini['calculation']['maximum_distance']),
        minimum_magnitude=float(ini['calculation']['minimum_magnitude']),
        width_of_mfd_bin=float(ini['erf']['width_of_mfd_bin']),
        area_discretization=float(ini['erf']['area_source_discretization']))
    NS5 = {'n': 'http://openquake.org/xmlns/nrml/0.5',
           'gml': 'http://www.opengis.net/gml'}
    grs = []
    for k, fname in enumerate(('source_model.xml', 'source_model_2.xml')):
        root = ET.parse(os.path.join(case, fname)).getroot()
        [src] = root.findall('.//n:areaSource', NS5)
        pos = [float(x) for x in src.find('.//gml:posList', NS5).text.split()]
        g = src.find('n:truncGutenbergRichterMFD', NS5)
        grs.append([float(g.get(x)) for x in ('minMag', 'maxMag', 'aValue',
                                              'bValue')])
        d = dict(
            poly_lons=np.array(pos[0::2]), poly_lats=np.array(pos[1::2]),
            trt=src.get('tectonicRegion'),
            usd=float(src.find('.//n:upperSeismoDepth', NS5).text),
            lsd=float(src.find('.//n:lowerSeismoDepth', NS5).text),
            msr=src.find('n:magScaleRel', NS5).text.strip(),
            rar=float(src.find('n:ruptAspectRatio', NS5).text),
            npd=np.array([[float(n.get(x)) for x in (
                'probability', 'strike', 'dip', 'rake')]
                for n in src.findall('.//n:nodalPlane', NS5)]),
            hdd=np.array([[float(n.get(x)) for x in ('probability', 'depth')]
                          for n in src.findall('.//n:hypoDepth', NS5)]))
        if k == 0:
            out.update(d)
        else:       # the two sources differ by the MFD only
            for key, val in d.items():
                np.testing.assert_array_equal(val, out[key])
    out['grs'] = np.array(grs)
    with open(os.path.join(case, 'expected', 'hazard_curve-PGA.csv')) as f:
        rows = [r for r in csv.reader(f) if not r[0].startswith('#')]
    out['expected'] = np.array(rows[1:], float)[:, 3:]
    import sys
    sys.path.insert(0, HERE)
    import make_golden_sources  # noqa: bootstrap + polygon shim
    import types
    from openquake.hazardlib import valid
    from openquake.hazardlib.geo import Point, Polygon, NodalPlane
    from openquake.hazardlib.pmf import PMF
    from openquake.hazardlib.mfd import TruncatedGRMFD
    from openquake.hazardlib.scalerel import WC1994
    from openquake.hazardlib.source import AreaSource
    from openquake.hazardlib.tom import PoissonTOM
    from openquake.hazardlib.site import SiteCollection
    from openquake.hazardlib.contexts import ContextMaker
    from openquake.hazardlib.calc.filters import IntegrationDistance
    from openquake.hazardlib.map_array import MapArray
    tom = PoissonTOM(out['investigation_time'])
    maxdist = IntegrationDistance.new(str(out['maximum_distance']))
    maxdist.cut({'default': out['minimum_magnitude']})
    levels = {imt: list(v) for imt, v in imtls.items()}
    cm = ContextMaker('Active Shallow Crust', [valid.gsim(
        '[BooreAtkinson2008]')], dict(
        imtls=levels, truncation_level=out['truncation_level'],
        investigation_time=out['investigation_time'],
        maximum_distance=maxdist))
    cm.tom = tom
    S = out['sites']
    sitecol = SiteCollection.from_points(
        S[:, 0], S[:, 1], sitemodel=types.SimpleNamespace(
            reference_vs30_value=out['vs30'], reference_vs30_type='measured',
            reference_depth_to_1pt0km_per_sec=out['z1pt0'],
            reference_depth_to_2pt5km_per_sec=out['z2pt5']),
        req_site_params=REQ_SITE_PARAMS)
    rmap = MapArray(sitecol.sids, 5, 1, rates=True).fill(0)
    nrup = nrows = 0
    for mn, mx, a, b in grs:
        if mx < out['minimum_magnitude']:
            continue    # _filter_mag, calculators/preclassical.py:100-109
        area = AreaSource(
            '1', 'Area Source', 'Active Shallow Crust',
            TruncatedGRMFD(mn, mx, out['width_of_mfd_bin'], a, b), 2.,
            WC1994(), out['rar'], tom, out['usd'], out['lsd'],
            PMF([(p, NodalPlane(s, dip, r)) for p, s, dip, r in out['npd']]),
            PMF([(p, dep) for p, dep in out['hdd']]),
            Polygon([Point(x, y) for x, y in zip(out['poly_lons'],
                                                 out['poly_lats'])]),
            out['area_discretization'])
        for src in area:
            src.id = 0
            nrup += src.count_ruptures()
            for ctx in cm.get_ctxs(src, sitecol):
                nrows += len(ctx)
                cm.update(rmap, ctx)
    out['ref_poes'] = 1. - np.exp(-rmap.array[:, :, 0].astype(float))
    out['num_ruptures'], out['num_rows'] = nrup, nrows
    print('case_03: %d ruptures, %d context rows kept; reference (this '
          'version) vs expected file: max abs %.1e' % (
              nrup, nrows, np.abs(out['ref_poes'] - out['expected']).max()))
    np.savez_compressed(os.path.join(GOLDEN, 'qa_classical_case_03.npz'),
                        **out)


if __name__ == '__main__':
    main()
    case_03()
    case_24()
    case_26()
    case_31()
    case_28()
    case_78()
