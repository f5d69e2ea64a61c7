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


if __name__ == '__main__':
    main()
