"""
Known answers of the reference for the Joyner-Boore distance of rectangular
meshes (RectangularMesh.get_joyner_boore_distance needs shapely, absent from
this image, so the reference cannot be run for it here): the expected values
and tolerances of the reference's OWN tests
(openquake/hazardlib/tests/geo/mesh_test.py:351-447 and _mesh_test_data.py),
with their input meshes, as a fixture:

    python oracle/make_golden_mesh_rjb.py  ->  tests/golden/mesh_rjb.npz

Each case: mesh lons / lats (2D), site lons / lats, expected distances (km)
and the absolute tolerance the reference's test grants.  The oracle
(oracle/geo_np.py mesh_rjb) is asserted against them here.
"""
import os
import sys
import importlib.util
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import geo_np  # noqa

REF = os.environ.get('OQ_REFERENCE', '/root/reference')
spec = importlib.util.spec_from_file_location(
    '_mesh_test_data', os.path.join(
        REF, 'openquake/hazardlib/tests/geo/_mesh_test_data.py'))
data = importlib.util.module_from_spec(spec)
spec.loader.exec_module(data)


def gdist(lon1, lat1, lon2, lat2):
    # Point.distance for points at depth 0: geodetic distance
    return float(geo_np._arc_length(lon1, lat1, lon2, lat2))


def cases():
    out = []
    # test_simple (mesh_test.py:352-372)
    lons = np.array([np.arange(-1, 1.2, 0.2)] * 11)
    lats = lons.transpose() + 1
    out.append(('simple', lons, lats,
                [(0, 0.5), (1, 1), (0.6, -1), (-0.8, 2.1), (0.75, 2.3)],
                [0., 0., gdist(0.6, -1, 0.6, 0), gdist(-0.8, 2.1, -0.8, 2),
                 gdist(0.75, 2.3, 0.75, 2)],
                # assertAlmostEqual(places=7) for the first two
                [1e-7, 1e-7, 0.1, 0.02, 0.04]))
    # test_vertical_mesh (:374-385)
    out.append(('vertical', np.array([[0, 0.5, 1, 2], [0, 0.5, 1, 2]], float),
                np.zeros((2, 4)), [(0.5, 0), (0.5, 1), (0.5, 5)],
                [0., gdist(0.5, 1, 0.5, 0), gdist(0.5, 5, 0.5, 0)],
                [1., 1., 1.]))
    # test_mesh_of_two_points (:387-396): one row of three points
    out.append(('one_row', np.array([[0, 0.5, 1]], float), np.zeros((1, 3)),
                [(0.5, 1), (0.5, 0)], [gdist(0.5, 1, 0.5, 0), 0.],
                [.01, .01]))
    # test_mesh_of_one_point (:398-406)
    out.append(('one_point', np.array([[1.]]), np.array([[0.]]),
                [(1, 0), (0.5, 0)], [0., gdist(0.5, 0, 1, 0)], [0.2, 0.2]))
    # test3 / test4 / test5 (:419-429), delta = 0.02
    for k in ('TEST3', 'TEST4', 'TEST5'):
        lo, la, _dep = np.array(getattr(data, k + '_MESH')).transpose()
        out.append((k.lower(), lo.transpose(), la.transpose(),
                    [tuple(getattr(data, k + '_SITE'))],
                    [getattr(data, k + '_JB_DISTANCE')], [0.02]))
    # test_version (:431-447): either of two shapely / libgeos versions
    lons = np.array([[-121.3956, -121.41050474, -121.42542273, -121.44035399,
                      -121.45529855, -121.47025643]] * 2)
    lats = np.array([[36.8257, 36.85963772, 36.89357357, 36.92750756,
                      36.96143968, 36.99536993]] * 2)
    out.append(('version', lons, lats, [(-121.76, 37.23)],
                [(36.61260128 + 36.61389245) / 2], [0.004]))
    return out


if __name__ == '__main__':
    dic = {'names': []}
    for name, lons, lats, sites, exp, tol in cases():
        slon, slat = np.array(sites, float).T
        got = geo_np.mesh_rjb(lons, lats, slon, slat)
        err = np.abs(got - np.array(exp))
        print('%-10s mesh %s expected %s oracle %s' % (
            name, lons.shape, np.round(exp, 4), np.round(got, 4)))
        assert (err <= np.array(tol)).all(), (name, got, exp, tol)
        dic['names'].append(name)
        dic[name + '_lons'], dic[name + '_lats'] = lons, lats
        dic[name + '_slon'], dic[name + '_slat'] = slon, slat
        dic[name + '_expected'] = np.array(exp, float)
        dic[name + '_tol'] = np.array(tol, float)
    dic['names'] = np.array(dic['names'])
    np.savez_compressed(os.path.join(HERE, '..', 'tests', 'golden',
                                     'mesh_rjb.npz'), **dic)
