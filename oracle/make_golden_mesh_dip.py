"""
Golden fixture for the mean dip, strike and width of rectangular meshes
(SURVEY.md section 8f row 2): the reference's
RectangularMesh.get_mean_inclination_and_azimuth (geo/mesh.py:728-783) and
get_mean_width (:861-885) on meshes of 2x2 .. 12x40 points: dipping,
vertical, overturned (inclination > 90 before the flip), flat at depth 0,
pointing west, across the date line.

    python oracle/make_golden_mesh_dip.py   -> tests/golden/mesh_dip.npz
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.geo.mesh import RectangularMesh  # noqa
from openquake.hazardlib.geo import geodetic  # noqa
import geo_np  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


def make_mesh(rng, rows, cols, lon0, lat0, strike, wiggle, dip, spacing,
              flat=False):
    lons, lats = [lon0], [lat0]
    az = strike
    for i in range(cols - 1):
        az = az + rng.uniform(-wiggle, wiggle)
        lo, la = geodetic.point_at(lons[-1], lats[-1], az,
                                   spacing * rng.uniform(0.8, 1.2))
        lons.append(float(lo))
        lats.append(float(la))
    lons, lats = np.array(lons), np.array(lats)
    LO, LA, DE = [lons], [lats], [np.full(cols, 0. if flat else 2.)]
    for r in range(1, rows):
        hor = spacing * np.cos(np.radians(dip))
        lo, la = geodetic.point_at(LO[-1], LA[-1], strike + 90., hor)
        LO.append(lo)
        LA.append(la)
        DE.append(DE[-1] + (0. if flat else spacing * np.sin(np.radians(dip))))
    return np.array(LO), np.array(LA), np.array(DE)


def main():
    rng = np.random.default_rng(21)
    specs = [  # rows, cols, lon0, lat0, strike, wiggle, dip, spacing
        (2, 2, 10., 45., 30., 0., 60., 5.),
        (2, 9, 10., 45., 200., 10., 45., 4.),      # pointing south-west
        (5, 7, -120., 35., 320., 20., 90., 3.),    # vertical
        (4, 12, 179.8, -20., 80., 15., 30., 6.),   # across the date line
        (12, 40, 30., 60., 275., 8., 75., 2.),     # pointing west
        (3, 5, 100., 0.5, 0., 5., 120., 4.),       # overturned: flipped
        (3, 6, -60., -40., 140., 10., 0., 5., True),   # flat on the surface
    ]
    lon, lat, dep, shape, off, res = [], [], [], [], [0], []
    for spec in specs:
        LO, LA, DE = make_mesh(rng, *spec)
        m = RectangularMesh(LO, LA, DE)
        incl, azim = m.get_mean_inclination_and_azimuth()
        width = m.get_mean_width()
        mine = geo_np.mean_inclination_and_azimuth(LO, LA, DE)
        np.testing.assert_allclose(mine, (incl, azim), rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(geo_np.mean_width(LO, LA, DE), width,
                                   rtol=1e-12)
        lon.append(LO.ravel())
        lat.append(LA.ravel())
        dep.append(DE.ravel())
        shape.append(LO.shape)
        off.append(off[-1] + LO.size)
        res.append((incl, azim, width))
        print(spec[:2], incl, azim, width)
    np.savez_compressed(
        os.path.join(GOLDEN, 'mesh_dip.npz'), lon=np.concatenate(lon),
        lat=np.concatenate(lat), dep=np.concatenate(dep),
        shape=np.array(shape), off=np.array(off), res=np.array(res))


if __name__ == '__main__':
    main()
