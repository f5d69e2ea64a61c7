"""
Golden fixture for Rx and Ry0 of mesh surfaces:
the reference's BaseSurface.get_rx_distance / get_ry0_distance
(openquake/hazardlib/geo/surface/base.py:319-384, :193-239) on top edges of
2 .. 40 points (straight, bent, across the date line), the mean strike from
RectangularMesh.get_mean_inclination_and_azimuth as SimpleFaultSurface
.get_strike does (geo/surface/simple_fault.py).

    python oracle/make_golden_mesh_rx.py   -> tests/golden/mesh_rx.npz
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.geo.mesh import Mesh, RectangularMesh  # noqa
from openquake.hazardlib.geo.surface.base import BaseSurface  # noqa
from openquake.hazardlib.geo import geodetic  # noqa
import geo_np  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


class Surface(BaseSurface):
    """the concrete minimum: a rectangular mesh of two rows (top and bottom
    edge), strike as in SimpleFaultSurface.get_strike"""
    def get_strike(self):
        return self.mesh.get_mean_inclination_and_azimuth()[1]

    def get_dip(self):
        return self.mesh.get_mean_inclination_and_azimuth()[0]

    def get_width(self):
        return self.mesh.get_mean_width()

    def get_area(self):
        raise NotImplementedError

    def get_bounding_box(self):
        raise NotImplementedError

    def get_middle_point(self):
        raise NotImplementedError

    def get_surface_boundaries(self):
        raise NotImplementedError

    def get_surface_boundaries_3d(self):
        raise NotImplementedError


def make_surface(rng, K, lon0, lat0, strike0, wiggle, dip, seg_km):
    lons, lats = [lon0], [lat0]
    az = strike0
    for i in range(K - 1):
        az = az + rng.uniform(-wiggle, wiggle)
        lo, la = geodetic.point_at(lons[-1], lats[-1], az,
                                   seg_km * rng.uniform(0.6, 1.4))
        lons.append(float(lo))
        lats.append(float(la))
    lons, lats = np.array(lons), np.array(lats)
    width = 12.
    hor = width * np.cos(np.radians(dip))
    blon, blat = geodetic.point_at(lons, lats, strike0 + 90., hor)
    top = np.full(K, 1.5)
    bot = top + width * np.sin(np.radians(dip))
    mesh = RectangularMesh(np.array([lons, blon]), np.array([lats, blat]),
                           np.array([top, bot]))
    return Surface(mesh)


def main():
    rng = np.random.default_rng(11)
    specs = [  # K, lon0, lat0, strike0, wiggle, dip, seg_km
        (2, 10., 45., 30., 0., 60., 20.),
        (3, 10., 45., 200., 15., 45., 10.),
        (4, -120., 35., 320., 25., 80., 8.),
        (9, 179.7, -20., 80., 20., 50., 7.),     # across the date line
        (40, 30., 60., 10., 12., 70., 3.),
        (17, -70., -33., 175., 30., 30., 5.),
        (2, 179.95, 5., 90., 0., 90., 30.),
    ]
    N = 500
    tl, ta, off, strikes = [], [], [0], []
    rx = np.zeros((len(specs), N))
    ry0 = np.zeros((len(specs), N))
    slon = np.zeros((len(specs), N))
    slat = np.zeros((len(specs), N))
    for s, spec in enumerate(specs):
        sfc = make_surface(rng, *spec)
        lons, lats = sfc.mesh.lons[0], sfc.mesh.lats[0]
        span = 0.6 + 0.012 * spec[0] * spec[-1]
        lo = lons[len(lons) // 2] + rng.uniform(-span, span, N)
        lo = np.where(lo > 180, lo - 360, np.where(lo < -180, lo + 360, lo))
        la = lats[len(lats) // 2] + rng.uniform(-span, span, N)
        # some sites exactly on the top edge vertices and on the trace
        lo[:2], la[:2] = lons[:2], lats[:2]
        lo[2], la[2] = lons[-1], lats[-1]
        sites = Mesh(lo, la)
        rx[s] = sfc.get_rx_distance(sites)
        ry0[s] = sfc.get_ry0_distance(sites)
        strike = sfc.get_strike()
        # pin the oracle restatement
        np.testing.assert_allclose(geo_np.rx_distance(lons, lats, lo, la),
                                   rx[s], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(
            geo_np.ry0_distance(lons, lats, strike, lo, la), ry0[s],
            rtol=1e-12, atol=1e-9)
        tl.append(lons)
        ta.append(lats)
        off.append(off[-1] + len(lons))
        strikes.append(strike)
        slon[s], slat[s] = lo, la
    np.savez_compressed(
        os.path.join(GOLDEN, 'mesh_rx.npz'), top_lon=np.concatenate(tl),
        top_lat=np.concatenate(ta), off=np.array(off),
        strike=np.array(strikes), site_lon=slon, site_lat=slat, rx=rx, ry0=ry0)
    print('mesh_rx.npz', rx.shape, np.abs(rx).max(), ry0.max(),
          (ry0 == 0).mean())


if __name__ == '__main__':
    main()
