"""
Golden fixture for the rupture distance of mesh surfaces (SURVEY.md section
8f row 2, first step): the reference's Mesh.get_min_distance
(openquake/hazardlib/geo/mesh.py:365-383) on meshes of various sizes, and the
multi-fault composition numpy.min over sections (contexts.py:969-971).

    python oracle/make_golden_mesh.py   -> tests/golden/mesh_rrup.npz
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.geo.mesh import Mesh  # noqa
import geo_np  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')


def main():
    rng = np.random.default_rng(8)
    sizes = [1, 7, 300, 1000, 513, 256, 257]
    lons, lats, deps, off = [], [], [], [0]
    for n in sizes:
        c = rng.uniform(-1, 1, 2) + (179.5, 10.)   # across the date line
        lo = c[0] + rng.uniform(-0.4, 0.4, n)
        lons.append(np.where(lo > 180, lo - 360, lo))
        lats.append(c[1] + rng.uniform(-0.4, 0.4, n))
        deps.append(rng.uniform(0, 25, n))
        off.append(off[-1] + n)
    N = 333
    slon = 179.5 + rng.uniform(-2, 2, N)
    slon = np.where(slon > 180, slon - 360, slon)
    slat = 10. + rng.uniform(-2, 2, N)
    sdep = np.where(rng.random(N) < 0.2, rng.uniform(-1, 2, N), 0.)
    sites = Mesh(slon, slat, sdep)
    out = np.zeros((len(sizes), N))
    for s in range(len(sizes)):
        mesh = Mesh(lons[s], lats[s], deps[s])
        out[s] = mesh.get_min_distance(sites)
        # pin the oracle restatement
        xyz = geo_np.spherical_to_cartesian(lons[s], lats[s], deps[s])
        sxyz = geo_np.spherical_to_cartesian(slon, slat, sdep)
        mine = np.sqrt(((xyz[:, None, :] - sxyz[None]) ** 2).sum(-1)).min(0)
        np.testing.assert_allclose(mine, out[s], rtol=1e-12, atol=1e-9)
    # multi-fault ruptures: lists of sections
    rups = [[0], [1, 2], [2, 3, 4, 5], [6, 0, 3]]
    rmin = np.array([out[r].min(axis=0) for r in rups])
    np.savez_compressed(
        os.path.join(GOLDEN, 'mesh_rrup.npz'), lon=np.concatenate(lons),
        lat=np.concatenate(lats), dep=np.concatenate(deps), off=np.array(off),
        site_lon=slon, site_lat=slat, site_dep=sdep, rrup=out,
        rup_off=np.cumsum([0] + [len(r) for r in rups]),
        rup_sec=np.concatenate(rups), rup_rrup=rmin)
    print('mesh_rrup.npz', out.shape, out.min(), out.max())


if __name__ == '__main__':
    main()
