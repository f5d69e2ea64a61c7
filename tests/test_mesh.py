"""Rupture distance of mesh surfaces (first
step): oqg_mesh_min_distance / oqg_section_min against the reference's
Mesh.get_min_distance (oracle/make_golden_mesh.py).  Tolerance: rtol 1e-10,
atol 1e-8 km (the atol of the planar distances): the device expands the
squared distance around the first mesh point, which costs <= 1e-10 km^2 of
cancellation, see k_mesh_min_distance."""
import os
import numpy as np
import pytest

import geo_np
from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'mesh_rrup.npz'))
OFF = D['off']


def test_oracle_matches_reference():
    sxyz = geo_np.spherical_to_cartesian(D['site_lon'], D['site_lat'],
                                         D['site_dep'])
    for s in range(len(OFF) - 1):
        sl = slice(OFF[s], OFF[s + 1])
        xyz = geo_np.spherical_to_cartesian(D['lon'][sl], D['lat'][sl],
                                            D['dep'][sl])
        mine = np.sqrt(((xyz[:, None, :] - sxyz[None]) ** 2).sum(-1)).min(0)
        np.testing.assert_allclose(mine, D['rrup'][s], rtol=1e-12, atol=1e-9)


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


@pytest.mark.gpu
def test_mesh_min_distance_golden(oq):
    meshes = [oq.Mesh(D['lon'][OFF[s]:OFF[s + 1]], D['lat'][OFF[s]:OFF[s + 1]],
                      D['dep'][OFF[s]:OFF[s + 1]])
              for s in range(len(OFF) - 1)]
    sites = oq.Mesh(D['site_lon'], D['site_lat'], D['site_dep'])
    got = oq.mesh_min_distance(meshes, sites)
    np.testing.assert_allclose(got, D['rrup'], rtol=1e-10, atol=1e-8)
    one = meshes[3].get_min_distance(sites)
    np.testing.assert_array_equal(one, got[3])
    # a site collection works as the target too
    sc = oq.SiteCollection.from_points(D['site_lon'], D['site_lat'],
                                       D['site_dep'], dict(vs30=760.))
    np.testing.assert_array_equal(oq.mesh_min_distance(meshes[:2], sc),
                                  got[:2])
    with pytest.raises(ValueError):
        oq.mesh_min_distance([], sites)


@pytest.mark.gpu
def test_section_min_golden(oq):
    roff = D['rup_off']
    rups = [D['rup_sec'][roff[u]:roff[u + 1]] for u in range(len(roff) - 1)]
    got = oq.section_min(D['rrup'], rups)
    np.testing.assert_array_equal(got, D['rup_rrup'])
    with pytest.raises(ValueError):
        oq.section_min(D['rrup'], [[99]])


# ---- Rx / Ry0 (oracle/make_golden_mesh_rx.py) ------------------------------
X = np.load(os.path.join(GOLDEN, 'mesh_rx.npz'))
XOFF = X['off']


def test_oracle_rx_ry0_matches_reference():
    for s in range(len(XOFF) - 1):
        sl = slice(XOFF[s], XOFF[s + 1])
        rx = geo_np.rx_distance(X['top_lon'][sl], X['top_lat'][sl],
                                X['site_lon'][s], X['site_lat'][s])
        np.testing.assert_allclose(rx, X['rx'][s], rtol=1e-12, atol=1e-9)
        ry0 = geo_np.ry0_distance(X['top_lon'][sl], X['top_lat'][sl],
                                  X['strike'][s], X['site_lon'][s],
                                  X['site_lat'][s])
        np.testing.assert_allclose(ry0, X['ry0'][s], rtol=1e-12, atol=1e-9)
    # the fixture exercises both signs, the zero band of ry0 and all the
    # branches (arc, semi-arc, segment)
    assert (X['rx'] > 0).any() and (X['rx'] < 0).any()
    assert (X['ry0'] == 0).any() and (X['ry0'] > 0).any()


@pytest.mark.gpu
def test_mesh_rx_ry0_golden(oq):
    """atol 1e-8 km: the device evaluates the reference's spherical
    triangles with unit vectors (k_mesh_rx_ry0), the same quantities up to
    rounding"""
    S = len(XOFF) - 1
    edges = [(X['top_lon'][XOFF[s]:XOFF[s + 1]],
              X['top_lat'][XOFF[s]:XOFF[s + 1]]) for s in range(S)]
    # each surface has its own sites in the fixture
    for s in range(S):
        sites = oq.Mesh(X['site_lon'][s], X['site_lat'][s])
        rx, ry0 = oq.mesh_rx_ry0(edges, X['strike'], sites)
        np.testing.assert_allclose(rx[s], X['rx'][s], rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(ry0[s], X['ry0'][s], rtol=1e-10, atol=1e-8)
        # the other surfaces against the oracle on these sites
        for t in range(S):
            if t != s:
                np.testing.assert_allclose(
                    rx[t], geo_np.rx_distance(*edges[t], X['site_lon'][s],
                                              X['site_lat'][s]),
                    rtol=1e-10, atol=1e-8)
                np.testing.assert_allclose(
                    ry0[t], geo_np.ry0_distance(
                        *edges[t], X['strike'][t], X['site_lon'][s],
                        X['site_lat'][s]), rtol=1e-10, atol=1e-8)
    sites = oq.Mesh(X['site_lon'][0], X['site_lat'][0])
    only_rx, none = oq.mesh_rx_ry0(edges[:3], None, sites, ry0=False)
    assert none is None
    np.testing.assert_allclose(only_rx[0], X['rx'][0], rtol=1e-10, atol=1e-8)
    with pytest.raises(ValueError):
        oq.mesh_rx_ry0([([1.], [2.])], [0.], sites)
    with pytest.raises(ValueError):
        oq.mesh_rx_ry0(edges, None, sites)


@pytest.mark.gpu
def test_mesh_rx_long_edge_vs_oracle(oq):
    """a top edge longer than one shared-memory chunk of segments"""
    rng = np.random.default_rng(3)
    K = 150
    lons = 20. + np.cumsum(rng.uniform(0.005, 0.02, K))
    lats = 40. + np.cumsum(rng.uniform(-0.01, 0.012, K))
    slon = rng.uniform(19.5, 23., 700)
    slat = rng.uniform(39., 41.5, 700)
    rx, ry0 = oq.mesh_rx_ry0([(lons, lats)], [75.], oq.Mesh(slon, slat))
    np.testing.assert_allclose(rx[0], geo_np.rx_distance(lons, lats, slon,
                                                         slat),
                               rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(
        ry0[0], geo_np.ry0_distance(lons, lats, 75., slon, slat),
        rtol=1e-10, atol=1e-8)


# ---- mean dip / strike / width (oracle/make_golden_mesh_dip.py) ------------
G = np.load(os.path.join(GOLDEN, 'mesh_dip.npz'))


def _dip_meshes():
    out = []
    for s, (r, c) in enumerate(G['shape']):
        sl = slice(G['off'][s], G['off'][s + 1])
        out.append(tuple(G[k][sl].reshape(r, c) for k in ('lon', 'lat', 'dep')))
    return out


def test_oracle_dip_strike_width_matches_reference():
    for s, m in enumerate(_dip_meshes()):
        incl, azim = geo_np.mean_inclination_and_azimuth(*m)
        np.testing.assert_allclose([incl, azim, geo_np.mean_width(*m)],
                                   G['res'][s], rtol=1e-12, atol=1e-10)


@pytest.mark.gpu
def test_mesh_dip_strike_width_golden(oq):
    """degrees and km; atol 1e-6 degrees: a vertical surface sits on the
    square-root singularity of sqrt(1 - cos^2) (the reference itself returns
    89.9999999999997 there)"""
    meshes = _dip_meshes()
    got = oq.mesh_dip_strike_width(meshes)
    np.testing.assert_allclose(np.delete(got, 2, axis=0),
                               np.delete(G['res'], 2, axis=0), rtol=1e-10,
                               atol=1e-9)
    # mesh 2 is exactly vertical: the sum of area * cos(inclination) is pure
    # rounding noise and its sign decides between (90 - 0, strike) and
    # (90 + 0 -> flipped to 90 - 0, strike + 180), two names of the same
    # plane; the reference's own answer depends on its summation order
    np.testing.assert_allclose(got[2, [0, 2]], G['res'][2, [0, 2]],
                               rtol=1e-10, atol=1e-6)
    assert min(abs(got[2, 1] - G['res'][2, 1]),
               abs(abs(got[2, 1] - G['res'][2, 1]) - 180.)) < 1e-6
    assert got[6, 0] == 0.       # flat on the Earth surface
    # SimpleFaultSurface.get_dip / get_strike: the top row of cells only
    top = [tuple(a[:2] for a in m) for m in meshes]
    got_top = oq.mesh_dip_strike_width(top)
    for s, m in enumerate(top):
        if s == 2:
            continue
        incl, azim = geo_np.mean_inclination_and_azimuth(*m)
        np.testing.assert_allclose(got_top[s, :2], [incl, azim], rtol=1e-10,
                                   atol=1e-6)
    # the strike feeds Ry0
    with pytest.raises(ValueError):
        oq.mesh_dip_strike_width([(m[0][:1], m[1][:1], m[2][:1])
                                  for m in meshes])
    with pytest.raises(ValueError):
        oq.mesh_dip_strike_width([tuple(a[::-1] for a in meshes[1])])
