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
