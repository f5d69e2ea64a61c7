"""CPU tests: the numpy oracle (oracle/*_np.py) against the golden fixtures
produced by running the reference itself (oracle/make_golden.py) and against
row samples of the reference's own GMPE verification tables."""
import os
import json
import numpy as np
import pytest

import geo_np
import gmpe_np
import poe_np
import hazard_np
from helpers import (GOLDEN, load, gmpe_cases, pipeline_cases, ctx_of, rup_of,
                     sites_of, read_kat, json_kwargs, wc1994_area)


@pytest.mark.parametrize('tag', gmpe_cases())
def test_gmpe_golden(tag):
    d = load('gmpe_' + tag)
    out = gmpe_np.MODELS[str(d['name'])](
        ctx_of(d), [str(i) for i in d['imts']], **json_kwargs(d))
    np.testing.assert_allclose(out, d['mean_stds'], rtol=1e-13, atol=1e-14)


KAT_MODEL = {'bssa14': 'BooreEtAl2014', 'cy14': 'ChiouYoungs2014',
             'ask14': 'AbrahamsonEtAl2014', 'cb14': 'CampbellBozorgnia2014',
             'ba08': 'BooreAtkinson2008', 'akkar14_rjb': 'AkkarEtAlRjb2014',
             'cy08': 'ChiouYoungs2008', 'toro02': 'ToroEtAl2002',
             'c03': 'Campbell2003', 'sadigh97': 'SadighEtAl1997'}
KAT_ROW = {'MEAN': 0, 'TOTAL_STDDEV': 1, 'INTER_EVENT_STDDEV': 2,
           'INTRA_EVENT_STDDEV': 3}
# the reference's own tolerances in percent
KAT_TOL = {'bssa14': 2.0, 'cy14': 0.05, 'ask14': 0.1, 'cb14': 0.1,
           'ba08': 0.6, 'akkar14_rjb': 0.1, 'cy08': 0.05, 'toro02': 0.1,
           'c03': 0.1, 'sadigh97': 0.5}


def kat_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith('kat_'))


def kat_defaults(ctx, n):
    """columns the tables leave out (same defaults as the reference's test
    reader, hazardlib/tests/gsim/utils.py)"""
    have = vars(ctx)
    for name in ('z1pt0', 'z2pt5'):
        have.setdefault(name, np.full(n, -999.))
    have.setdefault('vs30measured', np.ones(n, bool))
    for name in ('ry0', 'hypo_depth', 'width', 'ztor', 'dip', 'rake', 'rx',
                 'rjb', 'rrup'):
        have.setdefault(name, np.zeros(n))
    return ctx


@pytest.mark.parametrize('fname', kat_files())
def test_gmpe_reference_tables(fname):
    tag = fname[4:].split('__')[0]
    ctx, rtype, cols = read_kat(fname)
    n = len(ctx.mag)
    ctx = kat_defaults(ctx, n)
    imts = list(cols)
    out = gmpe_np.MODELS[KAT_MODEL[tag]](ctx, imts)
    got = out[KAT_ROW[rtype]]
    for m, imt in enumerate(imts):
        exp = cols[imt]
        val = np.exp(got[m]) if rtype == 'MEAN' else got[m]
        ok = exp != 0
        err = np.abs(val[ok] / exp[ok] - 1) * 100
        assert err.max() < KAT_TOL[tag], (fname, imt, err.max())


def test_wc1994_matches_fixture():
    d = load('planar')
    r = rup_of(d)
    np.testing.assert_allclose(wc1994_area(r['mag'], r['rake']), r['area'],
                               rtol=1e-14)


def planar_of(d):
    r = rup_of(d)
    return r, geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], float(d['usd']),
        float(d['lsd']), float(d['rar']))


def test_planar_golden():
    d = load('planar')
    r, pla = planar_of(d)
    for f in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr', 'sdr', 'hypo'):
        np.testing.assert_allclose(pla[f], d['planar_' + f], rtol=0,
                                   atol=1e-10, err_msg=f)
    xyz = hazard_np.site_xyz(sites_of(d))
    for name in ('rrup', 'rjb', 'rx', 'ry0', 'repi', 'rhypo', 'azimuth',
                 'clon', 'clat'):
        np.testing.assert_allclose(geo_np.DIST[name](pla, xyz),
                                   d['dist_' + name], rtol=1e-11, atol=1e-9,
                                   err_msg=name)


def test_reference_known_distances():
    # hazardlib/tests/contexts_test.py:347-360: PointSource M7 at (0, 0),
    # strike 90 dip 90: checks the same quantities through the oracle on a
    # vertical strike-90 rupture (signs and symmetry)
    pla = geo_np.build_planar(
        np.array([7.]), np.array([90.]), np.array([90.]), np.array([0.]),
        np.array([1.]), np.array([wc1994_area(7., 0.)]), np.array([0.]),
        np.array([0.]), np.array([10.]), 0., 20., 1.)
    xyz = geo_np.spherical_to_cartesian(
        np.array([0., 0., 0.3]), np.array([0.1, -0.2, 0.]), np.zeros(3))
    rjb = geo_np.get_rjb(pla, xyz)[0]
    rx = geo_np.get_rx(pla, xyz)[0]
    ry0 = geo_np.get_ry0(pla, xyz)[0]
    # sites north / south of an E-W vertical fault: |rx| == rjb, ry0 == 0
    np.testing.assert_allclose(np.abs(rx[:2]), rjb[:2], rtol=1e-9)
    assert rx[0] < 0 < rx[1]   # north of a strike-90 fault is the footwall
    np.testing.assert_allclose(ry0[:2], 0, atol=1e-9)
    assert ry0[2] > 0 and abs(rx[2]) < 1e-6


@pytest.mark.parametrize('tag', pipeline_cases())
def test_pipeline_golden(tag):
    d = load('pipeline_' + tag)
    gsims = [tuple(g) for g in json.loads(str(d['gsims']))]
    imts = [str(i) for i in d['imts']]
    imtls = {i: d['imls'] for i in imts}
    tl, itime = float(d['truncation_level']), float(d['investigation_time'])
    ctx = ctx_of(d)
    # distances of the fixture rows from the rupture / site parameters
    r, pla = planar_of(d)
    order = np.argsort(r['mag'], kind='stable')
    pla = {k: v[order] for k, v in pla.items()}
    mag = r['mag'][order]
    sites = sites_of(d)
    maxdist = float(d['maxdist'])
    nb = json.loads(str(d['negbin']))
    po = None
    if nb:
        from scipy.stats import nbinom
        # NegativeBinomialTOM.get_pmf, tom.py:300-319
        rate = pla['wlr'][:, 2]
        mu, alpha, t = nb['mu'], nb['alpha'], itime
        tau = 1 / alpha
        theta = tau / (tau + rate * t)
        po = np.array([nbinom.pmf(np.arange(10), tau, th) for th in theta])
    octx = hazard_np.get_ctx(gsims, pla, mag, sites,
                             np.full(len(mag), maxdist), float(d['mindist']),
                             po)
    assert len(octx) == len(ctx)
    for name in ('rrup', 'rjb', 'rx', 'repi', 'rhypo'):
        if hasattr(ctx, name):
            np.testing.assert_allclose(getattr(octx, name),
                                       getattr(ctx, name), rtol=1e-11,
                                       atol=1e-9, err_msg=name)
    if nb:
        np.testing.assert_allclose(octx.probs_occur, ctx.probs_occur,
                                   rtol=1e-10, atol=1e-15)
    poes, ms = hazard_np.get_poes(gsims, ctx, imtls, tl)
    np.testing.assert_allclose(ms, d['mean_stds'], rtol=1e-12, atol=1e-13)
    ref = d['poes']
    big = ref > 1e-8
    np.testing.assert_allclose(poes[big], ref[big], rtol=1e-5)
    N, L, G = d['pmap'].shape
    arr = np.ones((N, L, G), np.float32)
    sid_to_row = np.arange(N, dtype=np.uint32)
    hazard_np.update(arr, sid_to_row, gsims, ctx, imtls, tl, itime)
    np.testing.assert_allclose(arr, d['pmap'], rtol=2e-6, atol=1e-7)
    rat = np.zeros((N, L, G), np.float32)
    hazard_np.update(rat, sid_to_row, gsims, ctx, imtls, tl, itime,
                     rates_mode=True)
    np.testing.assert_allclose(rat, d['rmap'], rtol=2e-6, atol=1e-9)


def test_truncnorm_golden():
    with open(os.path.join(GOLDEN, 'truncnorm.json')) as f:
        cases = json.load(f)
    for c in cases:
        z = np.array(c['z'])
        exp = np.array([np.nan if x is None else x for x in c['sf']])
        got = poe_np.truncnorm_sf(c['phi_b'], z)
        ok = np.isfinite(exp)
        np.testing.assert_allclose(got[ok], exp[ok], rtol=1e-12, atol=1e-16)
        assert np.isnan(got[~ok]).all()
    # SURVEY.md section 8c edge cases
    phi_b = float(poe_np.ndtr(3.))
    np.testing.assert_allclose(
        poe_np.truncnorm_sf(phi_b, np.array([2.99, 2.9999999, 3., 3.5])),
        [4.51109943e-05, 4.44384681e-10, 0., 0.], rtol=1e-6, atol=1e-18)


def ask14_from_planar_rows(n=6):
    """the rows of contexts_test.py:370-387 (test_from_planar): M5, rake 45,
    vertical strike-slip plane, vs30 760 not measured, z1pt0 20"""
    from helpers import Bag
    rrup = np.linspace(1., 100., n)
    return Bag(mag=np.full(n, 5.), rake=np.full(n, 45.), dip=np.full(n, 90.),
               ztor=np.full(n, 3.5), width=np.full(n, 3.), rrup=rrup,
               rjb=rrup, rx=rrup, ry0=np.zeros(n), vs30=np.full(n, 760.),
               vs30measured=np.zeros(n, bool), z1pt0=np.full(n, 20.))


def test_reference_known_ask14_stddevs():
    """contexts_test.py:384-387: sig, tau, phi of AbrahamsonEtAl2014 for PGA
    are .79162428, .47, .637 on all the sites of that example"""
    ms = gmpe_np.mean_stds([('AbrahamsonEtAl2014', {})],
                           ask14_from_planar_rows(), ['PGA'])
    np.testing.assert_allclose(ms[1], .79162428, rtol=1e-8)
    np.testing.assert_allclose(ms[2], .47, rtol=1e-12)
    np.testing.assert_allclose(ms[3], .637, rtol=1e-12)
