"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE
(/root/reference, imported through oracle/refshim) and, in the same run, pin
the numpy oracle (oracle/*_np.py) against them.  Runs only in the build
container; the fixtures it writes travel with the repo.

    python oracle/make_golden.py

Fixtures
--------
gmpe_<tag>.npz      random context rows + reference (4, M, N) mean/std/tau/phi
planar.npz          rupture parameters, the reference's build_planar output,
                    sites and the reference's rrup/rjb/rx/ry0/repi/rhypo/azimuth
pipeline_<tag>.npz  planar ruptures x sites through the reference's
                    _get_ctx_planar -> get_mean_stds -> gen_poes -> update:
                    ctx columns, mean/stds, float32 PoEs, float32 pne map and
                    rate map
truncnorm.json      truncnorm_sf edge cases
kat_<gmpe>.csv      row subsamples of the reference's own GMPE verification
                    tables (openquake/hazardlib/tests/gsim/data)
"""
import os
import sys
import csv
import json
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib import valid  # noqa
from openquake.hazardlib.contexts import ContextMaker, _get_ctx_planar  # noqa
from openquake.baselib.general import RecordBuilder  # noqa
from openquake.hazardlib.calc.filters import IntegrationDistance  # noqa
from openquake.hazardlib.geo.surface import planar as P  # noqa
from openquake.hazardlib.site import SiteCollection  # noqa
from openquake.hazardlib.map_array import MapArray  # noqa
from openquake.hazardlib.tom import PoissonTOM, NegativeBinomialTOM  # noqa
from openquake.hazardlib.scalerel.wc1994 import WC1994  # noqa
from openquake.hazardlib.stats import truncnorm_sf, ndtr  # noqa
import geo_np  # noqa
import gmpe_np  # noqa
import poe_np  # noqa
import hazard_np  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
IMTS = ['PGA', 'SA(0.1)', 'SA(0.2)', 'SA(0.5)', 'SA(1.0)', 'SA(2.0)',
        'SA(3.0)']
IMTS_X = ['PGA', 'PGV', 'SA(0.05)', 'SA(0.1)', 'SA(0.2)', 'SA(0.27)',
          'SA(0.5)', 'SA(0.7)', 'SA(1.0)', 'SA(2.0)', 'SA(3.0)']

# tag -> (class name, toml, oracle kwargs)
GMPE_CASES = {
    'bssa14': ('BooreEtAl2014', '[BooreEtAl2014]', {}),
    'bssa14_cal': ('BooreEtAl2014', '[BooreEtAl2014]\nregion="CAL"',
                   dict(region='CAL')),
    'bssa14_jpn_nosof': ('BooreEtAl2014',
                         '[BooreEtAl2014]\nregion="JPN"\nsof=0',
                         dict(region='JPN', sof=0)),
    'bssa14_eps': ('BooreEtAl2014', '[BooreEtAl2014]\nsigma_mu_epsilon=1.5',
                   dict(sigma_mu_epsilon=1.5)),
    'cy14': ('ChiouYoungs2014', '[ChiouYoungs2014]', {}),
    'cy14_jpn': ('ChiouYoungs2014', '[ChiouYoungs2014]\nregion="JPN"',
                 dict(region='JPN')),
    'cy14_wen_nohw': ('ChiouYoungs2014',
                      '[ChiouYoungs2014]\nregion="WEN"\nuse_hw=false',
                      dict(region='WEN', use_hw=False)),
    'ask14': ('AbrahamsonEtAl2014', '[AbrahamsonEtAl2014]', {}),
    'ask14_jpn': ('AbrahamsonEtAl2014', '[AbrahamsonEtAl2014]\nregion="JPN"',
                  dict(region='JPN')),
    'ask14_twn': ('AbrahamsonEtAl2014', '[AbrahamsonEtAl2014]\nregion="TWN"',
                  dict(region='TWN')),
    'ask14_chn': ('AbrahamsonEtAl2014', '[AbrahamsonEtAl2014]\nregion="CHN"',
                  dict(region='CHN')),
    'cb14': ('CampbellBozorgnia2014', '[CampbellBozorgnia2014]', {}),
    'cb14_sj_eps': ('CampbellBozorgnia2014',
                    '[CampbellBozorgnia2014]\nSJ=true\nsigma_mu_epsilon=0.5',
                    dict(SJ=True, sigma_mu_epsilon=0.5)),
    'ba08': ('BooreAtkinson2008', '[BooreAtkinson2008]', {}),
    'ab06': ('AtkinsonBoore2006', '[AtkinsonBoore2006]', {}),
    'ab06_mblg87': ('AtkinsonBoore2006',
                    '[AtkinsonBoore2006]\nmag_eq="Mblg87"\nscale_fac=0.5146',
                    dict(mag_eq='Mblg87', scale_fac=0.5146)),
    'akkar14_rjb': ('AkkarEtAlRjb2014', '[AkkarEtAlRjb2014]', {}),
    'akkar14_repi': ('AkkarEtAlRepi2014', '[AkkarEtAlRepi2014]', {}),
    'akkar14_rhyp': ('AkkarEtAlRhyp2014',
                     '[AkkarEtAlRhyp2014]\nadjustment_factor=1.2',
                     dict(adjustment_factor=1.2)),
    # the models of the LogicTreeCase2ClassicalPSHA demo; a 4th element is
    # the IMT list when the model does not cover IMTS_X
    'cy08': ('ChiouYoungs2008', '[ChiouYoungs2008]', {}),
    'toro02': ('ToroEtAl2002', '[ToroEtAl2002]', {},
               ['PGA', 'SA(0.04)', 'SA(0.1)', 'SA(0.27)', 'SA(0.7)', 'SA(1.0)',
                'SA(2.0)', 'SA(3.0)', 'SA(4.0)']),
    'c03': ('Campbell2003', '[Campbell2003]', {},
            ['PGA', 'SA(0.02)', 'SA(0.1)', 'SA(0.27)', 'SA(0.7)', 'SA(1.0)',
             'SA(2.0)', 'SA(3.0)', 'SA(4.0)']),
    # the SHARE variants (style of faulting and rock adjustment); the first
    # is the model of the reference manual's ContextMaker example
    'toro02_share': ('ToroEtAl2002SHARE', '[ToroEtAl2002SHARE]', {},
                     ['PGA', 'SA(0.04)', 'SA(0.1)', 'SA(0.27)', 'SA(0.7)',
                      'SA(1.0)', 'SA(2.0)', 'SA(3.0)', 'SA(4.0)']),
    # the model of the PEER verification tests (acceptance/peer_test.py)
    'sadigh97': ('SadighEtAl1997', '[SadighEtAl1997]', {},
                 ['PGA', 'SA(0.075)', 'SA(0.1)', 'SA(0.27)', 'SA(0.7)',
                  'SA(1.0)', 'SA(2.0)', 'SA(3.0)', 'SA(4.0)']),
    'c03_share': ('Campbell2003SHARE', '[Campbell2003SHARE]', {},
                  ['PGA', 'SA(0.02)', 'SA(0.1)', 'SA(0.27)', 'SA(0.7)',
                   'SA(1.0)', 'SA(2.0)', 'SA(3.0)', 'SA(4.0)']),
}


def random_ctx(cm, n, rng):
    """Random context rows covering every branch of the models"""
    ctx = cm.new_ctx(n)
    names = ctx.dtype.names
    ctx['mag'] = rng.choice(np.round(np.arange(4.0, 8.3, 0.1), 1), n)
    if 'rake' in names:
        ctx['rake'] = rng.choice([-180, -150, -120, -90, -60, -45, -30, 0, 30,
                                  45, 60, 90, 135, 150, 180, 17.3], n)
    if 'dip' in names:
        ctx['dip'] = rng.choice([20, 30, 45, 60, 75, 90.], n)
    if 'ztor' in names:
        ctx['ztor'] = rng.uniform(0, 20, n)
    if 'width' in names:
        ctx['width'] = rng.uniform(1, 25, n)
    if 'hypo_depth' in names:
        ctx['hypo_depth'] = rng.uniform(2, 25, n)
    if 'vs30' in names:
        ctx['vs30'] = rng.choice([150, 180, 200, 225, 260, 300, 400, 560, 760,
                                  800, 1000, 1130, 1500, 2200], n
                                 ) * rng.uniform(0.95, 1.05, n)
    if 'vs30measured' in names:
        ctx['vs30measured'] = rng.random(n) < .5
    if 'z1pt0' in names:
        ctx['z1pt0'] = rng.uniform(0, 900, n)
        ctx['z1pt0'][rng.random(n) < .2] = -999
    if 'z2pt5' in names:
        ctx['z2pt5'] = rng.uniform(0.2, 6, n)
        ctx['z2pt5'][rng.random(n) < .2] = -999
    rjb = np.abs(rng.choice([0, 0.5, 5, 20, 50, 100, 200, 400], n) *
                 rng.uniform(0.5, 1.5, n))
    rrup = np.sqrt(rjb**2 + rng.uniform(0, 10, n)**2) + 0.01
    for d in ('rjb', 'repi'):
        if d in names:
            ctx[d] = rjb
    ctx['rrup'] = rrup
    if 'rhypo' in names:
        ctx['rhypo'] = rrup * 1.1
    if 'rx' in names:
        ctx['rx'] = rng.uniform(-1, 1, n) * (rjb + 5)
    if 'ry0' in names:
        ctx['ry0'] = np.where(rng.random(n) < .5, 0,
                              rng.uniform(0, 1, n) * rjb)
    ctx['sids'] = np.arange(n)
    return ctx


CTX_FIELDS = ('mag rake dip ztor width hypo_depth vs30 vs30measured z1pt0 '
              'z2pt5 rjb rrup rx ry0 repi rhypo').split()


def ctx_columns(ctx):
    return {'ctx_' + f: np.array(ctx[f]) for f in CTX_FIELDS
            if f in ctx.dtype.names}


def make_gmpe():
    rng = np.random.default_rng(7)
    # OQ_GOLDEN_ONLY=tag1,tag2 writes only those fixtures (the random
    # stream still runs through every case, so the others do not change)
    only = os.environ.get('OQ_GOLDEN_ONLY')
    only = only.split(',') if only else None
    for tag, case in GMPE_CASES.items():
        name, toml, kw = case[:3]
        imts = case[3] if len(case) > 3 else IMTS_X
        gs = valid.gsim(toml)
        cm = ContextMaker('*', [gs], dict(
            imtls={i: [0.1] for i in imts}, truncation_level=3.))
        ctx = random_ctx(cm, 384, rng)
        if only and tag not in only:
            continue
        ref = cm.get_mean_stds([ctx], split_by_mag=False)[:, 0]
        mine = gmpe_np.MODELS[name](ctx, imts, **kw)
        np.testing.assert_allclose(mine, ref, rtol=1e-13, atol=1e-14,
                                   err_msg=tag)
        np.savez_compressed(
            os.path.join(GOLDEN, 'gmpe_%s.npz' % tag), name=name,
            kwargs=json.dumps(kw), imts=np.array(imts), mean_stds=ref,
            **ctx_columns(ctx))
        print('gmpe', tag, 'oracle == reference to',
              np.abs(mine - ref).max())


def synth_ruptures(U, rng, box=1.5):
    """random planar ruptures in a box"""
    mag = rng.choice(np.round(np.arange(5.0, 8.0, 0.1), 1), U)
    strike = rng.uniform(0, 360, U)
    dip = rng.choice([30., 45., 60., 90.], U)
    rake = rng.choice([-90., 0., 90.], U)
    lon = rng.uniform(0, box, U)
    lat = rng.uniform(0, box, U)
    dep = rng.uniform(5, 15, U)
    msr = WC1994()
    area = np.array([msr.get_median_area(m, r) for m, r in zip(mag, rake)])
    rate = 10 ** (4 - mag) / U
    return dict(mag=mag, strike=strike, dip=dip, rake=rake, hypo_lon=lon,
                hypo_lat=lat, hypo_dep=dep, area=area, rate=rate)


def synth_sites(N, rng, box=1.5, sentinels=False):
    n_side = int(np.ceil(np.sqrt(N)))
    gx, gy = np.meshgrid(np.linspace(0, box, n_side),
                         np.linspace(0, box, n_side))
    sites = dict(sids=np.arange(N, dtype=np.uint32),
                 lon=gx.ravel()[:N], lat=gy.ravel()[:N], depth=np.zeros(N),
                 vs30=rng.uniform(180, 1500, N),
                 vs30measured=rng.random(N) < .5,
                 z1pt0=rng.uniform(5, 800, N), z2pt5=rng.uniform(0.3, 6, N))
    if sentinels:
        sites['z1pt0'][rng.random(N) < .1] = -999
        sites['z2pt5'][rng.random(N) < .1] = -999
    return sites


def ref_planar(r, usd=0., lsd=25., rar=1.5):
    """The reference's build_planar, one rupture at a time -> (U, 3) recarray"""
    out = []
    for u in range(len(r['mag'])):
        planin = np.zeros((1, 1), P.planin_dt).view(np.recarray)
        for k in ('mag', 'strike', 'dip', 'rake', 'rate', 'area'):
            planin[k][:] = r[k][u]
        hdd = np.array([[1.0, r['hypo_dep'][u]]])
        pla = P.build_planar(planin, hdd, r['hypo_lon'][u], r['hypo_lat'][u],
                             usd, lsd, rar)
        out.append(pla.reshape(-1, 3))
    return np.concatenate(out).view(np.recarray)


def ref_sitecol(sites):
    class SM:
        pass
    sm = SM()
    for k in ('vs30', 'vs30measured', 'z1pt0', 'z2pt5'):
        setattr(sm, k, sites[k])
    return SiteCollection.from_points(
        sites['lon'], sites['lat'], sites['depth'], sm,
        req_site_params=('vs30', 'vs30measured', 'z1pt0', 'z2pt5'))


PLANAR_FIELDS = ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr', 'sdr',
                 'hypo')


def make_planar():
    rng = np.random.default_rng(20240)
    r = synth_ruptures(96, rng)
    sites = synth_sites(64, rng)
    # a few sites right on top of / very close to ruptures
    sites['lon'][:8] = r['hypo_lon'][:8] + 1e-3
    sites['lat'][:8] = r['hypo_lat'][:8] - 1e-3
    ref = ref_planar(r)
    mine = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    for f in PLANAR_FIELDS:
        np.testing.assert_allclose(mine[f], ref[f], rtol=0, atol=1e-10,
                                   err_msg=f)
    sc = ref_sitecol(sites)
    dists = {}
    dists['rrup'] = P.project(ref, sc.xyz)[0]
    for name in ('rjb', 'rx', 'ry0', 'repi', 'rhypo', 'azimuth'):
        dists[name] = P.get_distances_planar(ref, sc, name)
    _rrup, xx, yy = P.project(ref, sc.xyz)
    closest = P.project_back(ref, xx, yy)        # (3, U, N)
    dists['clon'], dists['clat'] = closest[0], closest[1]
    for name, d in dists.items():
        m = geo_np.DIST[name](mine, sc.xyz)
        np.testing.assert_allclose(m, d, rtol=1e-11, atol=1e-9, err_msg=name)
    np.savez_compressed(
        os.path.join(GOLDEN, 'planar.npz'), usd=0., lsd=25., rar=1.5,
        **{'rup_' + k: v for k, v in r.items()},
        **{'site_' + k: v for k, v in sites.items()},
        **{'planar_' + f: np.array(ref[f]) for f in PLANAR_FIELDS},
        **{'dist_' + k: v for k, v in dists.items()})
    print('planar: oracle == reference')


PIPE_CASES = {
    # tag: (gsims [(name, toml, kw)], maxdist, mindist, negbin, U, N)
    'cfg3': (['bssa14', 'cy14'], '1000', 0, False, 48, 40),
    'cfg4': (['bssa14', 'cy14', 'ask14', 'cb14'], '1000', 0, False, 24, 30),
    'ba08_filter': (['ba08'], '60', 5., False, 40, 40),
    'akkar_ab06': (['akkar14_rjb', 'akkar14_repi', 'akkar14_rhyp', 'ab06'],
                   '200', 0, False, 24, 30),
    'negbin': (['bssa14', 'cy14'], '1000', 0, True, 24, 30),
}


def make_pipeline():
    imls = np.logspace(-2, 0.3, 20)
    imtls = {imt: imls for imt in IMTS}
    for tag, (gtags, maxdist, mindist, negbin, U, N) in PIPE_CASES.items():
        rng = np.random.default_rng(abs(hash(tag)) % 1000 + 5)
        rng = np.random.default_rng(sum(map(ord, tag)))
        r = synth_ruptures(U, rng)
        sites = synth_sites(N, rng, sentinels=True)
        ref = ref_planar(r)
        sc = ref_sitecol(sites)
        gsims = [valid.gsim(GMPE_CASES[t][1]) for t in gtags]
        ogsims = [(GMPE_CASES[t][0], GMPE_CASES[t][2]) for t in gtags]
        cm = ContextMaker('*', gsims, dict(
            imtls=imtls, truncation_level=3., investigation_time=50.,
            maximum_distance=IntegrationDistance.new(maxdist),
            minimum_distance=mindist))
        if negbin:
            tom = NegativeBinomialTOM(50., mu=1., alpha=0.5)
        else:
            tom = PoissonTOM(50.)
        dd = cm.defaultdict.copy()
        dd['probs_occur'] = np.zeros(10 if negbin else 0)
        builder = RecordBuilder(**dd)
        cm.siteparams = [par for par in sc.array.dtype.names if par in dd]
        cm.ruptparams = cm.REQUIRES_RUPTURE_PARAMETERS | {'occurrence_rate'}
        cm.fewsites = False
        ctxs = []
        mags = r['mag']
        for magi, mag in enumerate(np.unique(mags)):
            idx = np.flatnonzero(mags == mag)
            pla = ref[idx]
            ctx = _get_ctx_planar(cm, builder, mag, 1.0, magi, pla, sc,
                                  0, 0, tom)
            ctx = ctx[ctx.rrup < cm.maximum_distance(mag)]
            if len(ctx):
                ctxs.append(ctx)
        ctx = np.concatenate(ctxs).view(np.recarray)
        mean_stds = cm.get_mean_stds([ctx], split_by_mag=True)
        # get_mean_stds(split_by_mag=True) reorders rows by magnitude, which
        # is already the order of ctx here
        cm.tom = tom
        poes = np.concatenate(
            [p for p, *_ in cm.gen_poes(ctx.copy().view(np.recarray))])
        L, G = 20 * len(IMTS), len(gsims)
        pmap = MapArray(sc.sids, L, G).fill(1)
        cm.update(pmap, ctx.copy().view(np.recarray))
        rmap = MapArray(sc.sids, L, G, rates=True).fill(0)
        cm.update(rmap, ctx.copy().view(np.recarray))

        # ---- pin the oracle on the same inputs -------------------------
        mine = geo_np.build_planar(
            r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
            r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
        order = np.argsort(mags, kind='stable')  # the reference groups by mag
        mine = {k: v[order] for k, v in mine.items()}
        omag = mags[order]
        magdist = np.array([float(cm.maximum_distance(m)) for m in omag])
        po = None
        if negbin:
            po = tom.get_pmf(mine['wlr'][:, 2])
        octx = hazard_np.get_ctx(ogsims, mine, omag, sites, magdist,
                                 mindist, po)
        assert len(octx) == len(ctx), (len(octx), len(ctx))
        for d in cm.REQUIRES_DISTANCES | {'rrup'}:
            np.testing.assert_allclose(getattr(octx, d), ctx[d], rtol=1e-11,
                                       atol=1e-9, err_msg=d)
        opoes, oms = hazard_np.get_poes(ogsims, octx, imtls, 3.)
        np.testing.assert_allclose(oms, mean_stds, rtol=1e-9, atol=1e-10)
        big = poes > 1e-8
        np.testing.assert_allclose(opoes[big], poes[big], rtol=1e-5)
        oarr = np.ones((N, L, G), np.float32)
        sid_to_row = np.arange(N, dtype=np.uint32)
        hazard_np.update(oarr, sid_to_row, ogsims, octx, imtls, 3., 50.)
        np.testing.assert_allclose(oarr, pmap.array, rtol=2e-6, atol=1e-7)
        orat = np.zeros((N, L, G), np.float32)
        hazard_np.update(orat, sid_to_row, ogsims, octx, imtls, 3., 50.,
                         rates_mode=True)
        np.testing.assert_allclose(orat, rmap.array, rtol=2e-6, atol=1e-9)
        cols = {}
        for f in ctx.dtype.names:
            cols['ctx_' + f] = np.array(ctx[f])
        np.savez_compressed(
            os.path.join(GOLDEN, 'pipeline_%s.npz' % tag),
            gsims=json.dumps(ogsims), imts=np.array(IMTS), imls=imls,
            truncation_level=3., investigation_time=50.,
            maxdist=float(maxdist), mindist=float(mindist),
            negbin=json.dumps(dict(mu=1., alpha=0.5) if negbin else {}),
            usd=0., lsd=25., rar=1.5,
            **{'rup_' + k: v for k, v in r.items()},
            **{'site_' + k: v for k, v in sites.items()},
            mean_stds=mean_stds, poes=poes, pmap=pmap.array, rmap=rmap.array,
            **cols)
        print('pipeline', tag, 'rows', len(ctx), 'poes', poes.shape,
              'pne range', pmap.array.min(), pmap.array.max(),
              ': oracle == reference')


def make_truncnorm():
    out = []
    for tl in (3., 2., 2.5, 5., 6., 0., 1e-9, 99.):
        phi_b = float(ndtr(tl))
        z = np.array([-7., -tl, -tl + 1e-7, -2.5, -1., -1e-3, 0., 1e-3, .5,
                      1., 2., 2.5, 2.9, 2.99, 2.9999999, tl - 1e-7, tl,
                      tl + 1e-7, 3.5, 4.9, 5.5, 7.])
        with np.errstate(all='ignore'):
            sf = truncnorm_sf(phi_b, z)
            mine = poe_np.truncnorm_sf(float(poe_np.ndtr(tl)), z)
        ok = np.isfinite(sf)
        np.testing.assert_allclose(mine[ok], sf[ok], rtol=1e-12, atol=1e-16)
        out.append(dict(truncation_level=tl, phi_b=phi_b, z=z.tolist(),
                        sf=[None if np.isnan(x) else float(x) for x in sf]))
    with open(os.path.join(GOLDEN, 'truncnorm.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('truncnorm: oracle == reference')


KAT = {  # reference verification tables (hazardlib/tests/gsim/data) ->
         # (gmpe tag, files, rows kept per file)
    'bssa14': ('BSSA2014', ['BSSA_2014_MEAN.csv', 'BSSA_2014_TOTAL_STD.csv',
                            'BSSA_2014_INTER_STD.csv',
                            'BSSA_2014_INTRA_STD.csv']),
    'cy14': ('NGA/CY14', ['CY14_MEDIAN_MS_HW_NM.csv', 'CY14_MEDIAN_MS_HW_RV.csv',
                          'CY14_MEDIAN_MS_HW_SS.csv', 'CY14_TOTAL_EVENT_SIGMA.csv',
                          'CY14_INTER_EVENT_SIGMA.csv',
                          'CY14_INTRA_EVENT_SIGMA.csv']),
    'ask14': ('ASK14', ['ASK14_ResMEAN_RegCAL.csv', 'ASK14_ResStdTot_RegCAL.csv',
                        'ASK14_ResStdPhi_RegCAL.csv']),
    'cb14': ('CB14', ['CB2014_MEAN.csv', 'CB2014_STD_TOTAL.csv',
                      'CB2014_STD_INTER.csv', 'CB2014_STD_INTRA.csv']),
    'ba08': ('NGA/BA08', ['BA08_MEDIAN_NM.csv', 'BA08_MEDIAN_RV.csv',
                          'BA08_MEDIAN_SS.csv', 'BA08_STD_INTRA.csv',
                          'BA08_STD_INTER.csv', 'BA08_SIGTM_SS.csv']),
    'akkar14_rjb': ('AKKAR14', ['AKKAR_2014_RJB_MEAN.csv',
                                'AKKAR_2014_RJB_STD_TOTAL.csv']),
    'cy08': ('NGA/CY08', ['CY08_MEDIAN_MS_HW_NM.csv', 'CY08_MEDIAN_MS_HW_RV.csv',
                          'CY08_MEDIAN_MS_HW_SS.csv',
                          'CY08_SIGMEAS_MS_HW_SS.csv',
                          'CY08_SIGINFER_MS_HW_SS.csv',
                          'CY08_INTER_EVENT_SIGMA.csv',
                          'CY08_INTRA_EVENT_SIGMA.csv']),
    'toro02': ('TORO02', ['T02_MEAN.csv', 'T02_STD_TOTAL.csv']),
    'c03': ('C03', ['C03_MEAN.csv', 'C03_STD_TOTAL.csv']),
    'sadigh97': ('SADIGH97', ['SADIGH1997_ROCK_MEAN.csv',
                              'SADIGH1997_ROCK_STD_TOTAL.csv',
                              'SADIGH1997_SOIL_MEAN.csv',
                              'SADIGH1997_SOIL_STD_TOTAL.csv']),
}


def make_kat(keep=60):
    base = os.path.join(oq_bootstrap.REFERENCE, 'openquake', 'hazardlib',
                        'tests', 'gsim', 'data')
    rng = np.random.default_rng(3)
    only = os.environ.get('OQ_GOLDEN_ONLY')
    only = only.split(',') if only else None
    for tag, (sub, files) in KAT.items():
        for fname in files:
            src = os.path.join(base, sub, fname)
            if not os.path.exists(src):
                print('MISSING', src)
                continue
            with open(src) as f:
                rows = list(csv.reader(f))
            header, body = rows[0], rows[1:]
            idx = sorted(rng.choice(len(body), min(keep, len(body)),
                                    replace=False))
            if only and tag not in only:
                continue
            dest = os.path.join(GOLDEN, 'kat_%s__%s' % (tag, fname))
            with open(dest, 'w', newline='') as f:
                w = csv.writer(f)
                w.writerow(header)
                w.writerows(body[i] for i in idx)
        print('kat', tag, len(files), 'files')


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    what = sys.argv[1:] or ['truncnorm', 'gmpe', 'planar', 'pipeline', 'kat']
    for w in what:
        globals()['make_' + w]()
