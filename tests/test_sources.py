"""Rupture generation of point sources and
BASELINE.json configs[0] (the AreaSourceClassicalPSHA demo).

CPU part: the oracle (oracle/source_np.py) and the host-side MFD / batch
logic against the fixtures produced by the reference
(oracle/make_golden_sources.py).  GPU part (-m gpu): oqg_point_ruptures
against the reference's PointSource.get_planar output, and the demo's hazard
curves through ContextMaker.hazard_point_sources against the reference's
ContextMaker.get_ctxs -> update.
"""
import os
import numpy as np
import pytest

import source_np
import hazard_np
from helpers import GOLDEN

PS = np.load(os.path.join(GOLDEN, 'point_sources.npz'))
NAMES = [str(n) for n in PS['names']]


def case(name):
    g = lambda k: PS[name + '__' + k]  # noqa
    sc = g('scalars')
    d = dict(lon=sc[0], lat=sc[1], usd=sc[2], lsd=sc[3], rar=sc[4],
             shift=bool(sc[5]), msr=str(g('msr')), mags=g('mags'),
             mag_rates=g('mag_rates'), npd=[tuple(x) for x in g('npd')],
             hdd=[tuple(x) for x in g('hdd')],
             dip_fracs=(list(g('dip_fracs'))
                        if name + '__dip_fracs' in PS.files else None),
             planar=g('planar'), mag=g('mag'), rate=g('rate'))
    if name + '__gr' in PS.files:
        d['gr'] = g('gr')
    return d


def flat(pla):
    U = len(pla['wlr'])
    out = np.zeros((U, 3, 14))
    out[:, :, 0:4] = pla['corners']
    out[:, :, 4:8] = pla['xyz']
    for k, name in enumerate(('normal', 'uv1', 'uv2', 'wlr', 'sdr', 'hypo')):
        out[:, :, 8 + k] = pla[name]
    return out.reshape(U, 42)


def as_source(oq, d, sid='1'):
    mfd = oq.ArbitraryMFD(d['mags'], d['mag_rates'])
    return oq.PointSource(sid, d['lon'], d['lat'], mfd, d['msr'], d['rar'],
                          d['usd'], d['lsd'], d['npd'], d['hdd'],
                          d['dip_fracs'])


# ------------------------------------------------------------------ CPU
@pytest.mark.parametrize('name', NAMES)
def test_oracle_point_ruptures_golden(name):
    d = case(name)
    pla, mag, rate = source_np.point_source_ruptures(d, d['shift'])
    np.testing.assert_allclose(flat(pla), d['planar'], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(mag, d['mag'])
    np.testing.assert_allclose(rate, d['rate'], rtol=1e-15)


@pytest.mark.parametrize('name', [n for n in NAMES if n + '__gr' in PS.files])
def test_truncated_gr_matches_reference(name):
    import oq_engine_b200 as oq
    d = case(name)
    mfd = oq.TruncatedGRMFD(*d['gr'])
    mags, rates = zip(*mfd.get_annual_occurrence_rates())
    np.testing.assert_allclose(mags, d['mags'], rtol=1e-15)
    np.testing.assert_allclose(rates, d['mag_rates'], rtol=1e-15)
    m, r = source_np.truncated_gr_rates(*d['gr'])
    np.testing.assert_allclose(m, d['mags'], rtol=1e-15)
    np.testing.assert_allclose(r, d['mag_rates'], rtol=1e-15)


def test_mfd_validation_errors():
    import oq_engine_b200 as oq
    with pytest.raises(ValueError):
        oq.TruncatedGRMFD(5., 5.05, 0.1, 4., 1.)   # narrower than a bin
    with pytest.raises(ValueError):
        oq.TruncatedGRMFD(5., 6., 0., 4., 1.)
    with pytest.raises(ValueError):
        oq.EvenlyDiscretizedMFD(5., 0.1, [])
    e = oq.EvenlyDiscretizedMFD(5.05, 0.1, [1e-2, 1e-3])
    assert e.get_annual_occurrence_rates() == [(5.05, 1e-2),
                                               (5.05 + 1 * 0.1, 1e-3)]


def test_point_source_validation_errors():
    import oq_engine_b200 as oq
    mfd = oq.ArbitraryMFD([5.], [1e-2])
    ok = dict(nodal_plane_distribution=[(1., 0., 90., 0.)],
              hypocenter_distribution=[(1., 5.)])
    with pytest.raises(NotImplementedError):   # no device implementation
        oq.PointSource('1', 0., 0., mfd, 'Leonard2014_SCR', 1., 0., 10., **ok)
    with pytest.raises(ValueError):            # lsd above usd
        oq.PointSource('1', 0., 0., mfd, 'WC1994', 1., 10., 5., **ok)
    with pytest.raises(ValueError):            # depth below lsd
        oq.PointSource('1', 0., 0., mfd, 'WC1994', 1., 0., 4., **ok)


def test_batch_layout_and_expand():
    import oq_engine_b200 as oq
    srcs = [as_source(oq, case(n), n) for n in NAMES if n != 'pointmsr']
    # two sources sharing the same distributions share one slice
    d = case('multi')
    twin = oq.PointSource('twin', 1., 2., srcs[1].mfd, 'WC1994', 2., 0., 12.,
                          srcs[1].nodal_plane_distribution,
                          srcs[1].hypocenter_distribution)
    b = oq.PointSourceBatch(srcs + [twin])
    assert b.S == 4 and b.U == 8 + 54 + 32 + 54
    assert list(b.rup_off) == [0, 8, 62, 94, 148]
    assert b.beg['mag'][1] == b.beg['mag'][3]
    assert b.beg['np'][1] == b.beg['np'][3]
    assert len(b.mag) == 8 + 9 + 8   # the twin adds nothing
    mag, rate = b.expand()
    want_mag = np.concatenate([case(n)['mag'] for n in
                               ('single', 'multi', 'dipfrac_shift', 'multi')])
    want_rate = np.concatenate([case(n)['rate'] for n in
                                ('single', 'multi', 'dipfrac_shift', 'multi')])
    np.testing.assert_array_equal(mag, want_mag)
    np.testing.assert_allclose(rate, want_rate, rtol=1e-15)
    assert b.hd_frac is not None and b.hd_frac[-2] == 0.25
    with pytest.raises(ValueError):            # mixed scaling relationships
        oq.PointSourceBatch(srcs + [as_source(oq, case('pointmsr'))])


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_device_point_ruptures_golden(oq, name):
    d = case(name)
    h = oq.contexts.get_handle()
    b = oq.PointSourceBatch([as_source(oq, d)])
    md = np.linspace(100., 300., len(b.mag))
    t_pla, t_mag, t_md = oq.point_ruptures_device(b, h, md, d['shift'])[:3]
    pla = t_pla.cpu().numpy()
    want = d['planar'].copy()
    # wlr[2] of the device ruptures is the occurrence rate, of the
    # reference's planar array the (nodal plane x depth) weight
    got_rate, pla[:, 2 * 14 + 11] = pla[:, 2 * 14 + 11].copy(), 0.
    want[:, 2 * 14 + 11] = 0.
    # PointMSR ruptures are 10 m x 10 m: their unit vectors and sizes come
    # from differences of corner coordinates of magnitude 6371 km that agree
    # to 6 digits, conditioned to ~1e-9 relative in the reference itself
    tol = 5e-9 if d['msr'] == 'PointMSR' else 1e-11
    np.testing.assert_allclose(pla, want, rtol=tol, atol=tol)
    np.testing.assert_allclose(got_rate, d['rate'], rtol=1e-15)
    np.testing.assert_array_equal(t_mag.cpu().numpy(), d['mag'])
    nphc = len(d['npd']) * len(d['hdd'])
    np.testing.assert_array_equal(t_md.cpu().numpy(), np.repeat(md, nphc))


@pytest.mark.gpu
def test_device_point_ruptures_many_sources(oq):
    """a batch of sources with shared and own distributions against the
    oracle, source by source"""
    rng = np.random.default_rng(5)
    mfd = oq.TruncatedGRMFD(4.7, 7.2, 0.1, 3.9, 0.95)
    npd = [(0.25, 0., 90., 0.), (0.25, 45., 60., 90.), (0.5, 200., 35., -90.)]
    hdd = [(0.2, 3.), (0.5, 8.), (0.3, 17.)]
    srcs, descs = [], []
    for i in range(40):
        lon, lat = rng.uniform(-180, 180), rng.uniform(-70, 70)
        own = i % 5 == 0
        m = oq.EvenlyDiscretizedMFD(5.05, 0.2, list(rng.uniform(1e-4, 1e-2, 6))
                                    ) if own else mfd
        srcs.append(oq.PointSource(str(i), lon, lat, m, 'WC1994', 1.5, 0., 20.,
                                   npd, hdd))
        mags, rates = zip(*m.get_annual_occurrence_rates())
        descs.append(dict(lon=lon, lat=lat, usd=0., lsd=20., rar=1.5,
                          msr='WC1994', mags=mags, mag_rates=rates, npd=npd,
                          hdd=hdd))
    b = oq.PointSourceBatch(srcs)
    t_pla, t_mag, _ = oq.point_ruptures_device(b, oq.contexts.get_handle())
    pla, mag = t_pla.cpu().numpy(), t_mag.cpu().numpy()
    for s, d in enumerate(descs):
        opla, omag, orate = source_np.point_source_ruptures(d)
        want = flat(opla)
        want[:, 2 * 14 + 11] = orate
        sl = slice(b.rup_off[s], b.rup_off[s + 1])
        np.testing.assert_allclose(pla[sl], want, rtol=1e-11, atol=1e-11)
        np.testing.assert_array_equal(mag[sl], omag)


@pytest.mark.gpu
def test_cfg1_area_source_demo(oq):
    """BASELINE.json configs[0]: demos/hazard/AreaSourceClassicalPSHA
    (BooreAtkinson2008, 8 IMTs x 45 levels, 200 km) -- hazard curves of the
    reference's ContextMaker path on the same point sources and sites"""
    d = np.load(os.path.join(GOLDEN, 'cfg1_area_demo.npz'))
    # the area source spreads its MFD evenly over its points
    # (source/area.py: rates scaled by 1 / number of points)
    total = oq.TruncatedGRMFD(*d['gr'])
    mags, rates = zip(*total.get_annual_occurrence_rates())
    np.testing.assert_allclose(mags, d['mags'], rtol=1e-15)
    np.testing.assert_allclose(np.array(rates) / len(d['src_lon']),
                               d['mag_rates'], rtol=1e-14)
    mfd = oq.ArbitraryMFD(d['mags'], d['mag_rates'])
    npd = [tuple(x) for x in d['npd']]
    hdd = [tuple(x) for x in d['hdd']]
    # the point sources come from the area source itself (polygon
    # discretisation on the host, area.py), as in the demo's job.ini
    srcs = list(oq.AreaSource(
        '1', [-0.5, -0.3, 0.1, 0.3], [-0.5, -0.1, 0.2, -0.8], 10., total,
        'WC1994', float(d['rar']), float(d['usd']), float(d['lsd']), npd,
        hdd))
    assert len(srcs) == len(d['src_lon'])
    N = len(d['site_lon'])
    sites = dict(vs30=np.full(N, float(d['vs30'])))
    sc = oq.SiteCollection.from_points(d['site_lon'], d['site_lat'],
                                       np.zeros(N), sites)
    imtls = {str(imt): d['levels'] for imt in d['imts']}
    cm = oq.ContextMaker('Active Shallow Crust',
                         [oq.get_gsim('BooreAtkinson2008')], dict(
        imtls=imtls, truncation_level=float(d['truncation_level']),
        investigation_time=float(d['investigation_time']),
        maximum_distance=oq.IntegrationDistance.new(
            str(float(d['maximum_distance'])))))
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
    kept = cm.hazard_point_sources(pmap, srcs, sc, count=True)
    assert kept == int(d['nrows'])      # the same rows survive rrup < 200 km
    poes = (~pmap).array
    want = d['poes']
    assert poes.shape == want.shape == (N, 8 * 45, 1)
    # (1) against the reference's map.  The reference multiplies its
    # float32 pne map 416 times per site (one rounding of ~3e-8 each, order
    # dependent); the device rounds once.  In pne
    # space that is the rtol 2e-5 of the other pipeline tests; as PoE
    # = 1 - pne it is an absolute 1.5e-6.
    np.testing.assert_allclose(1. - poes, 1. - want, rtol=2e-5, atol=1.5e-7)
    np.testing.assert_allclose(poes, want, atol=1.5e-6)
    top = want >= 1e-2
    np.testing.assert_allclose(poes[top], want[top], rtol=1e-4)
    # hazard curves are non increasing in the level
    assert (np.diff(poes.reshape(N, 8, 45), axis=2) <= 1e-7).all()
    # (2) tight check in rate space (no float32 pne cancellation): the
    # device rate map against a float64 accumulation of the oracle's PoEs
    # over the oracle's own ruptures of the 52 point sources
    rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0.)
    cm.hazard_point_sources(rmap, srcs, sc)
    plas, omags = [], []
    for lon, lat in zip(d['src_lon'], d['src_lat']):
        desc = dict(lon=lon, lat=lat, usd=float(d['usd']), lsd=float(d['lsd']),
                    rar=float(d['rar']), msr='WC1994', mags=d['mags'],
                    mag_rates=d['mag_rates'], npd=npd, hdd=hdd)
        pla, mag, rate = source_np.point_source_ruptures(desc)
        pla['wlr'][:, 2] = rate
        plas.append(pla)
        omags.append(mag)
    pla = {k: np.concatenate([p[k] for p in plas]) for k in plas[0]}
    mag = np.concatenate(omags)
    osites = dict(sids=sc.sids, lon=d['site_lon'], lat=d['site_lat'],
                  depth=np.zeros(N), vs30=sites['vs30'])
    pairs = [('BooreAtkinson2008', {})]
    octx = hazard_np.get_ctx(pairs, pla, mag, osites,
                             np.full(len(mag), float(d['maximum_distance'])))
    assert len(octx) == kept
    acc = hazard_np.exact_rate_sum(pairs, octx, sc.sids, imtls, 3., 50.)
    big = acc > 1e-10
    np.testing.assert_allclose(rmap.array[big], acc[big], rtol=1e-5)
    np.testing.assert_allclose(rmap.array, acc, rtol=1e-5, atol=1e-10)


# ------------------------------------------- pointsource_distance (psdist)
PSD = np.load(os.path.join(GOLDEN, 'psdist.npz'))


def psd_sources(oq):
    s0 = oq.PointSource(
        '0', 12.5, 44.1, oq.EvenlyDiscretizedMFD(
            5.05, 0.3, [0.02, 0.011, 0.006, 0.002, 0.0007, 0.0002, 0.00008,
                        0.00003, 0.00001]),
        'WC1994', 2., 0., 12., [(0.3, 30., 30., 90.), (0.5, 140., 60., -90.),
                                (0.2, 270., 90., 180.)],
        [(0.4, 2.), (0.6, 11.)])
    s1 = oq.PointSource(
        '1', 12.9, 43.8, oq.TruncatedGRMFD(5.0, 7.0, 0.5, 3.2, 0.9), 'WC1994',
        1., 0., 20., [(0.5, 10., 45., 0.), (0.5, 100., 80., 45.)],
        [(1., 10.)])
    return [s0, s1]


def test_psdist_host_matches_reference():
    """PointSource.get_psdist / _get_max_rupture_projection_radius
    (point.py:215-256) and the rupture counts of the approximation"""
    import oq_engine_b200 as oq
    b = oq.PointSourceBatch(psd_sources(oq), float(PSD['psdist']))
    assert list(b.block) == [3 * 2 + 1, 2 * 1 + 1]
    assert list(b.rup_off) == [0, 9 * 7, 9 * 7 + 4 * 3]
    psd = b.psdist(np.full(len(b.mag), float(PSD['maxdist'])))
    np.testing.assert_allclose(psd[0], PSD['src0_psdist'], rtol=1e-14)
    np.testing.assert_allclose(psd[1], PSD['src1_psdist'], rtol=1e-14)
    # rates in device order: the average rupture carries the whole rate
    mag, rate = b.expand()
    blk = rate[:63].reshape(9, 7)
    np.testing.assert_allclose(blk[:, 6], PSD['src0_avg_rate'], rtol=1e-15)
    np.testing.assert_allclose(blk[:, :6].sum(axis=1), PSD['src0_avg_rate'],
                               rtol=1e-14)
    # no approximation: 1000 km (the hazardlib default) or nothing
    for off in (None, 0, 1000.):
        assert oq.PointSourceBatch(psd_sources(oq), off).U == 9 * 6 + 4 * 2


@pytest.mark.gpu
def test_psdist_average_ruptures_golden(oq):
    """the average rupture per magnitude against PointSource.iruptures"""
    h = oq.contexts.get_handle()
    for i, src in enumerate(psd_sources(oq)):
        b = oq.PointSourceBatch([src], float(PSD['psdist']))
        md = np.full(len(b.mag), float(PSD['maxdist']))
        t_pla, t_mag, t_md, t_aux = oq.point_ruptures_device(b, h, md)
        pla, aux = t_pla.cpu().numpy(), t_aux.cpu().numpy()
        B = int(b.block[0])
        avg = pla.reshape(-1, B, 42)[:, B - 1]
        want = PSD['src%d_avg_planar' % i].copy()
        got_rate, avg[:, 2 * 14 + 11] = avg[:, 2 * 14 + 11].copy(), 0.
        want[:, 2 * 14 + 11] = 0.
        np.testing.assert_allclose(avg, want, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(got_rate, PSD['src%d_avg_rate' % i],
                                   rtol=1e-15)
        aux = aux.reshape(-1, B, 5)
        np.testing.assert_allclose(aux[:, 0, 3], PSD['src%d_psdist' % i],
                                   rtol=1e-13)
        assert (aux[:, :B - 1, 4] == 1).all() and (aux[:, B - 1, 4] == 2).all()
        loc = oq.SiteCollection.from_points([src.lon], [src.lat]).xyz[0]
        np.testing.assert_allclose(aux[0, 0, :3], loc, rtol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['many', 'few'])
def test_psdist_hazard_golden(oq, tag):
    """hazard with pointsource_distance = 30 km against the reference's
    get_ctxs -> update: many sites (near / far split per site) and few
    sites (all-or-nothing per magnitude)"""
    N = len(PSD[tag + '_lon'])
    sc = oq.SiteCollection.from_points(
        PSD[tag + '_lon'], PSD[tag + '_lat'], np.zeros(N),
        dict(vs30=PSD[tag + '_vs30']))
    imtls = {str(imt): PSD['levels'] for imt in PSD['imts']}
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('200'),
             pointsource_distance=dict(default=float(PSD['psdist']))))
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
    kept = cm.hazard_point_sources(pmap, psd_sources(oq), sc, count=True)
    assert kept == int(PSD[tag + '_nrows'])
    poes = (~pmap).array
    np.testing.assert_allclose(1. - poes, 1. - PSD[tag + '_poes'], rtol=2e-5,
                               atol=1.5e-7)
    np.testing.assert_allclose(poes, PSD[tag + '_poes'], atol=1.5e-6)
    # without the approximation the map is different (and has more rows)
    cm0 = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('200')))
    pm0 = oq.MapArray(sc.sids, cm0.L, cm0.G).fill(1.)
    kept0 = cm0.hazard_point_sources(pm0, psd_sources(oq), sc, count=True)
    if tag == 'many':
        assert kept0 > kept
        assert np.abs((~pm0).array - poes).max() > 1e-5
    else:   # sites on top of both sources: every magnitude keeps all planes
        assert kept0 == kept
        np.testing.assert_allclose((~pm0).array, poes, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['many', 'few'])
def test_get_ctxs_of_point_sources(oq, tag):
    """the unfused operator on PointSource objects: get_ctxs yields the same
    rows as the reference (count) and get_pmap the same map as the fused
    pass, with the pointsource_distance split"""
    N = len(PSD[tag + '_lon'])
    sc = oq.SiteCollection.from_points(
        PSD[tag + '_lon'], PSD[tag + '_lat'], np.zeros(N),
        dict(vs30=PSD[tag + '_vs30']))
    imtls = {str(imt): PSD['levels'] for imt in PSD['imts']}
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('200'),
             pointsource_distance=float(PSD['psdist'])))
    ctxs = []
    for i, src in enumerate(psd_sources(oq)):
        got = list(cm.get_ctx_iter(src, sc, src_id=i))   # a generator
        assert all(len(np.unique(c.mag)) == 1 for c in got)
        ctxs.extend(got)
    ctx = cm.recarray(ctxs)
    assert len(ctx) == int(PSD[tag + '_nrows'])
    pmap = cm.get_pmap(ctx, oq.PoissonTOM(50.))
    np.testing.assert_allclose(1. - pmap.array, 1. - PSD[tag + '_poes'],
                               rtol=2e-5, atol=1.5e-7)
    # a rupture is seen either by the near or by the far sites, never both
    if tag == 'many':
        rates = {}
        for c in ctxs:
            for sid in np.unique(c.sids):
                tot = c.occurrence_rate[c.sids == sid].sum()
                rates.setdefault((int(c.src_id[0]), float(c.mag[0])), set()
                                 ).add(round(float(tot), 14))
        # every site sees the same total rate per (source, magnitude)
        assert all(len(v) == 1 for v in rates.values()), rates


# ------------------------------------------------------------ area sources
AP = np.load(os.path.join(GOLDEN, 'area_polygons.npz'))


@pytest.mark.parametrize('name', [str(n) for n in AP['names']])
def test_area_discretization_matches_reference(name):
    """Polygon.discretize (geo/polygon.py:232-264): the demo polygon, one
    with edges longer than 100 km (resampled), one across the date line and
    a concave one -- the same points, bit for bit"""
    from oq_engine_b200.area import discretize
    lons, lats = discretize(AP[name + '_poly'][0], AP[name + '_poly'][1],
                            float(AP[name + '_spacing']))
    np.testing.assert_array_equal(lons, AP[name + '_lons'])
    np.testing.assert_array_equal(lats, AP[name + '_lats'])


def test_area_source_yields_the_demo_point_sources():
    """source/area.py:232-290: the 52 point sources of the
    AreaSourceClassicalPSHA demo with the rescaled MFD (a - log10(52))"""
    import oq_engine_b200 as oq
    d = np.load(os.path.join(GOLDEN, 'cfg1_area_demo.npz'))
    area = oq.AreaSource(
        '1', [-0.5, -0.3, 0.1, 0.3], [-0.5, -0.1, 0.2, -0.8], 10.,
        oq.TruncatedGRMFD(*d['gr']), 'WC1994', float(d['rar']),
        float(d['usd']), float(d['lsd']), [tuple(x) for x in d['npd']],
        [tuple(x) for x in d['hdd']])
    pts = list(area)
    np.testing.assert_array_equal([p.lon for p in pts], d['src_lon'])
    np.testing.assert_array_equal([p.lat for p in pts], d['src_lat'])
    mags, rates = zip(*pts[0].get_annual_occurrence_rates())
    np.testing.assert_allclose(mags, d['mags'], rtol=1e-15)
    np.testing.assert_allclose(rates, d['mag_rates'], rtol=1e-15)
    assert all(p.mfd is pts[0].mfd for p in pts)   # one shared slice
    b = oq.PointSourceBatch(pts)
    assert b.U == 52 * 8 and len(b.mag) == 8
    with pytest.raises(ValueError):
        list(oq.AreaSource('2', [0., 0.01, 0.01], [0., 0., 0.01], 50.,
                           oq.TruncatedGRMFD(*d['gr']), 'WC1994', 1., 0., 10.,
                           [(1., 0., 90., 0.)], [(1., 5.)]))


@pytest.mark.gpu
def test_psdist_through_rmap_maker(oq):
    """pointsource_distance reaches the fused pass through RmapMaker (and so
    through classical / calc_hazard_curves) too: same rows and map as
    get_ctxs -> update of the reference with pointsource_distance = 30 km"""
    tag = 'many'
    N = len(PSD[tag + '_lon'])
    sc = oq.SiteCollection.from_points(
        PSD[tag + '_lon'], PSD[tag + '_lat'], np.zeros(N),
        dict(vs30=PSD[tag + '_vs30']))
    imtls = {str(imt): PSD['levels'] for imt in PSD['imts']}
    cm = oq.ContextMaker(
        'Active Shallow Crust',
        [oq.get_gsim('BooreEtAl2014'), oq.get_gsim('BooreAtkinson2008')],
        dict(imtls=imtls, truncation_level=3., investigation_time=50.,
             maximum_distance=oq.IntegrationDistance.new('200'),
             pointsource_distance=float(PSD['psdist'])))
    res = oq.RmapMaker(cm, sc, psd_sources(oq)).make()
    rates = res['rmap'].array
    # the reference's map is float32 pnes: an absolute 6e-8 on each of them
    np.testing.assert_allclose(np.exp(-rates.astype(np.float64)),
                               1. - PSD[tag + '_poes'], rtol=2e-5,
                               atol=1.5e-7)
    exp = -np.log(1. - PSD[tag + '_poes'].astype(np.float64))
    big = exp > 1e-3
    np.testing.assert_allclose(rates[big], exp[big], rtol=2e-4)


def test_recarray_of_no_contexts_is_empty():
    import oq_engine_b200 as oq
    cm = oq.ContextMaker('*', [oq.get_gsim('BooreEtAl2014')],
                         dict(imtls={'PGA': [0.1]}))
    ctx = cm.recarray([])
    assert len(ctx) == 0 and 'rjb' in ctx.dtype.names


@pytest.mark.gpu
def test_reference_manual_contextmaker_example(oq):
    """The ContextMaker example of the reference's manual
    (doc/contributing/developing-with-the-engine.rst:196-333), whose printed
    values are known answers of the whole path: area source 126 (incremental
    MFD of 15 magnitudes, one nodal plane, one depth) -> 47 point sources ->
    705 ruptures, one site at (15.0, 45.2), ToroEtAl2002SHARE, truncation
    level 99, 200 km, 50 years:
        len(ctx) == 705, ctx.mag[0] == 4.7, round(ctx.rrup[0], 1) == 106.4,
        round(ctx.rjb[0], 1) == 105.9, get_pmap(ctx).array == [[[0.00507998]]]
    Polygon discretisation on the host (area.py), everything else on the
    device: contexts through from_srcs, the map through get_pmap AND through
    the fused pass."""
    rates = [1.4731083E-02, 9.2946848E-03, 5.8645496E-03, 3.7002807E-03,
             2.3347193E-03, 1.4731083E-03, 9.2946848E-04, 5.8645496E-04,
             3.7002807E-04, 2.3347193E-04, 1.4731083E-04, 9.2946848E-05,
             1.7588460E-05, 1.1097568E-05, 2.3340307E-06]
    lons = [1.5026169E+01, 1.5650548E+01, 1.6273108E+01, 1.6398742E+01,
            1.5947759E+01, 1.5677179E+01]
    lats = [4.5773603E+01, 4.6176279E+01, 4.6083465E+01, 4.6024744E+01,
            4.5648318E+01, 4.5422577E+01]
    src = oq.AreaSource(
        '126', lons, lats, 10., oq.EvenlyDiscretizedMFD(4.7, .2, rates),
        'WC1994', 1., 0., 30., [(1., 6.9033586E+01, 5.7596810E+01, 0.)],
        [(1., 1.0200000E+01)])
    srcs = list(src)
    assert len(srcs) == 47
    assert len(src.mfd.get_annual_occurrence_rates()) == 15
    sc = oq.SiteCollection.from_points(
        [15.0], [45.2], [0.], dict(vs30=600., vs30measured=True))
    cm = oq.ContextMaker('*', [oq.get_gsim('ToroEtAl2002SHARE')], dict(
        imtls={'PGA': [0.1]}, truncation_level=99., investigation_time=50.,
        maximum_distance=oq.IntegrationDistance.new('200.0')))
    ctx = cm.from_srcs(srcs, sc)
    assert len(ctx) == 705
    assert float(ctx.mag[0]) == 4.7
    assert float(round(ctx.rrup[0], 1)) == 106.4
    assert float(round(ctx.rjb[0], 1)) == 105.9
    np.testing.assert_array_equal(np.unique(ctx.sids), [0])
    mean, sig, tau, phi = cm.get_mean_stds([ctx])
    assert mean.shape == sig.shape == (1, 1, 705)
    got = cm.get_pmap(ctx).array
    assert got.shape == (1, 1, 1) and got.dtype == np.float32
    np.testing.assert_allclose(got, [[[0.00507998]]], rtol=2e-6)
    # the fused pass over the same sources
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1.)
    assert cm.hazard_point_sources(pmap, srcs, sc, count=True) == 705
    np.testing.assert_allclose((~pmap).array, [[[0.00507998]]], rtol=2e-5)


@pytest.mark.gpu
def test_reference_known_pointsource_distance_rrups(oq):
    """The engine QA test classical/case_48 (calculators/tests/
    classical_test.py:647-704): a point source with 10 magnitudes, PeerMSR
    and 2 hypocentral depths, two sites at 55 and 111 km, maximum distance
    110 km.  The reference's test holds the rupture distances of the context
    rows, rounded to 0.1 km: the 20 exact ones per site, and the 15 left
    with pointsource_distance = 50 km (for the magnitudes whose psdist the
    site is beyond, the two depths collapse into the average rupture).
    Rupture generation, distances and the near / far selection on the
    device."""
    expect = np.array([
        [54.3, 109.8], [54.1, 109.7], [53.9, 109.4], [53.7, 109.3],
        [53.3, 108.9], [53.3, 108.8], [52.7, 108.3], [52.7, 108.2],
        [52.0, 107.5], [51.9, 107.5], [50.5, 106.1], [50.4, 106.0],
        [47.8, 103.2], [47.7, 103.1], [43.7, 98.8], [43.6, 98.6],
        [38.2, 92.0], [38.0, 91.9], [33.0, 82.3], [32.8, 82.2]])
    approx = np.array([
        [54.2, 109.7], [53.8, 109.3], [53.3, 108.8], [52.7, 108.2],
        [51.9, 107.5], [50.5, 106.1], [50.4, 106.0], [47.8, 103.2],
        [47.7, 103.1], [43.7, 98.8], [43.6, 98.6], [38.2, 92.0],
        [38.0, 91.9], [33.0, 82.3], [32.8, 82.2]])
    src = oq.PointSource(
        '1', 0., 0., oq.TruncatedGRMFD(5., 7., 0.2, 3., 1.), 'PeerMSR', 1.,
        0., 10., [(1., 144., 87., -179.)],
        [(.66666667, 3.), (.33333333, 6.)])
    sc = oq.SiteCollection.from_points(
        [0., 0.], [0.5, 1.], [0., 0.],
        dict(vs30=800., vs30measured=True, z1pt0=50., z2pt5=2.5))
    for psd, want in ((1000., expect), (50., approx)):
        cm = oq.ContextMaker('*', [oq.SadighEtAl1997()], dict(
            imtls={'PGA': [0.1]}, truncation_level=2., investigation_time=100.,
            maximum_distance=oq.IntegrationDistance.new('110.0'),
            pointsource_distance=psd))
        ctx = cm.recarray(cm.get_ctxs(src, sc))
        for sid in (0, 1):
            got = np.sort(ctx.rrup[ctx.sids == sid])[::-1]
            assert len(got) == len(want), (psd, sid, len(got))
            np.testing.assert_allclose(got, want[:, sid], atol=.05,
                                       err_msg='psd %s site %d' % (psd, sid))
