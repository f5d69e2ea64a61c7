"""GPU parity tests (run on a B200 with `-m gpu`): the CUDA path, called
through the C ABI (ctypes) and through the reference-like Python operators,
against the golden fixtures produced by the reference and against the numpy
oracle on seeded inputs.

Tolerances: mean/sig/tau/phi float64 rtol 1e-10 (the device uses CUDA's
libm, 1-2 ulp per call); distances rtol 1e-9 / atol 1e-8 km (the reference's
own formulas are conditioned to ~1e-12 relative); PoEs float32 rtol 1e-5
where PoE >= 1e-8 (BASELINE.json north_star); maps rtol 1e-5 against a
float64 re-accumulation of the oracle's PoEs and rtol 2e-5 against the
reference's float32 read-modify-write maps.
"""
import json
import ctypes
import numpy as np
import pytest

import geo_np
import hazard_np
import poe_np
from helpers import (GOLDEN, load, gmpe_cases, pipeline_cases, ctx_of, rup_of,
                     sites_of, json_kwargs, synth_ruptures, synth_sites, Bag)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()   # fails loudly if the extension is missing
    return oq_engine_b200


def make_gsims(oq, pairs):
    return [oq.get_gsim(name, **kw) for name, kw in pairs]


def to_recarray(bag):
    names = list(vars(bag))
    dt = [(n, vars(bag)[n].dtype, vars(bag)[n].shape[1:]) for n in names]
    arr = np.zeros(len(bag), dt)
    for n in names:
        arr[n] = vars(bag)[n]
    return arr.view(np.recarray)


def planar_of(oq, d):
    r = rup_of(d)
    pla = oq.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], float(d['usd']),
        float(d['lsd']), float(d['rar']))
    return r, pla


def sitecol_of(oq, d):
    s = sites_of(d)
    return oq.SiteCollection.from_points(s['lon'], s['lat'], s['depth'], s)


# ------------------------------------------------------------------ GMPEs
@pytest.mark.parametrize('tag', gmpe_cases())
def test_gmpe_golden(oq, tag):
    d = load('gmpe_' + tag)
    gsim = oq.get_gsim(str(d['name']), **json_kwargs(d))
    ctx = to_recarray(ctx_of(d))
    imts = [str(i) for i in d['imts']]
    M, N = len(imts), len(ctx)
    out = np.zeros((4, M, N))
    gsim.compute(ctx, imts, *out)       # the reference operator
    np.testing.assert_allclose(out, d['mean_stds'], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('tag', gmpe_cases())
def test_gmpe_staged_golden(oq, tag):
    """the per-site / per-rupture / per-pair / per-IMT staging used inside
    the fused kernel reproduces the reference's mean and total sigma"""
    d = load('gmpe_' + tag)
    gsim = oq.get_gsim(str(d['name']), **json_kwargs(d))
    ctx = to_recarray(ctx_of(d))
    imts = [str(i) for i in d['imts']]
    out = oq.contexts.mean_stds_staged_device([gsim], ctx, imts)
    np.testing.assert_allclose(out[:, 0], d['mean_stds'][:2], rtol=1e-10,
                               atol=1e-12)


def test_gmpe_unsupported_is_an_error(oq):
    with pytest.raises(NotImplementedError):
        oq.get_gsim('ZhaoEtAl2006Asc')
    with pytest.raises(NotImplementedError):
        oq.BooreEtAl2014(usgs_basin_scaling=True)
    with pytest.raises(KeyError):   # IMT outside the table
        oq.ContextMaker('*', [oq.AtkinsonBoore2006()],
                        dict(imtls={'SA(9.0)': [0.1]}))


# -------------------------------------------------------------- distances
def test_planar_golden(oq):
    d = load('planar')
    r, pla = planar_of(oq, d)
    for f in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr', 'sdr', 'hypo'):
        np.testing.assert_allclose(pla[f], d['planar_' + f], rtol=0,
                                   atol=1e-10, err_msg=f)
    sc = sitecol_of(oq, d)
    gsims = [oq.AbrahamsonEtAl2014(), oq.AkkarEtAlRepi2014(),
             oq.AkkarEtAlRhyp2014()]
    cm = oq.ContextMaker('*', gsims, dict(imtls={'PGA': [0.1]}))
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    ctxs = cm.get_ctxs(src, sc)
    ctx = cm.recarray(ctxs)
    assert len(ctx) == len(r['mag']) * len(sc)
    for name in ('rrup', 'rjb', 'rx', 'ry0', 'repi', 'rhypo'):
        got = np.zeros((len(r['mag']), len(sc)))
        got[ctx.rup_id, ctx.sids] = ctx[name]
        np.testing.assert_allclose(got, d['dist_' + name], rtol=1e-9,
                                   atol=1e-8, err_msg=name)


def test_distances_c_abi_all_metrics(oq):
    import torch
    d = load('planar')
    r, pla = planar_of(oq, d)
    s = sites_of(d)
    h = oq.contexts.get_handle()
    dev = 'cuda:%d' % h.device
    U, N = len(pla), len(s['lon'])
    t_pla = torch.from_numpy(oq.planar.as_flat(pla)).to(dev)
    st = oq._lib.OqgSites()
    keep = []
    for name in ('lon', 'lat', 'depth'):
        t = torch.from_numpy(np.ascontiguousarray(s[name])).to(dev)
        keep.append(t)
        setattr(st, name, t.data_ptr())
    out = torch.empty((9, U, N), dtype=torch.float64, device=dev)
    h.call('oqg_distances', ctypes.c_void_p(t_pla.data_ptr()), U,
           ctypes.byref(st), N, 0x1ff, ctypes.c_void_p(out.data_ptr()))
    out = out.cpu().numpy()
    for k, name in enumerate(oq._lib.DIST):
        exp = d['dist_' + name]
        if name == 'azimuth':   # circular quantity
            diff = np.abs((out[k] - exp + 180) % 360 - 180)
            assert diff.max() < 1e-6, diff.max()
        elif name in ('clon', 'clat'):   # degrees: 1e-10 deg = 1e-8 km
            np.testing.assert_allclose(out[k], exp, rtol=0, atol=1e-10,
                                       err_msg=name)
        else:
            np.testing.assert_allclose(out[k], exp, rtol=1e-9, atol=1e-8,
                                       err_msg=name)


def test_few_sites_rows_carry_the_closest_points(oq):
    """contexts.py:385-386, 425-427: with at most max_sites_disagg sites
    the context rows have the clon / clat columns (project_back)"""
    d = load('planar')
    r, pla = planar_of(oq, d)
    s = sites_of(d)
    few = {k: v[:6] for k, v in s.items()}
    sc = oq.SiteCollection.from_points(few['lon'], few['lat'], few['depth'],
                                       few)
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()], dict(imtls={'PGA': [0.1]}))
    ctx = cm.recarray(cm.get_ctxs(oq.contexts.PlanarRuptures(pla, r['mag']),
                                  sc))
    assert cm.fewsites and {'clon', 'clat'} <= set(ctx.dtype.names)
    for name in ('clon', 'clat', 'rrup', 'rjb'):
        got = np.zeros((len(r['mag']), 6))
        got[ctx.rup_id, ctx.sids] = ctx[name]
        np.testing.assert_allclose(got, d['dist_' + name][:, :6], rtol=1e-9,
                                   atol=1e-8, err_msg=name)
    # many sites: no such columns
    many = sitecol_of(oq, d)
    ctx = cm.recarray(cm.get_ctxs(oq.contexts.PlanarRuptures(pla, r['mag']),
                                  many))
    assert not cm.fewsites and 'clon' not in ctx.dtype.names


def test_distances_random_vs_oracle(oq):
    rng = np.random.default_rng(11)
    r = synth_ruptures(300, rng, box=3.0)
    sites = synth_sites(500, rng, box=3.0)
    sites['depth'] = rng.uniform(0, 2, 500) * (rng.random(500) < .3)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    opla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    for f in ('corners', 'xyz', 'normal', 'uv1', 'uv2', 'wlr'):
        np.testing.assert_allclose(pla[f], opla[f], rtol=0, atol=1e-9)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    cm = oq.ContextMaker('*', [oq.AbrahamsonEtAl2014(),
                               oq.AkkarEtAlRhyp2014(),
                               oq.AkkarEtAlRepi2014()],
                         dict(imtls={'PGA': [0.1]}))
    ctx = cm.recarray(cm.get_ctxs(oq.contexts.PlanarRuptures(pla, r['mag']),
                                  sc))
    xyz = hazard_np.site_xyz(sites)
    for name in ('rrup', 'rjb', 'rx', 'ry0', 'repi', 'rhypo'):
        exp = geo_np.DIST[name](opla, xyz)
        got = np.zeros_like(exp)
        got[ctx.rup_id, ctx.sids] = ctx[name]
        np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-8,
                                   err_msg=name)


# ------------------------------------------------- unfused pipeline pieces
def cmaker_of(oq, d, **extra):
    gsims = make_gsims(oq, [tuple(g) for g in json.loads(str(d['gsims']))])
    imtls = {str(i): d['imls'] for i in d['imts']}
    nb = json.loads(str(d['negbin']))
    cm = oq.ContextMaker('*', gsims, dict(
        imtls=imtls, truncation_level=float(d['truncation_level']),
        investigation_time=float(d['investigation_time']),
        maximum_distance=float(d['maxdist']),
        minimum_distance=float(d['mindist']), **extra))
    itime = float(d['investigation_time'])
    cm.tom = (oq.NegativeBinomialTOM(itime, **nb) if nb
              else oq.PoissonTOM(itime))
    return cm


@pytest.mark.parametrize('tag', pipeline_cases())
def test_pipeline_unfused_golden(oq, tag):
    d = load('pipeline_' + tag)
    cm = cmaker_of(oq, d)
    r, pla = planar_of(oq, d)
    sc = sitecol_of(oq, d)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    ctx = cm.recarray(cm.get_ctxs(src, sc))
    ref = to_recarray(ctx_of(d))
    assert len(ctx) == len(ref)
    # the lazy form with a small host budget: the same rows, block by block
    cm.GET_CTXS_BYTES = 4 * len(sc) * 8 * 5
    lazy = cm.recarray(list(cm.get_ctx_iter(src, sc)))
    del cm.GET_CTXS_BYTES
    assert lazy.tobytes() == ctx.tobytes()
    # same order: by magnitude, then rupture, then site
    np.testing.assert_array_equal(ctx.sids, ref.sids)
    for name in ctx.dtype.names:
        if name in ('src_id', 'rup_id', 'sids'):
            continue
        np.testing.assert_allclose(ctx[name], ref[name], rtol=1e-9,
                                   atol=1e-8, err_msg=name)
    # mean and stds on the reference's rows
    ms = cm.get_mean_stds([ref], split_by_mag=False)
    np.testing.assert_allclose(ms, d['mean_stds'], rtol=1e-10, atol=1e-12)
    # PoEs
    poes = np.concatenate([p for p, *_ in cm.gen_poes(ref.copy().view(
        np.recarray))])
    exp = d['poes']
    big = exp > 1e-8
    np.testing.assert_allclose(poes[big], exp[big], rtol=1e-5)
    assert np.abs(poes - exp).max() < 1e-7
    # gen_poes never holds more than GEN_POES_BYTES of PoEs: with a small
    # budget the same rows come in more, smaller blocks of one magnitude
    cm.GEN_POES_BYTES = 40 * cm.L * cm.G * 4
    blocks = list(cm.gen_poes(ref.copy().view(np.recarray)))
    assert max(len(b[0]) for b in blocks) <= 40
    assert all(len(np.unique(b[4].mag)) == 1 for b in blocks)
    assert all(b[1].shape == (cm.G, cm.M, len(b[0])) for b in blocks)
    np.testing.assert_array_equal(np.concatenate([b[0] for b in blocks]),
                                  poes)
    del cm.GEN_POES_BYTES
    # maps
    N, L, G = d['pmap'].shape
    pmap = oq.MapArray(sc.sids, L, G).fill(1)
    cm.update(pmap, ref)
    np.testing.assert_allclose(pmap.array, d['pmap'], rtol=2e-5, atol=1e-7)
    rmap = oq.MapArray(sc.sids, L, G, rates=True).fill(0)
    cm.update(rmap, ref)
    np.testing.assert_allclose(rmap.array, d['rmap'], rtol=2e-5, atol=1e-9)
    # get_pmap = 1 - pne map
    # 1 - pne in float32: absolute error 2e-5 * pne
    np.testing.assert_allclose(cm.get_pmap(ref).array, 1. - d['pmap'],
                               rtol=0, atol=2e-5)


@pytest.mark.parametrize('tag', pipeline_cases())
def test_pipeline_fused_golden(oq, tag):
    d = load('pipeline_' + tag)
    cm = cmaker_of(oq, d)
    r, pla = planar_of(oq, d)
    sc = sitecol_of(oq, d)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    N, L, G = d['pmap'].shape
    pmap = oq.MapArray(sc.sids, L, G).fill(1)
    kept = cm.hazard_planar(pmap, src, sc, count=True)
    assert kept == len(d['ctx_sids'])
    np.testing.assert_allclose(pmap.array, d['pmap'], rtol=2e-5, atol=1e-7)
    rmap = oq.MapArray(sc.sids, L, G, rates=True).fill(0)
    cm.hazard_planar(rmap, src, sc)
    np.testing.assert_allclose(rmap.array, d['rmap'], rtol=2e-5, atol=1e-9)
    if not json.loads(str(d['negbin'])):
        # order independent check against the float64 sum of the
        # reference's own float32 PoEs
        ref = ctx_of(d)
        acc = np.zeros((N, L, G))
        x = ref.occurrence_rate[:, None, None] * d['poes'] * float(
            d['investigation_time'])
        np.add.at(acc, ref.sids, x)
        big = acc > 1e-12
        np.testing.assert_allclose(rmap.array[big], acc[big], rtol=1e-5)


def test_truncnorm_edges(oq):
    """truncnorm_sf edge cases of the reference through oqg_poes"""
    import torch
    with open(GOLDEN + '/truncnorm.json') as f:
        cases = json.load(f)
    h = oq.contexts.get_handle()
    dev = 'cuda:%d' % h.device
    for c in cases:
        z = np.array(c['z'])
        exp = np.array([np.nan if x is None else x for x in c['sf']])
        n = len(z)
        cm = oq.ContextMaker('*', [oq.BooreAtkinson2008()], dict(
            imtls={'PGA': [1.0]}, truncation_level=c['truncation_level']))
        hh = cm.handle()
        # mean = -z, sig = 1, level log(1) = 0  ->  (0 - mean) / sig = z
        ms = np.zeros((4, 1, 1, n))
        ms[0, 0, 0] = -z
        ms[1, 0, 0] = 1.
        t_ms = torch.from_numpy(ms).to(dev)
        out = torch.empty((n, 1, 1), dtype=torch.float32, device=dev)
        hh.call('oqg_poes', ctypes.c_void_p(t_ms.data_ptr()), n,
                ctypes.c_void_p(out.data_ptr()))
        got = out.cpu().numpy().ravel()
        ok = np.isfinite(exp)
        np.testing.assert_allclose(got[ok], exp[ok].astype(np.float32),
                                   rtol=2e-6, atol=1e-30)
        assert np.isnan(got[~ok]).all()


# ------------------------------------------ random larger cases vs oracle
CFG = {
    'cfg3': [('BooreEtAl2014', {}), ('ChiouYoungs2014', {})],
    'cfg4': [('BooreEtAl2014', {}), ('ChiouYoungs2014', {}),
             ('AbrahamsonEtAl2014', {}), ('CampbellBozorgnia2014', {})],
    'mixed': [('BooreAtkinson2008', {}), ('AkkarEtAlRjb2014', {}),
              ('AtkinsonBoore2006', {}), ('AkkarEtAlRhyp2014', {}),
              ('AkkarEtAlRepi2014', {}),
              ('BooreEtAl2014', dict(region='CAL', sigma_mu_epsilon=1.)),
              ('ChiouYoungs2014', dict(region='JPN'))],
}
IMTS = ['PGA', 'SA(0.1)', 'SA(0.2)', 'SA(0.5)', 'SA(1.0)', 'SA(2.0)',
        'SA(3.0)']


@pytest.mark.parametrize('cfg,U,N,maxdist,mindist,L1', [
    ('cfg3', 257, 333, 1000., 0., 20),
    ('cfg4', 130, 190, 1000., 0., 20),
    ('cfg3', 200, 100, 80., 3., 20),      # distance filter + clamp
    ('mixed', 96, 65, 150., 0., 20),
    ('cfg3', 64, 70, 1000., 0., 45),      # three level chunks per IMT
    ('cfg3', 33, 31, 1000., 0., 7),       # short level grid, ragged tiles
    ('cfg4', 1, 1, 1000., 0., 20),        # single pair
])
def test_fused_vs_oracle(oq, cfg, U, N, maxdist, mindist, L1):
    rng = np.random.default_rng(U * 1000 + N)
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng, sentinels=True)
    imls = np.logspace(-2, 0.3, L1)
    imtls = {imt: imls for imt in IMTS}
    pairs = CFG[cfg]
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.,
        maximum_distance=maxdist, minimum_distance=mindist))
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    L, G = len(IMTS) * L1, len(pairs)
    rmap = oq.MapArray(sc.sids, L, G, rates=True).fill(0)
    kept = cm.hazard_planar(rmap, src, sc, count=True)
    # oracle
    opla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    octx = hazard_np.get_ctx(pairs, opla, r['mag'], sites,
                             np.full(U, maxdist), mindist)
    assert kept == len(octx)
    acc = hazard_np.exact_rate_sum(pairs, octx, sites['sids'], imtls, 3., 50.)
    big = acc > 1e-10
    np.testing.assert_allclose(rmap.array[big], acc[big], rtol=1e-5)
    assert np.abs(rmap.array - acc).max() <= 1e-5 * acc.max() + 1e-12
    # pne mode and host-buffer entry point give the same map
    pmap = oq.MapArray(sc.sids, L, G).fill(1)
    cm.hazard_planar(pmap, src, sc)
    np.testing.assert_allclose(pmap.array, np.exp(-acc), rtol=1e-5)
    # the unfused operators agree with the fused kernel
    ctx = cm.recarray(cm.get_ctxs(src, sc)) if len(octx) else None
    if ctx is not None:
        rmap2 = oq.MapArray(sc.sids, L, G, rates=True).fill(0)
        cm.update(rmap2, ctx)
        np.testing.assert_allclose(rmap2.array, rmap.array, rtol=2e-6,
                                   atol=1e-12)


def test_fused_host_entry_point(oq):
    """oqg_hazard_planar_host: host buffers in, host map out"""
    rng = np.random.default_rng(5)
    U, N = 100, 77
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    cm.hazard_planar(pmap, src, sc)
    h = cm.handle()
    flat = oq.planar.as_flat(pla)
    mag = np.ascontiguousarray(r['mag'])
    md = cm.magdist(r['mag'])
    st = oq._lib.OqgSites()
    keep = {}
    for name in ('lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5'):
        keep[name] = np.ascontiguousarray(sites[name], float)
        setattr(st, name, keep[name].ctypes.data)
    keep['m'] = np.ascontiguousarray(sites['vs30measured']).view(np.uint8)
    st.vs30measured = keep['m'].ctypes.data
    out = np.ones((N, cm.L, cm.G), np.float32)
    kept = ctypes.c_ulonglong(0)
    h.call('oqg_hazard_planar_host', flat.ctypes.data, mag.ctypes.data,
           md.ctypes.data, None, 0, U, ctypes.byref(st), N, 0., 50., 0,
           out.ctypes.data, ctypes.byref(kept))
    assert kept.value == U * N
    np.testing.assert_array_equal(out, pmap.array)
    # the same ruptures as a stream of 5 sources through one session
    # (oqg_host_begin / add / end: sites and map stay on the device, the
    # uploads are double-buffered): the calls one after the other on a map
    # kept by the caller give the same bits
    one = np.ones((N, cm.L, cm.G), np.float32)
    cuts = [0, 7, 40, 41, 90, U]
    for a, b in zip(cuts[:-1], cuts[1:]):
        h.call('oqg_hazard_planar_host', flat[a:].ctypes.data,
               mag[a:].ctypes.data, md[a:].ctypes.data, None, 0, b - a,
               ctypes.byref(st), N, 0., 50., 0, one.ctypes.data, None)
    for init in (None, np.ones((N, cm.L, cm.G), np.float32)):
        h.call('oqg_host_begin', ctypes.byref(st), N, 0,
               None if init is None else init.ctypes.data)
        with pytest.raises(RuntimeError):
            h.call('oqg_host_begin', ctypes.byref(st), N, 0, None)
        for a, b in zip(cuts[:-1], cuts[1:]):
            # a fresh copy per call, dropped right after: `add` returns when
            # the caller's buffers can be reused
            f, m, d = flat[a:b].copy(), mag[a:b].copy(), md[a:b].copy()
            h.call('oqg_host_add', f.ctypes.data, m.ctypes.data,
                   d.ctypes.data, None, 0, b - a, 0., 50.)
            f[:] = m[:] = d[:] = np.nan
        got = np.zeros((N, cm.L, cm.G), np.float32)
        kept = ctypes.c_ulonglong(0)
        h.call('oqg_host_end', got.ctypes.data, ctypes.byref(kept))
        assert kept.value == U * N
        np.testing.assert_array_equal(got, one)
    np.testing.assert_allclose(one, pmap.array, rtol=2e-6)
    with pytest.raises(RuntimeError):
        h.call('oqg_host_add', flat.ctypes.data, mag.ctypes.data,
               md.ctypes.data, None, 0, U, 0., 50.)


def test_fused_negbin_vs_oracle(oq):
    rng = np.random.default_rng(77)
    U, N = 60, 50
    r = synth_ruptures(U, rng)
    r['rate'] = r['rate'] * 50   # make the non-Poissonian part visible
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pairs = CFG['cfg3']
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    cm.tom = oq.NegativeBinomialTOM(50., mu=1., alpha=0.5)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    pmap = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    cm.hazard_planar(pmap, src, sc)
    opla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    po = cm.tom.get_pmf(opla['wlr'][:, 2])
    octx = hazard_np.get_ctx(pairs, opla, r['mag'], sites, None, 0, po)
    poes, _ = hazard_np.get_poes(pairs, octx, imtls, 3.)
    logp = np.zeros((N, cm.L, cm.G))
    for i in range(len(octx)):
        pne = poe_np.get_pnes(octx.occurrence_rate[i], octx.probs_occur[i],
                              poes[i].astype(float), 50.)
        logp[octx.sids[i]] += np.log(pne)
    np.testing.assert_allclose(pmap.array, np.exp(logp), rtol=1e-5)
    assert pmap.array.min() < 0.9


def test_map_on_device_and_sidx(oq):
    """MapArray on the device, sites given as a subset with their own sids"""
    rng = np.random.default_rng(3)
    U, N = 50, 64
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    full = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    cm.hazard_planar(full, src, sc)
    # two tiles (sids % 2) into a device-resident map
    dmap = oq.MapArray(sc.sids, cm.L, cm.G, device='cuda:0').fill(1)
    for tile in sc.split(2):
        cm.hazard_planar(dmap, src, tile)
    np.testing.assert_array_equal(dmap.numpy(), full.array)
    # two halves of the ruptures multiply
    half = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    for sl in (slice(0, 20), slice(20, U)):
        cm.hazard_planar(half, oq.contexts.PlanarRuptures(
            pla[sl], r['mag'][sl]), sc)
    np.testing.assert_allclose(half.array, full.array, rtol=1e-6)
    # the monitors: one fused pass per hazard_planar call so far; the unfused
    # operators count under the reference's names (contexts.py:1161, 1171)
    assert cm.fused_mon.counts == 5 and cm.fused_mon.duration > 0
    ctxs = cm.get_ctxs(src, sc)
    cm.get_pmap(cm.recarray(ctxs))
    assert cm.ir_mon.counts == 1
    assert cm.gmf_mon.counts == 1 and cm.poe_mon.counts == 1


def test_fast_poe_table_vs_exact(oq):
    """The float32 polynomial table of the truncated-normal survival
    function used by the fused kernel against the FP64 erf formula: the
    table is in use for truncation_level=3 with a measured error <= 2e-6,
    both paths agree on the map, and an untabulated truncation level (99)
    falls back to the exact formula"""
    rng = np.random.default_rng(99)
    U, N = 150, 130
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng, sentinels=True)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    L_ = oq._lib.lib()
    for tl, expect_table in ((3., True), (2., True), (99., False)):
        cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
            imtls=imtls, truncation_level=tl, investigation_time=50.))
        h = cm.handle()
        in_use, err = ctypes.c_int(0), ctypes.c_double(0)
        L_.oqg_poe_table_info(h._h, ctypes.byref(in_use), ctypes.byref(err))
        assert bool(in_use.value) == expect_table, (tl, err.value)
        if expect_table:
            assert err.value <= 2e-6
        fast = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
        cm.hazard_planar(fast, src, sc)
        L_.oqg_set_exact_poe(h._h, 1)
        try:
            exact = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
            cm.hazard_planar(exact, src, sc)
        finally:
            L_.oqg_set_exact_poe(h._h, 0)
        big = exact.array > 1e-12
        np.testing.assert_allclose(fast.array[big], exact.array[big],
                                   rtol=5e-6)


def test_sharding_on_one_gpu(oq):
    """The multi-GPU decompositions executed rank after rank on one GPU:
    site tiles assemble to the unsharded map; the product of per-rank pne
    maps (sum of rate maps) over source groups equals one pass over all
    the groups"""
    import torch
    rng = np.random.default_rng(21)
    N = 90
    sites = synth_sites(N, rng)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    srcs = []
    for g in range(4):
        r = synth_ruptures(40, np.random.default_rng(50 + g))
        pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                              r['rate'], r['area'], r['hypo_lon'],
                              r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
        srcs.append(oq.PlanarRuptures(pla, r['mag']))
    full = oq.MapArray(sc.sids, cm.L, cm.G).fill(1)
    for src in srcs:
        cm.hazard_planar(full, src, sc)
    world = 3
    # site sharding
    out = np.zeros_like(full.array)
    for rank in range(world):
        tmap = None
        for src in srcs:
            if tmap is None:
                tmap, tile = oq.sharding.hazard_site_sharded(
                    cm, src, sc, rank, world)
            else:
                cm.hazard_planar(tmap, src, tile)
        gathered = oq.sharding.gather_tiles(tmap.array, tile.sids, N,
                                            partial=True)
        out[tile.sids] = gathered.cpu().numpy()[tile.sids]
    np.testing.assert_array_equal(out, full.array)
    # source-group sharding (the all-reduce is the identity at world 1)
    prod = torch.ones_like(tmap.array.new_empty(full.array.shape))
    rates = torch.zeros_like(prod)
    for rank in range(world):
        for mode, acc in ((False, prod), (True, rates)):
            pm = oq.MapArray(sc.sids, cm.L, cm.G, rates=mode,
                             device='cuda:0').fill(0 if mode else 1)
            for g in oq.sharding.groups_of(rank, world, len(srcs)):
                cm.hazard_planar(pm, srcs[g], sc)
            if mode:
                acc += pm.array
            else:
                acc *= pm.array
    np.testing.assert_allclose(prod.cpu().numpy(), full.array, rtol=1e-6)
    np.testing.assert_allclose(np.exp(-rates.cpu().numpy()), full.array,
                               rtol=1e-5)
    pm = oq.sharding.hazard_group_sharded([cm] * 4, srcs, sc, 0, 1)
    np.testing.assert_allclose(pm.numpy(), full.array, rtol=1e-6)


def test_full_size_properties(oq):
    """Size-independent properties of the fused pass at a size the oracle
    cannot reach (20 000 ruptures x 10 000 sites, cfg3 models: 5.6e10 PoE
    evaluations): additivity over a split of the ruptures, linearity in the
    occurrence rates, invariance under a permutation of the sites,
    monotonicity of the curves, and a spot check of 24 sites against the
    oracle."""
    rng = np.random.default_rng(2024)
    U, N = 20_000, 10_000
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pairs = CFG['cfg3']
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)

    def planar(sel, scale=1.):
        pla = oq.build_planar(
            r['mag'][sel], r['strike'][sel], r['dip'][sel], r['rake'][sel],
            r['rate'][sel] * scale, r['area'][sel], r['hypo_lon'][sel],
            r['hypo_lat'][sel], r['hypo_dep'][sel], 0., 25., 1.5)
        return oq.PlanarRuptures(pla, r['mag'][sel])

    def rates(src, sitecol):
        m = oq.MapArray(sitecol.sids, cm.L, cm.G, rates=True).fill(0)
        kept = cm.hazard_planar(m, src, sitecol, count=True)
        return m.array, kept
    everything = slice(None)
    full, kept = rates(planar(everything), sc)
    assert kept == U * N
    assert np.isfinite(full).all() and (full >= 0).all()
    # additivity: ruptures split in two uneven parts
    a, _ = rates(planar(slice(0, 7_003)), sc)
    b, _ = rates(planar(slice(7_003, U)), sc)
    np.testing.assert_allclose(a + b, full, rtol=3e-6, atol=1e-12)
    # linearity in the rates
    twice, _ = rates(planar(everything, 2.), sc)
    np.testing.assert_allclose(twice, 2 * full, rtol=1e-6, atol=1e-12)
    # permutation of the sites
    perm = rng.permutation(N)
    psites = {k: v[perm] for k, v in sites.items() if k != 'sids'}
    psc = oq.SiteCollection.from_points(psites['lon'], psites['lat'],
                                        psites['depth'], psites)
    permuted, _ = rates(planar(everything), psc)
    np.testing.assert_allclose(permuted, full[perm], rtol=1e-6, atol=1e-12)
    # rates of exceedance do not increase with the level
    curves = full.reshape(N, len(IMTS), 20, cm.G)
    assert (np.diff(curves, axis=2) <= 1e-7 * curves[:, :, :-1]).all()
    # spot check against the oracle: 24 sites x the first 300 ruptures
    sel = rng.choice(N, 24, replace=False)
    sub = {k: v[sel] for k, v in sites.items()}
    sub['sids'] = np.arange(24, dtype=np.uint32)
    ssc = oq.SiteCollection.from_points(sub['lon'], sub['lat'], sub['depth'],
                                        sub)
    got, _ = rates(planar(slice(0, 300)), ssc)
    opla = geo_np.build_planar(
        r['mag'][:300], r['strike'][:300], r['dip'][:300], r['rake'][:300],
        r['rate'][:300], r['area'][:300], r['hypo_lon'][:300],
        r['hypo_lat'][:300], r['hypo_dep'][:300], 0., 25., 1.5)
    octx = hazard_np.get_ctx(pairs, opla, r['mag'][:300], sub)
    acc = hazard_np.exact_rate_sum(pairs, octx, sub['sids'], imtls, 3., 50.)
    big = acc > 1e-10
    np.testing.assert_allclose(got[big], acc[big], rtol=1e-5)


def test_get_mean_stds_keeps_the_row_order(oq):
    """split_by_mag evaluates runs of equal magnitude (split_array on ordered
    indices): the columns stay in the order of the rows"""
    d = load('gmpe_bssa14')
    ctx = to_recarray(ctx_of(d))[:50]
    assert len(np.unique(ctx.mag)) > 3 and (np.diff(ctx.mag) < 0).any()
    cm = oq.ContextMaker('*', [oq.BooreEtAl2014()], dict(
        imtls={'PGA': [0.1], 'SA(1.0)': [0.1]}))
    a = cm.get_mean_stds([ctx], split_by_mag=True)
    b = cm.get_mean_stds([ctx], split_by_mag=False)
    np.testing.assert_array_equal(a, b)
    rev = cm.get_mean_stds([ctx[::-1]])
    np.testing.assert_array_equal(rev[..., ::-1], a)


def test_reference_known_ask14_stddevs(oq):
    """contexts_test.py:384-387 (test_from_planar): sig, tau, phi of
    AbrahamsonEtAl2014 for PGA are .79162428, .47, .637"""
    from test_oracle import ask14_from_planar_rows
    ctx = to_recarray(ask14_from_planar_rows())
    cm = oq.ContextMaker('*', [oq.AbrahamsonEtAl2014()],
                         dict(imtls={'PGA': [0.1]}))
    mea, sig, tau, phi = cm.get_mean_stds([ctx])
    np.testing.assert_allclose(sig, .79162428, rtol=1e-8)
    np.testing.assert_allclose(tau, .47, rtol=1e-12)
    np.testing.assert_allclose(phi, .637, rtol=1e-12)
    ms = oq.contexts.mean_stds_staged_device(cm.gsims, ctx, cm.imts)
    np.testing.assert_allclose(ms[1], .79162428, rtol=1e-8)


def test_reference_known_planar_distances(oq):
    """The reference's own known answers for the planar distances,
    contexts_test.py:327-360 (PlanarDistancesTestCase): a point source of
    magnitude 7 at (0, 0), strike 90, dip 90, hypocentral depth 10 km, WC1994,
    aspect ratio 1, seismogenic layer 0 - 20 km, sites at (0.25, 0) and
    (0.35, 0), built and measured entirely on the device path"""
    src = oq.PointSource('ps', 0., 0., oq.ArbitraryMFD([7.0], [1.]), 'WC1994',
                         1.0, 0., 20., [(1.0, 90., 90., 90.)], [(1.0, 10.)])
    sc = oq.SiteCollection.from_points(
        [0.25, 0.35], [0., 0.], [0., 0.],
        dict(vs30=760., vs30measured=True, z1pt0=48., z2pt5=0.6))
    # models that together require every distance of the reference's test
    cm = oq.ContextMaker('*', [oq.AbrahamsonEtAl2014(), oq.AkkarEtAlRepi2014(),
                               oq.AkkarEtAlRhyp2014()],
                         dict(imtls={'PGA': [0.01]}, truncation_level=3.))
    [ctx] = cm.get_ctxs(src, sc)
    aac = np.testing.assert_allclose
    aac(ctx.rrup, [9.32409196, 20.44343079], rtol=1e-7)
    aac(ctx.rx, [0., 0.], atol=1e-9)
    aac(ctx.ry0, [9.26597563, 20.38546829], rtol=1e-7)
    aac(ctx.rjb, [9.26597481, 20.3854596], rtol=1e-7)
    aac(ctx.rhypo, [29.54267222, 40.18243627], rtol=1e-7)
    aac(ctx.repi, [27.79873166, 38.91822433], rtol=1e-7)


def test_sharding_groups_with_their_own_gsims(oq):
    """groups with different gsim sets occupy different columns
    of the map (gid): combining the ranks is a concatenation on the G axis,
    done by the same all-reduce with identities elsewhere"""
    import torch
    rng = np.random.default_rng(31)
    N = 70
    sites = synth_sites(N, rng)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    par = dict(imtls=imtls, truncation_level=3., investigation_time=50.)
    sets = [CFG['cfg3'], [('BooreAtkinson2008', {})],
            [('CampbellBozorgnia2014', {}), ('AbrahamsonEtAl2014', {})]]
    cms = [oq.ContextMaker('*', make_gsims(oq, s), par) for s in sets]
    srcs = []
    for g in range(3):
        r = synth_ruptures(30, np.random.default_rng(70 + g))
        pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                              r['rate'], r['area'], r['hypo_lon'],
                              r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
        srcs.append(oq.PlanarRuptures(pla, r['mag']))
    off = oq.sharding.gid_offsets(cms)
    assert list(off) == [0, 2, 3, 5]
    # reference: every group on its own
    want = np.zeros((N, cms[0].L, 5), np.float32)
    for g in range(3):
        pm = oq.MapArray(sc.sids, cms[g].L, cms[g].G, rates=True).fill(0)
        cms[g].hazard_planar(pm, srcs[g], sc)
        want[:, :, off[g]:off[g + 1]] = pm.array
    # two "ranks" executed one after the other: their maps add up to it
    tot = torch.zeros((N, cms[0].L, 5), device='cuda:0')
    for rank in range(2):
        big = oq.sharding.hazard_group_sharded_gid(cms, srcs, sc, rank, 2)
        assert big.array.shape == (N, cms[0].L, 5) and big.rates
        mine = oq.sharding.groups_of(rank, 2, 3)
        for g in range(3):
            cols = big.array[:, :, off[g]:off[g + 1]]
            assert bool((cols != 0).any()) == (g in mine)
        tot += big.array
    np.testing.assert_array_equal(tot.cpu().numpy(), want)
    one = oq.sharding.hazard_group_sharded_gid(cms, srcs, sc, 0, 1,
                                               rates=False)
    np.testing.assert_allclose(one.numpy(), np.exp(-want.astype(float)),
                               rtol=1e-6)


def test_fused_planar_alias(oq):
    """oqg_fused_planar (the name of SURVEY 8b) == oqg_hazard_planar"""
    import torch
    rng = np.random.default_rng(77)
    U, N = 40, 50
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls={imt: np.logspace(-2, 0.3, 20) for imt in IMTS},
        truncation_level=3., investigation_time=50.))
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    want = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
    cm.hazard_planar(want, oq.PlanarRuptures(pla, r['mag']), sc)
    h = cm.handle()
    from oq_engine_b200.contexts import _dev, _ptr
    t_pla = _dev(oq.planar.as_flat(pla), h.device)
    t_mag = _dev(r['mag'], h.device)
    t_md = _dev(cm.magdist(r['mag']), h.device)
    st, keep = cm._sites_struct(sc, h.device)
    t_map = torch.zeros((N, cm.L, cm.G), dtype=torch.float32,
                        device='cuda:%d' % h.device)
    h.call('oqg_fused_planar', _ptr(t_pla), _ptr(t_mag), U, ctypes.byref(st),
           N, _ptr(t_md), 0., 50., 1, _ptr(t_map), None)
    np.testing.assert_array_equal(t_map.cpu().numpy(), want.array)


# ------------------------------------------- per-PoE parity of the fused path
def _one_rupture(oq, rng, N):
    """One M6.5 rupture with rate * T = 1 exactly and N sites around it: the
    rates map of the fused kernel then holds the PoEs themselves"""
    r = synth_ruptures(1, rng)
    r['mag'][:] = 6.5
    r['rate'][:] = 2. ** -5
    r['area'] = np.array([400.])
    sites = synth_sites(N, rng, sentinels=True)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    return oq.contexts.PlanarRuptures(pla, r['mag']), sc


def _check_poes(got, exp, tag):
    """north_star tolerance: rtol 1e-5 where the PoE is >= 1e-8, 1e-7
    absolute below"""
    big = exp >= 1e-8
    rel = np.abs(got[big] / exp[big] - 1.)
    assert rel.max() < 1e-5, (tag, rel.max(), exp[big][rel.argmax()])
    assert np.abs(got[~big] - exp[~big]).max(initial=0.) < 1e-7, tag
    return int(big.sum())


@pytest.mark.parametrize('tl', [2., 2.5, 3., 5., 6.])
@pytest.mark.parametrize('shape', ['auto', 'plain'])
def test_fused_fast_path_per_poe(oq, tl, shape):
    """Every PoE of the table-driven fused path against truncnorm_sf
    (hazardlib/stats.py:38-73) in float64, not a sum over ruptures: one
    rupture with rate * T = 1 and 3000 sites give 3000 x 140 x 2 PoEs per
    truncation level with z spread over [-13, 13]; both CTA shapes of the
    fused kernel (sub-CTAs around the replicated table / plain table)"""
    rng = np.random.default_rng(int(tl * 10))
    N, L1 = 3000, 20
    src, sc = _one_rupture(oq, rng, N)
    imls = np.logspace(-4, 1.5, L1)
    imtls = {imt: imls for imt in IMTS}
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls=imtls, truncation_level=tl, investigation_time=32.))
    h = cm.handle()
    L_ = oq._lib.lib()
    in_use, err = ctypes.c_int(0), ctypes.c_double(0)
    L_.oqg_poe_table_info(h._h, ctypes.byref(in_use), ctypes.byref(err))
    assert in_use.value, 'the table path must be the one under test'
    h.call('oqg_set_fused_config', 0, int(shape == 'plain'))
    try:
        rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
        kept = cm.hazard_planar(rmap, src, sc, count=True)
        subs = L_.oqg_last_fused_subs(h._h)
    finally:
        h.call('oqg_set_fused_config', 0, 0)
    assert kept == N
    if shape == 'plain':
        assert subs == 1, subs
    ctx = cm.recarray(cm.get_ctxs(src, sc))
    np.testing.assert_array_equal(ctx.sids, sc.sids)
    ms = cm.get_mean_stds([ctx])        # (4, G, M, N) float64, device
    phi_b = float(poe_np.ndtr(tl))
    ll = np.log(imls)
    n = 0
    for g in range(cm.G):
        for m in range(cm.M):
            z = (ll[None, :] - ms[0, g, m][:, None]) / ms[1, g, m][:, None]
            exp = poe_np.truncnorm_sf(phi_b, z)
            got = rmap.array[:, m * L1:(m + 1) * L1, g].astype(float)
            n += _check_poes(got, exp, (tl, shape, g, m))
    assert n > 50_000, n     # PoEs checked at rtol 1e-5


@pytest.mark.parametrize('tl', [2., 3., 5.])
def test_fused_fast_path_edge_z(oq, tl):
    """The reference's edge cases through the fused table path, one site and
    levels placed at chosen z: z = +-tl, tl - 1e-7 ... tl - 1e-2 (relative
    accuracy next to the truncation point, where the textbook formula
    cancels), 2.99, 2.9999999, and just above tl (exactly 0)"""
    rng = np.random.default_rng(7)
    src, sc = _one_rupture(oq, rng, 1)
    zs = [-tl - 1e-7, -tl, -tl + 1e-7, -1., 0., 1., tl - 1e-2, tl - 1e-3,
          tl - 1e-4, tl - 1e-5, tl - 3e-6, tl - 1e-7, tl, tl + 1e-7]
    if tl == 3.:
        zs += [2.99, 2.9999999]
    zs = np.array(sorted(zs))
    for name in ('BooreEtAl2014', 'ChiouYoungs2014', 'AbrahamsonEtAl2014',
                 'CampbellBozorgnia2014'):
        prm = dict(imtls={imt: [1.] * len(zs) for imt in IMTS},
                   truncation_level=tl, investigation_time=32.)
        cm0 = oq.ContextMaker('*', [oq.get_gsim(name)], prm)
        ctx = cm0.recarray(cm0.get_ctxs(src, sc))
        ms = cm0.get_mean_stds([ctx])
        imtls = {imt: np.exp(ms[0, 0, m, 0] + zs * ms[1, 0, m, 0])
                 for m, imt in enumerate(IMTS)}
        prm['imtls'] = imtls
        cm = oq.ContextMaker('*', [oq.get_gsim(name)], prm)
        rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
        cm.hazard_planar(rmap, src, sc)
        phi_b = float(poe_np.ndtr(tl))
        for m, imt in enumerate(IMTS):
            z = (np.log(imtls[imt]) - ms[0, 0, m, 0]) / ms[1, 0, m, 0]
            exp = poe_np.truncnorm_sf(phi_b, z)
            got = rmap.array[0, m * len(zs):(m + 1) * len(zs), 0].astype(float)
            _check_poes(got, exp, (tl, name, imt))
            # at and beyond the truncation level the PoE is exactly 0 / 1
            assert (got[z >= tl + 5e-8] == 0.).all()
            assert (got[z <= -tl - 5e-8] == 1.).all()


def test_fused_unsorted_levels(oq):
    """A level grid that is not increasing (the level skipping of the fused
    kernel assumes sorted levels and must switch itself off)"""
    rng = np.random.default_rng(11)
    N, L1 = 500, 20
    src, sc = _one_rupture(oq, rng, N)
    imls = np.logspace(-3, 1., L1)
    imtls = {imt: rng.permutation(imls) for imt in IMTS}
    cm = oq.ContextMaker('*', make_gsims(oq, CFG['cfg3']), dict(
        imtls=imtls, truncation_level=3., investigation_time=32.))
    rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
    cm.hazard_planar(rmap, src, sc)
    ctx = cm.recarray(cm.get_ctxs(src, sc))
    ms = cm.get_mean_stds([ctx])
    phi_b = float(poe_np.ndtr(3.))
    for g in range(cm.G):
        for m, imt in enumerate(IMTS):
            ll = np.log(imtls[imt])
            z = (ll[None, :] - ms[0, g, m][:, None]) / ms[1, g, m][:, None]
            _check_poes(rmap.array[:, m * L1:(m + 1) * L1, g].astype(float),
                        poe_np.truncnorm_sf(phi_b, z), (g, imt))


def test_two_handles_two_streams(oq):
    """Two oqg handles on one device, each with its own stream, models and
    levels, used alternately without host synchronisation: the __constant__
    tables have one owner at a time (ConstGuard in csrc/oqg.cu), so every
    call must see its own handle's tables"""
    import torch
    from oq_engine_b200.contexts import _dev, _ptr
    rng = np.random.default_rng(3)
    U, N = 600, 400
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    src = oq.contexts.PlanarRuptures(pla, r['mag'])
    cms = [oq.ContextMaker('*', [oq.get_gsim('BooreEtAl2014')], dict(
               imtls={imt: np.logspace(-2, 0.3, 20) for imt in IMTS},
               truncation_level=3., investigation_time=50.)),
           oq.ContextMaker('*', [oq.get_gsim('ChiouYoungs2014'),
                                 oq.get_gsim('CampbellBozorgnia2014')], dict(
               imtls={imt: np.logspace(-3, 0., 15) for imt in IMTS[:4]},
               truncation_level=2.5, investigation_time=1.))]
    # solo results through the usual one-handle-per-device path
    solo = []
    for cm in cms:
        m = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
        cm.hazard_planar(m, src, sc)
        solo.append(m.array.copy())
    dev = 0
    hs = [oq._lib.Handle(dev) for _ in cms]      # own streams
    for cm, h in zip(cms, hs):
        cm.configure(h)
    flat = oq.planar.as_flat(pla)
    t_pla, t_mag = _dev(flat, dev), _dev(r['mag'], dev)
    st, keep = cms[0]._sites_struct(sc, dev)
    maps, mds = [], []
    for cm in cms:
        maps.append(torch.zeros((N, cm.L, cm.G), dtype=torch.float32,
                                device='cuda:%d' % dev))
        mds.append(_dev(cm.magdist(r['mag']), dev))
    torch.cuda.synchronize()
    reps = 6
    for _ in range(reps):
        for cm, h, t_map, t_md in zip(cms, hs, maps, mds):
            h.call('oqg_hazard_planar', _ptr(t_pla), _ptr(t_mag), _ptr(t_md),
                   None, 0, U, ctypes.byref(st), None, N, 0.,
                   float(cm.investigation_time), 1, _ptr(t_map), None)
    for h in hs:
        h.call('oqg_synchronize')
    for t_map, exp in zip(maps, solo):
        np.testing.assert_allclose(t_map.cpu().numpy(), reps * exp,
                                   rtol=2e-6, atol=1e-30)
    for h in hs:
        h.close()


def test_magnitude_outside_the_integration_distance_pairs(oq):
    """maximum_distance = [(5.5, 100), (7, 300)]: the reference's magdepdist
    (calc/filters.py:183-190) gives distance 0 outside [5.5, 7], so those
    ruptures contribute nothing"""
    rng = np.random.default_rng(21)
    U, N = 200, 64
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pairs = CFG['cfg3']
    md = oq.IntegrationDistance.new('[(5.5, 100), (7, 300)]')
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.,
        maximum_distance=md))
    mags = r['mag']
    inside = (mags >= 5.5) & (mags <= 7)
    assert inside.any() and (~inside).any()
    magdist = cm.magdist(mags)
    assert (magdist[~inside] == 0).all()
    np.testing.assert_allclose(magdist[inside],
                               np.interp(mags[inside], [5.5, 7], [100, 300]))
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
    kept = cm.hazard_planar(rmap, oq.contexts.PlanarRuptures(pla, mags), sc,
                            count=True)
    opla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    octx = hazard_np.get_ctx(pairs, opla, mags, sites, magdist, 0.)
    assert kept == len(octx) and set(np.unique(octx.mag)) <= set(
        mags[inside])
    acc = hazard_np.exact_rate_sum(pairs, octx, sites['sids'], imtls, 3., 50.)
    big = acc > 1e-10
    np.testing.assert_allclose(rmap.array[big], acc[big], rtol=1e-5)
    # the unfused operator drops the same rows
    ctx = cm.recarray(cm.get_ctxs(oq.contexts.PlanarRuptures(pla, mags), sc))
    assert len(ctx) == kept


def test_update_more_than_65535_sites(oq):
    """ContextMaker.update on a map with more rows than gridDim.y allows in
    one launch (the site dimension is launched in chunks)"""
    rng = np.random.default_rng(5)
    N = 70_000
    src, sc = _one_rupture(oq, rng, N)
    cm = oq.ContextMaker('*', [oq.get_gsim('BooreEtAl2014')], dict(
        imtls={'PGA': np.logspace(-2, 0, 4)}, truncation_level=3.,
        investigation_time=32.))
    ctx = cm.recarray(cm.get_ctxs(src, sc))
    a = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
    cm.update(a, ctx)
    b = oq.MapArray(sc.sids, cm.L, cm.G, rates=True).fill(0)
    cm.hazard_planar(b, src, sc)
    np.testing.assert_allclose(a.array, b.array, rtol=2e-6, atol=1e-30)
    assert (a.array[-1] > 0).any()


# ------------------------------------------- full-size spot rows (BASELINE 3)
def _oracle_rows(pairs, r, sites, sel, imtls):
    """float64 rate sums of the oracle for the sites `sel` and ALL the
    ruptures of r: len(sel) x U context rows"""
    sub = {k: v[sel] for k, v in sites.items()}
    sub['sids'] = np.arange(len(sel), dtype=np.uint32)
    opla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    octx = hazard_np.get_ctx(pairs, opla, r['mag'], sub)
    return hazard_np.exact_rate_sum(pairs, octx, sub['sids'], imtls, 3., 50.)


def test_cfg4_full_site_grid_spot_rows(oq):
    """cfg4: the four NGA-West2 models on the FULL 100 000-site grid (1000
    ruptures: 1e8 pairs, 5.6e10 evaluations), the map rows of 100 random
    sites against the oracle: 1e5 (rupture, site) rows x 4 models; then the
    same through the site sharding of SURVEY 8e (tiles by sids % 4, gathered)"""
    import torch
    rng = np.random.default_rng(44)
    U, N = 1000, 100_000
    r = synth_ruptures(U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pairs = CFG['cfg4']
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    src = oq.PlanarRuptures(pla, r['mag'])
    rmap = oq.MapArray(sc.sids, cm.L, cm.G, rates=True,
                       device='cuda:0').fill(0)
    kept = cm.hazard_planar(rmap, src, sc, count=True)
    assert kept == U * N
    sel = np.sort(rng.choice(N, 100, replace=False))
    acc = _oracle_rows(pairs, r, sites, sel, imtls)       # 1e5 rows
    got = rmap.array[sel].cpu().numpy()
    big = acc > 1e-10
    assert big.sum() > 30_000
    np.testing.assert_allclose(got[big], acc[big], rtol=1e-5)
    # site sharding: 4 tiles by sids % 4, assembled by gather_tiles
    world = 4
    whole = None
    for rank in range(world):
        tmap, tile = oq.sharding.hazard_site_sharded(cm, src, sc, rank, world,
                                                     rates=True)
        part = oq.sharding.gather_tiles(tmap.array, tile.sids, N,
                                        partial=True)
        whole = part if whole is None else whole + part
    # same kernel, same sites; another split of the ruptures into chunks
    # (float32 partial sums are flushed at other points)
    torch.testing.assert_close(whole, rmap.array, rtol=2e-6, atol=1e-30)


def test_cfg5_groups_product_spot_rows(oq):
    """cfg5: 8 source groups over the full 50 000-site grid, the pne maps of
    the ranks multiplied as the NCCL PRODUCT all-reduce does (2 emulated
    ranks of 4 groups), 50 random sites against the oracle: 8 x 250 x 50 =
    1e5 (rupture, site) rows"""
    import torch
    rng = np.random.default_rng(55)
    NG, U, N = 8, 250, 50_000
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    pairs = CFG['cfg3']
    cm = oq.ContextMaker('*', make_gsims(oq, pairs), dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    groups, srcs = [], []
    for g in range(NG):
        r = synth_ruptures(U, np.random.default_rng(20240 + g))
        groups.append(r)
        pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                              r['rate'], r['area'], r['hypo_lon'],
                              r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
        srcs.append(oq.PlanarRuptures(pla, r['mag']))
    world = 2
    prod = None
    for rank in range(world):
        pm = oq.MapArray(sc.sids, cm.L, cm.G, device='cuda:0').fill(1)
        for g in oq.sharding.groups_of(rank, world, NG):
            cm.hazard_planar(pm, srcs[g], sc)
        prod = pm.array if prod is None else prod * pm.array   # PRODUCT
    sel = np.sort(rng.choice(N, 50, replace=False))
    acc = sum(_oracle_rows(pairs, r, sites, sel, imtls) for r in groups)
    got = prod[torch.as_tensor(sel, device='cuda:0')].cpu().numpy()
    np.testing.assert_allclose(got, np.exp(-acc), rtol=1e-5)
    big = acc > 1e-3          # ... and where the hazard is: on the rates
    np.testing.assert_allclose(-np.log(got[big].astype(np.float64)),
                               acc[big], rtol=2e-4)
