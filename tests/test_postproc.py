"""Post-processing of the rates map.

CPU part: the oracle (oracle/post_np.py) against the fixture produced by the
reference's own functions (oracle/make_golden_post.py).  GPU part (-m gpu):
oqg_postclassical / oqg_hazard_maps through the Python mirror of
`postclassical` against the same fixture.

Tolerances: curves and statistics are float32 on output as in the reference
(MapArray arrays): rtol 2e-6 (the device reads the rates as float32, as
stored by MapArray), atol 1e-9 (float32 rounding of 1 - exp(-x) near 0 is
exact to 1e-7 relative; std of nearly equal curves cancels).  Hazard maps:
rtol 1e-5 -- the reference takes a float32 logarithm of the curve
(see oracle/post_np.py), the device a float64 one.
"""
import os
import itertools
import numpy as np
import pytest

import post_np
from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'post.npz'))
N, L, Gt = D['rates'].shape
R, M = int(D['R']), int(D['M'])
L1 = L // M
OFF = D['trt_rlzs_off']
TRT_RLZS = [D['trt_rlzs'][OFF[g]:OFF[g + 1]] for g in range(Gt)]
QUANT = list(D['quantiles'])
KINDS = [('mean', 0.), ('std', 0.)] + [('quantile', q) for q in QUANT] + [
    ('max', 0.)]
NAMES = ['mean', 'std'] + ['quantile-%s' % q for q in QUANT] + ['max']
IMTLS = {imt: D['imls'] for imt in ('PGA', 'SA(0.2)', 'SA(1.0)')}


def weights_of(name):
    return D['w1'][:, None] if name == 'w1' else D['wim']


# ------------------------------------------------------------------ CPU
def test_oracle_hcurves_golden():
    np.testing.assert_allclose(
        post_np.hcurves_from_rates(D['rates'], TRT_RLZS, R), D['hcurves'],
        rtol=1e-15)


@pytest.mark.parametrize('wname', ['w1', 'wim'])
@pytest.mark.parametrize('use_rates', [0, 1])
def test_oracle_stats_golden(wname, use_rates):
    want = D['stats_%s_%d' % (wname, use_rates)]
    w = weights_of(wname)
    for k, (kind, q) in enumerate(KINDS):
        for n in range(0, N, 3):
            got = post_np.stat_curve(D['hcurves'][n], kind, q, w, M,
                                     bool(use_rates))
            np.testing.assert_allclose(got, want[k, n], rtol=1e-12,
                                       atol=1e-300)


def test_oracle_hazard_maps_golden():
    mean = D['stats_w1_0'][0].reshape(N, M, L1).astype(np.float32)
    for m in range(M):
        np.testing.assert_allclose(
            post_np.compute_hazard_maps(mean[:, m], D['imls'], D['poes']),
            D['hmaps'][:, m], rtol=1e-13)


def test_oracle_to_rates_edges():
    np.testing.assert_array_equal(post_np.to_rates(D['to_rates_in']),
                                  D['to_rates_out'])


def test_quantile_tie_break_is_by_weight():
    # numpy's sort(order='c') breaks ties in the value with the weight
    c = np.array([[0.2], [0.1], [0.2], [0.3]])
    w = np.array([0.4, 0.1, 0.2, 0.3])
    got = post_np.quantile_curve(0.25, c, w)
    # sorted (c, w): (.1,.1) (.2,.2) (.2,.4) (.3,.3); cum .1 .3 .7 1.
    assert abs(got[0] - (0.1 + (0.25 - 0.1) * (0.2 - 0.1) / 0.2)) < 1e-15


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope='module')
def oq():
    import oq_engine_b200
    oq_engine_b200._lib.lib()
    return oq_engine_b200


def rmap_of(oq):
    rmap = oq.MapArray(np.arange(N, dtype=np.uint32) * 2 + 1, L, Gt,
                       rates=True)
    rmap.array = D['rates'].astype(np.float32)
    return rmap


def oracle32_hcurves():
    r32 = D['rates'].astype(np.float32).astype(float)
    return post_np.hcurves_from_rates(r32, TRT_RLZS, R)


@pytest.mark.gpu
@pytest.mark.parametrize('wname', ['w1', 'wim'])
@pytest.mark.parametrize('use_rates', [0, 1])
def test_postclassical_golden(oq, wname, use_rates):
    w = weights_of(wname)
    res = oq.postclassical(rmap_of(oq), TRT_RLZS, w if wname == 'wim'
                           else w[:, 0], IMTLS, NAMES, individual_rlzs=True,
                           poes=list(D['poes']), use_rates=bool(use_rates))
    assert set(res) == {'hcurves-rlzs', 'hcurves-stats', 'hmaps-rlzs',
                        'hmaps-stats'}
    assert len(res['hcurves-rlzs']) == R and len(res['hcurves-stats']) == 6
    assert res['hcurves-stats'][0].array.shape == (N, M, L1)
    assert (res['hcurves-stats'][0].sids == np.arange(N) * 2 + 1).all()
    # realizations: against the oracle on the float32 rates the device reads
    hc = oracle32_hcurves()
    got = np.stack([m.array.reshape(N, L) for m in res['hcurves-rlzs']], -1)
    np.testing.assert_allclose(got, hc, rtol=2e-7, atol=1e-30)
    # ... and against the reference (float64 rates): float32 input rounding
    np.testing.assert_allclose(got, D['hcurves'], rtol=2e-6, atol=1e-30)
    # statistics against the reference
    want = D['stats_%s_%d' % (wname, use_rates)]
    for k, name in enumerate(NAMES):
        got = res['hcurves-stats'][k].array.reshape(N, L)
        np.testing.assert_allclose(got, want[k], rtol=2e-6, atol=1e-9,
                                   err_msg=name)
        # and tightly against the oracle on the same float32 rates
        for n in (0, 3, 7, 9, 21):
            mine = post_np.stat_curve(hc[n], KINDS[k][0], KINDS[k][1], w, M,
                                      bool(use_rates))
            np.testing.assert_allclose(got[n], mine, rtol=3e-7, atol=1e-12,
                                       err_msg=name)
    if wname == 'wim':   # SA(0.2) has zero weights: no curve
        assert (res['hcurves-stats'][0].array[:, 1] == 0).all()


@pytest.mark.gpu
def test_hazard_maps_golden(oq):
    mean = D['stats_w1_0'][0].reshape(N, M, L1).astype(np.float32)
    got = oq.compute_hmaps(mean, IMTLS, list(D['poes']))
    assert got.shape == (N, M, len(D['poes']))
    np.testing.assert_allclose(got, D['hmaps'], rtol=1e-5, atol=0)
    assert ((got == 0) == (D['hmaps'] == 0)).all()   # PoE above the curve
    one = oq.compute_hazard_maps(mean[:, 0], D['imls'], D['poes'])
    np.testing.assert_array_equal(one, got[:, 0])
    with pytest.raises(ValueError):
        oq.compute_hazard_maps(mean[:, 0], D['imls'][:-1], D['poes'])


@pytest.mark.gpu
def test_postclassical_errors_and_single_rlz(oq):
    rmap = rmap_of(oq)
    with pytest.raises(NotImplementedError):
        oq.postclassical(rmap, TRT_RLZS, D['w1'], IMTLS, ['median'])
    with pytest.raises(ValueError):
        oq.postclassical(rmap, TRT_RLZS, D['w1'][:3], IMTLS, ['mean'])
    # R = 1: the realization curves are produced even without
    # individual_rlzs (classical.py:288), the mean equals the realization
    one = oq.postclassical(rmap, [[0]] * Gt, [1.], IMTLS, ['mean'])
    np.testing.assert_allclose(one['hcurves-stats'][0].array,
                               one['hcurves-rlzs'][0].array, rtol=1e-7)
    tot = D['rates'].astype(np.float32).astype(float).sum(axis=2)
    np.testing.assert_allclose(one['hcurves-rlzs'][0].array.reshape(N, L),
                               1 - np.exp(-tot), rtol=2e-7, atol=1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize('grouping', [True, False])
@pytest.mark.parametrize('R', [1, 4, 9, 33, 40, 130])
def test_postclassical_other_sizes_vs_oracle(oq, R, grouping):
    """every kernel variant (register path RP = 4 .. 32, shared-memory path
    for more groups) against the oracle on random rates, with ties between
    realizations sharing all their gsim columns: grouped (at most 21 distinct
    sets of 2 of 6 columns) and one group per realization"""
    rng = np.random.default_rng(R)
    n, m, l1, gt = 23, 2, 5, 6
    L = m * l1
    rates = (10 ** rng.uniform(-4, 0.3, (n, L, gt))).astype(np.float32)
    rates[2] = 0
    pick = rng.integers(0, gt, (R, 2))          # two columns per realization
    trt_rlzs = [np.where((pick == g).any(axis=1))[0] for g in range(gt)]
    w = rng.uniform(0.2, 1., R)
    w /= w.sum()
    names = ['mean', 'std', 'quantile-0.16', 'quantile-0.5', 'quantile-0.84',
             'max']
    imtls = {'PGA': np.logspace(-2, 0, l1), 'SA(1.0)': np.logspace(-2, 0, l1)}
    res = oq.postclassical(rates, trt_rlzs, w, imtls, names,
                           individual_rlzs=True, grouping=grouping)
    path, groups = oq.postproc.last_post_path()
    assert path == 0, path
    assert groups == (min(R, len(set(map(tuple, np.sort(pick, axis=1)))))
                      if grouping else R), groups
    hc = post_np.hcurves_from_rates(rates.astype(float), trt_rlzs, R)
    got = np.stack([x.array.reshape(n, L) for x in res['hcurves-rlzs']], -1)
    np.testing.assert_allclose(got, hc, rtol=2e-7, atol=1e-30)
    kinds = [('mean', 0.), ('std', 0.), ('quantile', .16), ('quantile', .5),
             ('quantile', .84), ('max', 0.)]
    for k, (kind, q) in enumerate(kinds):
        got = res['hcurves-stats'][k].array.reshape(n, L)
        for i in range(0, n, 4):
            mine = post_np.stat_curve(hc[i], kind, q, w[:, None], m)
            np.testing.assert_allclose(got[i], mine, rtol=5e-7, atol=1e-9,
                                       err_msg='%s R=%d site %d' % (
                                           names[k], R, i))


@pytest.mark.gpu
@pytest.mark.parametrize('grouping', [False, True])
@pytest.mark.parametrize('use_rates', [False, True])
def test_postclassical_many_rlzs_stream(oq, use_rates, grouping):
    """mean and std need no sort: the streaming kernel takes any number of
    realizations (one group per realization); grouped, the 600 realizations
    are at most 45 distinct sets of columns"""
    R = 600
    rng = np.random.default_rng(7)
    n, m, l1, gt = 17, 2, 5, 9
    L = m * l1
    rates = (10 ** rng.uniform(-4, 0.3, (n, L, gt))).astype(np.float32)
    rates[3] = 0
    pick = rng.integers(0, gt, (R, 2))
    trt_rlzs = [np.where((pick == g).any(axis=1))[0] for g in range(gt)]
    w = np.full(R, 1. / R)                       # sampling: equal weights
    imtls = {'PGA': np.logspace(-2, 0, l1), 'SA(1.0)': np.logspace(-2, 0, l1)}
    res = oq.postclassical(rates, trt_rlzs, w, imtls, ['mean', 'std', 'max'],
                           individual_rlzs=True, use_rates=use_rates,
                           grouping=grouping)
    assert oq.postproc.last_post_path()[0] == (0 if grouping else 1)
    hc = post_np.hcurves_from_rates(rates.astype(float), trt_rlzs, R)
    got = np.stack([x.array.reshape(n, L) for x in res['hcurves-rlzs']], -1)
    np.testing.assert_allclose(got, hc, rtol=2e-7, atol=1e-30)
    for k, kind in enumerate(['mean', 'std', 'max']):
        got = res['hcurves-stats'][k].array.reshape(n, L)
        for i in range(n):
            mine = post_np.stat_curve(hc[i], kind, 0., w[:, None], m,
                                      use_rates=use_rates)
            np.testing.assert_allclose(got[i], mine, rtol=5e-7, atol=1e-9,
                                       err_msg='%s site %d' % (kind, i))


@pytest.mark.gpu
@pytest.mark.parametrize('grouping', [False, True])
@pytest.mark.parametrize('R, use_rates', [(257, False), (600, True),
                                          (1500, False), (4096, False)])
def test_postclassical_many_rlzs_quantiles(oq, R, use_rates, grouping):
    """quantiles of 256 < R <= 4096 realizations against the oracle, with
    ties (realizations sharing their gsim columns) and unequal weights: one
    group per realization (the CTA-cooperative sort, k_postclassical_cta) and
    grouped (78 distinct sets of columns: the shared-memory sort path)"""
    rng = np.random.default_rng(R)
    n, m, l1, gt = 5, 2, 4, 12
    L = m * l1
    rates = (10 ** rng.uniform(-4, 0.3, (n, L, gt))).astype(np.float32)
    rates[1] = 0
    pick = rng.integers(0, gt, (R, 2))
    trt_rlzs = [np.where((pick == g).any(axis=1))[0] for g in range(gt)]
    w = rng.uniform(0.2, 1., R) if R != 600 else np.ones(R)
    w /= w.sum()
    names = ['mean', 'quantile-0.05', 'std', 'quantile-0.5', 'max',
             'quantile-0.95', 'quantile-1.0', 'quantile-0.0']
    kinds = [('mean', 0.), ('quantile', .05), ('std', 0.), ('quantile', .5),
             ('max', 0.), ('quantile', .95), ('quantile', 1.), ('quantile', 0.)]
    imtls = {'PGA': np.logspace(-2, 0, l1), 'SA(1.0)': np.logspace(-2, 0, l1)}
    res = oq.postclassical(rates, trt_rlzs, w, imtls, names,
                           individual_rlzs=True, use_rates=use_rates,
                           grouping=grouping)
    path, groups = oq.postproc.last_post_path()
    assert (path, groups <= 78) == ((0, True) if grouping else (2, False))
    hc = post_np.hcurves_from_rates(rates.astype(float), trt_rlzs, R)
    got = np.stack([x.array.reshape(n, L) for x in res['hcurves-rlzs']], -1)
    np.testing.assert_allclose(got, hc, rtol=2e-7, atol=1e-30)
    for k, (kind, q) in enumerate(kinds):
        got = res['hcurves-stats'][k].array.reshape(n, L)
        for i in range(n):
            mine = post_np.stat_curve(hc[i], kind, q, w[:, None], m,
                                      use_rates=use_rates)
            np.testing.assert_allclose(got[i], mine, rtol=5e-7, atol=1e-9,
                                       err_msg='%s R=%d site %d' % (
                                           names[k], R, i))
    if R == 4096 and not grouping:
        with pytest.raises(ValueError):
            oq.postclassical(np.zeros((1, L, 1), np.float32),
                             [np.arange(4097)], np.full(4097, 1 / 4097.),
                             imtls, ['quantile-0.5'], grouping=False)
    elif R == 4096:
        # one gsim column, any number of realizations: one group
        R2 = 5000
        w2 = rng.uniform(0.2, 1., R2)
        w2 /= w2.sum()
        r1 = (10 ** rng.uniform(-3, 0, (3, L, 1))).astype(np.float32)
        res = oq.postclassical(r1, [np.arange(R2)], w2, imtls,
                               ['quantile-0.5', 'mean'])
        assert oq.postproc.last_post_path() == (0, 1)
        exp = 1 - np.exp(-r1[:, :, 0].astype(float))
        for k in range(2):
            np.testing.assert_allclose(
                res['hcurves-stats'][k].array.reshape(3, L), exp, rtol=5e-7)


@pytest.mark.gpu
def test_postclassical_many_rlzs_imt_weights(oq):
    """IMT-dependent weights with more than 256 realizations (the weight
    column, its ranks and its sum change with the IMT; one IMT has no weight
    at all): CTA path with quantiles, streaming path without"""
    R = 300
    rng = np.random.default_rng(12)
    n, m, l1, gt = 6, 3, 4, 10
    L = m * l1
    rates = (10 ** rng.uniform(-4, 0.3, (n, L, gt))).astype(np.float32)
    pick = rng.integers(0, gt, (R, 2))
    trt_rlzs = [np.where((pick == g).any(axis=1))[0] for g in range(gt)]
    w = rng.uniform(0.2, 1., (R, m + 1))
    w[:, 1] = 0.
    w[:, [0, 2, 3]] /= w[:, [0, 2, 3]].sum(axis=0)
    imtls = {imt: np.logspace(-2, 0, l1) for imt in ('PGA', 'SA(0.3)',
                                                     'SA(1.0)')}
    hc = post_np.hcurves_from_rates(rates.astype(float), trt_rlzs, R)
    cases = ((['mean', 'quantile-0.1', 'quantile-0.9', 'std'],
              [('mean', 0.), ('quantile', .1), ('quantile', .9), ('std', 0.)]),
             (['std', 'max', 'mean'],
              [('std', 0.), ('max', 0.), ('mean', 0.)]))
    for (names, kinds), grouping in itertools.product(cases, (False, True)):
        res = oq.postclassical(rates, trt_rlzs, w, imtls, names,
                               grouping=grouping)
        path = oq.postproc.last_post_path()[0]
        assert path == (0 if grouping else 2 if 'quantile-0.1' in names
                        else 1), path
        assert 'hcurves-rlzs' not in res
        for k, (kind, q) in enumerate(kinds):
            got = res['hcurves-stats'][k].array.reshape(n, L)
            assert (got[:, l1:2 * l1] == 0).all()     # no weight: no curve
            for i in range(n):
                mine = post_np.stat_curve(hc[i], kind, q, w, m)
                np.testing.assert_allclose(got[i], mine, rtol=5e-7, atol=1e-9,
                                           err_msg='%s site %d' % (names[k], i))


@pytest.mark.gpu
def test_calc_mean_rates_golden(oq):
    """calc_mean_rates (calc/mean_rates.py:62-74) against the reference: one
    weight per gsim column and IMT-dependent weights"""
    imts = list(IMTLS)
    got = oq.calc_mean_rates(rmap_of(oq), D['gw1'], None, IMTLS)
    assert got.shape == (N, M, L1) and got.dtype == np.float64
    np.testing.assert_allclose(got, D['mean_rates_w1'], rtol=1e-13)
    got = oq.calc_mean_rates(D['rates'].astype(np.float32), D['gwim'],
                             lambda w, imt: w[:, imts.index(imt)], IMTLS)
    np.testing.assert_allclose(got, D['mean_rates_wim'], rtol=1e-13)
    assert (got[:, 1] == 0).all()
    with pytest.raises(ValueError):
        oq.calc_mean_rates(D['rates'], D['gw1'][:-1], None, IMTLS)
    # other numbers of gsim columns (rows per CTA, the direct kernel)
    rng = np.random.default_rng(5)
    for gt in (1, 7, 32, 100, 500):
        r = rng.uniform(0, 1e-2, (301, 6, gt)).astype(np.float32)
        w = rng.uniform(0.1, 1., gt)
        got = oq.calc_mean_rates(r, w, None, {'PGA': [1, 2, 3],
                                              'SA(1.0)': [1, 2, 3]})
        np.testing.assert_allclose(got.reshape(301, 6), r.astype(float) @ w,
                                   rtol=1e-13)
