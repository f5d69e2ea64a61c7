"""Ground motion fields given the epsilons:
oqg_gmf against the reference's GmfComputer._compute
(oracle/make_golden_gmf.py) in its three modes.  float32 arithmetic as in
the reference; numpy's and CUDA's float32 exp differ by an ulp: rtol 1e-6."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'gmf.npz'))
IMTS = [str(x) for x in D['imts']]


def test_numpy_restatement_matches_reference():
    ms, w, b = D['mean_stds'], D['within'], D['between']
    mean, sig, tau, phi = ms
    for tag in ('zero', 'total', 'both'):
        if tag == 'zero':
            v = np.repeat(mean[:, :, None], w.shape[2], axis=2)
        elif tag == 'total':
            v = mean[:, :, None] + sig[:, :, None] * w
        else:
            v = mean[:, :, None] + phi[:, :, None] * w + \
                tau[:, :, None] * b.T[:, None, :]
        out = np.where(np.array([i != 'MMI' for i in IMTS])[:, None, None],
                       np.exp(v), v)
        np.testing.assert_allclose(out, D['gmf_' + tag], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['zero', 'total', 'both'])
def test_gmf_golden(tag):
    import oq_engine_b200 as oq
    oq._lib.lib()
    kw = dict(zero=dict(truncation_within=0., truncation_between=0.),
              total=dict(total_stddev_only=True), both={})[tag]
    got = oq.compute_gmf(D['mean_stds'], IMTS, D['within'], D['between'], **kw)
    assert got.dtype == np.float32 and got.shape == D['gmf_' + tag].shape
    np.testing.assert_allclose(got, D['gmf_' + tag], rtol=1e-6)
    # MMI is not exponentiated
    if tag == 'zero':
        np.testing.assert_array_equal(
            got[2], np.repeat(D['mean_stds'][0, 2][:, None], got.shape[2], 1))
    if tag == 'both':
        with pytest.raises(ValueError):
            oq.compute_gmf(D['mean_stds'], IMTS, D['within'])


# ---------------------------------------------------------------------------
# the sampling side: GmfComputer.compute_all (oracle/make_golden_gmf_all.py)
# ---------------------------------------------------------------------------
A = np.load(os.path.join(GOLDEN, 'gmf_all.npz'))
AIMTS = [str(x) for x in A['imts']]
PERIODS = [0., 0.3, 1.0]
TAGS = ['plain', 'jb2009', 'minmax']


def test_pcg64_stream_is_numpys():
    """the generator the device reproduces: numpy's PCG64 uniforms from the
    (state, inc) numpy derives from the seed"""
    import gmf_np
    for tag in TAGS:
        st = [int(x) for x in A[tag + '_state']]
        u = gmf_np.pcg64_uniform((st[0] << 64) | st[1], (st[2] << 64) | st[3],
                                 64)
        np.testing.assert_array_equal(u, A[tag + '_uniform'])
        u2 = gmf_np.pcg64_uniform((st[0] << 64) | st[1],
                                  (st[2] << 64) | st[3], 10, offset=20)
        np.testing.assert_array_equal(u2, A[tag + '_uniform'][20:30])


def test_truncnorm_ppf_is_scipys_rvs():
    from scipy import stats
    import gmf_np
    for tl in (2., 3., 5., 99.):
        x = stats.truncnorm(-tl, tl).rvs((50, 40), np.random.default_rng(3))
        u = np.random.default_rng(3).uniform(size=2000)
        y = gmf_np.truncnorm_ppf(u, tl).reshape(50, 40)
        np.testing.assert_allclose(y, x, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('tag', TAGS)
def test_oracle_compute_all_matches_reference(tag):
    import gmf_np
    ms = [A['%s_ms_%d' % (tag, g)] for g in range(2)]
    rlzs = [A['%s_rlzs_%d' % (tag, g)] for g in range(2)]
    data, sig, eps = gmf_np.compute_all(
        ms, rlzs, AIMTS, PERIODS, A[tag + '_ctx_sids'], A[tag + '_lons'],
        A[tag + '_lats'], 12, int(A[tag + '_seed']),
        spatial=False if tag == 'jb2009' else None,
        min_iml=A[tag + '_min_iml'], max_iml=A[tag + '_max_iml'])
    np.testing.assert_array_equal(eps, A[tag + '_eps'])
    np.testing.assert_array_equal(sig, A[tag + '_sig'])
    for col in ('eid', 'sid', 'rlz'):
        np.testing.assert_array_equal(data[col], A['%s_df_%s' % (tag, col)])
    for imt in AIMTS:
        np.testing.assert_allclose(data[imt], A['%s_df_%s' % (tag, imt)],
                                   rtol=1e-6)


def _ctx_of(tag):
    names = [k[len(tag) + 5:] for k in A.files
             if k.startswith(tag + '_ctx_') and A[k].ndim == 1]
    ctx = np.zeros(len(A[tag + '_ctx_sids']), [
        (n, A['%s_ctx_%s' % (tag, n)].dtype) for n in names]).view(np.recarray)
    for n in names:
        ctx[n] = A['%s_ctx_%s' % (tag, n)]
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize('tl', [2., 3., 5., 99.])
def test_device_truncnorm_rvs_is_scipys(tl):
    """oqg_truncnorm_rvs against scipy.stats.truncnorm.rvs with numpy's
    default_rng: the uniforms are the same bit for bit, the inverse CDF is
    another float64 routine -> float32 values equal up to one rounding;
    a second call continues the stream (offset)"""
    import ctypes
    import torch
    from scipy import stats
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import get_handle, _ptr
    h = get_handle(0)
    seed, n1, n2 = 1234, 300_001, 70_000
    ref = stats.truncnorm(-tl, tl).rvs(n1 + n2, np.random.default_rng(seed)
                                       ).astype(np.float32)
    st = np.random.PCG64(seed).state['state']
    mask = (1 << 64) - 1
    state = (ctypes.c_uint64 * 2)(st['state'] >> 64, st['state'] & mask)
    inc = (ctypes.c_uint64 * 2)(st['inc'] >> 64, st['inc'] & mask)
    out = torch.empty(n1 + n2, dtype=torch.float32, device='cuda')
    h.call('oqg_truncnorm_rvs', state, inc, 0, n1, tl, _ptr(out))
    h.call('oqg_truncnorm_rvs', state, inc, n1, n2, tl, _ptr(out[n1:]))
    got = out.cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-7, atol=1e-9)
    assert (got != ref).mean() < 1e-3       # a different rounding is rare
    assert np.abs(got).max() <= min(tl, 9.)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', TAGS)
def test_gmf_computer_golden(tag):
    """GmfComputer.compute_all on the device against the reference's own run:
    the epsilons and sig exactly (up to a float32 rounding), the ground
    motion values within float32 arithmetic (rtol 2e-6; with the JB2009
    Cholesky factor from cuSOLVER instead of LAPACK 2e-5)"""
    import oq_engine_b200 as oq
    oq._lib.lib()
    cm = oq.ContextMaker('*', [oq.get_gsim('BooreAtkinson2008'),
                               oq.get_gsim('AkkarEtAlRjb2014')], dict(
        imtls={imt: [0.1] for imt in AIMTS}, truncation_level=3.))
    rlzs = [A['%s_rlzs_%d' % (tag, g)] for g in range(2)]
    corr = oq.JB2009CorrelationModel(False) if tag == 'jb2009' else None
    gc = oq.GmfComputer(_ctx_of(tag), A[tag + '_lons'], A[tag + '_lats'], cm,
                        rlzs, n_occ=12, seed=int(A[tag + '_seed']),
                        scenario=True, correlation_model=corr,
                        min_iml=A[tag + '_min_iml'],
                        max_iml=A[tag + '_max_iml'])
    data = gc.compute_all()
    np.testing.assert_array_equal(gc.eid, A[tag + '_eid'])
    np.testing.assert_array_equal(gc.rlz, A[tag + '_rlz'])
    np.testing.assert_allclose(gc.between_eps, A[tag + '_eps'], rtol=2e-7)
    np.testing.assert_allclose(gc.sig, A[tag + '_sig'], rtol=1e-6)
    for col in ('eid', 'sid', 'rlz'):
        np.testing.assert_array_equal(data[col], A['%s_df_%s' % (tag, col)])
    rtol = 2e-5 if tag == 'jb2009' else 2e-6
    for imt in AIMTS:
        np.testing.assert_allclose(data[imt], A['%s_df_%s' % (tag, imt)],
                                   rtol=rtol)
