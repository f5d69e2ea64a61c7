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
