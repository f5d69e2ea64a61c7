"""BASELINE.json configs[1], demos/hazard/LogicTreeCase2ClassicalPSHA (a
simple fault + an area source, 81 source-model paths x 4 GMPE paths = 324
realizations, mean / quantile curves and hazard maps), end to end on the
device path against the reference's own run of the same chain
(oracle/make_golden_cfg2.py -> tests/golden/cfg2_logic_tree.npz): the rates
of every (source variant, GMPE) column, the 324 hazard curves, the statistics
and the hazard maps."""
import os
import sys
import numpy as np
import pytest

from helpers import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, 'tools'))
D = np.load(os.path.join(GOLDEN, 'cfg2_logic_tree.npz'))


def test_realizations_of_the_demo():
    import demo_cfg2
    rlzs = demo_cfg2.realizations()
    assert len(rlzs) == 324 == len(D['weights'])
    np.testing.assert_allclose([r[4] for r in rlzs], D['weights'], rtol=1e-15)
    assert abs(sum(r[4] for r in rlzs) - 1.) < 1e-12


@pytest.mark.gpu
def test_cfg2_logic_tree_case2():
    import oq_engine_b200 as oq
    import demo_cfg2
    oq._lib.lib()
    rates, cols, trt_rlzs, weights, post, kept = demo_cfg2.run(oq, D)
    assert [c[0] for c in cols] == list(D['col_kind'])
    assert [c[1] for c in cols] == list(D['col_variant'])
    np.testing.assert_array_equal(np.concatenate(trt_rlzs), D['trt_rlzs'])
    # the same rows survive the distance filter and the pointsource_distance
    # collapse as in the reference
    np.testing.assert_array_equal([kept['fault', v] for v in range(9)],
                                  D['nrows_fault'])
    np.testing.assert_array_equal([kept['area', v] for v in range(9)],
                                  D['nrows_area'])
    # rates of the 36 columns: the reference accumulates float32 rates row by
    # row (update_pmap_r), the device in float64 and rounds once
    want = D['rates']
    big = want > 1e-7
    np.testing.assert_allclose(rates[big], want[big], rtol=2e-5)
    np.testing.assert_allclose(rates, want, rtol=2e-5, atol=1e-9)
    # 324 hazard curves
    hc = np.array([m.array[0, 0] for m in post['hcurves-rlzs']]).T   # (L, R)
    np.testing.assert_allclose(hc, D['hcurves'], rtol=2e-5, atol=1e-7)
    # mean and quantiles
    st = np.array([m.array[0, 0] for m in post['hcurves-stats']])
    np.testing.assert_allclose(st, D['stat_curves'], rtol=3e-5, atol=1e-7)
    # hazard maps at PoE 0.1 and 0.02 of the four statistics
    hm = np.array([m.array[0, 0] for m in post['hmaps-stats']])
    np.testing.assert_allclose(hm, D['hmaps'], rtol=1e-4)
