"""
Golden fixture for the ground-motion-field kernel (SURVEY.md section 8f row
4): the reference's GmfComputer._compute
(openquake/hazardlib/calc/gmf.py:476-525) called unbound on synthetic mean /
standard deviations and epsilons, in its three modes (no truncation: mean
only; total standard deviation only; between + within residuals).

    python oracle/make_golden_gmf.py   -> tests/golden/gmf.npz
"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, HERE)
import oq_bootstrap  # noqa
oq_bootstrap.install()

from openquake.hazardlib.calc.gmf import GmfComputer  # noqa
from openquake.hazardlib.const import StdDev  # noqa
from openquake.hazardlib.imt import from_string  # noqa

GOLDEN = os.path.join(HERE, '..', 'tests', 'golden')
F32 = np.float32


def main():
    rng = np.random.default_rng(33)
    M, N, E = 3, 57, 40
    imts = ['PGA', 'SA(0.3)', 'MMI']
    mean_stds = np.zeros((4, M, N))
    mean_stds[0] = rng.uniform(-6, 0, (M, N))
    mean_stds[0, 2] = rng.uniform(2, 9, N)            # MMI: not exponentiated
    mean_stds[2] = rng.uniform(0.2, 0.5, (M, N))
    mean_stds[3] = rng.uniform(0.4, 0.7, (M, N))
    mean_stds[1] = np.sqrt(mean_stds[2] ** 2 + mean_stds[3] ** 2)
    mean_stds = mean_stds.astype(F32)                 # gmf.py:406
    within = rng.standard_normal((M, N, E)).astype(F32)
    between = rng.standard_normal((E, M)).astype(F32)
    idxs = np.arange(E)
    out = dict(mean_stds=mean_stds, within=within, between=between,
               imts=np.array(imts))
    both = types.SimpleNamespace(DEFINED_FOR_STANDARD_DEVIATION_TYPES={
        StdDev.TOTAL, StdDev.INTER_EVENT, StdDev.INTRA_EVENT})
    total = types.SimpleNamespace(
        DEFINED_FOR_STANDARD_DEVIATION_TYPES={StdDev.TOTAL})
    for tag, tlw, tlb, gsim in (('zero', 0., 0., both), ('total', 3., 3., total),
                                ('both', 3., 3., both)):
        self = types.SimpleNamespace(
            cmaker=types.SimpleNamespace(oq=types.SimpleNamespace(
                mea_tau_phi=False)), tlw=tlw, tlb=tlb,
            within_event_model=None, between_eps=between,
            sig=np.zeros((E, M), F32))
        res = np.zeros((M, N, E), F32)
        for m, imt in enumerate(imts):
            res[m] = GmfComputer._compute(self, mean_stds[:, m], m,
                                          from_string(imt), gsim, within[m],
                                          idxs)
        out['gmf_' + tag] = res
        print(tag, res.shape, res.dtype, float(res[0].mean()))
    np.savez_compressed(os.path.join(GOLDEN, 'gmf.npz'), **out)


if __name__ == '__main__':
    main()
