"""Throughput of the post-processing kernels
against the HBM roofline: one JSON line per kernel.

    python tools/bench_post.py [--sites N] [--rlzs R] [--gsims Gt]

k_postclassical reads N*L*Gt float32 rates and writes (R + S) * N * L float32
curves: algorithmic bytes = 4 * N * L * (Gt + R + S).  k_hazard_maps reads
K*M*L1 float32 and writes K*M*P float32.
"""
import os
import sys
import json
import argparse
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sites', type=int, default=100_000)
    ap.add_argument('--rlzs', type=int, default=32)
    ap.add_argument('--gsims', type=int, default=8)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    import torch
    import ctypes
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import get_handle, _dev, _ptr
    h = get_handle(0)
    N, M, L1, Gt, R = a.sites, 7, 20, a.gsims, a.rlzs
    L = M * L1
    rng = np.random.default_rng(3)
    rates = torch.rand((N, L, Gt), dtype=torch.float32, device='cuda') * 1e-2
    # every realization picks one gsim column (one TRT)
    pick = rng.integers(0, Gt, R)
    rlzs = [np.where(pick == g)[0] for g in range(Gt)]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rlzs])]).astype(np.int32)
    t_off, t_rlz = _dev(off, 0), _dev(np.concatenate(rlzs).astype(np.int32), 0)
    w = rng.uniform(0.5, 1.5, R)
    w /= w.sum()
    t_w = _dev(w[None], 0)
    names = ['mean', 'std', 'quantile-0.05', 'quantile-0.5', 'quantile-0.95']
    S = len(names)
    kinds = (ctypes.c_int32 * S)(0, 1, 2, 2, 2)
    qs = (ctypes.c_double * S)(0, 0, .05, .5, .95)
    ws = (ctypes.c_double * 1)(1.)
    t_r = torch.empty((R, N, L), dtype=torch.float32, device='cuda')
    t_s = torch.empty((S, N, L), dtype=torch.float32, device='cuda')

    def post():
        h.call('oqg_postclassical', _ptr(rates), N, L, Gt, M, _ptr(t_off),
               _ptr(t_rlz), R, _ptr(t_w), 1, ws, S, kinds, qs, 0, _ptr(t_r),
               _ptr(t_s))

    def post_ungrouped():
        # one group per realization: what R distinct sets of columns cost
        h.call('oqg_set_post_grouping', 0)
        post()
        h.call('oqg_set_post_grouping', 1)
    P = 2
    t_i = _dev(np.tile(np.logspace(-2, 0.3, L1), (M, 1)), 0)
    t_p = _dev(np.array([0.1, 0.02]), 0)
    K = (R + S) * N
    curves = torch.cat([t_r, t_s])
    t_o = torch.empty((K, M, P), dtype=torch.float32, device='cuda')

    def hmaps():
        h.call('oqg_hazard_maps', _ptr(curves), K, M, L1, _ptr(t_i), P,
               _ptr(t_p), _ptr(t_o))
    t_gw = _dev(rng.uniform(0.1, 1., (1, Gt)), 0)
    t_mr = torch.empty((N, L), dtype=torch.float64, device='cuda')

    def mean_rates():
        h.call('oqg_mean_rates', _ptr(rates), N, L, Gt, M, _ptr(t_gw), 1,
               _ptr(t_mr))
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
        src = 'MEASURED_PEAKS.json'
    except Exception:
        peak, src = 6650., 'fallback'
    for name, fn, nbytes in (
            ('k_postclassical', post, 4. * N * L * (Gt + R + S)),
            ('k_postclassical_ungrouped', post_ungrouped,
             4. * N * L * (Gt + R + S)),
            ('k_hazard_maps', hmaps, 4. * K * M * (L1 + P)),
            ('k_mean_rates', mean_rates, 4. * N * L * Gt + 8. * N * L)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({
            'kernel': name, 'ms': ms, 'sites': N, 'levels': L, 'gsims': Gt,
            'rlzs': R, 'stats': names,
            'site_curves_per_s': N * (R + S) / (ms * 1e-3),
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak,
                         'unit': 'GB/s', 'frac': gbs / peak,
                         'peak_source': src,
                         'algorithmic_bytes': nbytes}}))


if __name__ == '__main__':
    main()
