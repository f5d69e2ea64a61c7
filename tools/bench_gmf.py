"""Throughput of the sampling side of the ground motion fields (SURVEY.md
section 8f row 4): oqg_truncnorm_rvs (numpy's PCG64 stream + truncated-normal
inverse CDF) against the HBM roofline of its float32 output, and one
GmfComputer.compute_all with and without the JB2009 spatial correlation.
One JSON line per measurement.

    python tools/bench_gmf.py [--samples S] [--sites N] [--events E]
"""
import os
import sys
import json
import time
import ctypes
import argparse
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--samples', type=int, default=200_000_000)
    ap.add_argument('--sites', type=int, default=4000)
    ap.add_argument('--events', type=int, default=1000)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    import torch
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import get_handle, _ptr
    h = get_handle(0)
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
        src = 'MEASURED_PEAKS.json'
    except Exception:
        peak, src = 6650., 'fallback'
    st = np.random.PCG64(42).state['state']
    mask = (1 << 64) - 1
    state = (ctypes.c_uint64 * 2)(st['state'] >> 64, st['state'] & mask)
    inc = (ctypes.c_uint64 * 2)(st['inc'] >> 64, st['inc'] & mask)
    n = a.samples
    out = torch.empty(n, dtype=torch.float32, device='cuda')

    def rvs():
        h.call('oqg_truncnorm_rvs', state, inc, 0, n, 3., _ptr(out))
    for _ in range(3):
        rvs()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps):
        rvs()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # scipy on one host core, on a bounded sample
    from scipy import stats
    m = 2_000_000
    t0 = time.perf_counter()
    ref = stats.truncnorm(-3., 3.).rvs(m, np.random.default_rng(42)
                                       ).astype(np.float32)
    dt = time.perf_counter() - t0
    same = float((out[:m].cpu().numpy() == ref).mean())
    gbs = 4. * n / (ms * 1e-3) / 1e9
    print(json.dumps({
        'kernel': 'k_truncnorm_rvs', 'samples': n, 'ms': ms,
        'samples_per_s': n / (ms * 1e-3),
        'scipy_one_core_samples_per_s': m / dt,
        'fraction_bitwise_equal_to_scipy': same,
        'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak,
                     'unit': 'GB/s', 'frac': gbs / peak, 'peak_source': src,
                     'algorithmic_bytes': 4. * n,
                     'note': 'the float32 samples written; the kernel is '
                             'bound by the FP64 inverse CDF, not by HBM'}}))
    # one rupture, N sites, E events, 3 IMTs, 2 gsims
    N, E = a.sites, a.events
    rng = np.random.default_rng(1)
    lons, lats = rng.uniform(0, 1., N), rng.uniform(0, 1., N)
    imts = ['PGA', 'SA(0.3)', 'SA(1.0)']
    cm = oq.ContextMaker('*', [oq.get_gsim('BooreAtkinson2008'),
                               oq.get_gsim('AkkarEtAlRjb2014')], dict(
        imtls={i: [0.1] for i in imts}, truncation_level=3.))
    ctx = np.zeros(N, [('mag', float), ('rake', float), ('vs30', float),
                       ('rjb', float), ('rrup', float), ('sids', np.uint32)]
                   ).view(np.recarray)
    ctx.mag, ctx.rake, ctx.vs30 = 6.5, 0., 760.
    ctx.rjb = rng.uniform(1, 100, N)
    ctx.rrup = ctx.rjb + 1.
    ctx.sids = np.arange(N)
    for tag, corr in (('plain', None),
                      ('jb2009', oq.JB2009CorrelationModel(False))):
        def run():
            gc = oq.GmfComputer(ctx, lons, lats, cm,
                                [np.array([0]), np.array([1])], n_occ=E,
                                seed=7, scenario=True, correlation_model=corr)
            return gc.compute_all()
        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data = run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({
            'bench': 'GmfComputer.compute_all', 'correlation': tag,
            'sites': N, 'events': E, 'imts': len(imts), 'gsims': 2,
            'seconds': dt, 'gmvs': int(len(data['eid'])) * len(imts),
            'gmvs_per_s': len(data['eid']) * len(imts) / dt,
            'note': 'wall time of the second call (Cholesky factors of the '
                    'first call are NOT reused: a new model object per call '
                    'would be; here the cache of the model is warm), '
                    'including the device -> host copy of the data frame '
                    'columns'}))


if __name__ == '__main__':
    main()
