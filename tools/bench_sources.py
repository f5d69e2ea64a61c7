"""Throughput of the rupture generator k_point_ruptures (SURVEY.md section
8f row 1) against the HBM roofline: one JSON line.

    python tools/bench_sources.py [--sources S]

An area-source-like batch: S point sources sharing one MFD (30 magnitudes),
4 nodal planes and 3 hypocentral depths -> 360 ruptures per source (390 with
the pointsource_distance approximation).  Algorithmic bytes: 336 (planar
record) + 8 (magnitude) + 8 (integration distance) [+ 40 (aux)] written per
rupture; the source description read is 64 B per source.
"""
import os
import sys
import json
import ctypes
import argparse
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sources', type=int, default=20_000)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    import torch
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import get_handle
    h = get_handle(0)
    rng = np.random.default_rng(2)
    mfd = oq.TruncatedGRMFD(5.0, 8.0, 0.1, 4.0, 1.0)
    npd = [(0.25, 0., 90., 0.), (0.25, 90., 60., 90.), (0.25, 180., 45., -90.),
           (0.25, 270., 30., 90.)]
    hdd = [(0.3, 5.), (0.4, 10.), (0.3, 15.)]
    srcs = [oq.PointSource(str(i), lon, lat, mfd, 'WC1994', 1.5, 0., 25., npd,
                           hdd)
            for i, (lon, lat) in enumerate(zip(rng.uniform(0, 3, a.sources),
                                               rng.uniform(0, 3, a.sources)))]
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
        src = 'MEASURED_PEAKS.json'
    except Exception:
        peak, src = 6650., 'fallback'
    for psd in (0., 100.):
        b = oq.PointSourceBatch(srcs, psd)
        md = np.full(len(b.mag), 300.)
        st, keep = b.to_device(0, md)
        U = b.U
        t_pla = torch.empty((U, 42), dtype=torch.float64, device='cuda')
        t_mag = torch.empty(U, dtype=torch.float64, device='cuda')
        t_md = torch.empty(U, dtype=torch.float64, device='cuda')
        t_aux = torch.empty((U, 5), dtype=torch.float64, device='cuda')

        def run():
            h.call('oqg_point_ruptures', ctypes.byref(st), b.S, U,
                   ctypes.c_void_p(t_pla.data_ptr()),
                   ctypes.c_void_p(t_mag.data_ptr()),
                   ctypes.c_void_p(t_md.data_ptr()),
                   ctypes.c_void_p(t_aux.data_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        nbytes = U * (336 + 8 + 8 + 40.) + b.S * 64.
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({
            'kernel': 'k_point_ruptures', 'pointsource_distance': psd,
            'sources': b.S, 'ruptures': U, 'ms': ms,
            'ruptures_per_s': U / (ms * 1e-3),
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': peak,
                         'unit': 'GB/s', 'frac': gbs / peak,
                         'peak_source': src, 'algorithmic_bytes': nbytes}}))


if __name__ == '__main__':
    main()
