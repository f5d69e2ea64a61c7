"""Throughput of k_mesh_min_distance against
the FP64 roofline: one JSON line.

    python tools/bench_mesh.py [--sections S] [--points P] [--sites N]

Algorithmic work: 8 FP64 flop per (mesh point, site) pair (3 subtractions,
3 multiply-adds counted as 5 flop with the first multiply, 1 minimum is not
counted); bytes are negligible (24 B per point and per site, 8 B per output).
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
    ap.add_argument('--sections', type=int, default=2000)
    ap.add_argument('--points', type=int, default=500)
    ap.add_argument('--sites', type=int, default=100_000)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    import torch
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import get_handle
    rng = np.random.default_rng(1)
    meshes = [oq.Mesh(rng.uniform(0, 3, a.points), rng.uniform(0, 3, a.points),
                      rng.uniform(0, 20, a.points)) for _ in range(a.sections)]
    sites = oq.Mesh(rng.uniform(0, 3, a.sites), rng.uniform(0, 3, a.sites))
    h = get_handle(0)
    for _ in range(2):
        oq.mesh_min_distance(meshes[:50], sites, as_tensor=True)
    torch.cuda.synchronize()
    # time the kernel itself: inputs resident, outputs on the device
    from oq_engine_b200.contexts import _dev, _ptr
    from oq_engine_b200.mesh import _sites_struct
    off = np.arange(a.sections + 1, dtype=np.int64) * a.points
    t_lon = _dev(np.concatenate([m.lons for m in meshes]), 0)
    t_lat = _dev(np.concatenate([m.lats for m in meshes]), 0)
    t_dep = _dev(np.concatenate([m.depths for m in meshes]), 0)
    t_off = _dev(off, 0)
    st, keep, N = _sites_struct(sites, 0)
    out = torch.empty((a.sections, N), dtype=torch.float64, device='cuda')

    def run():
        h.call('oqg_mesh_min_distance', _ptr(t_lon), _ptr(t_lat), _ptr(t_dep),
               int(off[-1]), _ptr(t_off), a.sections, ctypes.byref(st), N,
               _ptr(out))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    pairs = float(a.sections) * a.points * a.sites
    peak = ctypes.c_double(0)
    h.call('oqg_measure_fp64_peak', ctypes.byref(peak))
    tf = pairs * 8 / (ms * 1e-3) / 1e12
    print(json.dumps({
        'kernel': 'k_mesh_min_distance', 'ms': ms, 'sections': a.sections,
        'points_per_section': a.points, 'sites': a.sites,
        'point_site_pairs_per_s': pairs / (ms * 1e-3),
        'roofline': {'bound': 'fp64', 'achieved': tf, 'peak': peak.value,
                     'unit': 'TFLOP/s', 'frac': tf / peak.value,
                     'peak_source': 'DFMA microbenchmark in this process',
                     'flop_per_pair': 8}}))

    # ---- k_mesh_rx_ry0: S top edges of V vertices x N sites -----------------
    V = 20
    S2 = min(a.sections, 1000)
    elon = rng.uniform(0, 3, (S2, 1)) + np.cumsum(
        rng.uniform(0.01, 0.03, (S2, V)), axis=1)
    elat = rng.uniform(0, 3, (S2, 1)) + np.cumsum(
        rng.uniform(-0.01, 0.02, (S2, V)), axis=1)
    t_elon, t_elat = _dev(elon.ravel(), 0), _dev(elat.ravel(), 0)
    t_eoff = _dev(np.arange(S2 + 1, dtype=np.int64) * V, 0)
    t_strike = _dev(rng.uniform(0, 360, S2), 0)
    rx = torch.empty((S2, N), dtype=torch.float64, device='cuda')
    ry0 = torch.empty((S2, N), dtype=torch.float64, device='cuda')

    def run2():
        h.call('oqg_mesh_rx_ry0', _ptr(t_elon), _ptr(t_elat), _ptr(t_eoff),
               _ptr(t_strike), S2, ctypes.byref(st), N, _ptr(rx), _ptr(ry0))
    run2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        run2()
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / a.steps
    segs = float(S2) * (V - 1) * N
    out_bytes = 2. * S2 * N * 8
    print(json.dumps({
        'kernel': 'k_mesh_rx_ry0', 'ms': ms2, 'surfaces': S2,
        'vertices_per_edge': V, 'sites': a.sites,
        'segment_site_pairs_per_s': segs / (ms2 * 1e-3),
        'output_GBps': out_bytes / (ms2 * 1e-3) / 1e9}))


if __name__ == '__main__':
    main()
