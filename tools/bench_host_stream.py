"""A stream of small sources from host memory (the reference's loop over the
sources of a group, contexts.py:1548-1579): oqg_hazard_planar_host called per
source (sites + map up, map down, a synchronisation per call) against one
session oqg_host_begin / add / end (sites and map resident, double-buffered
uploads).  One JSON line.

    python tools/bench_host_stream.py [--sources S] [--ruptures U] [--sites N]
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
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sources', type=int, default=64)
    ap.add_argument('--ruptures', type=int, default=400)
    ap.add_argument('--sites', type=int, default=10_000)
    ap.add_argument('--steps', type=int, default=3)
    a = ap.parse_args()
    import torch
    import oq_engine_b200 as oq
    from helpers import synth_ruptures, synth_sites, IMTS
    S, U, N = a.sources, a.ruptures, a.sites
    rng = np.random.default_rng(11)
    r = synth_ruptures(S * U, rng)
    sites = synth_sites(N, rng)
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    cm = oq.ContextMaker('*', [oq.get_gsim('BooreEtAl2014'),
                               oq.get_gsim('ChiouYoungs2014')], dict(
        imtls=imtls, truncation_level=3., investigation_time=50.))
    h = cm.handle()
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)

    def pinned(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        return t
    flat = pinned(oq.planar.as_flat(pla))
    mag = pinned(r['mag'])
    md = pinned(cm.magdist(r['mag']))
    st = oq._lib.OqgSites()
    keep = {}
    for name in ('lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5'):
        keep[name] = pinned(np.asarray(sites[name], float))
        setattr(st, name, keep[name].data_ptr())
    keep['m'] = pinned(np.ascontiguousarray(sites['vs30measured']).view(
        np.uint8))
    st.vs30measured = keep['m'].data_ptr()
    out = torch.ones((N, cm.L, cm.G), dtype=torch.float32).pin_memory()

    def serial():
        out.fill_(1.)
        for s in range(S):
            h.call('oqg_hazard_planar_host', flat[s * U:].data_ptr(),
                   mag[s * U:].data_ptr(), md[s * U:].data_ptr(), None, 0, U,
                   ctypes.byref(st), N, 0., 50., 0, out.data_ptr(), None)
        return out.numpy().copy()

    def session():
        h.call('oqg_host_begin', ctypes.byref(st), N, 0, None)
        for s in range(S):
            h.call('oqg_host_add', flat[s * U:].data_ptr(),
                   mag[s * U:].data_ptr(), md[s * U:].data_ptr(), None, 0, U,
                   0., 50.)
        h.call('oqg_host_end', out.data_ptr(), None)
        return out.numpy().copy()

    res = {}
    for name, fn in (('per_call', serial), ('session', session)):
        ref = fn()
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            fn()
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0) / a.steps * 1e3, ref
    same = bool((res['per_call'][1] == res['session'][1]).all())
    evals = float(S) * U * N * cm.L * cm.G
    print(json.dumps({
        'bench': 'host_stream', 'sources': S, 'ruptures_per_source': U,
        'sites': N, 'gsims': cm.G, 'levels': cm.L,
        'per_call_ms': res['per_call'][0], 'session_ms': res['session'][0],
        'per_call_evals_per_s': evals / res['per_call'][0] * 1e3,
        'session_evals_per_s': evals / res['session'][0] * 1e3,
        'identical_maps': same,
        'h2d_bytes_per_source': U * (42 + 2) * 8,
        'map_bytes': int(out.numel() * 4)}))


if __name__ == '__main__':
    main()
