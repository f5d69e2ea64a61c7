"""BASELINE.json configs[0], demos/hazard/AreaSourceClassicalPSHA, end to
end on the device: area source -> point sources -> ruptures -> hazard curves
-> hazard maps at the job.ini's poes; prints one JSON line with the wall
times.  (The reference's hot loop for the same inputs, ContextMaker.get_ctxs
+ update through hazardlib with numba warm, took 7.3 s on one core of the
build container: DESIGN.md section 8.)

    python tools/demo_cfg1.py
"""
import os
import sys
import json
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def logscale(x_min, x_max, n):
    # valid.logscale of the reference
    return np.exp(np.linspace(np.log(x_min), np.log(x_max), n))


def main():
    import torch
    import oq_engine_b200 as oq
    imts = ['PGA', 'SA(0.025)', 'SA(0.05)', 'SA(0.1)', 'SA(0.2)', 'SA(0.5)',
            'SA(1.0)', 'SA(2.0)']
    imtls = {imt: logscale(0.005, 2.13, 45) for imt in imts}
    poes = [0.1, 0.02]

    def job():
        t = {}
        t0 = time.time()
        area = oq.AreaSource(
            '1', [-0.5, -0.3, 0.1, 0.3], [-0.5, -0.1, 0.2, -0.8], 10.,
            oq.TruncatedGRMFD(5.0, 6.5, 0.2, 4.5, 1.0), 'WC1994', 1., 0., 10.,
            [(1., 0., 90., 0.)], [(1., 5.)],
            tectonic_region_type='Active Shallow Crust')
        srcs = list(area)
        lons, lats = np.meshgrid(np.arange(-1.0, 1.0001, 0.09),
                                 np.arange(-1.5, 0.7001, 0.09))
        N = lons.size
        sc = oq.SiteCollection.from_points(
            lons.ravel(), lats.ravel(), np.zeros(N),
            dict(vs30=600., vs30measured=True, z1pt0=100., z2pt5=5.))
        t['sources_and_sites_s'] = time.time() - t0
        t0 = time.time()
        curves = oq.calc_hazard_curves(
            srcs, sc, imtls, {'Active Shallow Crust':
                              oq.get_gsim('BooreAtkinson2008')},
            truncation_level=3., investigation_time=50.,
            maximum_distance=oq.IntegrationDistance.new('200'))
        torch.cuda.synchronize()
        t['hazard_curves_s'] = time.time() - t0
        t0 = time.time()
        arr = np.stack([curves[imt] for imt in imts], axis=1)   # (N, M, L1)
        hmaps = oq.compute_hmaps(arr, imtls, poes)
        t['hazard_maps_s'] = time.time() - t0
        return t, len(srcs), N, curves, hmaps
    job()                      # warm up: library load, handle, first launch
    t, S, N, curves, hmaps = job()
    evals = S * 8 * N * 8 * 45
    print(json.dumps({
        'demo': 'AreaSourceClassicalPSHA', 'point_sources': S, 'ruptures': S * 8,
        'sites': N, 'imts': 8, 'levels': 45, 'timing': t,
        'total_s': sum(t.values()),
        'poe_evals_upper_bound': evals,
        'max_poe_PGA': float(curves['PGA'].max()),
        'hazard_map_PGA_10pct_max_g': float(hmaps[:, 0, 0].max()),
        'reference_hot_loop_s': 7.3,
        'reference_note': 'hazardlib get_ctxs + update, one core of the '
                          'build container, numba warm (DESIGN.md section 8)'}))


if __name__ == '__main__':
    main()
