#!/usr/bin/env python
"""Benchmark of the classical-PSHA hot path (BASELINE.json metric:
rupture x site x IMT x level PoE evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload cfg3|cfg4|cfg5] [--ruptures R] [--sites N]

A step is one fused pass ruptures x sites -> pne map over the whole workload
(cfg3: 100k planar ruptures x 10k sites, BooreEtAl2014 + ChiouYoungs2014,
PGA + 6 SA x 20 levels = 2.8e11 PoE evaluations).  With N > 1 GPUs (torchrun,
one rank per GPU) the sites are sharded: every rank owns its own 10k-site
tile and streams all the ruptures, no data-path collective (weak scaling).
`value` is measured with inputs resident in HBM; `e2e` goes through the
C-ABI entry point taking HOST buffers (oqg_hazard_planar_host: H2D of
ruptures and sites, map in and out).  `--impl reference` times the CPU
restatement of the reference path (oracle/, numpy, all host cores) on a
bounded sample of the same workload.
"""
import os
import sys
import json
import time
import ctypes
import argparse
import threading
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

IMTS = ['PGA', 'SA(0.1)', 'SA(0.2)', 'SA(0.5)', 'SA(1.0)', 'SA(2.0)',
        'SA(3.0)']
WORKLOADS = {
    # name: (ruptures, sites per GPU, gsims)
    'cfg3': (100_000, 10_000, ['BooreEtAl2014', 'ChiouYoungs2014']),
    'cfg4': (2_000_000, 100_000, ['BooreEtAl2014', 'ChiouYoungs2014',
                                  'AbrahamsonEtAl2014',
                                  'CampbellBozorgnia2014']),
    'cfg5': (500_000, 50_000, ['BooreEtAl2014', 'ChiouYoungs2014']),
}
# algorithmic FP64 flop convention of SURVEY.md section 8d
FLOP_GMPE = {'BooreEtAl2014': 220, 'ChiouYoungs2014': 420,
             'AbrahamsonEtAl2014': 560, 'CampbellBozorgnia2014': 520,
             'BooreAtkinson2008': 200}
FLOP_DIST = {'rrup': 40, 'rjb': 1300, 'rx': 330, 'ry0': 660}
FLOP_POE = 64
DISTS = {'BooreEtAl2014': ('rrup', 'rjb'),
         'ChiouYoungs2014': ('rrup', 'rjb', 'rx'),
         'AbrahamsonEtAl2014': ('rrup', 'rjb', 'rx', 'ry0'),
         'CampbellBozorgnia2014': ('rrup', 'rjb', 'rx')}
FP64_NOMINAL_TFLOPS = 37.0   # B200 (HGX) FP64 vector peak, for context


def synth(U, N, seed):
    """SURVEY.md section 8d synthetic ruptures and sites"""
    from helpers import synth_ruptures, synth_sites
    rng = np.random.default_rng(seed)
    return synth_ruptures(U, rng), synth_sites(N, rng)


def flops_per_row(gsim, M, L1):
    return (sum(FLOP_DIST[d] for d in DISTS[gsim]) + M * FLOP_GMPE[gsim] +
            M * L1 * FLOP_POE)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], False

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index),
                     '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop = True
        rows = [r for r in self.rows if len(r) >= 8]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        sm = sorted(float(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        reasons = [n for k, n in enumerate(names)
                   if any(r[4 + k] == 'Active' for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][1]),
                'power_w_max': max(float(r[2]) for r in rows),
                'reasons': reasons, 'samples': len(rows)}


# ------------------------------------------------------------ reference arm
def _cpu_worker(args):
    (pairs, r, sites, idx, imtls, tl, itime) = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import geo_np
    import hazard_np
    sub = {k: v[idx] for k, v in sites.items()}
    pla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    t0 = time.time()
    ctx = hazard_np.get_ctx(pairs, pla, r['mag'], sub, None, 0)
    L = sum(len(v) for v in imtls.values())
    arr = np.ones((len(idx), L, len(pairs)), np.float32)
    sid_to_row = np.zeros(int(sub['sids'].max()) + 1, np.uint32)
    sid_to_row[sub['sids']] = np.arange(len(idx))
    hazard_np.update(arr, sid_to_row, pairs, ctx, imtls, tl, itime,
                     chunk=50_000)
    return len(ctx), time.time() - t0, float(arr.sum())


def cpu_pass(pairs, r, sites, imtls, cores, pool):
    """One pass of the oracle (CPU restatement of get_ctxs + update) over
    ruptures x sites, site tiles over `cores` processes; returns rows, secs"""
    N = len(sites['sids'])
    tiles = [np.flatnonzero(sites['sids'] % cores == k) for k in range(cores)]
    tiles = [t for t in tiles if len(t)]
    t0 = time.time()
    res = pool.map(_cpu_worker, [(pairs, r, sites, t, imtls, 3., 50.)
                                 for t in tiles])
    dt = time.time() - t0
    return sum(x[0] for x in res), dt


def cpu_baseline(workload, U_sample, N, steps, warmup, seed=20240):
    import multiprocessing as mp
    _, _, names = WORKLOADS[workload]
    pairs = [(n, {}) for n in names]
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    r, sites = synth(U_sample, N, seed)
    cores = os.cpu_count() or 1
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            cpu_pass(pairs, {k: v[:2] for k, v in r.items()}, sites, imtls,
                     cores, pool)
        times, rows = [], 0
        for _ in range(steps):
            rows, dt = cpu_pass(pairs, r, sites, imtls, cores, pool)
            times.append(dt)
    evals = rows * len(IMTS) * 20 * len(pairs)
    dt = sum(times) / len(times)
    return evals / dt, cores, dt, evals


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    U, N, names = WORKLOADS[args.workload]
    N = args.sites or N
    U_sample = args.cpu_ruptures
    value, cores, dt, evals = cpu_baseline(args.workload, U_sample, N,
                                           args.steps, min(args.warmup, 1))
    sample = ('first %d of %d ruptures x %d sites (%.3g evals per step), '
              'numpy restatement of get_ctxs + update, site tiles '
              '(sids %% %d) over %d processes'
              % (U_sample, args.ruptures or U, N, evals, cores, cores))
    print(json.dumps({
        'impl': 'reference', 'metric': 'PoE evaluations per second',
        'value': value, 'unit': 'evals/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': workload_name(args), 'gsims': names,
                   'imts': len(IMTS), 'levels': 20},
        'cpu_baseline': {'value': value, 'unit': 'evals/s', 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


def workload_name(args):
    U, N, names = WORKLOADS[args.workload]
    return '%s: %d planar ruptures x %d sites/GPU, %s, %d IMTs x 20 levels' % (
        args.workload, args.ruptures or U, args.sites or N, '+'.join(names),
        len(IMTS))


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import oq_engine_b200 as oq
    from oq_engine_b200.contexts import PlanarRuptures, _dev, _ptr
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    U, N, names = WORKLOADS[args.workload]
    U, N = args.ruptures or U, args.sites or N
    M, L1, G = len(IMTS), 20, len(names)
    L = M * L1
    imtls = {imt: np.logspace(-2, 0.3, L1) for imt in IMTS}
    # cfg3 / cfg4: every rank the same ruptures, its own tile of N sites
    # (site sharding, no collective, weak scaling).  cfg5: 8 source groups
    # of U ruptures over the SAME N sites, rank k folds groups k, k + world,
    # ... into its own pne map and the maps are multiplied with one NCCL
    # all-reduce (PRODUCT) inside the timed region (strong scaling).
    by_groups = args.workload == 'cfg5'
    NGROUPS = 8
    my_groups = list(range(rank, NGROUPS, world)) if by_groups else [0]
    r, _ = synth(U, 4, 20240 + my_groups[0])
    _, sites = synth(4, N, 20240 + (1000 if by_groups else 1000 * (rank + 1)))
    cm = oq.ContextMaker('*', [oq.get_gsim(n) for n in names], dict(
        imtls=imtls, truncation_level=3., investigation_time=50.),
        device=local)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    h = cm.handle()
    if args.fused_subs or args.plain_table:
        h.call('oqg_set_fused_config', args.fused_subs, int(args.plain_table))
    dev = 'cuda:%d' % local
    L_ = oq._lib.lib()
    # ---- device resident inputs
    flat = oq.planar.as_flat(pla)
    mag = np.ascontiguousarray(r['mag'])
    md = cm.magdist(r['mag'])
    t_pla, t_mag, t_md = _dev(flat, local), _dev(mag, local), _dev(md, local)
    st, keep = cm._sites_struct(sc, local)
    t_map = torch.ones((N, L, G), dtype=torch.float32, device=dev)
    t_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    group_inputs = [(t_pla, t_mag, t_md)]
    for g in my_groups[1:]:
        rg, _ = synth(U, 4, 20240 + g)
        pg = oq.build_planar(rg['mag'], rg['strike'], rg['dip'], rg['rake'],
                             rg['rate'], rg['area'], rg['hypo_lon'],
                             rg['hypo_lat'], rg['hypo_dep'], 0., 25., 1.5)
        group_inputs.append((_dev(oq.planar.as_flat(pg), local),
                             _dev(np.ascontiguousarray(rg['mag']), local),
                             _dev(cm.magdist(rg['mag']), local)))

    def step_device():
        if by_groups:
            t_map.fill_(1.)
        for a, b, c in group_inputs:
            h.call('oqg_hazard_planar', _ptr(a), _ptr(b), _ptr(c),
                   None, 0, U, ctypes.byref(st), None, N, 0., 50., 0,
                   _ptr(t_map), _ptr(t_cnt))
        if by_groups and world > 1:
            dist.all_reduce(t_map, op=dist.ReduceOp.PRODUCT)

    # ---- pinned host inputs for the end-to-end arm
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    p_pla, p_mag, p_md = pinned(flat), pinned(mag), pinned(md)
    hs = oq._lib.OqgSites()
    hkeep = []
    for name in ('lon', 'lat', 'depth', 'vs30', 'z1pt0', 'z2pt5'):
        t = pinned(np.asarray(sc.array[name], float))
        hkeep.append(t)
        setattr(hs, name, t.data_ptr())
    t = pinned(np.asarray(sc.array['vs30measured']).view(np.uint8))
    hkeep.append(t)
    hs.vs30measured = t.data_ptr()
    p_map = torch.ones((N, L, G), dtype=torch.float32).pin_memory()
    kept = ctypes.c_ulonglong(0)
    h2d = (flat.nbytes + mag.nbytes + md.nbytes + 6 * N * 8 + N +
           p_map.numel() * 4)
    d2h = p_map.numel() * 4 + 8

    def step_host():
        h.call('oqg_hazard_planar_host', ctypes.c_void_p(p_pla.data_ptr()),
               ctypes.c_void_p(p_mag.data_ptr()),
               ctypes.c_void_p(p_md.data_ptr()), None, 0, U,
               ctypes.byref(hs), N, 0., 50., 0,
               ctypes.c_void_p(p_map.data_ptr()), ctypes.byref(kept))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 is flushed between the timed iterations by overwriting a buffer of
    # twice its size (126 MB on B200); the fill runs inside the timed region
    # (0.05 ms per step at the measured HBM bandwidth, 0.01 % of a step)
    l2_flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            l2_flush.zero_()
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L_.oqg_launch_count(h._h)
    t_cnt.zero_()
    ms = timed(step_device, args.steps)
    launches = L_.oqg_launch_count(h._h) - launches0
    pairs_kept = int(t_cnt[0].item()) // args.steps
    # per-kernel durations of the last step (CUDA events on the launching
    # stream, recorded inside oqg_hazard_planar around each fused launch)
    kms = (ctypes.c_float * 16)()
    nk = L_.oqg_last_fused_ms(h._h, kms, 16)
    kernel_ms = [float(kms[g]) for g in range(max(nk, 0))]
    clocks = sampler.summary() if rank == 0 else None
    # ---- end to end through the host entry point
    step_host()
    ms_e2e = timed(step_host, args.steps)
    assert kept.value * len(my_groups) == pairs_kept, (kept.value, pairs_kept)
    evals_step = pairs_kept * L * G          # per rank
    if by_groups:                            # total over the ranks
        tot = torch.tensor([evals_step], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        value = float(tot.item()) / (ms / args.steps * 1e-3)
        # the host entry point takes one group: scale to the same units
        e2e = pairs_kept / len(my_groups) * L * G * world / (
            ms_e2e / args.steps * 1e-3)
    else:
        value = evals_step * world / (ms / args.steps * 1e-3)
        e2e = evals_step * world / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (the slowest fused launch)
    peak = ctypes.c_double(0)
    h.call('oqg_measure_fp64_peak', ctypes.byref(peak))
    gdom = int(np.argmax(kernel_ms)) if kernel_ms else 0
    fl = flops_per_row(names[gdom], M, L1) * pairs_kept / len(group_inputs)
    ach = fl / (kernel_ms[gdom] * 1e-3) / 1e12 if kernel_ms else None
    alg_bytes = U * 65 * 8 + N * (7 * 8 + 33) + N * L * 8 * 2
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            hbm_peak = json.load(f)['hbm_gbs']
        hbm_src = 'MEASURED_PEAKS.json'
    except Exception:
        hbm_peak, hbm_src = 6650., 'fallback'
    # DRAM traffic and pipe utilisation of the same kernel from the committed
    # ncu capture (profiles/r1_v12_traffic.json; cfg3 full size), if the
    # workload is the one that was profiled
    traffic, pipes = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_v12_traffic.json')) as f:
            prof = json.load(f)
        k = prof['kernels'].get('k_fused<%s>' % names[gdom])
        if k and args.workload == 'cfg3' and not args.ruptures and \
                not args.sites:
            traffic = k['dram_bytes_read'] + k['dram_bytes_write']
            pipes = {'source': 'ncu, profiles/r1_v12_traffic.json',
                     'l1tex_lsu_data_pipe_pct': k['lsu_data_pipe_pct'],
                     'issue_slots_pct': k['issue_active_pct'],
                     'fp64_pipe_pct': k['fp64_pipe_pct']}
    except Exception:
        pass
    roofline = {
        'bound': 'fp64', 'kernel': 'k_fused<%s>' % names[gdom],
        'achieved': ach, 'peak': peak.value, 'unit': 'TFLOP/s',
        'frac': ach / peak.value if ach else None, 'traffic': traffic,
        'pipes': pipes,
        'convention': 'achieved = SURVEY 8d algorithmic FP64 flop of the '
                      'reference operation mix (erf = 50 flop, GMPE 220-560, '
                      'distances 40-1300 per row) / launch time; the kernel '
                      'executes fewer operations than that (table-driven PoE '
                      'and log/exp), so frac > 1 is possible: it measures '
                      'speed against "the reference arithmetic at the FP64 '
                      'peak"; the hardware limit of the kernel is the '
                      'L1/shared-memory data pipe (pipes)',
        'peak_source': 'DFMA microbenchmark (oqg_measure_fp64_peak) run in '
                       'this process; nominal %.0f' % FP64_NOMINAL_TFLOPS,
        'flop_per_row': flops_per_row(names[gdom], M, L1),
        'kernel_ms': kernel_ms,
        'hbm': {'achieved': alg_bytes / (kernel_ms[gdom] * 1e-3) / 1e9
                if kernel_ms else None, 'peak': hbm_peak, 'unit': 'GB/s',
                'peak_source': hbm_src,
                'note': 'algorithmic bytes per launch; not the bound'},
    }
    # ---- CPU baseline on a bounded sample, rank 0 only
    cpu = None
    if not args.no_cpu and world == 1:      # rank 0 at N = 1 only
        v, cores, dt, ev = cpu_baseline(args.workload, args.cpu_ruptures, N,
                                        1, 1)
        cpu = {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port',
               'sample': 'first %d ruptures x %d sites (%.3g evals, %.1f s)'
               % (args.cpu_ruptures, N, ev, dt)}
    print(json.dumps({
        'metric': 'PoE evaluations per second', 'value': value,
        'unit': 'evals/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
        'higher_is_better': True,
        'scaling': 'strong' if by_groups else 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'gsims': names,
                   'imts': M, 'levels': L1,
                   'sharding': ('source groups: %d groups, NCCL all-reduce '
                                '(PRODUCT) of the pne maps, %d bytes' % (
                                    NGROUPS, N * L * G * 4)
                                if by_groups else 'sites'),
                   'pairs_per_step_per_gpu': pairs_kept,
                   'l2': 'flushed between the timed iterations by '
                         'overwriting a 256 MB buffer (inside the timed '
                         'region); the kernel is compute bound: per step '
                         'it streams %.0f MB of ruptures per site tile '
                         'and rewrites the map and the partial sums'
                         % (U * 65 * 8 / 1e6)},
        'e2e': {'value': e2e, 'unit': 'evals/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=list(WORKLOADS))
    ap.add_argument('--ruptures', type=int, default=0)
    ap.add_argument('--sites', type=int, default=0)
    ap.add_argument('--cpu-ruptures', type=int, default=768,
                    help='ruptures of the bounded CPU sample (x all sites)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--fused-subs', type=int, default=0,
                    help='A/B: cap of the sub-CTAs per CTA (0 = automatic)')
    ap.add_argument('--plain-table', action='store_true',
                    help='A/B: one work item per CTA, unreplicated PoE table')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
