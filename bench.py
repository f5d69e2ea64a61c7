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
    r, sites = synth_ruptures(U, rng), synth_sites(N, rng)
    if os.environ.get('OQG_BENCH_SITE_PATCH'):
        # experiment: the caller hands the grid in 8 x 4 patches, not rows
        n_side = int(np.ceil(np.sqrt(N)))
        iy, ix = np.divmod(np.arange(N), n_side)
        order = np.lexsort((ix % 8, iy % 4, ix // 8, iy // 4))
        sites = {k: (v if k == 'sids' else v[order]) for k, v in sites.items()}
    return r, sites


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


def _ref_worker(args):
    # the REAL reference (baseline/_ref), loaded and JIT-compiled in the
    # parent before the fork
    (names, imtls, planar, mags, sites, idx) = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_runner
    sub = {k: v[idx] for k, v in sites.items()}
    rows, dt, arr = ref_runner.hazard_pass(names, imtls, planar, mags, sub)
    return rows, dt, float(arr.sum())


def reference_baseline(workload, U_sample, N, steps, warmup, seed=20240):
    """The unmodified reference (openquake.hazardlib from baseline/_ref) on
    the first U_sample ruptures of the workload x all the sites: its own
    get_ctx + ContextMaker.update, site tiles (sids % cores, the reference's
    tiling rule) over one process per host core.  Returns evals/s, cores,
    seconds per step, evaluations per step."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_runner
    _, _, names = WORKLOADS[workload]
    imtls = {imt: np.logspace(-2, 0.3, 20) for imt in IMTS}
    r, sites = synth(U_sample, N, seed)
    planar = ref_runner.build_planar(r)          # untimed, as on the GPU arm
    cores = os.cpu_count() or 1
    tiles = [np.flatnonzero(sites['sids'] % cores == k) for k in range(cores)]
    tiles = [t for t in tiles if len(t)]
    # JIT-compile everything in this process; the workers are forked from it
    ref_runner.hazard_pass(names, imtls, planar[:2], r['mag'][:2],
                           {k: v[tiles[0][:8]] for k, v in sites.items()})
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        jobs = [(names, imtls, planar, r['mag'], sites, t) for t in tiles]
        for _ in range(warmup):
            pool.map(_ref_worker, [(names, imtls, planar[:2], r['mag'][:2],
                                    sites, t[:8]) for t in tiles])
        times, rows = [], 0
        for _ in range(steps):
            t0 = time.time()
            res = pool.map(_ref_worker, jobs)
            times.append(time.time() - t0)
            rows = sum(x[0] for x in res)
    evals = rows * len(IMTS) * 20 * len(names)
    dt = sum(times) / len(times)
    return evals / dt, cores, dt, evals


def have_reference():
    return os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'openquake',
                                      'hazardlib'))


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
    if have_reference() and not args.port:
        kind = 'reference'
        value, cores, dt, evals = reference_baseline(
            args.workload, U_sample, N, args.steps, min(args.warmup, 1))
        what = ('the unmodified reference (openquake.hazardlib 3.27 from '
                'baseline/_ref): _get_ctx_planar + ContextMaker.update')
    else:
        kind = 'port'
        value, cores, dt, evals = cpu_baseline(
            args.workload, U_sample, N, args.steps, min(args.warmup, 1))
        what = 'numpy restatement of get_ctxs + update (oracle/)'
    sample = ('first %d of %d ruptures x %d sites (%.3g evals per step), '
              '%s, site tiles (sids %% %d) over %d processes'
              % (U_sample, args.ruptures or U, N, evals, what, cores, cores))
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
                         'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


def workload_name(args):
    U, N, names = WORKLOADS[args.workload]
    return '%s: %d planar ruptures x %d sites/GPU, %s, %d IMTs x 20 levels' % (
        args.workload, args.ruptures or U, args.sites or N, '+'.join(names),
        len(IMTS))



# ------------------------------------------------- strong-scaling legs
# Fixed total workloads split over the ranks, run inside the same bench
# process after the headline measurement, so that the driver's per-N lines
# carry both shardings of SURVEY 8e with their collectives timed:
#   strong_sites   cfg4 models, STRONG_SITES ruptures x sites in total, the
#                  sites split by `sids % world` (the reference's tiling rule);
#                  the tiles are gathered into the whole map on every rank
#                  (all_gather, timed separately)
#   strong_groups  cfg5 shape: 8 source groups of STRONG_GROUPS ruptures over
#                  the same sites, rank k folds groups k, k + world, ...; the
#                  pne maps are multiplied with one NCCL all-reduce (PRODUCT)
#                  inside the timed region (timed separately as well)
STRONG_SITES = (100_000, 80_000)       # ruptures, sites in total
STRONG_GROUPS = (50_000, 50_000, 8)    # ruptures per group, sites, groups


def _planar_inputs(oq, cm, U, seed, local):
    from oq_engine_b200.contexts import _dev
    r, _ = synth(U, 4, seed)
    pla = oq.build_planar(r['mag'], r['strike'], r['dip'], r['rake'],
                          r['rate'], r['area'], r['hypo_lon'], r['hypo_lat'],
                          r['hypo_dep'], 0., 25., 1.5)
    return (_dev(oq.planar.as_flat(pla), local),
            _dev(np.ascontiguousarray(r['mag']), local),
            _dev(cm.magdist(r['mag']), local))


def _events_ms(pairs):
    return [a.elapsed_time(b) for a, b in pairs]


def strong_legs(args, oq, rank, world, local, timed_max):
    """The two strong-scaling legs; returns {'strong_sites': {...},
    'strong_groups': {...}} on every rank (values are whole-job)"""
    import torch
    import torch.distributed as dist
    from oq_engine_b200.contexts import _ptr
    dev = 'cuda:%d' % local
    M, L1 = len(IMTS), 20
    L = M * L1
    imtls = {imt: np.logspace(-2, 0.3, L1) for imt in IMTS}
    out = {}
    steps, warm = args.strong_steps, 1

    def cmaker(names):
        return oq.ContextMaker('*', [oq.get_gsim(n) for n in names], dict(
            imtls=imtls, truncation_level=3., investigation_time=50.),
            device=local)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- sites: cfg4 models, fixed grid split by sids % world ------------
    U, N = STRONG_SITES
    U, N = int(U * args.strong_scale), N
    names = WORKLOADS['cfg4'][2]
    G = len(names)
    cm = cmaker(names)
    h = cm.handle()
    inp = _planar_inputs(oq, cm, U, 20240, local)
    _, sites = synth(4, N, 21240)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    tile = oq.sharding.site_tile(sc, rank, world)
    n_k = len(tile)
    st, keep = cm._sites_struct(tile, local)
    t_map = torch.ones((n_k, L, G), dtype=torch.float32, device=dev)
    t_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    ev, ev_wait = [], []
    whole = [None]
    sync_token = torch.zeros(1, dtype=torch.float32, device=dev)

    def step_sites():
        # drop the previous assembled map first: the caching allocator then
        # reuses its block instead of calling cudaMalloc for a second 179 MB
        # buffer inside the timed region (seen as 40 ms 'gather' outliers)
        whole[0] = None
        t_map.fill_(1.)
        h.call('oqg_hazard_planar', _ptr(inp[0]), _ptr(inp[1]), _ptr(inp[2]),
               None, 0, U, ctypes.byref(st), None, n_k, 0., 50., 0,
               _ptr(t_map), _ptr(t_cnt))
        # the ranks meet first (a 4-byte all-reduce on the stream), so that
        # the events time the exchange and not the wait for the slowest rank
        w0 = torch.cuda.Event(True)
        w0.record()
        if world > 1:
            dist.all_reduce(sync_token)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        whole[0] = oq.sharding.gather_tiles(t_map, tile.sids, N)
        b.record()
        ev.append((a, b))
        ev_wait.append((w0, a))

    for _ in range(warm):
        step_sites()
    ev.clear()
    ev_wait.clear()
    t_cnt.zero_()
    ms = timed_max(step_sites, steps)
    pairs = torch.tensor([float(t_cnt[0].item()) / steps],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pairs)
    evals = float(pairs.item()) * L * G
    gather_ms = allmax(sum(_events_ms(ev)) / steps)
    wait_sites_ms = allmax(sum(_events_ms(ev_wait)) / steps)
    # parity spot check: the first sites of the gathered map against a
    # single-rank pass over exactly those sites
    nchk = min(256, N)
    sub = sc.filter(np.arange(N) < nchk)
    st2, keep2 = cm._sites_struct(sub, local)
    chk = torch.ones((nchk, L, G), dtype=torch.float32, device=dev)
    h.call('oqg_hazard_planar', _ptr(inp[0]), _ptr(inp[1]), _ptr(inp[2]),
           None, 0, U, ctypes.byref(st2), None, nchk, 0., 50., 0, _ptr(chk),
           None)
    diff = float((whole[0][:nchk] - chk).abs().max().item())
    out['strong_sites'] = {
        'value': evals / (ms / steps * 1e-3), 'unit': 'evals/s',
        'ms_per_step': ms / steps, 'gather_ms': gather_ms,
        'wait_for_slowest_rank_ms': wait_sites_ms,
        'gather_bytes': N * L * G * 4, 'steps': steps,
        'workload': 'cfg4 models, %d ruptures x %d sites in total, sites '
                    'split by sids %% %d, tiles all-gathered into the whole '
                    'map on every rank' % (U, N, world),
        'check_max_abs_diff_vs_single_rank': diff}
    del t_map, whole, chk, inp
    torch.cuda.empty_cache()

    # ---- source groups: cfg5 shape, NCCL PRODUCT all-reduce ---------------
    U, N, NG = STRONG_GROUPS
    U = int(U * args.strong_scale)
    names = WORKLOADS['cfg5'][2]
    G = len(names)
    cm = cmaker(names)
    h = cm.handle()
    mine = oq.sharding.groups_of(rank, world, NG)
    inps = [_planar_inputs(oq, cm, U, 20240 + g, local) for g in mine]
    _, sites = synth(4, N, 21240)
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    st, keep = cm._sites_struct(sc, local)
    t_map = torch.ones((N, L, G), dtype=torch.float32, device=dev)
    t_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    ev, ev_wait = [], []

    def step_groups():
        t_map.fill_(1.)
        for a, b, c in inps:
            h.call('oqg_hazard_planar', _ptr(a), _ptr(b), _ptr(c), None, 0, U,
                   ctypes.byref(st), None, N, 0., 50., 0, _ptr(t_map),
                   _ptr(t_cnt))
        w0 = torch.cuda.Event(True)
        w0.record()
        if world > 1:
            dist.all_reduce(sync_token)     # meet first, see step_sites
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        oq.sharding.combine_maps(t_map, rates=False)
        e1.record()
        ev.append((e0, e1))
        ev_wait.append((w0, e0))

    for _ in range(warm):
        step_groups()
    ev.clear()
    ev_wait.clear()
    t_cnt.zero_()
    ms = timed_max(step_groups, steps)
    pairs = torch.tensor([float(t_cnt[0].item()) / steps],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pairs)
    evals = float(pairs.item()) * L * G
    ar_ms = allmax(sum(_events_ms(ev)) / steps)
    wait_groups_ms = allmax(sum(_events_ms(ev_wait)) / steps)
    nbytes = N * L * G * 4
    bus = (2. * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9
           if world > 1 and ar_ms > 0 else None)
    # parity spot check: all the groups folded by THIS rank alone over the
    # first sites, against the all-reduced map
    nchk = min(128, N)
    sub = sc.filter(np.arange(N) < nchk)
    st2, keep2 = cm._sites_struct(sub, local)
    chk = torch.ones((nchk, L, G), dtype=torch.float32, device=dev)
    for g in range(NG):
        a, b, c = (inps[mine.index(g)] if g in mine
                   else _planar_inputs(oq, cm, U, 20240 + g, local))
        h.call('oqg_hazard_planar', _ptr(a), _ptr(b), _ptr(c), None, 0, U,
               ctypes.byref(st2), None, nchk, 0., 50., 0, _ptr(chk), None)
    rel = float(((t_map[:nchk] - chk).abs() / chk.clamp_min(1e-30)).max())
    out['strong_groups'] = {
        'value': evals / (ms / steps * 1e-3), 'unit': 'evals/s',
        'ms_per_step': ms / steps, 'allreduce_ms': ar_ms,
        'wait_for_slowest_rank_ms': wait_groups_ms,
        'allreduce_bytes': nbytes, 'allreduce_bus_gbs': bus,
        'allreduce_bus_reference_gbs': 725.0, 'steps': steps,
        'workload': 'cfg5 shape: %d source groups x %d ruptures x %d sites, '
                    'groups round-robin over %d ranks, NCCL all-reduce '
                    '(PRODUCT) of the %d-byte pne maps inside the timed '
                    'region' % (NG, U, N, world, nbytes),
        'check_max_rel_diff_vs_single_rank': rel}
    return out


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
    # ---- strong-scaling legs (every rank takes part in their collectives)
    strong = {}
    if not args.no_strong and args.workload == 'cfg3' and not args.ruptures \
            and not args.sites:
        del t_pla, t_mag, t_md, t_map, group_inputs
        torch.cuda.empty_cache()
        strong = strong_legs(args, oq, rank, world, local, timed)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (the slowest fused launch)
    peak = ctypes.c_double(0)
    h.call('oqg_measure_fp64_peak', ctypes.byref(peak))
    # second, independent FP64 probe: cuBLAS DGEMM through torch
    a64 = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
    b64 = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
    torch.matmul(a64, b64)
    dg = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        torch.matmul(a64, b64)
        e1.record()
        torch.cuda.synchronize()
        dg.append(e0.elapsed_time(e1))
    dgemm_tflops = 2 * 4096 ** 3 / (min(dg) * 1e-3) / 1e12
    del a64, b64
    gdom = int(np.argmax(kernel_ms)) if kernel_ms else 0
    ngrp = len(my_groups)
    fl = flops_per_row(names[gdom], M, L1) * pairs_kept / ngrp
    ach = fl / (kernel_ms[gdom] * 1e-3) / 1e12 if kernel_ms else None
    alg_bytes = U * 65 * 8 + N * (7 * 8 + 33) + N * L * 8 * 2
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            hbm_peak = json.load(f)['hbm_gbs']
        hbm_src = 'MEASURED_PEAKS.json'
    except Exception:
        hbm_peak, hbm_src = 6650., 'fallback'
    # Per-launch figures that only a profiler can give (DRAM traffic, warp
    # instructions, pipe utilisation) come from the ncu capture of THIS
    # kernel version committed under profiles/ (tools/ncu_metrics.py writes
    # the JSON from the .ncu-rep); they are null when the capture is absent
    # or was taken with another build of the kernels.
    prof, prof_src = None, None
    try:
        from oq_engine_b200 import _build
        with open(os.path.join(ROOT, 'profiles', 'r2_fused_metrics.json')) as f:
            allprof = json.load(f)
        if allprof.get('kernel_fingerprint') == _build.kernel_fingerprint():
            prof = allprof['kernels'].get(names[gdom])
            prof_src = 'profiles/r2_fused_metrics.json (%s)' % allprof['report']
    except Exception:
        pass
    sm_hz = (clocks or {}).get('sm_mhz') or 1965.
    sm_hz *= 1e6
    issue = None
    if prof and kernel_ms:
        # warp instructions scale with the rupture-site pairs of a launch
        inst = prof['warp_inst_per_pair'] * pairs_kept / ngrp
        slots = kernel_ms[gdom] * 1e-3 * sm_hz * 148 * 4
        issue = {'warp_inst_per_pair': prof['warp_inst_per_pair'],
                 'warp_inst_per_launch': inst,
                 'issue_slots_per_launch': slots, 'frac': inst / slots,
                 'definition': 'warp instructions of the launch (ncu '
                               'smsp__inst_executed.sum per pair x pairs) / '
                               '(launch time x 148 SMs x 4 schedulers x SM '
                               'clock sampled in this run)',
                 'ncu_issue_active_pct': prof.get('issue_active_pct'),
                 'ncu_lsu_data_pipe_pct': prof.get('lsu_data_pipe_pct'),
                 'ncu_fp64_pipe_pct': prof.get('fp64_pipe_pct'),
                 'source': prof_src}
    # executed evaluations: one extra, untimed statistics pass
    executed = None
    if not by_groups:
        h = cm.handle()     # the strong legs configured other models
        L_.oqg_set_count_evals(h._h, 1)
        t_cnt.zero_()
        t_map2 = torch.ones((N, L, G), dtype=torch.float32, device=dev)
        a, b, c = _planar_inputs(oq, cm, U, 20240, local)
        h.call('oqg_hazard_planar', _ptr(a), _ptr(b), _ptr(c), None, 0, U,
               ctypes.byref(st), None, N, 0., 50., 0, _ptr(t_map2),
               _ptr(t_cnt))
        L_.oqg_set_count_evals(h._h, 0)
        ex = int(t_cnt[1].item())
        nominal = pairs_kept * L * G
        executed = {
            'evals_per_step': ex, 'nominal_evals_per_step': nominal,
            'level_skip_fraction': 1. - ex / nominal,
            'executed_evals_per_s': ex / (ms / args.steps * 1e-3),
            'note': '`value` counts every (pair, gsim, IMT, level) as the '
                    'metric defines it; the kernel skips the levels above '
                    'mean + truncation_level * sigma of a whole warp, whose '
                    'PoEs are exactly 0, and executes this many'}
        del t_map2
    roofline = {
        'bound': 'fp64', 'kernel': 'k_fused<%s>' % names[gdom],
        'achieved': ach, 'peak': peak.value, 'unit': 'TFLOP/s',
        'frac': ach / peak.value if ach else None,
        'traffic': prof['dram_bytes'] if prof else None,
        'traffic_source': prof_src,
        'convention': 'achieved = SURVEY 8d ALGORITHMIC FP64 flop of the '
                      'reference operation mix (erf = 50 flop, GMPE 220-560, '
                      'distances 40-1300 per row) / launch time.  The kernel '
                      'does that work with far fewer operations (float32 '
                      'cubic table for the PoE, table-driven log / exp, level '
                      'skipping), so frac > 1: it is speed relative to "the '
                      'reference arithmetic at the FP64 peak", NOT a hardware '
                      'utilisation.  The hardware fractions are `issue` '
                      '(instruction-issue roofline) and the ncu pipe '
                      'percentages inside it; `executed` gives the '
                      'evaluations really performed.',
        'peak_source': 'DFMA microbenchmark (oqg_measure_fp64_peak) in this '
                       'process; cross-checked by `peak_probes`',
        'peak_probes': {'dfma_tflops': peak.value,
                        'cublas_dgemm_4096_tflops': dgemm_tflops,
                        'nominal_tflops': FP64_NOMINAL_TFLOPS},
        'flop_per_row': flops_per_row(names[gdom], M, L1),
        'kernel_ms': kernel_ms,
        'issue': issue, 'executed': executed,
        'hbm': {'achieved': alg_bytes / (kernel_ms[gdom] * 1e-3) / 1e9
                if kernel_ms else None, 'peak': hbm_peak, 'unit': 'GB/s',
                'peak_source': hbm_src, 'algorithmic_bytes': alg_bytes,
                'note': 'algorithmic bytes per launch; not the bound'},
    }
    # ---- CPU baseline on a bounded sample, rank 0 only
    cpu = None
    if not args.no_cpu and world == 1:      # rank 0 at N = 1 only
        if have_reference() and not args.port:
            v, cores, dt, ev = reference_baseline(
                args.workload, args.cpu_ruptures, N, 1, 1)
            kind = 'reference'
        else:
            v, cores, dt, ev = cpu_baseline(args.workload, args.cpu_ruptures,
                                            N, 1, 1)
            kind = 'port'
        cpu = {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': kind,
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
        **strong,
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
    ap.add_argument('--port', action='store_true',
                    help='CPU arm: the numpy port of oracle/ even when the '
                         'reference is installed under baseline/_ref')
    ap.add_argument('--no-strong', action='store_true',
                    help='skip the strong-scaling legs')
    ap.add_argument('--strong-steps', type=int, default=2)
    ap.add_argument('--strong-scale', type=float, default=1.0,
                    help='scales the ruptures of the strong-scaling legs')
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
