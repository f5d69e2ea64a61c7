"""world_size-2 tests of the multi-GPU decomposition on CPU (gloo): the
host logic of site sharding (tiles + gather) and of source-group sharding
(prod / sum all-reduce of the maps), with the oracle standing in for the
per-rank kernels."""
import os
import sys
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, synth_ruptures, synth_sites, IMTS

PAIRS = [('BooreEtAl2014', {}), ('ChiouYoungs2014', {})]
IMLS = np.logspace(-2, 0.3, 5)
N, U = 24, 12


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def oracle_map(r, sites, rates):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import geo_np
    import hazard_np
    pla = geo_np.build_planar(
        r['mag'], r['strike'], r['dip'], r['rake'], r['rate'], r['area'],
        r['hypo_lon'], r['hypo_lat'], r['hypo_dep'], 0., 25., 1.5)
    ctx = hazard_np.get_ctx(PAIRS, pla, r['mag'], sites, None, 0)
    imtls = {imt: IMLS for imt in IMTS[:2]}
    L = 2 * len(IMLS)
    n = len(sites['sids'])
    arr = (np.zeros if rates else np.ones)((n, L, 2), np.float32)
    sid_to_row = np.zeros(int(sites['sids'].max()) + 1, np.uint32)
    sid_to_row[sites['sids']] = np.arange(n)
    hazard_np.update(arr, sid_to_row, PAIRS, ctx, imtls, 3., 50.,
                     rates_mode=rates)
    return arr


def groups():
    out = []
    for g in range(4):
        rng = np.random.default_rng(100 + g)
        out.append(synth_ruptures(U, rng))
    return out, synth_sites(N, np.random.default_rng(7))


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oq_engine_b200 as oq
    grps, sites = groups()
    sc = oq.SiteCollection.from_points(sites['lon'], sites['lat'],
                                       sites['depth'], sites)
    res = {}
    # ---- source-group sharding: prod of pnes / sum of rates
    for rates in (False, True):
        arr = (np.zeros if rates else np.ones)((N, 2 * len(IMLS), 2),
                                               np.float32)
        for g in oq.sharding.groups_of(rank, world, len(grps)):
            m = oracle_map(grps[g], sites, rates)
            arr = arr + m if rates else arr * m
        t = torch.from_numpy(arr)
        oq.sharding.combine_maps(t, rates)
        res['groups_rates' if rates else 'groups_pne'] = t.numpy()
    # ---- site sharding: own tile, no collective on the data path
    tile = oq.sharding.site_tile(sc, rank, world)
    sub = {k: v[tile.sids] for k, v in sites.items()}
    m = oracle_map(grps[0], sub, False)
    full = oq.sharding.gather_tiles(torch.from_numpy(m), tile.sids, N)
    res['sites'] = full.numpy()
    # ---- a filtered site collection: site IDs with gaps, ragged tiles (rank
    # 0 owns 3 sites, rank 1 owns 8) -> rows in the order of the given sids
    fsids = np.array([3, 40, 7, 9, 22, 5, 11, 13, 17, 19, 2], np.int64)
    mine = fsids[:3] if rank == 0 else fsids[3:]
    tile_arr = torch.from_numpy(
        (mine[:, None, None] * np.ones((1, 2, 2))).astype(np.float32))
    res['ragged'] = oq.sharding.gather_tiles(tile_arr, mine, fsids).numpy()
    # a site nobody owns is an error on every rank, not an empty row
    try:
        oq.sharding.gather_tiles(tile_arr[:2], mine[:2], fsids)
        res['unowned'] = 'no error'
    except ValueError as exc:
        res['unowned'] = str(exc)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world_size_2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    grps, sites = groups()
    exp_pne = np.ones((N, 2 * len(IMLS), 2), np.float32)
    exp_rat = np.zeros((N, 2 * len(IMLS), 2), np.float32)
    for r in grps:
        exp_pne = exp_pne * oracle_map(r, sites, False)
        exp_rat = exp_rat + oracle_map(r, sites, True)
    single = oracle_map(grps[0], sites, False)
    for rank in range(world):
        res = results[rank]
        np.testing.assert_allclose(res['groups_pne'], exp_pne, rtol=1e-6)
        np.testing.assert_allclose(res['groups_rates'], exp_rat, rtol=1e-6,
                                   atol=1e-12)
        np.testing.assert_array_equal(res['sites'], single)
        fsids = np.array([3, 40, 7, 9, 22, 5, 11, 13, 17, 19, 2])
        np.testing.assert_array_equal(
            res['ragged'], fsids[:, None, None] * np.ones((1, 2, 2)))
        assert 'no owner' in res['unowned'], res['unowned']
    # pne product and rate sum describe the same hazard
    np.testing.assert_allclose(np.exp(-exp_rat), exp_pne, rtol=2e-5)
