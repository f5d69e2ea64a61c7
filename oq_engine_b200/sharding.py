"""Multi-GPU decomposition of the hot path: one process per GPU.

The reference parallelises the classical calculator as a map-reduce over
(source block x site tile) tasks whose results are summed in the master
(openquake/calculators/classical.py:386-432, 591-650; site tiles by
``sids % ntiles``, openquake/hazardlib/site.py:133-161).  Here

* **site sharding** gives every rank a site tile and all the ruptures: map
  rows are independent, so there is NO data-path collective; the tiles are
  only gathered at the end if the caller wants the whole map;
* **source-group sharding** gives every rank whole source groups and a map
  over all the sites; the maps are combined once, pne maps multiplicatively
  (``all_reduce(PRODUCT)``) and rate maps additively (``all_reduce(SUM)``),
  exactly the reductions the reference master performs
  (classical.py:431, hazard_curve.py:156).  Over NCCL this is one
  NVLink/NVSwitch all-reduce of N x L x G float32.

torch.distributed provides the process group (nccl on GPUs, gloo in the CPU
tests); nothing here launches kernels by itself.
"""
import numpy


def site_tile(sitecol, rank, world, mode='modulo'):
    """
    The site tile of `rank`: ``sids % world == rank`` like the reference's
    TileGetter, or a contiguous block (mode='block')
    """
    if world == 1:
        return sitecol
    if mode == 'modulo':
        return sitecol.filter(sitecol.sids % world == rank)
    n = len(sitecol)
    lo, hi = rank * n // world, (rank + 1) * n // world
    mask = numpy.zeros(n, bool)
    mask[lo:hi] = True
    return sitecol.filter(mask)


def groups_of(rank, world, n_groups):
    """Round-robin assignment of source groups to ranks"""
    return [g for g in range(n_groups) if g % world == rank]


def combine_maps(array, rates, group=None):
    """
    Combine per-rank maps over ALL the sites in place: product of the
    probabilities of no exceedance, or sum of the rates.

    :param array: torch tensor (N, L, G) float32 (CUDA with nccl, CPU with
        gloo)
    """
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return array
    op = dist.ReduceOp.SUM if rates else dist.ReduceOp.PRODUCT
    dist.all_reduce(array, op=op, group=group)
    return array


def gather_tiles(tile_array, tile_sids, sids, group=None, partial=False):
    """
    Assemble the whole (N, L, G) map from the site tiles of the ranks;
    returns a tensor on every rank, rows in the order of `sids`.

    :param tile_array: (n_k, L, G) tensor of this rank
    :param tile_sids: the site IDs of its rows
    :param sids: the site IDs of the complete site collection (any unique
        integers: a filtered collection has gaps), or their number N when
        they are 0 .. N-1

    The tiles may have any sizes: their lengths are all-gathered first and
    the tiles padded to the longest.  Every site must be owned by exactly one
    rank; a site nobody owns is an error, not an uninitialised row (unless
    `partial`: then such rows are zero, for callers that assemble the tiles
    of several ranks themselves).
    """
    import torch
    import torch.distributed as dist
    dev = tile_array.device
    if numpy.isscalar(sids):
        sids = numpy.arange(int(sids), dtype=numpy.int64)
    sids = numpy.asarray(sids, numpy.int64)
    N = len(sids)
    tile_sids = numpy.asarray(tile_sids, numpy.int64)
    assert len(tile_sids) == len(tile_array), (len(tile_sids), len(tile_array))
    # site ID -> row of the assembled map
    order = numpy.argsort(sids, kind='stable')
    pos = numpy.searchsorted(sids[order], tile_sids)
    if len(tile_sids) and (pos.max(initial=0) >= N or
                           (sids[order][pos] != tile_sids).any()):
        raise ValueError('gather_tiles: a tile holds site IDs that are not '
                         'in the site collection')
    rows = torch.as_tensor(order[pos], dtype=torch.int64, device=dev)
    shape = tuple(tile_array.shape[1:])
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi:
        allrows, tiles = rows, tile_array
    else:
        # one all_gather of the tile lengths, then one of the padded rows and
        # one of the padded tiles into contiguous buffers; a padding row
        # points at the extra row N, dropped at the end
        world = dist.get_world_size(group)
        n = torch.tensor([len(tile_array)], dtype=torch.int64, device=dev)
        lens = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(lens, n, group=group)
        nmax = int(lens.max().item())
        pad = torch.zeros((nmax,) + shape, dtype=tile_array.dtype, device=dev)
        pad[:len(tile_array)] = tile_array
        prow = torch.full((nmax,), N, dtype=torch.int64, device=dev)
        prow[:len(rows)] = rows
        tiles = torch.empty((world * nmax,) + shape, dtype=tile_array.dtype,
                            device=dev)
        allrows = torch.empty(world * nmax, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(tiles, pad, group=group)
        dist.all_gather_into_tensor(allrows, prow, group=group)
    out = torch.zeros((N + 1,) + shape, dtype=tile_array.dtype, device=dev)
    out[allrows] = tiles    # (index_copy_ took 60 ms for 179 MB on B200)
    if not partial:
        seen = torch.zeros(N + 1, dtype=torch.int32, device=dev)
        seen.index_add_(0, allrows, torch.ones_like(allrows,
                                                   dtype=torch.int32))
        seen = seen[:N]
        if not bool((seen == 1).all()):
            raise ValueError('gather_tiles: %d sites have no owner and %d '
                             'have several' % (int((seen == 0).sum()),
                                               int((seen > 1).sum())))
    return out[:N]


def hazard_site_sharded(cmaker, src, sitecol, rank, world, rates=False,
                        mode='modulo'):
    """
    Fused pass of all the ruptures of `src` over this rank's site tile;
    returns (MapArray of the tile on the device, tile)
    """
    from .map_array import MapArray
    tile = site_tile(sitecol, rank, world, mode)
    pmap = MapArray(tile.sids, cmaker.L, cmaker.G, rates=rates,
                    device='cuda:%d' % cmaker.handle().device)
    pmap.fill(0 if rates else 1)
    cmaker.hazard_planar(pmap, src, tile)
    return pmap, tile


def hazard_group_sharded(cmakers, srcs, sitecol, rank, world, rates=False,
                         group=None):
    """
    Source-group sharding: `cmakers[g]`, `srcs[g]` describe group g; this
    rank folds its groups into a map over all the sites, then the maps are
    combined with one all-reduce (PRODUCT of pnes / SUM of rates).
    All the groups must share L and G (same `g` columns).
    """
    from .map_array import MapArray
    cm0 = cmakers[0]
    pmap = MapArray(sitecol.sids, cm0.L, cm0.G, rates=rates,
                    device='cuda:%d' % cm0.handle().device)
    pmap.fill(0 if rates else 1)
    for g in groups_of(rank, world, len(srcs)):
        cmakers[g].hazard_planar(pmap, srcs[g], sitecol)
    combine_maps(pmap.array, rates, group)
    return pmap


def gid_offsets(cmakers):
    """Column offsets of groups with their own gsims in the global map:
    group g owns the columns off[g] .. off[g + 1] (the `gid` of the
    reference's cmakers, contexts.py:1988-1991)"""
    return numpy.concatenate([[0], numpy.cumsum([cm.G for cm in cmakers])])


def hazard_group_sharded_gid(cmakers, srcs, sitecol, rank, world, rates=True,
                             group=None):
    """
    Source-group sharding when the groups have DIFFERENT gsim sets: group g
    owns its own columns of the (N, L, Gt) map (Gt = sum of the G of the
    groups), so combining the ranks is a concatenation on the G axis.  Every
    rank fills the columns of its groups in a map initialised with the
    identity of the reduction (0 rates / 1 pne) and the same all-reduce as in
    `hazard_group_sharded` assembles the whole map on every rank: a column is
    written by exactly one rank, so SUM / PRODUCT with identities IS the
    concatenation.  All the groups must share the levels (L).
    """
    import torch
    from .map_array import MapArray
    off = gid_offsets(cmakers)
    L = cmakers[0].L
    assert all(cm.L == L for cm in cmakers), 'the groups must share L'
    dev = 'cuda:%d' % cmakers[0].handle().device
    big = MapArray(sitecol.sids, L, int(off[-1]), rates=rates, device=dev)
    big.fill(0 if rates else 1)
    for g in groups_of(rank, world, len(srcs)):
        cm = cmakers[g]
        pm = MapArray(sitecol.sids, L, cm.G, rates=rates, device=dev)
        pm.fill(0 if rates else 1)
        cm.hazard_planar(pm, srcs[g], sitecol)
        big.array[:, :, off[g]:off[g + 1]] = pm.array
    combine_maps(big.array, rates, group)
    big.gid = numpy.arange(off[-1], dtype=numpy.uint16)
    return big
