"""Multi-GPU plumbing (one process per GPU, torch.distributed).

* Fingerprinting / augmentation: queries are independent -> contiguous query slices per
  rank, no data-path collective (`query_slice`).
* Matching: the reference index is sharded by hash range (`hash_range`); per-track raw
  counts are summed with one reduce-scatter that also assigns each query an owner rank, the
  candidates' (track, delta-t) hit lists go to the owners with one all-to-all, results are
  all-gathered (`match_sharded`).
The functions take the process group's backend as it comes: NCCL on GPUs, gloo in the CPU
tests (which exercise the partitioning/collective logic with numpy stand-ins for the kernels).
"""
from __future__ import annotations


def query_slice(n_queries: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the batch owned by `rank`."""
    per = -(-n_queries // world)
    lo = min(n_queries, rank * per)
    return lo, min(n_queries, lo + per)


def hash_range(rank: int, world: int, hashbits: int = 20):
    """Bucket range [lo, hi) of the index shard held by `rank`."""
    nb = 1 << hashbits
    per = -(-nb // world)
    lo = min(nb, rank * per)
    return lo, min(nb, lo + per)


def _reduce_scatter_rows(x, world, group):
    """Sum x [world * own, ...] over the ranks; rank r keeps rows [r * own, (r + 1) * own)."""
    import torch
    import torch.distributed as dist

    own = x.shape[0] // world
    out = torch.empty(own, *x.shape[1:], dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM, group=group)
    return out


def _all_gather_rows(x, world, group):
    import torch
    import torch.distributed as dist

    out = torch.empty(world * x.shape[0], *x.shape[1:], dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _to_owners(x, world, group):
    """x [world * own, ...] holds this shard's rows for every query; returns [world (shard), own, ...]:
    for the queries this rank owns, the rows every shard produced."""
    import torch
    import torch.distributed as dist

    out = torch.empty_like(x)
    dist.all_to_all_single(out, x.contiguous(), group=group)
    return out.view(world, x.shape[0] // world, *x.shape[1:])


def match_sharded(ctx, hashes, nh, params=None, max_rows: int = 16, list_cap: int = 2048, sub_batch: int = 256,
                  group=None):
    """match_hashes for a batch of queries that every rank holds, against an index sharded by
    hash range (each rank's `ctx` holds its shard).  Returns (results [B,max_rows,7], nrows [B])
    identical on every rank.

    Per sub-batch: every rank counts its shard's hits for all queries; one **reduce-scatter** sums the
    per-track histograms and leaves each rank the totals of the queries it owns (a contiguous 1/world
    slice); the owner selects candidates (Matcher._best_count_ids) and all-gathers them (small); every
    rank collects the candidates' (track, delta-t) hits from its shard and an all-to-all sends them to the
    owners, which align (Matcher._approx_match_counts) and all-gather the result rows."""
    import torch
    import torch.distributed as dist

    from . import lib

    params = params or lib.match_defaults()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = hashes.shape[0]
    res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=hashes.device)
    nrows = torch.empty(B, dtype=torch.int32, device=hashes.device)
    if world > 1:
        sub_batch = max(world, sub_batch // world * world)
    for q0 in range(0, B, sub_batch):
        hq, nq = hashes[q0:q0 + sub_batch], nh[q0:q0 + sub_batch]
        n_real = hq.shape[0]
        if world > 1 and n_real % world:  # pad the last sub-batch with empty queries so it splits evenly
            pad = world - n_real % world
            hq = torch.cat([hq, torch.zeros(pad, *hq.shape[1:], dtype=hq.dtype, device=hq.device)])
            nq = torch.cat([nq, torch.zeros(pad, dtype=nq.dtype, device=nq.device)])
        counts = ctx.match_counts(hq, nq)
        if world == 1:
            cand, ncand = ctx.match_select(counts, params)
            lst, nlist = ctx.match_collect(hq, nq, cand, ncand, params, list_cap)
            r, n = ctx.match_align(lst[None], nlist[None], cand, ncand, params, max_rows)
        else:
            mine = _reduce_scatter_rows(counts, world, group)                 # per-track match histograms
            cand_m, ncand_m = ctx.match_select(mine, params)
            cand, ncand = _all_gather_rows(cand_m, world, group), _all_gather_rows(ncand_m, world, group)
            lst, nlist = ctx.match_collect(hq, nq, cand, ncand, params, list_cap)
            lists, nlists = _to_owners(lst, world, group), _to_owners(nlist, world, group)
            r_m, n_m = ctx.match_align(lists.contiguous(), nlists.contiguous(), cand_m, ncand_m, params, max_rows)
            r, n = _all_gather_rows(r_m, world, group), _all_gather_rows(n_m, world, group)
        res[q0:q0 + n_real], nrows[q0:q0 + n_real] = r[:n_real], n[:n_real]
    return res, nrows


def match_replicated(ctx, hashes, nh, params=None, max_rows: int = 16, group=None):
    """match_hashes for a batch every rank holds, against an index REPLICATED on every rank (each `ctx` holds
    the whole table: the reference's 2^20 x 100 table is 419 MB).  Queries are the independent units: rank r
    matches its `query_slice` with the one-kernel matcher and the result rows are all-gathered - no collective
    on the data path.  This is the throughput mode; `match_sharded` (hash-range shards, histograms summed over
    NCCL) is the capacity mode for indexes that do not fit one GPU.  Returns (results [B,max_rows,7],
    nrows [B]) identical on every rank."""
    import torch
    import torch.distributed as dist

    from . import lib

    params = params or lib.match_defaults()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = hashes.shape[0]
    if world == 1:
        return ctx.match(hashes, nh, params, max_rows)
    per = -(-B // world)
    lo, hi = query_slice(B, rank, world)
    res = torch.zeros(per, max_rows, 7, dtype=torch.int32, device=hashes.device)
    nrows = torch.zeros(per, dtype=torch.int32, device=hashes.device)
    if hi > lo:
        r, n = ctx.match(hashes[lo:hi], nh[lo:hi], params, max_rows)
        res[: hi - lo], nrows[: hi - lo] = r, n
    res_all, nrows_all = _all_gather_rows(res, world, group), _all_gather_rows(nrows, world, group)
    return res_all[:B], nrows_all[:B]

