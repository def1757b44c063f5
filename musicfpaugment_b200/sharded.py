"""Multi-GPU plumbing (one process per GPU, torch.distributed).

* Fingerprinting / augmentation: queries are independent -> contiguous query slices per
  rank, no data-path collective (`query_slice`).
* Matching: the reference index is sharded by hash range (`hash_range`); per-track raw
  counts are summed with one all-reduce, the candidates' (track, delta-t) hit lists are
  exchanged with one all-gather, everything else is local (`match_sharded`).
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


def match_sharded(ctx, hashes, nh, params=None, max_rows: int = 16, list_cap: int = 2048, sub_batch: int = 256,
                  group=None):
    """match_hashes for a batch of queries that every rank holds, against an index sharded by
    hash range (each rank's `ctx` holds its shard).  Returns (results [B,max_rows,7], nrows [B])
    identical on every rank."""
    import torch
    import torch.distributed as dist

    from . import lib

    params = params or lib.match_defaults()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B = hashes.shape[0]
    res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=hashes.device)
    nrows = torch.empty(B, dtype=torch.int32, device=hashes.device)
    for q0 in range(0, B, sub_batch):
        hq, nq = hashes[q0:q0 + sub_batch], nh[q0:q0 + sub_batch]
        counts = ctx.match_counts(hq, nq)
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)      # per-track match histograms
        cand, ncand = ctx.match_select(counts, params)
        lst, nlist = ctx.match_collect(hq, nq, cand, ncand, params, list_cap)
        if world > 1:
            # outputs are the inputs concatenated along dim 0 (the only shape every backend accepts)
            lists = torch.empty(world * lst.shape[0], lst.shape[1], dtype=lst.dtype, device=lst.device)
            nlists = torch.empty(world * nlist.shape[0], dtype=nlist.dtype, device=nlist.device)
            dist.all_gather_into_tensor(lists, lst, group=group)            # candidates' (track, delta-t) hits
            dist.all_gather_into_tensor(nlists, nlist, group=group)
            lists, nlists = lists.view(world, *lst.shape), nlists.view(world, *nlist.shape)
        else:
            lists, nlists = lst[None], nlist[None]
        r, n = ctx.match_align(lists, nlists, cand, ncand, params, max_rows)
        res[q0:q0 + sub_batch], nrows[q0:q0 + sub_batch] = r, n
    return res, nrows
