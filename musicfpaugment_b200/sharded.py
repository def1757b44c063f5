"""Multi-GPU plumbing (one process per GPU, torch.distributed).

* Fingerprinting / augmentation: queries are independent -> contiguous query slices per
  rank, no data-path collective (`query_slice`).
* Matching: the reference index is sharded by hash range (`hash_range`); every rank lists the
  hits of its buckets as packed (track, delta-t) words, ONE all-to-all sends them to the query's
  owner rank, which runs the whole matcher on them; result rows are all-gathered (`match_sharded`).
  `match_sharded_dense` is the round-1 exchange (reduce-scatter of dense per-track histograms).
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


def default_words_cap(cap_hashes: int, depth: int, world: int, bound: bool = False) -> int:
    """Capacity of one query's hit-word list on one shard, rounded up to 256 words.

    bound=True: cap_hashes * depth, what a query whose hashes ALL fall into one shard produces - it cannot overflow.
    The peer exchange uses it: capacity there is only address space in the receive buffer, the words that exist
    are all that crosses the link.  bound=False (the all-to-all, which moves the capacity): the expected share
    cap_hashes / world of the query's buckets plus six standard deviations of that binomial count, every bucket
    taken as full - an overflow (flagged per query, never silent) is then a < 1e-8 event per query and shard."""
    if bound or world == 1:
        rows = cap_hashes
    else:
        mean = cap_hashes / world
        rows = min(cap_hashes, mean + 6.0 * (mean * (1.0 - 1.0 / world)) ** 0.5 + 1.0)
    return (int(rows * depth) + 255) // 256 * 256


class PeerExchange:
    """Receive buffers of the fused sparse exchange, one set per (context, world, own, words_cap): two
    [world, own, words_cap] word buffers (consecutive sub-batches alternate, so a rank may sweep sub-batch i + 1
    into its peers while they still read sub-batch i), their [world, own] counts and a row of barrier flags, all in
    ONE allocation made by `mfpa_peer_alloc`; the other ranks' allocations are mapped with `mfpa_peer_open` from the
    IPC handles all-gathered here.  `turn` counts the sub-batches of every call on this exchange (the same on every
    rank): buffer = turn % 2, barrier epoch = turn + 1."""

    N_BUF = 2

    def __init__(self, ctx, world: int, rank: int, own: int, words_cap: int, group=None):
        import torch.distributed as dist

        from . import lib

        if world > lib.MAX_PEERS:
            raise lib.MfpaError(f"peer exchange: {world} ranks, the peer set holds {lib.MAX_PEERS}")
        self.ctx, self.world, self.rank, self.own, self.words_cap = ctx, world, rank, own, words_cap
        rows = world * own
        self.flag_bytes = 256
        self.n_bytes = (rows * 4 + 255) // 256 * 256
        self.w_bytes = (rows * words_cap * 4 + 255) // 256 * 256
        total = self.flag_bytes + self.N_BUF * (self.n_bytes + self.w_bytes)
        self.base, handle = ctx.peer_alloc(total)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.bases = [self.base if r == rank else ctx.peer_open(handles[r]) for r in range(world)]
        self.sets = []
        for b in range(self.N_BUF):
            ps = lib.PeerSet()
            ps.world, ps.rank = world, rank
            for r, base in enumerate(self.bases):
                ps.flags[r] = base
                ps.nwords[r] = base + self.flag_bytes + b * (self.n_bytes + self.w_bytes)
                ps.words[r] = ps.nwords[r] + self.n_bytes
            self.sets.append(ps)
        self.turn = 0
        dist.barrier(group=group)     # every rank has mapped every buffer before anyone stores into one

    def buffer(self, b: int):
        """(words address, counts address) of this rank's receive buffer b."""
        n = self.base + self.flag_bytes + b * (self.n_bytes + self.w_bytes)
        return n + self.n_bytes, n

    def close(self, group=None):
        """Collective: unmap the peers' buffers, wait until every rank has done so, then free this rank's."""
        import torch
        import torch.distributed as dist

        if not self.bases:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for r, base in enumerate(self.bases):
            if r != self.rank:
                self.ctx.peer_close(base)
        dist.barrier(group=group)
        self.ctx.peer_free(self.base)
        self.bases = []


def release_peer_exchanges(ctx, group=None):
    """Collective: free every cached PeerExchange of `ctx` (call it on all ranks, before the contexts close)."""
    cache = ctx.__dict__.pop("_peer_exchanges", {})
    for key in sorted(cache, key=lambda k: k[:4]):
        cache[key].close(group)


def _peer_exchange(ctx, world, rank, own, words_cap, group):
    """The context's cached PeerExchange for this list capacity (set up once: IPC mapping costs milliseconds).  One
    that was built for at least `own` queries per rank serves smaller sub-batches too (a sub-batch's lists fill a
    prefix of the buffer); a larger request replaces it - collectively, every rank takes the same path."""
    cache = ctx.__dict__.setdefault("_peer_exchanges", {})
    key = (world, rank, 0, words_cap, id(group))
    px = cache.get(key)
    if px is not None and px.own < own:
        px.close(group)
        px = None
    if px is None:
        px = cache[key] = PeerExchange(ctx, world, rank, own, words_cap, group)
    return px


def peer_supported(ctx, group=None) -> bool:
    """Collective, cached on the context: can every rank map every other rank's memory (CUDA IPC + peer access; all
    ranks on one node)?  Probed once with a 4 KiB buffer.  When some rank cannot, `match_sharded` keeps the NCCL
    all-to-all - another transport for the same words, not another implementation."""
    import sys

    import torch
    import torch.distributed as dist

    if "_peer_ok" in ctx.__dict__:
        return ctx._peer_ok
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ok, base, opened = 1, None, []
    try:
        base, handle = ctx.peer_alloc(4096)
    except Exception as e:  # noqa: BLE001 - any failure means "not here"
        ok, handle = 0, b""
        print(f"rank {rank}: peer memory unavailable ({e})", file=sys.stderr)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    if ok and all(len(h) == 64 for h in handles):
        try:
            opened = [ctx.peer_open(handles[r]) for r in range(world) if r != rank]
        except Exception as e:  # noqa: BLE001
            ok = 0
            print(f"rank {rank}: cannot map a peer's memory ({e}); sharded matching uses the NCCL all-to-all", file=sys.stderr)
    else:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", ctx.device))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    for a in opened:
        ctx.peer_close(a)
    dist.barrier(group=group)
    if base is not None:
        ctx.peer_free(base)
    ctx._peer_ok = bool(flag.item())
    return ctx._peer_ok


PEER_BUFFER_BYTES = 4 << 30   # both receive buffers of a PeerExchange together (the sub-batch is sized to fit)


def match_sharded(ctx, hashes, nh, params=None, max_rows: int = 16, words_cap: int | None = None,
                  sub_batch: int | None = None, group=None, list_cap=None, exchange: str | None = None):
    """match_hashes for a batch of queries that every rank holds, against an index sharded by
    hash range (each rank's `ctx` holds its shard).  Returns (results [B,max_rows,7], nrows [B])
    identical on every rank.

    Sparse exchange: per sub-batch every rank sweeps ITS buckets once and lists the hits of
    every query as packed (track, delta-t) words (`match_emit`: a few thousand words per query and shard,
    where the dense per-track histogram is 200 KB); ONE all-to-all keyed by the query's owner rank (a
    contiguous 1/world slice of the sub-batch) moves them; the owner runs the whole of match_hashes -
    shared-memory histogram, candidate selection, delta-t alignment - on the words of all shards
    (`match_owner`).  The all-to-all of sub-batch i is in flight while the ranks sweep sub-batch i + 1; the
    owners' result rows are all-gathered once at the end.  `list_cap` is accepted for backward compatibility.

    exchange = "peer" fuses the exchange into the sweep: the emit kernel stores each query's words straight into
    its owner rank's memory over NVLink (`match_emit_peer`, buffers of a cached `PeerExchange`), a one-block barrier
    kernel in peer memory separates it from the owner step - no all-to-all, no staging list, only the words that
    exist cross the link.  It is the default on GPUs of one node (contexts that have `peer_alloc`); "nccl" keeps
    the all-to-all (other transports, the gloo tests)."""
    import torch
    import torch.distributed as dist

    from . import lib

    params = params or lib.match_defaults()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, cap, _ = hashes.shape
    if world == 1:
        return ctx.match(hashes, nh, params, max_rows)
    dev = hashes.device
    if exchange is None:
        exchange = "peer" if hasattr(ctx, "peer_alloc") and dev.type == "cuda" and peer_supported(ctx, group) else "nccl"
    if words_cap is None:
        words_cap = default_words_cap(cap, getattr(ctx, "depth", 100), world, bound=exchange == "peer")
    if sub_batch is None:
        # peer: nothing overlaps between sub-batches, so take the batch whole if the receive buffers allow;
        # all-to-all: sub-batches of 2048 pipeline the collective behind the next sweep
        sub_batch = max(world, PEER_BUFFER_BYTES // (PeerExchange.N_BUF * words_cap * 4)) if exchange == "peer" else 2048
    sub = max(world, min(sub_batch, -(-B // world) * world) // world * world)
    own = sub // world
    n_sub = -(-B // sub)
    res_m = torch.zeros(n_sub * own, max_rows, 7, dtype=torch.int32, device=dev)
    nrows_m = torch.zeros(n_sub * own, dtype=torch.int32, device=dev)
    if exchange == "peer":
        rank = dist.get_rank(group)
        px = _peer_exchange(ctx, world, rank, own, words_cap, group)
        for i in range(n_sub):
            hq, nq = _sub_batch(hashes, nh, i, sub)
            b = px.turn % px.N_BUF
            ctx.match_emit_peer(hq, nq, px.sets[b], words_cap)
            px.turn += 1
            ctx.peer_barrier(px.sets[b], px.turn)
            w_addr, n_addr = px.buffer(b)
            r, n = ctx.match_owner_at(w_addr, n_addr, world, own, words_cap, params, max_rows)
            res_m[i * own:(i + 1) * own], nrows_m[i * own:(i + 1) * own] = r, n
        return _gather_owned(res_m, nrows_m, world, n_sub, own, sub, B, max_rows, group)

    def finish(job):
        i, works, recv_w, recv_n = job
        for w in works:
            w.wait()
        r, n = ctx.match_owner(recv_w.view(world, own, words_cap), recv_n.view(world, own), params, max_rows)
        res_m[i * own:(i + 1) * own], nrows_m[i * own:(i + 1) * own] = r, n

    pending = None
    for i in range(n_sub):
        hq, nq = _sub_batch(hashes, nh, i, sub)
        words, nwords = ctx.match_emit(hq, nq, words_cap)
        recv_w, recv_n = torch.empty_like(words), torch.empty_like(nwords)
        works = [dist.all_to_all_single(recv_w, words, group=group, async_op=True),
                 dist.all_to_all_single(recv_n, nwords, group=group, async_op=True)]
        if pending is not None:
            finish(pending)       # owner step of sub-batch i - 1, behind the sweep of sub-batch i
        pending = (i, works, recv_w, recv_n)
    finish(pending)
    return _gather_owned(res_m, nrows_m, world, n_sub, own, sub, B, max_rows, group)


def _sub_batch(hashes, nh, i, sub):
    """Sub-batch i, padded with empty queries to `sub` rows so that it splits evenly over the ranks."""
    import torch

    hq, nq = hashes[i * sub:(i + 1) * sub], nh[i * sub:(i + 1) * sub]
    if hq.shape[0] < sub:
        pad = sub - hq.shape[0]
        hq = torch.cat([hq, torch.zeros(pad, *hq.shape[1:], dtype=hq.dtype, device=hq.device)])
        nq = torch.cat([nq, torch.zeros(pad, dtype=nq.dtype, device=nq.device)])
    return hq.contiguous(), nq.contiguous()


def _gather_owned(res_m, nrows_m, world, n_sub, own, sub, B, max_rows, group):
    """All-gather the owners' rows: gathered [owner][sub-batch][own] -> query order [sub-batch][owner][own]."""
    res_all, nrows_all = _all_gather_rows(res_m, world, group), _all_gather_rows(nrows_m, world, group)
    res = res_all.view(world, n_sub, own, max_rows, 7).permute(1, 0, 2, 3, 4).reshape(n_sub * sub, max_rows, 7)[:B]
    nrows = nrows_all.view(world, n_sub, own).permute(1, 0, 2).reshape(n_sub * sub)[:B]
    return res.contiguous(), nrows.contiguous()


def match_sharded_dense(ctx, hashes, nh, params=None, max_rows: int = 16, list_cap: int = 2048, sub_batch: int = 256,
                        group=None):
    """The round-1 exchange, kept for indexes beyond the sparse path's 110 000 tracks: per sub-batch every rank
    counts its shard's hits; one **reduce-scatter** sums the dense per-track histograms and leaves each rank the
    totals of the queries it owns; the owner selects candidates and all-gathers them; every rank collects the
    candidates' (track, delta-t) hits from its shard and an all-to-all sends them to the owners, which align and
    all-gather the result rows.  Moves 200-400 KB per query and shard; measured 2.4 x slower on 8 GPUs than one."""
    import torch
    import torch.distributed as dist

    from . import lib

    params = params or lib.match_defaults()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
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

