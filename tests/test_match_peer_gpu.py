"""The exchange fused into the sweep (mfpa_match_emit_peer / mfpa_peer_barrier, include/mfpa.h): words stored
straight into the owner ranks' buffers.  On one GPU the "ranks" run one after the other against buffers of the same
device (the kernel's addressing is what is checked); with >= 2 GPUs a real two-process run maps the buffers through
CUDA IPC and compares `sharded.match_sharded(exchange="peer")` with the NCCL exchange and the single-shard matcher."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emit_peer_addresses_owner_buffers(mfpa_ctx):
    from musicfpaugment_b200 import lib, sharded, synth

    world, B = 3, 24
    own = B // world
    table, counts, hpid, th = synth.hash_index(20000, 1000, seed=5000)
    q, nq, _ = synth.planted_queries(th, B, n_hashes=400, frac=0.3, seed=6100)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    p = lib.match_defaults()
    mfpa_ctx.index_load(table, counts, hpid)
    res_w, nrows_w = mfpa_ctx.match(h, n)
    cap = sharded.default_words_cap(h.shape[1], table.shape[1], world)
    # owner o's receive buffer: [world (shard)][own][cap] words + [world][own] counts
    bufs = [torch.full((world, own, cap), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
    cnts = [torch.full((world, own), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
    flags = torch.zeros(lib.MAX_PEERS, dtype=torch.int32, device="cuda")
    for r in range(world):
        lo, hi = sharded.hash_range(r, world)
        mfpa_ctx.index_load(table[lo:hi], counts[lo:hi], hpid, hash_lo=lo)
        ps = lib.PeerSet()
        ps.world, ps.rank = world, r
        for o in range(world):
            ps.words[o], ps.nwords[o], ps.flags[o] = bufs[o].data_ptr(), cnts[o].data_ptr(), flags.data_ptr()
        mfpa_ctx.match_emit_peer(h, n, ps, cap)
        # the plain emit of the same shard lists the same words (as sets: the order inside a list is free)
        w, nw = mfpa_ctx.match_emit(h, n, cap)
        for qi in (0, own, B - 1):
            o, j = divmod(qi, own)
            k = int(nw[qi])
            assert int(cnts[o][r, j]) == k
            assert torch.equal(torch.sort(bufs[o][r, j, :k]).values, torch.sort(w[qi, :k]).values)
    torch.cuda.synchronize()
    for o in range(world):
        assert int(cnts[o].min()) >= 0
        res_o, nrows_o = mfpa_ctx.match_owner_at(bufs[o].data_ptr(), cnts[o].data_ptr(), world, own, cap, p)
        assert torch.equal(nrows_o, nrows_w[o * own:(o + 1) * own]) and torch.equal(res_o, res_w[o * own:(o + 1) * own])


def test_peer_barrier_single_rank_and_alloc(mfpa_ctx):
    from musicfpaugment_b200 import lib

    addr, handle = mfpa_ctx.peer_alloc(4096)
    assert addr and len(handle) == 64 and handle != bytes(64)
    ps = lib.PeerSet()
    ps.world, ps.rank = 1, 0
    ps.flags[0] = addr
    for epoch in (1, 2, 3):
        mfpa_ctx.peer_barrier(ps, epoch)
    torch.cuda.synchronize()
    bad = lib.PeerSet()
    bad.world, bad.rank = 2, 0
    bad.flags[0] = addr            # rank 1 missing
    with pytest.raises(lib.MfpaError):
        mfpa_ctx.peer_barrier(bad, 1)
    mfpa_ctx.peer_free(addr)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from musicfpaugment_b200 import lib, sharded, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = lib.Context(rank)
    table, counts, hpid, th = synth.hash_index(20000, 1000, seed=5000)
    q, nq, _ = synth.planted_queries(th, 301, n_hashes=400, frac=0.3, seed=6100)      # 301: a padded last sub-batch
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    p = lib.match_defaults()
    ctx.index_load(table, counts, hpid)
    res_w, nrows_w = ctx.match(h, n, p)
    lo, hi = sharded.hash_range(rank, world)
    ctx.index_load(table[lo:hi], counts[lo:hi], hpid, hash_lo=lo)
    ok = True
    for _ in range(3):                                                                 # buffers and epochs carry over
        r_p, n_p = sharded.match_sharded(ctx, h, n, p, sub_batch=64, exchange="peer")
        ok = ok and torch.equal(n_p, nrows_w) and torch.equal(r_p, res_w)
    r_n, n_n = sharded.match_sharded(ctx, h, n, p, sub_batch=64, exchange="nccl")
    ok = ok and torch.equal(n_n, nrows_w) and torch.equal(r_n, res_w)
    torch.cuda.synchronize()
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node")
def test_peer_exchange_two_processes():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out.get(0) is True and out.get(1) is True
