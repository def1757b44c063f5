"""CPU tests of the multi-GPU plumbing (musicfpaugment_b200/sharded.py) under gloo, world size 2.

The four matching kernels are replaced by a numpy stand-in built from the oracle (test
infrastructure only), so what is exercised here is the host logic a real run depends on:
the hash-range partition of the index, the one all-to-all of every shard's packed (track, delta-t)
hit words to the query's owner rank with the next sub-batch's sweep behind it, the owners' result
rows gathered back into query order, the dense round-1 exchange (reduce-scatter of per-track counts,
all-gather of candidates, all-to-all of candidate hit lists), the peer-memory exchange (`PeerExchange`: buffer layout,
alternating receive buffers and barrier epochs, over POSIX shared memory instead of CUDA IPC) and the contiguous
query partition.
The same `match_sharded` drives the CUDA kernels over NCCL on the GPU box.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from musicfpaugment_b200 import sharded, synth
from oracle import audfprint_np as O


class _P:  # mfpa_match_params stand-in
    window, threshcount, search_depth, max_alignments_per_id = 2, 5, 100, 100


class NumpyShardCtx:
    """The step-wise matching interface of lib.Context over one hash-range shard, in numpy."""

    def __init__(self, table, counts, hpid, lo, hi):
        self.ht = O.HashTable()
        self.ht.table = np.zeros_like(table)
        self.ht.counts = np.zeros_like(counts)
        self.ht.table[lo:hi] = table[lo:hi]
        self.ht.counts[lo:hi] = counts[lo:hi]
        self.ht.hashesperid = hpid
        self.n_tracks = len(hpid)
        self.depth = table.shape[1]

    def _hits(self, hq, n):
        return self.ht.get_hits(hq[:n].numpy())

    def match_counts(self, hashes, nh):
        out = torch.zeros(hashes.shape[0], self.n_tracks, dtype=torch.int32)
        for i in range(hashes.shape[0]):
            ids = self._hits(hashes[i], int(nh[i]))[:, 0]
            out[i] = torch.from_numpy(np.bincount(ids, minlength=self.n_tracks).astype(np.int32))
        return out

    def match_select(self, counts, p):
        B = counts.shape[0]
        cand = torch.zeros(B, p.search_depth, 2, dtype=torch.int32)
        ncand = torch.zeros(B, dtype=torch.int32)
        for i in range(B):
            c = counts[i].numpy()
            ids = np.nonzero(c)[0]
            raw = c[ids]
            wtd = raw / self.ht.hashesperid[ids].astype(float)
            order = np.argsort(wtd)[::-1][: min(int(np.count_nonzero(raw > p.threshcount)), p.search_depth)]
            ncand[i] = len(order)
            cand[i, : len(order), 0] = torch.from_numpy(ids[order].astype(np.int32))
            cand[i, : len(order), 1] = torch.from_numpy(raw[order].astype(np.int32))
        return cand, ncand

    def match_collect(self, hashes, nh, cand, ncand, p, list_cap):
        B = hashes.shape[0]
        lst = torch.zeros(B, list_cap, dtype=torch.int32)
        nlist = torch.zeros(B, dtype=torch.int32)
        for i in range(B):
            hits = self._hits(hashes[i], int(nh[i]))
            ids = cand[i, : int(ncand[i]), 0].numpy()
            rank = {int(t): k for k, t in enumerate(ids)}
            rows = [(rank[int(h[0])] << 16) | (int(h[1]) + 16384) for h in hits if int(h[0]) in rank]
            nlist[i] = len(rows)
            lst[i, : len(rows)] = torch.tensor(rows[:list_cap], dtype=torch.int32)
        return lst, nlist

    def match_align(self, lists, nlists, cand, ncand, p, max_rows):
        n_lists, B, _ = lists.shape
        res = torch.zeros(B, max_rows, 7, dtype=torch.int32)
        nrows = torch.zeros(B, dtype=torch.int32)
        for i in range(B):
            ent = np.concatenate([lists[s, i, : int(nlists[s, i])].numpy() for s in range(n_lists)]).astype(np.int64)
            rows = []
            mint = int((ent & 0xFFFF).min()) - 16384 if len(ent) else 0
            for rank in range(int(ncand[i])):
                dt = (ent[(ent >> 16) == rank] & 0xFFFF) - 16384 - mint
                bc = np.bincount(dt)
                f = np.zeros(bc.shape, np.float32)
                lm = np.nonzero(O.locmax(bc))[0]
                f[lm] = bc[lm]
                found = 0
                while True:
                    mode = int(np.argmax(f))
                    if f[mode] <= p.threshcount:
                        break
                    lo, hi = max(0, mode - p.window), mode + p.window + 1
                    rows.append([int(cand[i, rank, 0]), int(bc[lo:hi].sum()), mode + mint, int(cand[i, rank, 1]), rank, 0, 0])
                    f[lo:hi] = 0
                    found += 1
                    if found > p.max_alignments_per_id:
                        break
            rows = np.asarray(rows, np.int32).reshape(-1, 7)
            rows = rows[(-rows[:, 1]).argsort(kind="stable")][:max_rows]
            nrows[i] = len(rows)
            res[i, : len(rows)] = torch.from_numpy(rows)
        return res, nrows


    def match_emit(self, hashes, nh, words_cap):
        """lib.Context.match_emit: this shard's hits as (track << 15) | (delta-t + 16384) words."""
        B = hashes.shape[0]
        words = torch.zeros(B, words_cap, dtype=torch.int32)
        nwords = torch.zeros(B, dtype=torch.int32)
        for i in range(B):
            hits = self._hits(hashes[i], int(nh[i]))
            w = ((hits[:, 0].astype(np.int64) << 15) | (hits[:, 1].astype(np.int64) + 16384)).astype(np.uint32)
            nwords[i] = len(w)
            words[i, : min(len(w), words_cap)] = torch.from_numpy(w[:words_cap].view(np.int32))
        return words, nwords

    def match_owner(self, words, nwords, p, max_rows):
        """lib.Context.match_owner: counts, select, collect and align from the hit words of all shards."""
        n_shards, B, _ = words.shape
        counts = torch.zeros(B, self.n_tracks, dtype=torch.int32)
        lists = []
        for i in range(B):
            w = np.concatenate([words[s, i, : int(nwords[s, i])].numpy().view(np.uint32) for s in range(n_shards)]).astype(np.int64)
            lists.append(w)
            counts[i] = torch.from_numpy(np.bincount(w >> 15, minlength=self.n_tracks).astype(np.int32))
        cand, ncand = self.match_select(counts, p)
        list_cap = max(1, max(len(w) for w in lists))
        lst = torch.zeros(1, B, list_cap, dtype=torch.int32)
        nlist = torch.zeros(1, B, dtype=torch.int32)
        for i, w in enumerate(lists):
            rank = {int(t): k for k, t in enumerate(cand[i, : int(ncand[i]), 0].numpy())}
            rows = [(rank[int(x >> 15)] << 16) | int(x & 0x7FFF) for x in w if int(x >> 15) in rank]
            nlist[0, i] = len(rows)
            lst[0, i, : len(rows)] = torch.tensor(rows, dtype=torch.int32)
        return self.match_align(lst, nlist, cand, ncand, p, max_rows)

    def match(self, hashes, nh, p, max_rows):
        """lib.Context.match (one shard holds the whole index): the four steps back to back."""
        counts = self.match_counts(hashes, nh)
        cand, ncand = self.match_select(counts, p)
        lst, nlist = self.match_collect(hashes, nh, cand, ncand, p, 4096)
        return self.match_align(lst[None], nlist[None], cand, ncand, p, max_rows)


class PeerShardCtx(NumpyShardCtx):
    """NumpyShardCtx + the peer-memory entry points of lib.Context (peer_alloc / peer_open / match_emit_peer /
    peer_barrier / match_owner_at) over POSIX shared memory: what `sharded.PeerExchange` and the "peer" exchange of
    `match_sharded` drive on GPUs through CUDA IPC.  Addresses are (segment number << 40) + byte offset."""

    def __init__(self, *a):
        super().__init__(*a)
        self._segs = {}

    def _attach(self, shm):
        k = len(self._segs) + 1
        self._segs[k] = shm
        return k << 40

    def _view(self, addr, n, dtype=np.int32):
        shm = self._segs[addr >> 40]
        return np.ndarray((n,), dtype=dtype, buffer=shm.buf, offset=addr & ((1 << 40) - 1))

    def peer_alloc(self, nbytes):
        from multiprocessing import shared_memory

        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        shm.buf[:nbytes] = bytes(nbytes)
        return self._attach(shm), shm.name.encode().ljust(64, b"\0")

    def peer_open(self, handle):
        from multiprocessing import shared_memory

        return self._attach(shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode()))

    def peer_close(self, addr):
        self._segs.pop(addr >> 40).close()

    def peer_free(self, addr):
        shm = self._segs.pop(addr >> 40)
        shm.close()
        shm.unlink()

    def match_emit_peer(self, hashes, nh, peers, words_cap):
        B = hashes.shape[0]
        own = B // peers.world
        words, nwords = self.match_emit(hashes, nh, words_cap)
        for q in range(B):
            owner, j = divmod(q, own)
            slot = peers.rank * own + j
            self._view(peers.words[owner] + slot * words_cap * 4, words_cap)[:] = words[q].numpy()
            self._view(peers.nwords[owner] + slot * 4, 1)[0] = int(nwords[q])

    def peer_barrier(self, peers, epoch):
        self._view(peers.flags[peers.rank] + 4 * peers.rank, 1)[0] = epoch   # the kernel's flag word, for the record
        dist.barrier()

    def match_owner_at(self, w_addr, n_addr, n_shards, B, words_cap, p, max_rows):
        words = torch.from_numpy(self._view(w_addr, n_shards * B * words_cap).copy()).view(n_shards, B, words_cap)
        nwords = torch.from_numpy(self._view(n_addr, n_shards * B).copy()).view(n_shards, B)
        return self.match_owner(words, nwords, p, max_rows)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table, counts, hpid, th = synth.hash_index(300, 400, seed=21, depth=20)
        q, nq, truth = synth.planted_queries(th, 7, n_hashes=200, frac=0.4, seed=22)  # 7: the last sub-batch needs padding
        lo, hi = sharded.hash_range(rank, world)
        ctx = NumpyShardCtx(table, counts, hpid, lo, hi)
        res, nrows = sharded.match_sharded(ctx, torch.from_numpy(q), torch.from_numpy(nq), params=_P, max_rows=8,
                                           sub_batch=4)
        # the round-1 exchange (dense histograms through a reduce-scatter) must give the same rows
        res_d, nrows_d = sharded.match_sharded_dense(ctx, torch.from_numpy(q), torch.from_numpy(nq), params=_P, max_rows=8,
                                                     list_cap=4096, sub_batch=4)
        assert torch.equal(nrows_d, nrows) and torch.equal(res_d, res)
        # the exchange fused into the sweep: words written straight into the owners' (shared-memory) buffers; three
        # calls so that the two receive buffers and the barrier epochs carry over from call to call
        pctx = PeerShardCtx(table, counts, hpid, lo, hi)
        for sub in (4, 4, 1000):
            res_p, nrows_p = sharded.match_sharded(pctx, torch.from_numpy(q), torch.from_numpy(nq), params=_P, max_rows=8,
                                                   sub_batch=sub, exchange="peer")
            assert torch.equal(nrows_p, nrows) and torch.equal(res_p, res)
        sharded.release_peer_exchanges(pctx)
        assert not pctx._segs
        # every rank must hold identical results
        gathered = [torch.zeros_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        # throughput mode: the whole index on every rank, queries sharded, rows all-gathered - same rows
        full = NumpyShardCtx(table, counts, hpid, 0, 1 << 20)
        res_r, nrows_r = sharded.match_replicated(full, torch.from_numpy(q), torch.from_numpy(nq), params=_P, max_rows=8)
        assert res_r.shape == res.shape and torch.equal(nrows_r, nrows) and torch.equal(res_r, res)
        # query-sharded fingerprint slices cover the batch exactly once
        mine = torch.zeros(11, dtype=torch.int32)
        a, b = sharded.query_slice(11, rank, world)
        mine[a:b] = 1
        dist.all_reduce(mine)
        assert torch.equal(mine, torch.ones(11, dtype=torch.int32))
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), res=res.numpy(), nrows=nrows.numpy(), q=q, nq=nq, truth=truth)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_match_sharded_equals_single_table(tmp_path, world):
    """world 3: uneven hash ranges (2^20 / 3), 7 queries padded to sub-batches of 3, owners of 1-2 queries."""
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = np.load(tmp_path / "out.npz")
    table, counts, hpid, _ = synth.hash_index(300, 400, seed=21, depth=20)
    ht = O.HashTable(depth=20)
    ht.table, ht.counts, ht.hashesperid = table, counts, hpid
    top1 = 0
    for i in range(len(g["q"])):
        want = O.match_hashes(ht, g["q"][i, : g["nq"][i]])[:8]
        n = int(g["nrows"][i])
        assert n == len(want), i
        got = g["res"][i, :n]
        assert np.array_equal(got[:, 1], want[:, 1])
        assert sorted(map(tuple, got[:, :4].tolist())) == sorted(map(tuple, want[:, :4].tolist()))
        top1 += int(n > 0 and got[0, 0] == g["truth"][i])
    assert top1 >= len(g["q"]) - 1


def test_partitions():
    for world in (1, 2, 3, 4, 8):
        edges = [sharded.hash_range(r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1 << 20
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        for n in (0, 1, 7, 10000):
            sl = [sharded.query_slice(n, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(sl[r][1] == sl[r + 1][0] for r in range(world - 1))
