"""GPU parity: matching kernels (C ABI) vs the oracle's match_hashes / get_hits, on the golden
index that the reference HashTable.store built and on a synthetic saturated index."""
import os

import numpy as np
import pytest
import torch

from oracle import audfprint_np as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _lib():
    from musicfpaugment_b200 import lib

    return lib


def _golden():
    g = np.load(os.path.join(GOLD, "match.npz"))
    ht = O.HashTable()
    idx = g["counts_nonzero_idx"]
    ht.counts[idx] = g["counts_nonzero"]
    ht.table[idx] = g["table_rows"]
    ht.hashesperid = g["hashesperid"]
    return g, ht


def _batch(queries, cap=None):
    cap = cap or max(len(q) for q in queries)
    h = np.zeros((len(queries), cap, 2), np.int32)
    n = np.zeros(len(queries), np.int32)
    for i, q in enumerate(queries):
        h[i, : len(q)] = q
        n[i] = len(q)
    return torch.from_numpy(h).cuda(), torch.from_numpy(n).cuda()


def _rows_equal(got, want):
    """Same rows; order may differ among equal filtered counts (numpy's argsort is unstable there)."""
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got[:, 1], want[:, 1])
    key = lambda r: sorted(map(tuple, r[:, [0, 1, 2, 3]].tolist()))
    assert key(got) == key(want)


def test_golden_index_match_and_hits(mfpa_ctx):
    g, ht = _golden()
    mfpa_ctx.index_load(ht.table, ht.counts, ht.hashesperid)
    queries = [g[f"q{i}"] for i in range(int(g["n_queries"]))]
    hits = mfpa_ctx.get_hits(torch.from_numpy(queries[0]).cuda()).cpu().numpy()
    assert np.array_equal(hits, g["hits0"])
    h, n = _batch(queries)
    res, nrows = mfpa_ctx.match(h, n)
    res, nrows = res.cpu().numpy(), nrows.cpu().numpy()
    for i in range(len(queries)):
        want = g[f"res{i}"]
        assert nrows[i] == len(want), i
        _rows_equal(res[i, : nrows[i]], want)


def test_synthetic_saturated_index(mfpa_ctx):
    from musicfpaugment_b200 import sharded, synth

    lib = _lib()
    table, counts, hpid, th = synth.hash_index(20000, 1000, seed=5000)
    assert counts.max() > 30
    mfpa_ctx.index_load(table, counts, hpid)
    q, nq, truth = synth.planted_queries(th, 48, n_hashes=400, frac=0.3, seed=6000)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    res, nrows = mfpa_ctx.match(h, n)
    res2, nrows2 = sharded.match_sharded_dense(mfpa_ctx, h, n, sub_batch=20)  # step-wise path, world size 1
    assert torch.equal(res, res2) and torch.equal(nrows, nrows2)
    res, nrows = res.cpu().numpy(), nrows.cpu().numpy()
    ht = O.HashTable()
    ht.table, ht.counts, ht.hashesperid = table, counts, hpid
    top1 = 0
    for i in range(len(q)):
        want = O.match_hashes(ht, q[i, : nq[i]])
        assert nrows[i] == len(want), i
        if len(want):
            _rows_equal(res[i, : nrows[i]], want)
            top1 += int(res[i, 0, 0] == truth[i])
    assert top1 >= 46  # planted tracks are recovered


def test_hash_range_shards_sum_to_the_whole(mfpa_ctx):
    """Counts from two half-range shards add up to the single-shard counts (what the all-reduce does)."""
    from musicfpaugment_b200 import lib, sharded, synth

    table, counts, hpid, th = synth.hash_index(5000, 600, seed=11)
    q, nq, _ = synth.planted_queries(th, 8, n_hashes=300, seed=12)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    mfpa_ctx.index_load(table, counts, hpid)
    whole = mfpa_ctx.match_counts(h, n).clone()
    total = torch.zeros_like(whole)
    lists, nlists = [], []
    p = lib.match_defaults()
    cand, ncand = mfpa_ctx.match_select(whole, p)
    for r in range(2):
        lo, hi = sharded.hash_range(r, 2)
        mfpa_ctx.index_load(table[lo:hi], counts[lo:hi], hpid, hash_lo=lo)
        total += mfpa_ctx.match_counts(h, n)
        l, nl = mfpa_ctx.match_collect(h, n, cand, ncand, p, 2048)
        lists.append(l); nlists.append(nl)
    assert torch.equal(total, whole)
    res_s, nrows_s = mfpa_ctx.match_align(torch.stack(lists), torch.stack(nlists), cand, ncand, p)
    mfpa_ctx.index_load(table, counts, hpid)
    res_w, nrows_w = mfpa_ctx.match(h, n)
    assert torch.equal(nrows_s, nrows_w) and torch.equal(res_s, res_w)


def test_sparse_exchange_equals_single_shard(mfpa_ctx):
    """match_emit on three hash-range shards + match_owner on the stacked (track, delta-t) words (what the one
    all-to-all of sharded.match_sharded delivers) = the single-shard matcher, row for row; an undersized word
    list and an out-of-range query time are flagged, not mis-matched."""
    from musicfpaugment_b200 import lib, sharded, synth

    table, counts, hpid, th = synth.hash_index(20000, 1000, seed=5000)
    q, nq, _ = synth.planted_queries(th, 24, n_hashes=400, frac=0.3, seed=6100)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    p = lib.match_defaults()
    mfpa_ctx.index_load(table, counts, hpid)
    res_w, nrows_w = mfpa_ctx.match(h, n)
    cap = sharded.default_words_cap(h.shape[1], table.shape[1], 3)
    words, nwords = [], []
    for r in range(3):
        lo, hi = sharded.hash_range(r, 3)
        mfpa_ctx.index_load(table[lo:hi], counts[lo:hi], hpid, hash_lo=lo)
        w, nw = mfpa_ctx.match_emit(h, n, cap)
        words.append(w.clone()); nwords.append(nw.clone())
    assert int(torch.stack(nwords).max()) <= cap
    res_s, nrows_s = mfpa_ctx.match_owner(torch.stack(words), torch.stack(nwords), p)
    assert torch.equal(nrows_s, nrows_w) and torch.equal(res_s, res_w)
    # a list that does not hold a shard's hits: the owner reports -1 for exactly those queries
    small = int(torch.stack(nwords).max()) - 1
    w_s = torch.stack([w[:, :small].contiguous() for w in words])
    res_o, nrows_o = mfpa_ctx.match_owner(w_s, torch.stack(nwords), p)
    over = (torch.stack(nwords) > small).any(dim=0)
    assert bool(over.any()) and torch.equal(nrows_o[over], torch.full_like(nrows_o[over], -1))
    assert torch.equal(nrows_o[~over], nrows_w[~over])
    # query times beyond the table's 14-bit time field cannot be packed into a hit key: flagged with -5
    mfpa_ctx.index_load(table, counts, hpid)
    h_bad = h.clone()
    h_bad[0, 0, 0] = 20000
    _, nrows_b = mfpa_ctx.match(h_bad, n)
    assert int(nrows_b[0]) == -5 and torch.equal(nrows_b[1:], nrows_w[1:])


def test_packed_counts_option(mfpa_ctx):
    """MFPA_OPT_MATCH_PACKED: 16-bit counter pairs (what the sharded path reduce-scatters) give the same
    candidates; adding two shards' packed rows as uint32 equals the packed whole."""
    from musicfpaugment_b200 import lib, sharded, synth

    table, counts, hpid, th = synth.hash_index(5001, 600, seed=13)  # odd track count: last word half used
    q, nq, _ = synth.planted_queries(th, 6, n_hashes=300, seed=14)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    p = lib.match_defaults()
    mfpa_ctx.index_load(table, counts, hpid)
    plain = mfpa_ctx.match_counts(h, n)
    cand0, ncand0 = mfpa_ctx.match_select(plain, p)
    try:
        mfpa_ctx.set_option(lib.OPT_MATCH_PACKED, 1)
        packed = mfpa_ctx.match_counts(h, n)
        assert packed.shape == (6, 2501)
        w = packed.cpu().numpy().view(np.uint32)
        un = np.stack([w & 0xFFFF, w >> 16], axis=-1).reshape(6, -1)[:, :5001]
        assert np.array_equal(un.astype(np.int32), plain.cpu().numpy())
        cand1, ncand1 = mfpa_ctx.match_select(packed, p)
        assert torch.equal(cand0, cand1) and torch.equal(ncand0, ncand1)
        total = torch.zeros_like(packed)
        for r in range(2):
            lo, hi = sharded.hash_range(r, 2)
            mfpa_ctx.index_load(table[lo:hi], counts[lo:hi], hpid, hash_lo=lo)
            total += mfpa_ctx.match_counts(h, n)
        assert torch.equal(total, packed)
    finally:
        mfpa_ctx.set_option(lib.OPT_MATCH_PACKED, 0)
        mfpa_ctx.index_load(table, counts, hpid)


@pytest.mark.parametrize("n_tracks,per_track,n_hashes", [(301, 3000, 2000), (12000, 1500, 4000)])
def test_fused_match_equals_four_step_path(mfpa_ctx, n_tracks, per_track, n_hashes):
    """mfpa_match's one-kernel counts+select+collect (histogram kept in shared memory) gives the rows of
    the four-step path: long queries (more rows than the kernel's row cache), a small index where many
    tracks pass the count threshold (deep candidate lists), an odd track count, empty queries; the second
    case has more contenders (thousands of tracks above the count threshold) than either kernel's
    shared-memory list holds, so both rank by repeated scans."""
    from musicfpaugment_b200 import lib, synth

    table, counts, hpid, th = synth.hash_index(n_tracks, per_track, seed=21)
    q, nq, _ = synth.planted_queries(th, 12 if n_tracks < 1000 else 5, n_hashes=n_hashes, frac=0.2, seed=22)
    nq[3] = 0
    nq[-1] = 1
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    mfpa_ctx.index_load(table, counts, hpid)
    p = lib.match_defaults()
    res_f, nrows_f = mfpa_ctx.match(h, n, p, max_rows=128)
    try:
        mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 1)
        res_u, nrows_u = mfpa_ctx.match(h, n, p, max_rows=128)
    finally:
        mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 0)
    assert torch.equal(nrows_f, nrows_u)
    assert int(nrows_f[3]) == 0 and int(nrows_f[0]) >= 1
    _, ncand = mfpa_ctx.match_select(mfpa_ctx.match_counts(h, n), p)
    assert int(ncand.max()) == p.search_depth   # the candidate lists are as deep as they get
    for i in range(len(q)):
        k = int(nrows_f[i])
        assert torch.equal(res_f[i, :k], res_u[i, :k]), i
    ht = O.HashTable()
    ht.table, ht.counts, ht.hashesperid = table, counts, hpid
    for i in (0, 2):
        want = O.match_hashes(ht, q[i, : nq[i]])
        _rows_equal(res_f[i, : int(nrows_f[i])].cpu().numpy(), want)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_candidate_ranking_with_uneven_hashesperid(mfpa_ctx, seed):
    """_best_count_ids ranks by raw / hashesperid over ALL tracks (audfprint_match.py:110-129): with track
    lengths spread over two orders of magnitude a short track with few hits outranks long tracks with many,
    which is what the contender-list shortcut (quotient >= the smallest quotient among tracks above the
    count threshold; tracks skipped by raw count < m * min(hashesperid)) must get right.  One-kernel
    matcher, four-step path and the oracle must agree on candidates and result rows."""
    from musicfpaugment_b200 import lib, synth

    r = np.random.default_rng(100 + seed)
    n_tracks = [257, 1500, 4001][seed - 1]
    table, counts, hpid, th = synth.hash_index(n_tracks, 2500, seed=40 + seed)
    hpid = r.integers(1, 6000, size=n_tracks).astype(np.uint32)        # weights only: any positive value is legal
    hpid[r.integers(0, n_tracks, size=8)] = 1                           # a few extremely "short" tracks
    q, nq, _ = synth.planted_queries(th, 6, n_hashes=1200, frac=0.25, seed=50 + seed)
    h, n = torch.from_numpy(q).cuda(), torch.from_numpy(nq).cuda()
    mfpa_ctx.index_load(table, counts, hpid)
    p = lib.match_defaults()
    res_f, nrows_f = mfpa_ctx.match(h, n, p, max_rows=128)
    cand_d, ncand_d = mfpa_ctx.match_select(mfpa_ctx.match_counts(h, n), p)   # dense-row selection
    try:
        mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 1)
        res_u, nrows_u = mfpa_ctx.match(h, n, p, max_rows=128)
    finally:
        mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 0)
    assert torch.equal(nrows_f, nrows_u)
    ht = O.HashTable()
    ht.table, ht.counts, ht.hashesperid = table, counts, hpid
    for i in range(len(q)):
        k = int(nrows_f[i])
        assert torch.equal(res_f[i, :k], res_u[i, :k]), i
        # candidates: exact rational order, ties -> higher id first
        hits = ht.get_hits(q[i, : nq[i]])
        raw = np.bincount(hits[:, 0], minlength=n_tracks).astype(np.int64)
        ids = np.nonzero(raw)[0]
        depth = min(int((raw > p.threshcount).sum()), p.search_depth)
        assert int(ncand_d[i]) == depth
        from fractions import Fraction
        order = sorted(ids.tolist(), key=lambda t: (Fraction(int(raw[t]), int(hpid[t])), t), reverse=True)[:depth]
        assert cand_d[i, :depth, 0].cpu().tolist() == order, i
        assert cand_d[i, :depth, 1].cpu().tolist() == [int(raw[t]) for t in order]
        want = O.match_hashes(ht, q[i, : nq[i]])
        assert k == len(want), i
        if k:
            _rows_equal(res_f[i, :k].cpu().numpy(), want)


def test_match_without_index_raises():
    lib = _lib()
    ctx = lib.Context(0)
    with pytest.raises(lib.MfpaError):
        ctx.match(torch.zeros(1, 4, 2, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda"))
    ctx.close()


def test_packed_counter_overflow_is_detected_not_silent(mfpa_ctx):
    """One track collecting >= 65536 hits from one query overflows the packed 16-bit counters of the one-kernel matcher:
    the histogram checksum flags the query (-6); the int32 counters (option value 2) give the exact raw counts."""
    lib = _lib()
    depth, n_tracks = 100, 8
    table = np.zeros((1 << 20, depth), np.uint32)
    counts = np.zeros(1 << 20, np.int32)
    n_hashes = 700                                     # 700 buckets x 100 entries of track 3 = 70 000 hits
    hashes = np.arange(n_hashes, dtype=np.int64) * 1499 + 17
    table[hashes] = ((3 + 1) << 14) + 50
    counts[hashes] = depth
    table[hashes[:40], 0] = ((5 + 1) << 14) + 70       # a second track with a few hits
    hpid = np.full(n_tracks, 1000, np.uint32)
    mfpa_ctx.index_load(table, counts, hpid)
    q = np.stack([np.full(n_hashes, 20, np.int32), hashes.astype(np.int32)], axis=1)[None]
    h, n = torch.from_numpy(q).cuda(), torch.tensor([n_hashes], dtype=torch.int32).cuda()
    _, nrows = mfpa_ctx.match(h, n)
    assert int(nrows[0]) == -6
    mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 2)    # four-step path, int32 counters in global memory
    try:
        raw = mfpa_ctx.match_counts(h, n)
        _, nrows = mfpa_ctx.match(h, n)
    finally:
        mfpa_ctx.set_option(lib.OPT_MATCH_UNFUSED, 0)
    raw = raw[0].cpu().numpy()
    assert raw[3] == 70000 - 40 and raw[5] == 40 and raw.sum() == 70000     # exact, no wrap-around
    assert int(nrows[0]) == -1                         # 70 000 candidate hits: past the alignment capacity, flagged
