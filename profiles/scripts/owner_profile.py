"""One GPU standing in for rank 0 of 2: emit both hash-range shards of 2368 owned queries, then the owner step.
Run under ncu to profile match_emit_kernel / match_owner_kernel / match_align_kernel on the bench's index shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from musicfpaugment_b200 import lib, sharded, synth  # noqa: E402

dev = torch.device("cuda", 0)
ctx = lib.Context(0)
world, own, n_tracks = 2, 2368, 100000
mp = lib.match_defaults()
words, nwords = [], []
q = nq = None
for r in range(world):
    lo, hi = sharded.hash_range(r, world)
    table, counts, hpid, tt, th = synth.hash_index_device(n_tracks, 1000, seed=5000, device=dev, hash_lo=lo, hash_hi=hi)
    ctx.index_load(table.cpu().numpy().view("uint32"), counts.cpu().numpy(), hpid.cpu().numpy().astype("uint32"), hash_lo=lo)
    del table
    if q is None:
        q, nq, truth = synth.planted_queries_device(tt, th, own, n_hashes=400, frac=0.3, seed=6000)
    wc = sharded.default_words_cap(q.shape[1], ctx.depth, world)
    for _ in range(2):
        w, nw = ctx.match_emit(q, nq, wc)
    words.append(w)
    nwords.append(nw)
W, NW = torch.stack(words), torch.stack(nwords)
for _ in range(3):
    res, nrows = ctx.match_owner(W, NW, mp, 4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    res, nrows = ctx.match_owner(W, NW, mp, 4)
e1.record()
torch.cuda.synchronize()
print("owner ms for", own, "queries:", e0.elapsed_time(e1) / 5, "top1", float(((nrows > 0) & (res[:, 0, 0] == truth)).float().mean()))
