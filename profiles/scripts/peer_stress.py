"""Back-to-back calls of the peer-memory sharded matcher against the NCCL exchange (torchrun, one rank per GPU):
prints how many queries differ per call, what their row counts are, and the fullest word list vs its capacity."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from musicfpaugment_b200 import lib, sharded, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ctx = lib.Context(local)
    B, n_tracks = 10000, 100000
    lo, hi = sharded.hash_range(rank, world)
    table, counts, hpid, tt, th = synth.hash_index_device(n_tracks, 1000, seed=5000, device=dev, hash_lo=lo, hash_hi=hi)
    ctx.index_load(table.cpu().numpy().view("uint32"), counts.cpu().numpy(), hpid.cpu().numpy().astype("uint32"), hash_lo=lo)
    del table
    q, nq, truth = synth.planted_queries_device(tt, th, B, n_hashes=400, frac=0.3, seed=6000)
    mp = lib.match_defaults()
    wc = sharded.default_words_cap(q.shape[1], ctx.depth, world)
    _, nw = ctx.match_emit(q, nq, wc)
    fullest = torch.tensor([int(nw.max())], device=dev)
    dist.all_reduce(fullest, op=dist.ReduceOp.MAX)
    r_ref, n_ref = sharded.match_sharded(ctx, q, nq, mp, max_rows=4, exchange="nccl")
    outs = [sharded.match_sharded(ctx, q, nq, mp, max_rows=4, exchange="peer") for _ in range(int(os.environ.get("CALLS", "30")))]
    torch.cuda.synchronize()
    bad = []
    for k, (r, n) in enumerate(outs):
        diff = (n != n_ref) | (r[:, 0, :4] != r_ref[:, 0, :4]).any(dim=1)
        if bool(diff.any()):
            idx = diff.nonzero().flatten()[:6]
            bad.append((k, int(diff.sum()), idx.tolist(), n[idx].tolist(), n_ref[idx].tolist()))
    if rank == 0:
        print("fence", "off" if os.environ.get("MFPA_PEER_NO_FENCE") else "on", "world", world, "words_cap", wc, "fullest list", int(fullest),
              "ref top1", float(((n_ref > 0) & (r_ref[:, 0, 0] == truth)).float().mean()), "ref flagged", int((n_ref < 0).sum()),
              "calls with differences", len(bad), bad[:8], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
