"""Phase timings of the hash-range-sharded matcher (run under torchrun, one rank per GPU).
Prints, on rank 0, CUDA-event milliseconds (max over ranks) for 10 000 planted queries vs a 100 000-track index:
emit only, exchange only, owner only, and the pipelined whole for a few sub-batch sizes."""
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
    cap = q.shape[1]
    flag = torch.zeros(1, device=dev)

    def timed(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        dist.all_reduce(flag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {}
    for sub in [int(v) for v in os.environ.get("MFPA_PHASE_SUBS", "1024,2048,5000,10000").split(",")]:
        sub = sub // world * world
        own = sub // world
        wc = sharded.default_words_cap(cap, ctx.depth, world)
        n_sub = -(-B // sub)
        hq, nh = q[:sub].contiguous(), nq[:sub].contiguous()
        words, nwords = ctx.match_emit(hq, nh, wc)
        recv_w, recv_n = torch.empty_like(words), torch.empty_like(nwords)

        def f_emit():
            for _ in range(n_sub):
                ctx.match_emit(hq, nh, wc, words, nwords)

        def f_xchg():
            for _ in range(n_sub):
                dist.all_to_all_single(recv_w, words)
                dist.all_to_all_single(recv_n, nwords)

        f_xchg()

        def f_owner():
            for _ in range(n_sub):
                ctx.match_owner(recv_w.view(world, own, wc), recv_n.view(world, own), mp, 4)

        def f_all():
            sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=sub, exchange="nccl")

        def f_peer():
            sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=sub, exchange="peer")

        px = sharded._peer_exchange(ctx, world, rank, own, wc, None)

        def f_emit_peer():
            for _ in range(n_sub):
                b = px.turn % 2
                ctx.match_emit_peer(hq, nh, px.sets[b], wc)
                px.turn += 1
                ctx.peer_barrier(px.sets[b], px.turn)

        def f_barrier():
            for _ in range(n_sub):
                px.turn += 1
                ctx.peer_barrier(px.sets[0], px.turn)

        out[sub] = {"words_cap": wc, "mean_words": float(nwords.float().mean().item()), "emit": timed(f_emit),
                    "exchange": timed(f_xchg), "owner": timed(f_owner), "nccl_pipelined": timed(f_all),
                    "emit_peer+barrier": timed(f_emit_peer), "barrier": timed(f_barrier), "peer_whole": timed(f_peer)}
        r_n, n_n = sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=sub, exchange="nccl")
        r_p, n_p = sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=sub, exchange="peer")
        out[sub]["peer_equals_nccl"] = bool(torch.equal(n_n, n_p) and torch.equal(r_n, r_p))
        out[sub]["top1"] = float(((n_p > 0) & (r_p[:, 0, 0] == truth)).float().mean().item())
        del words, nwords, recv_w, recv_n
        torch.cuda.empty_cache()
    if rank == 0:
        for k, v in out.items():
            print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()}, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
