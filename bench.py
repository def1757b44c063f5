#!/usr/bin/env python
"""Benchmark of the query-side hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload fingerprint] [--queries 10000]

One "step" = one pass of the hot path over one batch of synthetic 8 s / 8 kHz
queries.  At N = 1 the workload is BASELINE.json configs[1]: 10 k queries
through batched STFT (n_fft 512, hop 256) + audfprint peak picking + 20-bit
landmark hashes.  N > 1 (torchrun, one rank per GPU) shards queries with no
data-path collective: every rank runs its own 10 k batch (weak scaling).

* value    device-timed queries/s, inputs resident in HBM (CUDA events, max over ranks)
* e2e      the same batch through the C-ABI host entry point mfpa_fingerprint_host:
           pinned host waveforms in, CSR hash rows out, copies inside the timed region
* roofline dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
* cpu_baseline  the numpy oracle (port of the reference CPU path) on the host cores,
           bounded sample, rank 0 / N = 1 only
--impl reference times that CPU path alone (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its version banner to stdout when NCCL_DEBUG is set in the environment; stdout carries
# exactly one JSON line, so anything NCCL has to say goes to stderr.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
# ... and because the torch-bundled NCCL still prints its banner with a bare printf on some boxes, file
# descriptor 1 itself points at stderr for the whole run; the JSON line is written to the saved descriptor.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


METRIC = "8s_query_fingerprints_per_sec"
UNIT = "queries/s"
T_QUERY = 64000
N_FRAMES = 251
# SURVEY.md §8(d) algorithmic bytes per query (each stage reads its input once, writes its output once)
BYTES_STFT = 256_000 + 257 * N_FRAMES * 4          # S2: waveform in, magnitudes out
BYTES_PEAKS = 257 * N_FRAMES * 4 + 256 * N_FRAMES  # S3: magnitudes in, peak mask (u8-equivalent) out
BYTES_FUSED = 256_000                              # S2-S4 fused: waveform in (+ 8 B per hash out)
BYTES_CHAIN = 544_000                              # S1-S4 fused: x + noise + IR in (+ 8 B per hash out)
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 10 000 queries, shifts=1, from the
# `ncu --set full` capture summarised in profiles/r01k_summary.txt (scaled by items/10000 for other sizes)
NCU_TRAFFIC_10K = {"stft_mag": 2.692782e9 + 2.607007e9, "audfprint_peaks": 2.963450e9 + 0.020312e9,
                   "landmark_hashes(+merge)": 0.020157e9 + 0.000076e9}


def _measured(keys, default):
    """First of `keys` found (at any depth) in the driver-written MEASURED_PEAKS.json, else the profiling
    recipe's fallback - a missing or re-shaped file must not take the bench down."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
    except Exception:
        return default, "fallback"

    def find(o, k):
        if isinstance(o, dict):
            if k in o and isinstance(o[k], (int, float)):
                return float(o[k])
            for v in o.values():
                r = find(v, k)
                if r is not None:
                    return r
        return None

    for k in keys:
        v = find(d, k)
        if v:
            return v, "measured"
    return default, "fallback"


def _peaks():
    return _measured(("hbm_gbs", "hbm_copy_gbs", "hbm_gb_s", "hbm_bw_gbs"), 6650.0)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._halt = gpu_index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    seed, n, shifts = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch

    torch.set_num_threads(1)
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O

    X = synth.music_like(min(n, 8), seed=seed).numpy()  # 8 distinct queries per worker, cycled
    t0 = time.perf_counter()
    nh = 0
    for i in range(n):
        nh += len(O.wave2hashes(X[i % len(X)], shifts))
    return time.perf_counter() - t0, n, nh


def cpu_fingerprint_rate(per_core: int, shifts: int = 1, cores: int | None = None):
    """queries/s of the oracle (port of afp/audfprint wavfile2hashes) using all host cores."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 1, shifts)] * cores)  # warm the workers (imports)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(1000 + i, per_core, shifts) for i in range(cores)])
        wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    busy = max(r[0] for r in res)
    return n / busy, cores, n, wall


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = max(4, args.cpu_queries_per_core)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, cores, n, wall = cpu_fingerprint_rate(per_core, args.shifts, cores)
        if i >= args.warmup:
            rates.append((rate, n, wall))
    value = sum(r[0] for r in rates) / len(rates)
    n = rates[0][1]
    sample = f"{n} synthetic 8 s queries per step ({per_core} per core), oracle port of wavfile2hashes, shifts={args.shifts}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "queries_per_step": n, "n_samples": T_QUERY, "shifts": args.shifts},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(args):
    return (f"{args.queries} synthetic 8 s 8 kHz mono queries per GPU: batched STFT (n_fft 512, hop 256) + audfprint "
            f"peak picking + 20-bit landmark hashes, shifts={args.shifts} (BASELINE.json configs[1])")


# ------------------------------------------------------------------ GPU arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    # libmfpa.so is a build artefact: (re)build it from the sources if it is missing or stale - rank 0 of the
    # node does it, the others wait at the rendezvous below (a no-op when the source stamp is current)
    if local_rank == 0:
        from musicfpaugment_b200 import build as _build

        _build.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    from musicfpaugment_b200 import lib, synth

    ctx = lib.Context(local_rank, spread_table=np.exp(-0.5 * ((np.arange(-256, 257) / 30.0) ** 2)))
    p = lib.afp_defaults()
    B, S = args.queries, args.shifts
    items = B * S

    x = synth.music_like(B, seed=1234 + rank, device=dev, chunk=32)
    n = lib.num_frames(T_QUERY)
    mag = torch.empty(items, n, lib.MAG_PITCH, dtype=torch.float32, device=dev)
    qmax = torch.empty(items, dtype=torch.float32, device=dev)
    rec = torch.empty(items, n, dtype=torch.int64, device=dev)
    cap = lib.HASHES_PER_FRAME * n
    hashes = torch.empty(items, cap, 2, dtype=torch.int32, device=dev)
    nh = torch.empty(items, dtype=torch.int32, device=dev)
    out = torch.empty(B, cap * S, 2, dtype=torch.int32, device=dev) if S > 1 else hashes
    nout = torch.empty(B, dtype=torch.int32, device=dev) if S > 1 else nh
    L = lib.raw()
    C = lib.C if hasattr(lib, "C") else __import__("ctypes")
    h = ctx.handle
    ptr = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    launches_per_step = 3 + (1 if S > 1 else 0)

    def step(ev=None):
        if ev:
            ev[0].record()
        lib.check(L.mfpa_stft_mag(h, ptr(x), B, T_QUERY, x.stride(0), S, ptr(mag), ptr(qmax), stream))
        if ev:
            ev[1].record()
        lib.check(L.mfpa_audfprint_peaks(h, ptr(mag), ptr(qmax), B, T_QUERY, S, C.byref(p), ptr(rec), None, stream))
        if ev:
            ev[2].record()
        lib.check(L.mfpa_landmark_hashes(h, ptr(rec), items, n, C.byref(p), 1, ptr(hashes), cap, ptr(nh), stream))
        if S > 1:
            lib.check(L.mfpa_merge_shifts(h, ptr(hashes), ptr(nh), B, S, cap, n, ptr(out), cap * S, ptr(nout), stream))
        if ev:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    t_stft = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    t_peaks = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    t_lm = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    tot_hashes = int(nout.sum().item())

    # ---- end to end through the host C-ABI entry point (pinned host memory in, CSR rows out)
    x_host = torch.empty(B, T_QUERY, dtype=torch.float32).pin_memory()
    x_host.copy_(x)
    rows_host = torch.empty(max(tot_hashes * 2, 1024), 2, dtype=torch.int32).pin_memory()
    offs_host = torch.empty(B + 1, dtype=torch.int64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        lib.check(L.mfpa_fingerprint_host(h, ptr(x_host), B, T_QUERY, S, C.byref(p), ptr(rows_host), rows_host.shape[0],
                                          ptr(offs_host)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    assert int(offs_host[-1]) == tot_hashes, (int(offs_host[-1]), tot_hashes)
    # the same call on 16-bit PCM (what a decoded audio file holds): half the bytes over PCIe
    x16_host = torch.empty(B, T_QUERY, dtype=torch.int16).pin_memory()
    x16_host.copy_((x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))

    def e2e16_step():
        lib.check(L.mfpa_fingerprint_host_pcm16(h, ptr(x16_host), B, T_QUERY, S, C.byref(p), ptr(rows_host), rows_host.shape[0],
                                                ptr(offs_host)))

    e2e16_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e16_step()
    torch.cuda.synchronize()
    e2e16_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    n_rows16 = int(offs_host[-1])
    clocks = sampler.stop()   # sampled across the three timed regions above (device-resident, e2e, e2e PCM16)

    times = torch.tensor([ms_total, e2e_ms, e2e16_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e16_ms = times.tolist()
    if rank == 0:
        hbm, how = _peaks()
        ms_step = ms_total / args.steps
        value = world * B / (ms_step * 1e-3)
        stage_ms = {"stft_mag": t_stft, "audfprint_peaks": t_peaks, "landmark_hashes(+merge)": t_lm}
        dom = max(stage_ms, key=stage_ms.get)
        dom_bytes = {"stft_mag": BYTES_STFT, "audfprint_peaks": BYTES_PEAKS, "landmark_hashes(+merge)": 8000}[dom] * items
        achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "queries_per_gpu": B, "n_samples": T_QUERY, "shifts": S,
                       "l2_policy": "inputs larger than L2 (2.56 GB waveforms + 2.65 GB magnitudes per step)",
                       "hashes_per_query": tot_hashes / B, "parallelism": f"query-sharded x{world}, no collective"},
            "stage_ms": stage_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": NCU_TRAFFIC_10K[dom] * items / 10000,
                         "traffic_source": "profiles/r01k_summary.txt (ncu --set full, dram read+write)", "peak_source": how,
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0,   # also against the spec-sheet figure
                         "whole_path_frac": (BYTES_FUSED * B + 8 * tot_hashes) / (ms_step * 1e-3) / 1e9 / hbm},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": B * T_QUERY * 4, "d2h_bytes_per_step": tot_hashes * 8 + (B + 1) * 8,
                    "api": "mfpa_fingerprint_host (pinned host buffers, chunked copy/compute overlap)"},
            "e2e_pcm16": {"value": world * B / (e2e16_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16_ms,
                          "h2d_bytes_per_step": B * T_QUERY * 2, "d2h_bytes_per_step": n_rows16 * 8 + (B + 1) * 8,
                          "api": "mfpa_fingerprint_host_pcm16 (int16 PCM host buffers, converted to float32 on the device)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, nq, wall = cpu_fingerprint_rate(args.cpu_queries_per_core, S)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{nq} of the same synthetic queries ({args.cpu_queries_per_core} per core), "
                                              f"numpy oracle of wavfile2hashes, {wall:.1f} s wall"}
    del x, mag, rec, hashes, out, x_host, rows_host, x16_host
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs, device-timed (extra keys of the same JSON line)
    extras = {}
    if "chain" in args.also:
        ms, nhash = bench_full_chain(ctx, lib, dev, rank, B, 1, max(2, args.steps // 2), 3, barrier)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        hbm, _ = _peaks()
        extras["full_chain"] = {
            "workload": f"{B} queries per GPU through the AugmentFP chain (HPF, 1 s IR conv, noise at random SNR, gain, "
                        "clipping, LPF, HPF) fused with STFT + peaks + hashes, shifts=1 (BASELINE.json configs[2]); "
                        "loudspeaker cut-off clamped to >= 20 Hz",
            "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "hashes_per_query": nhash / B,
            "algorithmic_bytes_per_query": BYTES_CHAIN,
            "hbm_frac": BYTES_CHAIN * B / (ms * 1e-3) / 1e9 / hbm}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, cores, nq = cpu_chain_rate(2, 1)
            extras["full_chain"]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                                    "sample": f"{nq} queries (2 per core), numpy oracle of the chain + wavfile2hashes"}
    if "match" in args.also:
        ms, top1, nqh, rep = bench_match(ctx, lib, dev, rank, world, B, args.tracks, max(2, args.steps // 2), 3, barrier)
        t = torch.tensor([ms, rep[0] if rep else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_rep = float(t[0].item()), float(t[1].item())
        extras["match"] = {
            "workload": f"{B} planted 400-hash queries (the same batch on every rank) vs a synthetic {args.tracks}-track index "
                        f"(1000 hashes/track, depth 100) sharded by hash range over {world} GPU(s) (BASELINE.json configs[4])",
            "value": B / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "scaling": "strong",
            "top1_equals_planted_track": top1, "query_hashes": nqh,
            "collective": "none" if world == 1 else "NCCL reduce-scatter of packed per-track counts (query owners), all-gather of "
                                                                  "candidates, all-to-all of candidate hit lists, all-gather of result rows"}
        if rep:
            extras["match"]["replicated_index"] = {
                "workload": "the same queries against the whole index replicated on every rank, queries sharded "
                            "(sharded.match_replicated: no data-path collective, result rows all-gathered)",
                "value": B / (ms_rep * 1e-3), "unit": "queries/s", "ms_per_step": ms_rep, "scaling": "strong",
                "rows_equal_hash_range_mode": rep[1]}
    if "unet" in args.also:
        Bu = args.unet_queries
        ms, finite = bench_unet(ctx, lib, dev, rank, Bu, max(2, args.steps // 3), 2, barrier, args.unet_chunk)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tf = UNET_GFLOP * Bu / ms  # GFLOP / ms = TFLOP/s
        sustained, sustained_src = _measured(("bf16_tflops_sustained", "bf16_sustained_tflops", "bf16_tflops"), 1400.0)
        extras["unet"] = {
            "workload": f"UNet(1,1) denoiser forward on {Bu} normalised 257x251 magnitude spectrograms per GPU, random init, "
                        f"bf16 operands / fp32 accumulate, tcgen05 implicit GEMM, chunks of {args.unet_chunk} (BASELINE.json configs[3])",
            "value": world * Bu / (ms * 1e-3), "unit": "spectrograms/s", "ms_per_step": ms, "dtype": "bf16",
            "gflop_per_spectrogram": UNET_GFLOP, "finite": finite,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained,
                         "peak_source": "measured sustained bf16" if sustained_src == "measured" else "fallback", "traffic": None}}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, threads = _cpu_unet_rate(3)
            extras["unet"]["cpu_baseline"] = {"value": rate, "unit": "spectrograms/s", "cores": threads, "kind": "port",
                                              "sample": "3 forwards of the fp32 torch oracle UNet, 1x1x257x251"}
    if rank == 0:
        line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()



# ------------------------------------------------------------------ other BASELINE configs (device-timed)
def _cpu_chain_worker(args):
    """configs[0]/[2] on the CPU: oracle AugmentFP chain + wavfile2hashes on one core."""
    seed, n, shifts = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    import torch

    torch.set_num_threads(1)
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O
    from oracle import augment_np as A

    x = synth.music_like(n, seed=seed).numpy()
    ir = synth.impulse_responses(n, seed=seed + 1).numpy()
    noise = synth.rms_noise(n, seed=seed + 2).numpy()
    pr = synth.augment_params(n, seed=seed + 3)
    t0 = time.perf_counter()
    for i in range(n):
        prm = {"fc1": float(pr["fc1"][i]), "ir": ir[i], "noise": noise[i], "snr_db": float(pr["snr_db"][i]),
               "gain_factor": float(np.float32(10.0) ** (pr["gain_db"][i] / np.float32(20.0))),
               "clip_p": float(pr["clip_p"][i]), "fc2": float(pr["fc2"][i]), "fc3": float(pr["fc3"][i])}
        y = A.augment_chain(x[i], prm)
        O.wave2hashes(np.asarray(y, dtype=np.float32), shifts)
    return time.perf_counter() - t0, n


def cpu_chain_rate(per_core: int, shifts: int, cores: int | None = None):
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_chain_worker, [(1, 1, shifts)] * cores)
        res = pool.map(_cpu_chain_worker, [(2000 + 10 * i, per_core, shifts) for i in range(cores)])
    return sum(r[1] for r in res) / max(r[0] for r in res), cores, sum(r[1] for r in res)


def aug_param_array(lib, n, seed, ir_len):
    import numpy as np

    from musicfpaugment_b200 import synth

    pr = synth.augment_params(n, seed=seed)
    arr = np.zeros(n, dtype=lib.AUG_DTYPE)
    arr["apply"] = lib.AUG_ALL
    arr["fc1_hz"], arr["fc2_hz"], arr["fc3_hz"] = pr["fc1"], pr["fc2"], pr["fc3"]
    arr["snr_db"], arr["clip_p"] = pr["snr_db"], pr["clip_p"]
    arr["gain_factor"] = np.float32(10.0) ** (pr["gain_db"] / np.float32(20.0))
    arr["ir_len"] = ir_len
    return arr


def bench_full_chain(ctx, lib, dev, rank, B, shifts, steps, warmup, barrier):
    """BASELINE configs[2]: AugmentFP chain (1 s IR, noise at random SNR, filters, clipping) fused with
    STFT + peaks + hashes; x, noise and IR resident in HBM, hashes left in HBM."""
    import torch

    from musicfpaugment_b200 import synth

    x = synth.music_like(B, seed=1234 + rank, device=dev, chunk=32)
    ir = synth.impulse_responses(B, seed=2000 + rank, device=dev)
    noise = synth.rms_noise(B, seed=3000 + rank, device=dev)
    prm = aug_param_array(lib, B, 4000 + rank, ir.shape[1])
    p = lib.afp_defaults()
    hashes = nh = None
    for _ in range(warmup):
        hashes, nh = ctx.augment_fingerprint(x, prm, ir, noise, shifts, p)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        hashes, nh = ctx.augment_fingerprint(x, prm, ir, noise, shifts, p)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    return ms, int(nh.sum().item())


def bench_match(ctx, lib, dev, rank, world, B, n_tracks, steps, warmup, barrier):
    """BASELINE configs[4]: match B planted 400-hash queries against a synthetic n_tracks-track index
    sharded by hash range over the ranks; per-track histograms summed with an NCCL reduce-scatter; at N > 1
    the replicated-index mode is timed as well."""
    import torch

    from musicfpaugment_b200 import sharded, synth

    lo, hi = sharded.hash_range(rank, world)
    table, counts, hpid, tt, th = synth.hash_index_device(n_tracks, 1000, seed=5000, device=dev, hash_lo=lo, hash_hi=hi)
    ctx.index_load(table.cpu().numpy().view("uint32"), counts.cpu().numpy(), hpid.cpu().numpy().astype("uint32"), hash_lo=lo)
    del table
    q, nq, truth = synth.planted_queries_device(tt, th, B, n_hashes=400, frac=0.3, seed=6000)
    del tt, th
    torch.cuda.empty_cache()
    mp = lib.match_defaults()
    if world > 1:
        ctx.set_option(lib.OPT_MATCH_PACKED, 1)  # 16-bit counter pairs: half the bytes through the reduce-scatter

    def step():
        if world == 1:
            return ctx.match(q, nq, mp, max_rows=4)
        return sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=int(os.environ.get("MFPA_MATCH_SUB", "2048")))

    for _ in range(warmup):
        res, nrows = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res, nrows = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    top1 = float(((nrows > 0) & (res[:, 0, 0] == truth)).float().mean().item())
    rep = None
    if world > 1:
        # throughput mode: the whole index on every rank, queries sharded, rows all-gathered
        ctx.set_option(lib.OPT_MATCH_PACKED, 0)
        table, counts, hpid, _, _ = synth.hash_index_device(n_tracks, 1000, seed=5000, device=dev)
        ctx.index_load(table.cpu().numpy().view("uint32"), counts.cpu().numpy(), hpid.cpu().numpy().astype("uint32"))
        del table
        torch.cuda.empty_cache()
        for _ in range(warmup):
            res2, nrows2 = sharded.match_replicated(ctx, q, nq, mp, max_rows=4)
        barrier()
        e0.record()
        for _ in range(steps):
            res2, nrows2 = sharded.match_replicated(ctx, q, nq, mp, max_rows=4)
        e1.record()
        barrier()
        same = bool(torch.equal(nrows2, nrows) and torch.equal(res2[:, 0, :4], res[:, 0, :4]))
        rep = (e0.elapsed_time(e1) / steps, same)
    return ms, top1, int(nq.sum().item()), rep


UNET_GFLOP = 93.40  # per 257x251 spectrogram, SURVEY.md App. A.8


def _cpu_unet_rate(n_images: int):
    """torch fp32 forward of the oracle UNet on all host cores (the reference's CPU path for a19)."""
    import torch

    from oracle.unet_torch import seeded_unet

    net = seeded_unet(0)
    x = torch.rand(1, 1, 257, 251)
    with torch.no_grad():
        net(x)
        t0 = time.perf_counter()
        for _ in range(n_images):
            net(x)
    return n_images / (time.perf_counter() - t0), torch.get_num_threads()


def bench_unet(ctx, lib, dev, rank, B, steps, warmup, barrier, chunk):
    """BASELINE configs[3]: UNet denoiser forward on B normalised 257x251 magnitude spectrograms
    (random init, bf16, tcgen05 implicit GEMM), in place between the STFT and the picker."""
    import torch

    from musicfpaugment_b200 import synth

    den = lib.UNetDenoiser(ctx, max_chunk=chunk)
    den.load(synth.unet_random_params(0))
    x = synth.music_like(B, seed=1234 + rank, device=dev, chunk=32)
    mag, qmax = ctx.stft_mag(x, 1)
    ref = mag.clone()
    for _ in range(warmup):
        mag.copy_(ref)
        den.denoise_mag(mag, qmax)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        mag.copy_(ref)
        e0.record()
        den.denoise_mag(mag, qmax)
        e1.record()
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps
    finite = bool(torch.isfinite(mag[:, :, :257]).all().item())
    den.close()
    return ms, finite


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--shifts", type=int, default=1)
    ap.add_argument("--cpu-queries-per-core", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default="chain,match,unet",
                    help="comma list of the other BASELINE configs to time after the headline: chain, match, unet, or none")
    ap.add_argument("--unet-queries", type=int, default=148, help="spectrograms per step of the UNet leg")
    ap.add_argument("--unet-chunk", type=int, default=37, help="images per pass through the UNet (activation arena size)")
    ap.add_argument("--tracks", type=int, default=100000, help="tracks in the synthetic index of the match workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
