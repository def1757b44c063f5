#!/usr/bin/env python
"""Benchmark of the query-side hot path (BASELINE.json metric: 8 s-query fingerprints/s, aug + STFT + peaks + hash).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--queries 10000]

One "step" = one pass of the hot path over one batch of synthetic 8 s / 8 kHz queries: the FULL AugmentFP
degradation chain (loudspeaker high-pass, 1 s impulse-response convolution, background noise at a random
SNR, gain, clipping, low-pass, microphone high-pass) -> magnitude STFT (n_fft 512, hop 256) -> audfprint
peak picking -> 20-bit landmark hashes, 10 k queries per GPU (BASELINE.json configs[2], the configuration
the metric is quoted on).  N > 1 (torchrun, one rank per GPU) shards queries with no data-path collective:
every rank runs its own 10 k batch (weak scaling).

* value    device-timed queries/s: queries, impulse responses, noise rows and parameters resident in HBM
           (CUDA events on the launching stream, max over ranks)
* e2e      the same batch through the C-ABI host entry point mfpa_augment_fingerprint_host: queries in PINNED
           HOST memory in, CSR hash rows in host memory out, copies inside the timed region (impulse
           responses and the noise bank are device-resident sources, like an index; the noise rows are
           assembled from bank + per-query pieces inside the timed region)
* roofline dominant kernel (live CUDA-event stage times from the library, MFPA_OPT_STAGE_TIMES): algorithmic
           bytes / duration vs MEASURED_PEAKS.json
* parity   hash agreement of the GPU rows with the oracle rows over the whole cpu_baseline sample
* cpu_baseline  the numpy oracle (port of the reference CPU path: augment chain + wavfile2hashes) on the host
           cores, on the first queries of the same batch, rank 0 / N = 1 only
Extra keys time the other BASELINE configs (STFT+peaks+hashes only, matching, UNet).
--impl reference times the CPU path alone on the same config (rank 0; other ranks exit).
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
IR_LEN = 8000                                      # 1 s synthetic impulse response (configs[2])
MIN_FC1_HZ = 20.0                                  # loudspeaker cut-off clamp of the throughput run
# algorithmic bytes per query of each chain stage (input read once, output written once)
STAGE_BYTES = {
    "hpf1_filter": 65_536, "hpf1_conv": 512_000, "ir_filter": 4 * IR_LEN + 65_536, "ir_conv": 512_000 + 4 * IR_LEN,
    "mix": 768_000, "clip_lpf": 512_000, "hpf3_filter": 65_536, "hpf3_conv": 512_000, "stft": BYTES_STFT,
    "peaks": BYTES_PEAKS, "landmarks": 8_000,
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 10 000 queries from the `ncu --set full` capture of one step
# of the chain (profiles/r02z_chain_summary.txt)
NCU_TRAFFIC_10K = {"hpf1_filter": 0.000768e9 + 0.596881e9, "hpf1_conv": 3.216819e9 + 2.534520e9,
                   "ir_filter": 0.320955e9 + 0.600439e9, "ir_conv": 3.217038e9 + 2.526299e9,
                   "mix": (1.312196e9 + 0.004278e9) + (5.122590e9 + 2.662157e9) + (0.133051e9 + 0.003880e9),   # sample + mix + finish
                   "clip_lpf": 2.561486e9 + 2.518094e9, "hpf3_filter": 0.000876e9 + 0.596374e9,
                   "hpf3_conv": 3.216898e9 + 2.531945e9, "stft": 2.693002e9 + 2.607849e9,
                   "peaks": 2.890073e9 + 0.020177e9, "landmarks": 0.020131e9 + 0.000042e9}
NCU_TRAFFIC_SOURCE = "profiles/r02z_chain_summary.txt (ncu --set full of one step of the chain, dram read + write)"


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


# ------------------------------------------------------------------ CPU arm (oracle port of the reference path)
def _rms_normalize(v):
    """augmentation/utils.py:189-205 (float32): x / (rms + 1e-8)."""
    import numpy as np

    rms = np.sqrt(np.mean(np.square(v, dtype=np.float32), dtype=np.float32))
    return (v / (rms + np.float32(1e-8))).astype(np.float32)


def _chain_prm(pr, i, ir, noise):
    import numpy as np

    return {"fc1": float(pr["fc1"][i]), "ir": ir, "noise": noise, "snr_db": float(pr["snr_db"][i]),
            "gain_factor": float(np.float32(10.0) ** (pr["gain_db"][i] / np.float32(20.0))),
            "clip_p": float(pr["clip_p"][i]), "fc2": float(pr["fc2"][i]), "fc3": float(pr["fc3"][i])}


def _cpu_chain_worker(args):
    """One core: AugmentFP chain (oracle/augment_np.augment_chain) + wavfile2hashes (oracle/audfprint_np.wave2hashes)
    per query.  Inputs either come from .npy files written by the GPU arm (the first queries of its batch: the
    parity sample) or are generated here (reference arm: 8 distinct signals per worker, fresh parameters per query)."""
    files, lo, hi, seed, shifts, want_rows = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    import torch

    torch.set_num_threads(1)
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O
    from oracle import augment_np as A

    n = hi - lo
    if files:
        x = np.load(files["x"], mmap_mode="r")
        ir = np.load(files["ir"], mmap_mode="r")
        bank = np.load(files["bank"], mmap_mode="r")
        src = np.load(files["src"])
        pr = {k: np.load(files[k]) for k in ("fc1", "snr_db", "gain_db", "clip_p", "fc2", "fc3")}
        get = lambda i: (np.asarray(x[i]), np.asarray(ir[i]), np.asarray(bank[src[i]: src[i] + x.shape[1]]), i)
    else:
        k = min(n, 8)
        xs = synth.music_like(k, seed=seed).numpy()
        irs = synth.impulse_responses(k, length=IR_LEN, seed=seed + 1).numpy()
        nz = np.random.default_rng(seed + 2).standard_normal((k, xs.shape[1])).astype(np.float32)
        pr = synth.augment_params(n, seed=seed + 3, min_fc1_hz=MIN_FC1_HZ)
        get = lambda i: (xs[(i - lo) % k], irs[(i - lo) % k], nz[(i - lo) % k], i - lo)
    rows = []
    t0 = time.perf_counter()
    for i in range(lo, hi):
        xi, iri, nzi, pi = get(i)
        noise = _rms_normalize(_rms_normalize(nzi))   # random_background: the piece, then the row (background_noise.py:64-141)
        y = A.augment_chain(xi, _chain_prm(pr, pi, iri, noise))
        h = O.wave2hashes(np.asarray(y, dtype=np.float32), shifts)
        if want_rows:
            rows.append(np.asarray(h, dtype=np.int32).reshape(-1, 2))
    return time.perf_counter() - t0, n, rows


def cpu_chain_rate(per_core: int, shifts: int, cores: int | None = None, files=None, n_total: int | None = None):
    """queries/s of the oracle chain + fingerprint on all host cores (one process per core, 1 thread each).
    With `files`: queries [0, n_total) of the GPU arm's batch, hash rows returned for the parity check."""
    import multiprocessing as mp

    cores = cores or os.cpu_count() or 1
    n_total = n_total if n_total is not None else per_core * cores
    per = -(-n_total // cores)
    jobs = [(files, min(n_total, i * per), min(n_total, (i + 1) * per), 2000 + 10 * i, shifts, files is not None) for i in range(cores)]
    jobs = [j for j in jobs if j[2] > j[1]]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        pool.map(_cpu_chain_worker, [(None, 0, 1, 1, shifts, False)] * len(jobs))   # warm the workers (imports)
        t0 = time.perf_counter()
        res = pool.map(_cpu_chain_worker, jobs)
        wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    rows = [h for r in res for h in r[2]]
    return n / max(r[0] for r in res), len(jobs), n, wall, rows


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
    """queries/s of the oracle (port of afp/audfprint wavfile2hashes alone, BASELINE configs[1]) using all host cores."""
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


def config_of(args, world):
    """The workload description both arms print (BASELINE.json configs[2])."""
    return {
        "workload": (f"{args.queries} synthetic 8 s 8 kHz mono queries per GPU through the full AugmentFP chain (loudspeaker "
                     "high-pass, 1 s synthetic impulse-response convolution, noise mix at random SNR, gain, clipping, low-pass, "
                     "microphone high-pass) fused with batched STFT (n_fft 512, hop 256) + audfprint peak picking + 20-bit "
                     f"landmark hashes, shifts={args.shifts} (BASELINE.json configs[2], the metric's configuration)"),
        "queries_per_gpu": args.queries, "n_samples": T_QUERY, "shifts": args.shifts, "ir_len": IR_LEN,
        "min_loudspeaker_cutoff_hz": MIN_FC1_HZ, "augment_parameters": "default_parameters, every transform applied (p = 1)",
        "l2_policy": "inputs larger than L2 (2.56 GB queries + 2.56 GB noise rows + 0.32 GB responses per step)",
        "parallelism": f"query-sharded x{world}, no collective",
    }


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = max(2, args.cpu_queries_per_core)
    rates = []
    for i in range(args.warmup + args.steps):
        rate, used, n, wall, _ = cpu_chain_rate(per_core, args.shifts, cores)
        if i >= args.warmup:
            rates.append((rate, n, wall))
    value = sum(r[0] for r in rates) / len(rates)
    n = rates[0][1]
    sample = (f"{n} synthetic queries per step ({per_core} per core, 8 distinct signals per core, fresh chain parameters per query) "
              f"of the {args.queries}-query workload; numpy oracle = port of AugmentFP.__call__ + Audfprint_peaks.wavfile2hashes "
              "(the reference is pure Python and cannot travel to the GPU box; the port convolves long FIRs by FFT where julius "
              "runs a direct conv1d, so it is if anything faster than the reference)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(r[2] for r in rates) / len(rates), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_full_step_extrapolated": 1e3 * args.queries / value,
        "gpu_launches": 0,
    }
    emit(line)


def aug_param_array(lib, n, seed, ir_len):
    import numpy as np

    from musicfpaugment_b200 import synth

    pr = synth.augment_params(n, seed=seed, min_fc1_hz=MIN_FC1_HZ)
    arr = np.zeros(n, dtype=lib.AUG_DTYPE)
    arr["apply"] = lib.AUG_ALL
    arr["fc1_hz"], arr["fc2_hz"], arr["fc3_hz"] = pr["fc1"], pr["fc2"], pr["fc3"]
    arr["snr_db"], arr["clip_p"] = pr["snr_db"], pr["clip_p"]
    arr["gain_factor"] = np.float32(10.0) ** (pr["gain_db"] / np.float32(20.0))
    arr["ir_len"] = ir_len
    return arr, pr


def hash_agreement(got_rows, want_rows):
    """(sum |A & B| / sum |A | B|, fraction of queries whose row sets are identical)."""
    inter = union = same = 0
    for g, w in zip(got_rows, want_rows):
        a = {(int(t), int(h)) for t, h in g}
        b = {(int(t), int(h)) for t, h in w}
        inter += len(a & b)
        union += len(a | b)
        same += a == b
    return (inter / union if union else 1.0), same / max(1, len(want_rows))


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
    C = __import__("ctypes")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs, resident in HBM
    x = synth.music_like(B, seed=1234 + rank, device=dev, chunk=32)
    ir = synth.impulse_responses(B, length=IR_LEN, seed=2000 + rank, device=dev)
    prm, pr = aug_param_array(lib, B, 4000 + rank, IR_LEN)
    # noise: a device-resident bank of background audio; each query draws one T-sample excerpt (random_background's
    # pieces, background_noise.py:64-141), RMS-normalised twice by mfpa_noise_assemble
    g = torch.Generator(device=dev)
    g.manual_seed(3000 + rank)
    bank = torch.randn(args.noise_bank_samples, generator=g, device=dev)
    src = np.random.default_rng(3100 + rank).integers(0, args.noise_bank_samples - T_QUERY, B)
    pieces = np.zeros(B, dtype=lib.NOISE_PIECE_DTYPE)
    pieces["src_a"], pieces["src_b"], pieces["query"], pieces["dst"], pieces["len"] = src, -1, np.arange(B), 0, T_QUERY
    noise = ctx.noise_assemble(bank, pieces, B, T_QUERY)

    hashes = nh = None

    def step():
        nonlocal hashes, nh
        hashes, nh = ctx.augment_fingerprint(x, prm, ir, noise, S, p)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.set_option(lib.OPT_STAGE_TIMES, 1)   # CUDA events at every stage start, on the launching stream
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step()
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    stage_ms, stage_calls = ctx.stage_times()
    ctx.set_option(lib.OPT_STAGE_TIMES, 0)
    tot_hashes = int(nh.sum().item())
    launches_per_step = 13 + (1 if S > 1 else 0)   # 3 x (filter_spectrum + fftconv), clip_sample, mix, clip_finish, clip_lpf, stft, peaks, landmarks

    # ---- end to end through the host C-ABI entry point (pinned host queries in, CSR rows out)
    x_host = torch.empty(B, T_QUERY, dtype=torch.float32).pin_memory()
    x_host.copy_(x)
    rows_host = torch.empty(max(tot_hashes * 2, 1024), 2, dtype=torch.int32).pin_memory()
    offs_host = torch.empty(B + 1, dtype=torch.int64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step(xh):
        return ctx.augment_fingerprint_host(xh, prm, S, p, ir=ir, noise_bank=bank, pieces=pieces, rows=rows_host, offsets=offs_host)

    e2e_step(x_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step(x_host)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    n_rows_e2e = int(offs_host[-1])
    assert n_rows_e2e == tot_hashes, (n_rows_e2e, tot_hashes)
    # the same call on 16-bit PCM queries (what a decoded audio file holds, peak_extractor.py:348-389): half the bytes over PCIe
    x16_host = torch.empty(B, T_QUERY, dtype=torch.int16).pin_memory()
    x16_host.copy_((x * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
    e2e_step(x16_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step(x16_host)
    torch.cuda.synchronize()
    e2e16_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    n_rows16 = int(offs_host[-1])
    # the ceiling of the e2e path on this host: a plain pinned -> device copy of the same bytes, all ranks at once
    barrier()
    t0 = time.perf_counter()
    xd = torch.empty_like(x)
    for _ in range(e2e_steps):
        xd.copy_(x_host, non_blocking=True)
    torch.cuda.synchronize()
    h2d_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    del xd
    clocks = sampler.stop()   # sampled across the timed regions above (device-resident, e2e, e2e PCM16, copy ceiling)

    times = torch.tensor([ms_total, e2e_ms, e2e16_ms, h2d_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e16_ms, h2d_ms = times.tolist()
    line = None
    if rank == 0:
        hbm, how = _peaks()
        ms_step = ms_total / args.steps
        value = world * B / (ms_step * 1e-3)
        run = {k: v for k, v in stage_ms.items() if v > 0}
        dom = max(run, key=run.get)
        dom_bytes = STAGE_BYTES[dom] * B * (S if dom in ("stft", "peaks", "landmarks") else 1)
        achieved = dom_bytes / (run[dom] * 1e-3) / 1e9
        x_bytes = B * T_QUERY * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args, world),
            "hashes_per_query": tot_hashes / B,
            "stage_ms": stage_ms, "stage_ms_source": f"CUDA events inside the timed region (mean of the last {stage_calls} steps)",
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": NCU_TRAFFIC_10K.get(dom, 0) * B / 10000 or None,
                         "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": how,
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0,   # also against the spec-sheet figure
                         "whole_path_frac": (BYTES_CHAIN * B + 8 * tot_hashes) / (ms_step * 1e-3) / 1e9 / hbm,
                         "stage_frac": {k: STAGE_BYTES[k] * B / (v * 1e-3) / 1e9 / hbm for k, v in run.items()}},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": x_bytes + prm.nbytes + pieces.nbytes, "d2h_bytes_per_step": n_rows_e2e * 8 + (B + 1) * 8,
                    "h2d_gbs_per_rank": x_bytes / (e2e_ms * 1e-3) / 1e9,
                    "api": "mfpa_augment_fingerprint_host (pinned host queries, chunked copy/compute overlap; impulse responses and "
                           "noise bank device-resident, noise rows assembled per chunk from per-query pieces)"},
            "e2e_pcm16": {"value": world * B / (e2e16_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16_ms,
                          "h2d_bytes_per_step": x_bytes // 2 + prm.nbytes + pieces.nbytes, "d2h_bytes_per_step": n_rows16 * 8 + (B + 1) * 8,
                          "h2d_gbs_per_rank": x_bytes / 2 / (e2e16_ms * 1e-3) / 1e9,
                          "api": "the same entry with int16 PCM queries (what a decoded audio file holds), converted on the device"},
            "h2d_ceiling": {"ms_per_step": h2d_ms, "gbs_per_rank": x_bytes / (h2d_ms * 1e-3) / 1e9,
                            "what": f"plain pinned->device copy of the {x_bytes / 1e9:.2f} GB of float32 queries, {world} rank(s) at once: "
                                    "the host-side bound of the float32 e2e number"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            # the oracle on the first queries of the SAME batch: timing (cpu_baseline) and hash agreement (parity)
            cores = os.cpu_count() or 1
            n_cpu = min(B, args.cpu_queries_per_core * cores)
            tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            files = {k: os.path.join(tmp, f"mfpa_bench_{os.getpid()}_{k}.npy") for k in
                     ("x", "ir", "bank", "src", "fc1", "snr_db", "gain_db", "clip_p", "fc2", "fc3")}
            try:
                np.save(files["x"], x[:n_cpu].cpu().numpy())
                np.save(files["ir"], ir[:n_cpu].cpu().numpy())
                np.save(files["bank"], bank.cpu().numpy())
                np.save(files["src"], src[:n_cpu])
                for k in ("fc1", "snr_db", "gain_db", "clip_p", "fc2", "fc3"):
                    np.save(files[k], pr[k][:n_cpu])
                rate, used, nq, wall, want = cpu_chain_rate(0, S, cores, files=files, n_total=n_cpu)
            finally:
                for f in files.values():
                    if os.path.exists(f):
                        os.remove(f)
            hh, nn = hashes[:n_cpu].cpu().numpy(), nh[:n_cpu].cpu().numpy()
            agree, same = hash_agreement([hh[i, : nn[i]] for i in range(n_cpu)], want)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
                                    "sample": f"the first {nq} queries of the same batch ({args.cpu_queries_per_core} per core), numpy oracle "
                                              f"of the AugmentFP chain + wavfile2hashes, {wall:.1f} s wall"}
            line["parity"] = {"hash_agreement": agree, "queries_identical": same, "queries": nq,
                              "what": "sum |GPU & oracle| / sum |GPU | oracle| over the (time, hash) rows of the cpu_baseline sample; "
                                      "north_star gate: >= 0.999 end to end"}
    del x_host, rows_host, x16_host, noise, bank
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs (extra keys of the same JSON line)
    extras = {}
    if "fingerprint" in args.also:
        extras["fingerprint_only"] = bench_fingerprint_only(ctx, lib, dev, x, B, S, p, max(2, args.steps // 2), 3, barrier, world)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, cores, nq, wall = cpu_fingerprint_rate(max(8, args.cpu_queries_per_core), S)
            extras["fingerprint_only"]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                                          "sample": f"{nq} synthetic queries, numpy oracle of wavfile2hashes, {wall:.1f} s wall"}
    del x, ir
    torch.cuda.empty_cache()
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
            "collective": "none" if world == 1 else MATCH_COLLECTIVE}
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


MATCH_COLLECTIVE = ("none on the data path: the sweep kernel stores each query's packed (track, time skew) hit words straight into "
                    "its owner rank's memory over NVLink (CUDA IPC peer buffers, mfpa_match_emit_peer), a one-block barrier kernel "
                    "in peer memory separates it from the owner step; NCCL only all-gathers the result rows "
                    "(MFPA_MATCH_EXCHANGE=nccl: one all-to-all per sub-batch instead)")


# ------------------------------------------------------------------ other BASELINE configs (device-timed)
def bench_fingerprint_only(ctx, lib, dev, x, B, S, p, steps, warmup, barrier, world):
    """BASELINE configs[1]: batched STFT + audfprint peak picking + 20-bit landmark hashes on clean queries,
    device-timed with the queries resident in HBM, and end to end through mfpa_fingerprint_host."""
    import torch
    import torch.distributed as dist

    for _ in range(warmup):
        hashes, nh = ctx.fingerprint(x, S, p)
    barrier()
    ctx.set_option(lib.OPT_STAGE_TIMES, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        hashes, nh = ctx.fingerprint(x, S, p)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    stage_ms, _ = ctx.stage_times()
    ctx.set_option(lib.OPT_STAGE_TIMES, 0)
    tot = int(nh.sum().item())
    x_host = torch.empty(B, T_QUERY, dtype=torch.float32).pin_memory()
    x_host.copy_(x)
    rows_host = torch.empty(max(tot * 2, 1024), 2, dtype=torch.int32).pin_memory()
    offs_host = torch.empty(B + 1, dtype=torch.int64).pin_memory()
    ctx.fingerprint_host(x_host, S, p, rows=rows_host, offsets=offs_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.fingerprint_host(x_host, S, p, rows=rows_host, offsets=offs_host)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    hbm, _ = _peaks()
    return {
        "workload": f"{B} clean queries per GPU: batched STFT + audfprint peak picking + 20-bit landmark hashes, shifts={S} "
                    "(BASELINE.json configs[1]; round 1's headline)",
        "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "hashes_per_query": tot / B,
        "stage_ms": {k: v for k, v in stage_ms.items() if v > 0},
        "stft_hbm_frac": BYTES_STFT * B * S / (max(stage_ms["stft"], 1e-9) * 1e-3) / 1e9 / hbm,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": B * T_QUERY * 4,
                "d2h_bytes_per_step": tot * 8 + (B + 1) * 8, "api": "mfpa_fingerprint_host"}}


def bench_match(ctx, lib, dev, rank, world, B, n_tracks, steps, warmup, barrier):
    """BASELINE configs[4]: match B planted 400-hash queries against a synthetic n_tracks-track index
    sharded by hash range over the ranks; the shards' packed (track, time skew) hit words are stored by the sweep kernel
    straight into the query owners' memory over NVLink (sharded.match_sharded, exchange "peer"); at N > 1 the
    replicated-index mode is timed as well."""
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

    def step():
        if world == 1:
            return ctx.match(q, nq, mp, max_rows=4)
        sub = os.environ.get("MFPA_MATCH_SUB")
        return sharded.match_sharded(ctx, q, nq, mp, max_rows=4, sub_batch=int(sub) if sub else None,
                                     exchange=os.environ.get("MFPA_MATCH_EXCHANGE") or None)

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
    if world > 1:
        sharded.release_peer_exchanges(ctx)
    rep = None
    if world > 1:
        # throughput mode: the whole index on every rank, queries sharded, rows all-gathered
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
    ap.add_argument("--cpu-queries-per-core", type=int, default=64,
                    help="queries per host core of the CPU legs (cpu_baseline / parity sample; --impl reference step)")
    ap.add_argument("--noise-bank-samples", type=int, default=1 << 24, help="samples of the device-resident background-noise bank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default="fingerprint,match,unet",
                    help="comma list of the other BASELINE configs to time after the headline: fingerprint, match, unet, or none")
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
