"""GPU parity: CUDA AugmentFP chain (C ABI) vs the numpy oracle and the golden vectors the
reference's own transform classes produced.  Tolerance: 1e-4 relative (north star)."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_np as A

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


def _lib():
    from musicfpaugment_b200 import lib

    return lib


def _pack(lib, prms, T):
    """list of oracle-style dicts -> (params array, ir tensor, noise tensor)"""
    B = len(prms)
    arr = np.zeros(B, dtype=lib.AUG_DTYPE)
    lmax = max([len(p["ir"]) for p in prms if p.get("ir") is not None] + [1])
    ir = np.zeros((B, lmax), np.float32)
    noise = np.zeros((B, T), np.float32)
    for i, p in enumerate(prms):
        ap = lib.AUG_NORM
        if p.get("fc1") is not None:
            ap |= lib.AUG_HPF1; arr["fc1_hz"][i] = p["fc1"]
        if p.get("ir") is not None:
            ap |= lib.AUG_IR; ir[i, : len(p["ir"])] = p["ir"]; arr["ir_len"][i] = len(p["ir"])
        if p.get("noise") is not None:
            ap |= lib.AUG_NOISE; noise[i] = p["noise"]; arr["snr_db"][i] = p["snr_db"]
        if p.get("gain_factor") is not None:
            ap |= lib.AUG_GAIN; arr["gain_factor"][i] = p["gain_factor"]
        if p.get("clip_p") is not None:
            ap |= lib.AUG_CLIP; arr["clip_p"][i] = p["clip_p"]
        if p.get("fc2") is not None:
            ap |= lib.AUG_LPF; arr["fc2_hz"][i] = p["fc2"]
        if p.get("fc3") is not None:
            ap |= lib.AUG_HPF3; arr["fc3_hz"][i] = p["fc3"]
        arr["apply"][i] = ap
    return arr, torch.from_numpy(ir).cuda(), torch.from_numpy(noise).cuda()


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _golden_cases():
    g = np.load(os.path.join(GOLD, "augment.npz"))
    keys = ("fc1", "ir", "noise", "snr_db", "gain_factor", "clip_p", "fc2", "fc3")
    for i in range(int(g["n_cases"])):
        prm = {k: (g[f"p{i}_{k}"] if g[f"p{i}_{k}"].ndim else float(g[f"p{i}_{k}"])) for k in keys if f"p{i}_{k}" in g}
        yield i, g[f"x{i}"], prm, g[f"out{i}"]


def test_golden_cases(mfpa_ctx):
    lib = _lib()
    for i, x, prm, ref in _golden_cases():
        arr, ir, noise = _pack(lib, [prm], len(x))
        out = mfpa_ctx.augment(torch.from_numpy(x[None]).cuda(), arr, ir, noise).cpu().numpy()[0]
        assert _rel(out, ref) < TOL, (i, _rel(out, ref))


def test_random_chains_vs_oracle(mfpa_ctx):
    """Batch of 12 queries with different transform subsets, filter lengths, IR lengths and levels."""
    from musicfpaugment_b200 import synth

    lib = _lib()
    B, T = 12, 24000
    x = synth.music_like(B, n_samples=T, seed=31).numpy()
    irs = synth.impulse_responses(B, length=8192, seed=32).numpy()
    nz = synth.rms_noise(B, n_samples=T, seed=33).numpy()
    r = np.random.default_rng(34)
    prms = []
    for i in range(B):
        p = {}
        full = i < 4
        if full or r.random() < 0.6:
            p["fc1"] = float(r.uniform(8.0, 150.0))
        if full or r.random() < 0.6:
            p["ir"] = irs[i][: int(r.integers(1, 8193))]
        if full or r.random() < 0.6:
            p["noise"], p["snr_db"] = nz[i], float(r.uniform(-10, 10))
        if full or r.random() < 0.6:
            p["gain_factor"] = float(10 ** (r.uniform(-5, 5) / 20))
        if full or r.random() < 0.6:
            p["clip_p"] = float(r.uniform(0, 0.01))
        if full or r.random() < 0.6:
            p["fc2"] = float(r.uniform(3000, 3999))
        if full or r.random() < 0.6:
            p["fc3"] = float(r.uniform(30, 150))
        prms.append(p)
    prms[5] = {}                                  # nothing but the final normalisation
    prms[6] = {"fc2": 300.0, "clip_p": 0.004}     # long "low-pass" -> FFT route for stage 5
    arr, ir, noise = _pack(lib, prms, T)
    out = mfpa_ctx.augment(torch.from_numpy(x).cuda(), arr, ir, noise).cpu().numpy()
    for i in range(B):
        ref = A.augment_chain(x[i], prms[i])
        assert _rel(out[i], ref) < TOL, (i, _rel(out[i], ref), sorted(prms[i]))


def test_long_filters_and_long_impulse_responses(mfpa_ctx):
    """Filters longer than one overlap-save block (the partitioned path): loudspeaker cut-offs of 5, 2.5 and
    0.7 Hz (12 801, 25 601 and 91 429 taps - the last one longer than the signal), a 2 Hz microphone
    high-pass, impulse responses of 8193, 20 000 and 70 000 samples (longer than the signal), a long
    "low-pass", all mixed in one batch with ordinary queries so both paths run in the same launch."""
    from musicfpaugment_b200 import synth

    lib = _lib()
    B, T = 8, 32000
    x = synth.music_like(B, n_samples=T, seed=51).numpy()
    r = np.random.default_rng(52)   # slowly decaying responses: the tail past 8192 samples matters
    irs = (r.standard_normal((B, 70000)) * (np.exp(-np.arange(70000) / 9000.0) + 0.02)[None]).astype(np.float32)
    nz = synth.rms_noise(B, n_samples=T, seed=53).numpy()
    prms = [
        {"fc1": 5.0, "ir": irs[0][:8193], "noise": nz[0], "snr_db": 3.0, "gain_factor": 1.2, "clip_p": 0.004, "fc2": 3500.0, "fc3": 60.0},
        {"fc1": 2.5, "ir": irs[1][:20000], "fc3": 2.0},
        {"fc1": 0.7, "noise": nz[2], "snr_db": -4.0, "clip_p": 0.009},
        {"fc1": 90.0, "ir": irs[3][:70000], "fc2": 3900.0},
        {"fc1": 40.0, "ir": irs[4][:4000], "fc3": 100.0},            # ordinary query between the long ones
        {"ir": irs[5][:16385]},
        {"fc2": 6.0, "clip_p": 0.002},                                # 10 667-tap "low-pass"
        {},
    ]
    arr, ir, noise = _pack(lib, prms, T)
    out = mfpa_ctx.augment(torch.from_numpy(x).cuda(), arr, ir, noise).cpu().numpy()
    for i in range(B):
        ref = A.augment_chain(x[i], prms[i])
        assert _rel(out[i], ref) < TOL, (i, _rel(out[i], ref), sorted(prms[i]))
    # a 1 MiB scratch budget makes the long queries run one per group: same numbers
    try:
        mfpa_ctx.set_option(lib.OPT_PART_BUDGET_MB, 1)
        out1 = mfpa_ctx.augment(torch.from_numpy(x).cuda(), arr, ir, noise).cpu().numpy()
    finally:
        mfpa_ctx.set_option(lib.OPT_PART_BUDGET_MB, 2048)
    assert _rel(out1, out) < 1e-6   # (block sums arrive through atomics: the last bits are not ordered)


def test_partition_boundaries_and_short_signals(mfpa_ctx):
    """Either side of the fast/partitioned switch (8193 vs 8205 taps; 8192- vs 8193-sample responses) and
    signals much shorter than one partition with filters many times their length."""
    lib = _lib()
    for T, seed in ((20000, 61), (3000, 62), (257, 63)):
        r = np.random.default_rng(seed)
        x = r.standard_normal((6, T)).astype(np.float32) * 0.3
        ir = (r.standard_normal((6, 8193)) * np.exp(-np.arange(8193) / 3000.0)[None]).astype(np.float32)
        prms = [
            {"fc1": 7.8125},                      # half = 4096: 8193 taps, the longest single-block filter
            {"fc1": 7.8},                         # half = 4102: 8205 taps, two partitions
            {"ir": ir[2][:8192]},
            {"ir": ir[3][:8193]},
            {"fc1": 2.0, "ir": ir[4][:8193], "fc3": 3.0, "clip_p": 0.005},
            {"fc3": 0.5},                         # 128 001 taps
        ]
        arr, ird, noise = _pack(lib, prms, T)
        out = mfpa_ctx.augment(torch.from_numpy(x).cuda(), arr, ird, noise).cpu().numpy()
        for i in range(6):
            ref = A.augment_chain(x[i], prms[i])
            assert _rel(out[i], ref) < TOL, (T, i, _rel(out[i], ref), sorted(prms[i]))


def test_too_long_filter_and_bad_cutoffs_are_rejected(mfpa_ctx):
    lib = _lib()
    x = torch.zeros(1, 8000, device="cuda")
    for fc in (0.05, 0.0, -5.0, 4100.0):
        arr = np.zeros(1, dtype=lib.AUG_DTYPE)
        arr["apply"], arr["fc1_hz"] = lib.AUG_HPF1 | lib.AUG_NORM, fc
        with pytest.raises(lib.MfpaError):
            mfpa_ctx.augment(x, arr, None, None)


def test_augment_then_fingerprint_matches_two_step(mfpa_ctx):
    """Fused S1-S4 == augment() followed by fingerprint() (peak normalisation is irrelevant to hashes)."""
    from musicfpaugment_b200 import synth

    lib = _lib()
    B, T = 6, 64000
    x = synth.music_like(B, seed=41)
    irs = synth.impulse_responses(B, seed=42)
    nz = synth.rms_noise(B, seed=43)
    pr = synth.augment_params(B, seed=44)
    arr = np.zeros(B, dtype=lib.AUG_DTYPE)
    arr["apply"] = lib.AUG_ALL
    arr["fc1_hz"], arr["fc2_hz"], arr["fc3_hz"] = pr["fc1"], pr["fc2"], pr["fc3"]
    arr["snr_db"], arr["gain_factor"], arr["clip_p"], arr["ir_len"] = pr["snr_db"], 10 ** (pr["gain_db"] / 20), pr["clip_p"], 8000
    afp = lib.afp_defaults()
    y = mfpa_ctx.augment(x.cuda(), arr, irs.cuda(), nz.cuda())
    h2, n2 = mfpa_ctx.fingerprint(y, 1, afp)
    h1, n1 = mfpa_ctx.augment_fingerprint(x.cuda(), arr, irs.cuda(), nz.cuda(), 1, afp)
    agree = total = 0
    for i in range(B):
        a = {tuple(r) for r in h1[i, : int(n1[i])].cpu().numpy().tolist()}
        b = {tuple(r) for r in h2[i, : int(n2[i])].cpu().numpy().tolist()}
        agree += len(a & b); total += len(a | b)
    assert total > 0 and agree / total >= 0.999, (agree, total)
    # and the degraded waveform itself matches the oracle chain
    for i in range(2):
        prm = dict(fc1=float(pr["fc1"][i]), ir=irs[i].numpy(), noise=nz[i].numpy(), snr_db=float(pr["snr_db"][i]),
                   gain_factor=float(arr["gain_factor"][i]), clip_p=float(pr["clip_p"][i]), fc2=float(pr["fc2"][i]),
                   fc3=float(pr["fc3"][i]))
        ref = A.augment_chain(x[i].numpy(), prm)
        assert _rel(y[i].cpu().numpy(), ref) < TOL


def test_chain_end_to_end_hash_agreement(mfpa_ctx):
    """North star: at least 99.9 % hash agreement END TO END - the fused GPU path (AugmentFP chain -> STFT ->
    picker -> landmarks -> hashes, float32) against the oracle's chain followed by the oracle's float64
    wavfile2hashes, on the same inputs and dumped parameters (measured: 5646 / 5646 on 48 queries)."""
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O

    lib = _lib()
    B = 24
    x = synth.music_like(B, seed=141)
    irs = synth.impulse_responses(B, seed=142)
    nz = synth.rms_noise(B, seed=143)
    pr = synth.augment_params(B, seed=144)
    arr = np.zeros(B, dtype=lib.AUG_DTYPE)
    arr["apply"] = lib.AUG_ALL
    arr["fc1_hz"], arr["fc2_hz"], arr["fc3_hz"] = pr["fc1"], pr["fc2"], pr["fc3"]
    arr["snr_db"], arr["gain_factor"], arr["clip_p"], arr["ir_len"] = pr["snr_db"], 10 ** (pr["gain_db"] / 20), pr["clip_p"], 8000
    h, n = mfpa_ctx.augment_fingerprint(x.cuda(), arr, irs.cuda(), nz.cuda(), 1, lib.afp_defaults())
    agree = total = 0
    for i in range(B):
        prm = dict(fc1=float(pr["fc1"][i]), ir=irs[i].numpy(), noise=nz[i].numpy(), snr_db=float(pr["snr_db"][i]),
                   gain_factor=float(arr["gain_factor"][i]), clip_p=float(pr["clip_p"][i]), fc2=float(pr["fc2"][i]),
                   fc3=float(pr["fc3"][i]))
        want = {tuple(r) for r in O.wave2hashes(A.augment_chain(x[i].numpy(), prm), 1).tolist()}
        got = {tuple(r) for r in h[i, : int(n[i])].cpu().numpy().tolist()}
        agree += len(want & got)
        total += len(want | got)
    assert total > 1000 and agree / total >= 0.999, (agree, total)



def test_pooled_clipping_of_a_batch_equals_the_reference(mfpa_ctx):
    """B > 1: torch.quantile(samples[:, 0, :], q_vector) flattens the selected sub-batch, so each row is clipped to the
    quantiles of its own percentile over the POOLED rows (clipping.py:76-90).  MFPA_OPT_CLIP_POOLED reproduces it; the
    golden is the reference's Gain + Clipping on 4 rows, one of them left out by the gate.  Without the option the
    per-query rule (what B = 1 gives) runs - and differs."""
    lib = _lib()
    g = np.load(os.path.join(GOLD, "clip_pooled.npz"))
    x = torch.from_numpy(g["x"]).cuda()
    arr = np.zeros(4, dtype=lib.AUG_DTYPE)
    arr["apply"] = lib.AUG_GAIN
    arr["apply"][g["selected"]] |= lib.AUG_CLIP
    arr["gain_factor"], arr["clip_p"] = g["gain_factor"], g["clip_p"]
    mfpa_ctx.set_option(lib.OPT_CLIP_POOLED, 1)
    try:
        pooled = mfpa_ctx.augment(x, arr).cpu().numpy()
    finally:
        mfpa_ctx.set_option(lib.OPT_CLIP_POOLED, 0)
    per_row = mfpa_ctx.augment(x, arr).cpu().numpy()
    for i in range(4):
        assert _rel(pooled[i], g["out"][i]) <= 5e-6, i
    assert _rel(per_row[0], g["out"][0]) > 1e-3        # row 0 is the loud one: pooled thresholds clip it harder
    assert np.array_equal(per_row[2], pooled[2])       # the row Clipping skipped is untouched by either rule
