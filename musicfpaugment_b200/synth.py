"""Seeded synthetic inputs for tests and benchmarks.

There is no dataset in the image, so every workload is synthetic:
"music-like" 8 s / 8 kHz mono queries (noise floor + decaying harmonic
notes), exponentially decaying 1 s impulse responses, RMS-normalised noise,
AugmentFP parameter draws, and a synthetic landmark-hash index.  The
generators are torch so the 10 k-query benchmark batches can be produced
directly in HBM; the small CPU test cases use the same code on ``cpu``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SR = 8000
T_QUERY = 64000


def music_like(n_queries: int, n_samples: int = T_QUERY, seed: int = 1234, device="cpu",
               notes: int = 96, harmonics: int = 4, chunk: int = 16) -> torch.Tensor:
    """[n_queries, n_samples] float32, peak-normalised to 1."""
    dev = torch.device(device)
    out = torch.empty(n_queries, n_samples, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = torch.arange(n_samples, device=dev, dtype=torch.float32) / SR
    for lo in range(0, n_queries, chunk):
        b = min(chunk, n_queries - lo)
        x = 0.05 * torch.randn(b, n_samples, generator=g, device=dev)
        f0 = 80 + 1720 * torch.rand(b, notes, 1, generator=g, device=dev)
        onset = torch.rand(b, notes, 1, generator=g, device=dev) * (n_samples / SR)
        amp = 0.2 + 0.8 * torch.rand(b, notes, 1, generator=g, device=dev)
        tau = 0.05 + 0.35 * torch.rand(b, notes, 1, generator=g, device=dev)
        dt = t[None, None, :] - onset
        env = torch.where(dt >= 0, torch.exp(-dt.clamp(min=0) / tau), torch.zeros((), device=dev)) * amp
        for h in range(1, harmonics + 1):
            x += (env * torch.sin((2 * math.pi * h) * f0 * dt) / h).sum(dim=1)
        x /= x.abs().amax(dim=1, keepdim=True)
        out[lo:lo + b] = x
    return out


def white_noise(n_queries: int, n_samples: int = T_QUERY, seed: int = 1235, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    return torch.randn(n_queries, n_samples, generator=g, device=device, dtype=torch.float32)


def impulse_responses(n: int, length: int = SR, seed: int = 2000, device="cpu") -> torch.Tensor:
    """[n, length] float32: N(0,1)·exp(−t/0.15 s)."""
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    t = torch.arange(length, device=device, dtype=torch.float32)
    return torch.randn(n, length, generator=g, device=device) * torch.exp(-t / (0.15 * SR))


def rms_noise(n: int, n_samples: int = T_QUERY, seed: int = 3000, device="cpu") -> torch.Tensor:
    """[n, n_samples] float32 white noise with unit RMS per row
    (what augmentation/utils.py:189-205 `rms_normalize` leaves behind)."""
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    x = torch.randn(n, n_samples, generator=g, device=device)
    return x / torch.sqrt(torch.mean(x * x, dim=1, keepdim=True))


def _mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def _imel(m):
    return 700.0 * (10 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def augment_params(n: int, seed: int = 4000, min_fc1_hz: float = 20.0) -> dict:
    """Per-query AugmentFP draws under `default_parameters`
    (testing/parameters.py:248-267), every transform applied.
    Cut-offs are mel-uniform (pass_filters.py:58-82).  ``min_fc1_hz`` clamps the
    loudspeaker high-pass so its FIR stays ≤ 3201 taps."""
    r = np.random.default_rng(seed)

    def mel_u(lo, hi):
        return _imel(r.uniform(np.ceil(_mel(lo)), np.floor(_mel(hi)), n)).astype(np.float32)

    return {
        "fc1": np.maximum(mel_u(0.0, 150.0), np.float32(min_fc1_hz)),
        "snr_db": r.uniform(-10.0, 10.0, n).astype(np.float32),
        "gain_db": r.uniform(-5.0, 5.0, n).astype(np.float32),
        "clip_p": r.uniform(0.0, 0.01, n).astype(np.float32),
        "fc2": mel_u(3000.0, 3999.0),
        "fc3": mel_u(30.0, 150.0),
    }


def hash_index(n_tracks: int, hashes_per_track: int = 1000, seed: int = 5000, depth: int = 100,
               hashbits: int = 20, maxtimebits: int = 14, t_max: int = 940):
    """Synthetic landmark index in the reference's table layout
    (afp/audfprint/hash_table.py:53-68,82-100): buckets filled in track order,
    first `depth` entries kept (no random overwrite), counts saturating above.
    Returns (table uint32 [2^hashbits, depth], counts int32, hashesperid uint32,
    track_hashes int32 [n_tracks, hashes_per_track, 2])."""
    r = np.random.default_rng(seed)
    nb = 1 << hashbits
    h = r.integers(0, nb, size=(n_tracks, hashes_per_track), dtype=np.int64)
    t = r.integers(0, t_max, size=(n_tracks, hashes_per_track), dtype=np.int64)
    ids = np.repeat(np.arange(n_tracks, dtype=np.int64), hashes_per_track)
    hf, tf = h.ravel(), t.ravel()
    order = np.argsort(hf, kind="stable")  # stable: keeps track order inside a bucket
    hs = hf[order]
    counts = np.bincount(hs, minlength=nb).astype(np.int32)
    start = np.cumsum(counts) - counts
    slot = np.arange(len(hs)) - start[hs]
    keep = slot < depth
    table = np.zeros((nb, depth), dtype=np.uint32)
    val = ((ids[order] + 1) << maxtimebits) + (tf[order] & ((1 << maxtimebits) - 1))
    table[hs[keep], slot[keep]] = val[keep].astype(np.uint32)
    hashesperid = np.full(n_tracks, hashes_per_track, dtype=np.uint32)
    th = np.stack([t, h], axis=2).astype(np.int32)
    return table, counts, hashesperid, th


def planted_queries(track_hashes: np.ndarray, n_queries: int, n_hashes: int = 400, frac: float = 0.3,
                    seed: int = 6000, hashbits: int = 20, t_q_max: int = 251):
    """Query hash lists [n_queries, n_hashes, 2] (time, hash): `frac` of each
    list is copied from track q mod n_tracks at a constant time offset, the
    rest is uniform random.  Rows are unique and sorted by (time, hash) like
    wavfile2hashes output.  Returns (hashes int32, nh int32[n_queries], truth int32)."""
    r = np.random.default_rng(seed)
    n_tracks = track_hashes.shape[0]
    out = np.zeros((n_queries, n_hashes, 2), dtype=np.int32)
    nh = np.zeros(n_queries, dtype=np.int32)
    truth = np.arange(n_queries, dtype=np.int32) % n_tracks
    n_pl = int(frac * n_hashes)
    for q in range(n_queries):
        th = track_hashes[truth[q]]
        off = int(r.integers(0, 600))
        sel = th[(th[:, 0] >= off) & (th[:, 0] < off + t_q_max)]
        sel = sel[r.permutation(len(sel))[:n_pl]]
        rows = np.concatenate([
            np.stack([sel[:, 0] - off, sel[:, 1]], axis=1),
            np.stack([r.integers(0, t_q_max, n_hashes - len(sel)),
                      r.integers(0, 1 << hashbits, n_hashes - len(sel))], axis=1),
        ]).astype(np.int64)
        key = np.unique((rows[:, 0] << 32) + rows[:, 1])
        k = len(key)
        out[q, :k, 0] = key >> 32
        out[q, :k, 1] = key & 0xFFFFFFFF
        nh[q] = k
    return out, nh, truth


def hash_index_device(n_tracks: int, hashes_per_track: int = 1000, seed: int = 5000, depth: int = 100,
                      hashbits: int = 20, maxtimebits: int = 14, t_max: int = 940, device="cuda",
                      hash_lo: int = 0, hash_hi: int | None = None):
    """`hash_index` built with torch on `device` (the 100 k-track benchmark index is 10^8 entries;
    numpy's argsort of that takes most of a minute).  Same layout and fill rule as `hash_index`,
    different random stream.  Only buckets [hash_lo, hash_hi) are materialised (a hash-range shard:
    every rank draws the same stream from the same seed and keeps its own rows).
    Returns (table uint32-as-int32 [hash_hi-hash_lo, depth], counts int32, hashesperid int32,
    track_t int16 [n_tracks, hpt], track_h int32 [n_tracks, hpt])."""
    dev = torch.device(device)
    nb = 1 << hashbits
    hash_hi = nb if hash_hi is None else hash_hi
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = n_tracks * hashes_per_track
    h = torch.randint(0, nb, (n,), generator=g, device=dev, dtype=torch.int32)
    t = torch.randint(0, t_max, (n,), generator=g, device=dev, dtype=torch.int32)
    track_h, track_t = h.view(n_tracks, hashes_per_track), t.view(n_tracks, hashes_per_track).to(torch.int16)
    counts_all = torch.bincount(h, minlength=nb)
    hs, order = torch.sort(h, stable=True)  # stable: track order inside a bucket
    start = torch.cumsum(counts_all, 0) - counts_all
    slot = torch.arange(n, device=dev, dtype=torch.int64) - start[hs.long()]
    keep = (slot < depth) & (hs >= hash_lo) & (hs < hash_hi)
    ids = (order // hashes_per_track)
    val = ((ids + 1) << maxtimebits) + (t[order].long() & ((1 << maxtimebits) - 1))
    table = torch.zeros(hash_hi - hash_lo, depth, dtype=torch.int64, device=dev)
    table[(hs[keep] - hash_lo).long(), slot[keep]] = val[keep]
    table = (table & 0xFFFFFFFF).to(torch.int64)
    table = torch.where(table >= (1 << 31), table - (1 << 32), table).to(torch.int32)  # uint32 bit pattern
    counts = counts_all[hash_lo:hash_hi].to(torch.int32)
    hpid = torch.full((n_tracks,), hashes_per_track, dtype=torch.int32, device=dev)
    return table, counts, hpid, track_t, track_h


def planted_queries_device(track_t, track_h, n_queries: int, n_hashes: int = 400, frac: float = 0.3,
                           seed: int = 6000, hashbits: int = 20, t_q_max: int = 251):
    """`planted_queries` with torch on the tracks' device.  Query q plants ≈frac·n_hashes hashes of track
    q mod n_tracks at one time offset; the rest is uniform random.  Rows unique, sorted by (time, hash).
    Returns (hashes int32 [n_queries, n_hashes, 2], nh int32 [n_queries], truth int32 [n_queries])."""
    dev = track_h.device
    n_tracks, hpt = track_h.shape
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    truth = torch.arange(n_queries, device=dev) % n_tracks
    n_pl = int(frac * n_hashes)
    off = torch.randint(0, 600, (n_queries, 1), generator=g, device=dev)
    tt, hh = track_t[truth].long(), track_h[truth].long()
    inwin = (tt >= off) & (tt < off + t_q_max)
    score = torch.rand(n_queries, hpt, generator=g, device=dev) + inwin.float()  # in-window entries first
    sel = torch.topk(score, n_pl, dim=1).indices
    ok = torch.gather(inwin, 1, sel)
    pt = torch.gather(tt, 1, sel) - off
    ph = torch.gather(hh, 1, sel)
    rt = torch.randint(0, t_q_max, (n_queries, n_hashes), generator=g, device=dev)
    rh = torch.randint(0, 1 << hashbits, (n_queries, n_hashes), generator=g, device=dev)
    rt[:, :n_pl] = torch.where(ok, pt, rt[:, :n_pl])
    rh[:, :n_pl] = torch.where(ok, ph, rh[:, :n_pl])
    key, _ = torch.sort((rt << 32) + rh, dim=1)
    dup = torch.zeros_like(key, dtype=torch.bool)
    dup[:, 1:] = key[:, 1:] == key[:, :-1]
    big = torch.iinfo(torch.int64).max
    key, _ = torch.sort(torch.where(dup, torch.full_like(key, big), key), dim=1)
    nh = (~dup).sum(dim=1).to(torch.int32)
    valid = key != big
    out = torch.stack([torch.where(valid, key >> 32, 0), torch.where(valid, key & 0xFFFFFFFF, 0)], dim=2).to(torch.int32)
    return out.contiguous(), nh, truth.to(torch.int32)


def unet_random_params(seed: int = 0):
    """Random-init parameter vector for ``mfpa_unet_load`` (BASELINE.json configs[3]: "random init"):
    the layout of the reference ``UNet(1, 1)`` state_dict (training/unet.py:75-95) with
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) convolution weights (torch's default bound) and
    lightly perturbed BatchNorm statistics."""
    import numpy as np

    rng = np.random.default_rng(seed)
    widths = (64, 128, 256, 512, 1024)
    parts = []

    def conv(cin, cout, k, bias):
        b = 1.0 / np.sqrt(cin * k * k)
        parts.append(rng.uniform(-b, b, cout * cin * k * k).astype(np.float32))
        if bias:
            parts.append(rng.uniform(-b, b, cout).astype(np.float32))

    def bn(c):
        parts.append(rng.uniform(0.8, 1.2, c).astype(np.float32))          # weight
        parts.append((0.1 * rng.standard_normal(c)).astype(np.float32))    # bias
        parts.append((0.1 * rng.standard_normal(c)).astype(np.float32))    # running_mean
        parts.append(rng.uniform(0.8, 1.2, c).astype(np.float32))          # running_var

    def block(cin, cout):
        conv(cin, cout, 3, False)
        bn(cout)
        conv(cout, cout, 3, False)
        bn(cout)

    block(1, widths[0])
    for i in range(1, 5):
        block(widths[i - 1], widths[i])
    for i in range(4):
        cin, cout = widths[4 - i], widths[3 - i]
        b = 1.0 / np.sqrt(cout * 4)  # ConvTranspose2d fan_in = weight.size(1) * k * k
        parts.append(rng.uniform(-b, b, cin * cout * 4).astype(np.float32))
        parts.append(rng.uniform(-b, b, cout).astype(np.float32))
        block(cin, cout)
    conv(widths[0], 1, 1, True)
    return np.concatenate(parts)
