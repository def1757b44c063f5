"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REAL
reference (/root/reference, via oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container:  python -m oracle.make_golden
The reference ships no tests or golden vectors; these files are
the pin for both the numpy oracle and the CUDA path.  Library versions used to
produce them are recorded in each file's `meta` entry ("reference code on this
container's numpy/scipy/torch").
"""
from __future__ import annotations

import json
import os
import pickle
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _meta():
    import scipy
    import torch

    return json.dumps({"numpy": np.__version__, "scipy": scipy.__version__, "torch": torch.__version__,
                       "python": sys.version.split()[0], "generator": "oracle/make_golden.py"})


def _wavfile2hashes(ns, analyzer, x):
    """Call the reference's wavfile2hashes on a temporary .pkl query
    (the on-disk query format of testing/generate_queries.py:88-90)."""
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump(np.asarray(x, dtype=np.float32), f)
        name = f.name
    try:
        return analyzer.wavfile2hashes(name)
    finally:
        os.unlink(name)


def make_audfprint(ns):
    from musicfpaugment_b200 import synth

    music = synth.music_like(4).numpy()
    noise = synth.white_noise(2).numpy()
    cases = [music[0], music[1], music[2], noise[0], music[3][:12345], noise[1][:2000],
             (music[0] * 0.25 + 0.02 * noise[1])[:32000], music[1][:700]]
    settings = dict(ns.parameters.afp_settings["audfprint"])
    out = {"meta": _meta(), "n_cases": len(cases)}
    for i, x in enumerate(cases):
        x = np.ascontiguousarray(x, dtype=np.float32)
        an = ns.peak_extractor.Audfprint_peaks(settings)
        pk, mask, spec = an.find_peaks(x)
        lm = an.peaks2landmarks(pk)
        h = ns.peak_extractor.landmarks2hashes(lm)
        an1 = ns.peak_extractor.Audfprint_peaks(dict(settings, shifts=1))
        an4 = ns.peak_extractor.Audfprint_peaks(dict(settings, shifts=4))
        out[f"x{i}"] = x
        out[f"peaks{i}"] = np.asarray(pk, dtype=np.int32).reshape(-1, 2)
        out[f"landmarks{i}"] = np.asarray(lm, dtype=np.int32).reshape(-1, 4)
        out[f"hashes{i}"] = h
        try:
            out[f"wf2h_s1_{i}"] = _wavfile2hashes(ns, an1, x)
            out[f"wf2h_s4_{i}"] = _wavfile2hashes(ns, an4, x)
        except ValueError:  # np.hstack([]) on zero peaks
            out[f"wf2h_s1_{i}"] = np.zeros((0, 2), np.int32)
            out[f"wf2h_s4_{i}"] = np.zeros((0, 2), np.int32)
        if i in (4, 5):  # short cases: keep the float64 spectrogram and filtered sgram too
            out[f"spec{i}"] = spec
            import scipy.signal

            s = np.log(np.maximum(spec, spec.max() / 1e6))
            s = s - s.mean()
            out[f"sgram{i}"] = np.array([scipy.signal.lfilter([1, -1], [1, -0.98], r) for r in s])[:-1]
            out[f"mask{i}"] = mask
    np.savez_compressed(os.path.join(GOLD, "audfprint.npz"), **out)
    print("audfprint.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("peaks", "wf2h"))})


def make_match(ns):
    """Small index built by the reference HashTable.store (with bucket overflow and its
    random overwrite, seeded), queried through the reference Matcher.match_hashes."""
    r = np.random.default_rng(77)
    n_tracks, per_track, hot = 400, 600, 64
    ht = ns.hash_table.HashTable()
    ht.depth = 100
    random.seed(1234)
    tracks = []
    for t in range(n_tracks):
        # 15 % of hashes fall on 64 "hot" buckets so several buckets exceed depth 100
        h = np.where(r.random(per_track) < 0.15, r.integers(0, hot, per_track) * 1021 + 7,
                     r.integers(0, 1 << 20, per_track))
        tm = np.sort(r.integers(0, 900, per_track))
        rows = np.stack([tm, h], axis=1).astype(np.int32)
        tracks.append(rows)
        ht.store(f"track{t:04d}", rows)
    m = ns.match.Matcher()
    queries, results, hits_n = [], [], []
    for q in range(24):
        tid = int(r.integers(0, n_tracks))
        rows = tracks[tid]
        off = int(r.integers(0, 500))
        sel = rows[(rows[:, 0] >= off) & (rows[:, 0] < off + 251)]
        keep = sel[r.random(len(sel)) < (0.5 if q % 3 else 0.05)]
        noise_rows = np.stack([r.integers(0, 251, 120), r.integers(0, 1 << 20, 120)], axis=1)
        qrows = np.concatenate([np.stack([keep[:, 0] - off, keep[:, 1]], axis=1), noise_rows]).astype(np.int64)
        key = np.unique((qrows[:, 0] << 32) + qrows[:, 1])
        qh = np.stack([key >> 32, key & 0xFFFFFFFF], axis=1).astype(np.int32)
        res, _ = m.match_hashes(ht, qh)
        queries.append(qh)
        results.append(np.asarray(res, dtype=np.int32).reshape(-1, 7))
        hits_n.append(len(ht.get_hits(qh)))
    out = {"meta": _meta(), "n_tracks": n_tracks, "n_queries": len(queries),
           "counts_nonzero_idx": np.nonzero(ht.counts)[0].astype(np.int32),
           "counts_nonzero": ht.counts[np.nonzero(ht.counts)[0]],
           "hashesperid": ht.hashesperid, "hits_n": np.asarray(hits_n, np.int32)}
    nz = np.nonzero(ht.counts)[0]
    out["table_rows"] = ht.table[nz]  # only the non-empty buckets (the full table is 419 MB)
    for i, (qh, res) in enumerate(zip(queries, results)):
        out[f"q{i}"] = qh
        out[f"res{i}"] = res
    out["hits0"] = ht.get_hits(queries[0])
    np.savez_compressed(os.path.join(GOLD, "match.npz"), **out)
    print("match.npz: buckets", len(nz), "max count", int(ht.counts.max()), "results", [len(x) for x in results])


def make_dejavu(ns):
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O
    from oracle import dejavu_np as D

    x = synth.music_like(2, seed=99).numpy()
    out = {"meta": _meta()}
    cases = []
    for i in range(2):
        spec = O.normalise(O.stft_mag(x[i])[:, :120] ** 2)
        cases.append(D.log_spectrogram(spec))
    q = np.round(cases[0] / 8.0) * 8.0  # plateaus: many exact ties
    q[40:90, 30:70] = 0.0               # a zero "background" block (erosion XOR branch)
    cases.append(q)
    cases.append(cases[1][:23, :17])    # smaller than the 21 x 21 footprint in one dim
    for i, a in enumerate(cases):
        pk, mask = ns.dejavu_fingerprint.get_2D_peaks(a.copy(), plot=False, amp_min=50 if i != 3 else 5)
        out[f"arr{i}"] = a
        out[f"peaks{i}"] = np.asarray([(int(f), int(t)) for f, t in pk], dtype=np.int32).reshape(-1, 2)
        out[f"mask{i}"] = mask.astype(np.uint8)
        out[f"amp_min{i}"] = 50 if i != 3 else 5
    np.savez_compressed(os.path.join(GOLD, "dejavu.npz"), **out)
    print("dejavu.npz:", [len(out[f"peaks{i}"]) for i in range(len(cases))])


def main(which=None):
    from oracle import ref_loader

    ns = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    todo = {"audfprint": make_audfprint, "match": make_match, "dejavu": make_dejavu}
    try:
        from oracle.make_golden_augment import make_augment

        todo["augment"] = make_augment
    except ImportError:
        pass
    for name, fn in todo.items():
        if which and name not in which:
            continue
        fn(ns)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
