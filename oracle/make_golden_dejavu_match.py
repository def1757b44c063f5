"""TEST INFRASTRUCTURE ONLY — tests/golden/dejavu_match.npz and tests/golden/metrics.npz from the REAL reference.

Run in the build container:  python -m oracle.make_golden_dejavu_match

* dejavu_match.npz: the reference's own `CommonDatabase.return_matches` (afp/dejavu/postgres_database.py:180-229)
  and `Dejavu.align_matches` (afp/dejavu/dejavu.py:312-378) executed on a synthetic fingerprints table.  Postgres
  is replaced by a dict-backed cursor that answers the one query return_matches issues (SELECT hash, song_id,
  offset WHERE hash IN (...)) and `get_song_by_id`; psycopg2 (absent) is stubbed at import.
* metrics.npz: the reference's `Recall`, `Precision`, `F1score` (testing/metrics.py:10-192) on random peak masks;
  torchmetrics (absent) is stubbed at import, so `psnr` is NOT pinned by it.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def synthetic_table(n_songs=40, per_song=300, seed=7):
    """rows (hash, song id, offset): SHA-1[:20] of small (f1, f2, dt) triples like generate_hashes makes, from a value
    range small enough that many songs share hashes."""
    r = np.random.default_rng(seed)
    rows = []
    for sid in range(1, n_songs + 1):
        f1, f2, dt = r.integers(0, 40, per_song), r.integers(0, 40, per_song), r.integers(0, 12, per_song)
        off = np.sort(r.integers(0, 900, per_song))
        for a, b, c, o in zip(f1, f2, dt, off):
            h = hashlib.sha1(f"{a}|{b}|{c}".encode("utf-8")).hexdigest()[:20]
            rows.append((h, sid, int(o)))
    return rows


def synthetic_queries(rows, n_queries=12, seed=8):
    r = np.random.default_rng(seed)
    by_song = {}
    for h, sid, off in rows:
        by_song.setdefault(sid, []).append((h, off))
    queries = []
    for q in range(n_queries):
        sid = int(r.integers(1, len(by_song) + 1))
        t0 = int(r.integers(0, 600))
        own = [(h, off - t0) for h, off in by_song[sid] if t0 <= off < t0 + 250]
        own = [own[i] for i in r.permutation(len(own))[: max(1, int(0.6 * len(own)))]]
        noise = [(hashlib.sha1(f"{a}|{b}|{c}".encode()).hexdigest()[:20], int(o)) for a, b, c, o in
                 zip(r.integers(0, 40, 60), r.integers(0, 40, 60), r.integers(0, 12, 60), r.integers(0, 250, 60))]
        hashes = list(set(own + noise))          # FileRecognizer passes a set of (hash, offset)
        if q == n_queries - 1:
            hashes = [(h.lower(), o) for h, o in hashes[:5]] + [("f" * 20, 3)]   # lower case + a hash that is nowhere
        queries.append(hashes)
    return queries


def _reference_classes():
    from oracle import ref_loader

    ref_loader._install_stubs()
    for name in ("psycopg2", "psycopg2.extras"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["psycopg2.extras"].DictCursor = object
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.PeakSignalNoiseRatio = lambda **k: None
        sys.modules["torchmetrics"] = tm
    ns = ref_loader.load()          # registers the `dejavu` aliases, patches set_gpus / torch.load
    import afp.dejavu.database as r_db

    sys.modules["dejavu.database"] = r_db
    import afp.dejavu.postgres_database as r_pg

    sys.modules["dejavu.postgres_database"] = r_pg
    sys.modules["dejavu.fingerprint"] = ns.dejavu_fingerprint
    import afp.dejavu.dejavu as r_dj
    import testing.metrics as r_metrics

    return r_pg, r_dj, r_metrics, ns


class _Cursor:
    """Stand-in for the psycopg2 cursor behind CommonDatabase.return_matches."""

    def __init__(self, table):
        self.table, self.rows = table, []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def execute(self, query, values=()):
        self.rows = [(h, sid, off) for h in values for sid, off in self.table.get(h.upper(), [])]

    def __iter__(self):
        return iter(self.rows)


def main():
    r_pg, r_dj, r_metrics, ns = _reference_classes()
    rows = synthetic_table()
    queries = synthetic_queries(rows)
    table = {}
    for h, sid, off in rows:
        table.setdefault(h.upper(), []).append((sid, off))

    db = r_pg.PostgreSQLDatabase.__new__(r_pg.PostgreSQLDatabase)
    db.cursor = lambda **k: _Cursor(table)
    songs = {sid: {"song_name": f"song{sid}", "total_hashes": 300, "file_sha1": "AB" * 20} for sid in range(1, 41)}
    db.get_song_by_id = lambda sid: songs[sid]
    djv = r_dj.Dejavu.__new__(r_dj.Dejavu)
    djv.db = db
    djv.settings = dict(ns.parameters.afp_settings["dejavu"])

    out = {"meta": json.dumps({"generator": "oracle/make_golden_dejavu_match.py", "numpy": np.__version__}),
           "table_hash": np.array([h for h, _, _ in rows]), "table_song": np.array([s for _, s, _ in rows], np.int32),
           "table_offset": np.array([o for _, _, o in rows], np.int32), "n_queries": len(queries)}
    for i, hashes in enumerate(queries):
        matches, dedup = db.return_matches(hashes)
        res = djv.align_matches(matches, dedup, len(hashes))
        out[f"q{i}_hash"] = np.array([h for h, _ in hashes])
        out[f"q{i}_offset"] = np.array([o for _, o in hashes], np.int32)
        out[f"q{i}_matches"] = np.array(sorted(matches), np.int32).reshape(-1, 2)
        out[f"q{i}_dedup"] = np.array(sorted(dedup.items()), np.int32).reshape(-1, 2)
        if res:
            r0 = res[0]
            out[f"q{i}_top"] = np.array([r0["song_id"], r0["offset"], r0["nb_matches_with_offset"], r0["hashes_matched_in_input"],
                                         r0["input_total_hashes"]], np.int64)
            out[f"q{i}_conf"] = np.array([r0["input_confidence"], r0["input_confidence_2"], r0["fingerprinted_confidence"],
                                          r0["offset_seconds"]], np.float64)
        else:
            out[f"q{i}_top"] = np.zeros(0, np.int64)
    np.savez_compressed(os.path.join(GOLD, "dejavu_match.npz"), **out)

    import torch

    r = np.random.default_rng(11)
    m = {"meta": out["meta"], "n_cases": 5}
    for i in range(5):
        shape = (1, 251, 256) if i < 3 else (1, 17, 9)
        gt = (r.random(shape) < (0.02 if i != 2 else 0.0)).astype(np.float32)
        pred = np.where(r.random(shape) < 0.6, gt, (r.random(shape) < 0.02)).astype(np.float32)
        if i == 4:
            pred[:] = 0
        m[f"gt{i}"], m[f"pred{i}"] = gt, pred
        tp, tg = torch.from_numpy(pred), torch.from_numpy(gt)
        m[f"res{i}"] = np.array([r_metrics.Recall()(tp, tg), r_metrics.Precision()(tp, tg), r_metrics.F1score()(tp, tg)], np.float64)
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **m)
    print("wrote dejavu_match.npz, metrics.npz")


if __name__ == "__main__":
    main()
