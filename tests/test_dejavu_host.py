"""CPU: the in-memory stand-in for the Postgres tables (host logic only) and the oracle restatement of the Dejavu
matching semantics against golden vectors made by the reference's own return_matches / align_matches."""
import os
import sys

import numpy as np
import pytest

from oracle import dejavu_np as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_queries():
    g = np.load(os.path.join(GOLD, "dejavu_match.npz"))
    rows = list(zip(g["table_hash"].tolist(), g["table_song"].tolist(), g["table_offset"].tolist()))
    qs = [list(zip(g[f"q{i}_hash"].tolist(), g[f"q{i}_offset"].tolist())) for i in range(int(g["n_queries"]))]
    return g, rows, qs


def test_oracle_matches_reference_golden():
    g, rows, qs = golden_queries()
    for i, hashes in enumerate(qs):
        matches, dedup = D.return_matches(rows, hashes)
        assert np.array_equal(np.array(sorted(matches), np.int32).reshape(-1, 2), g[f"q{i}_matches"]), i
        assert np.array_equal(np.array(sorted(dedup.items()), np.int32).reshape(-1, 2), g[f"q{i}_dedup"]), i
        top = D.align_top(matches)
        if len(g[f"q{i}_top"]) == 0:
            assert top is None
        else:
            assert (top[0], top[1], top[2]) == tuple(int(v) for v in g[f"q{i}_top"][:3]), i
            assert dedup[top[0]] == int(g[f"q{i}_top"][3])


def test_metrics_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "metrics.npz"))
    for i in range(int(g["n_cases"])):
        got = [D.recall(g[f"pred{i}"], g[f"gt{i}"]), D.precision(g[f"pred{i}"], g[f"gt{i}"]), D.f1score(g[f"pred{i}"], g[f"gt{i}"])]
        assert np.allclose(got, g[f"res{i}"], rtol=0, atol=1e-12), (i, got, g[f"res{i}"])


@pytest.fixture()
def memdb():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("afp", "dejavu")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        import afp.dejavu.postgres_database as pg

        yield pg
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "dejavu")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_memory_database_tables(memdb, tmp_path):
    """Songs / fingerprints bookkeeping with the reference's SQL semantics: SERIAL ids from 1, only fingerprinted songs
    listed (SELECT_SONGS, postgres_database.py:334-343), duplicate (song, offset, hash) rows ignored (:288-296),
    unfingerprinted songs dropped by setup() (:32-39), ON DELETE CASCADE."""
    db = memdb.PostgreSQLDatabase(database="unit")
    db.empty()
    a = db.insert_song("a", "ab12", 3)
    b = db.insert_song("b", "cd34", 2)
    assert (a, b) == (1, 2) and db.get_songs() == [] and db.get_num_songs() == 0
    db.insert_hashes(a, [("00ff" * 5, 4), ("00ff" * 5, 4), ("00FF" * 5, 9), ("1234abcd" + "0" * 12, 1)])
    db.insert_hashes(b, [("00ff" * 5, 7)])
    db.set_song_fingerprinted(a)
    assert db.get_num_fingerprints() == 4            # one duplicate dropped, case-insensitive hex
    assert [s["song_name"] for s in db.get_songs()] == ["a"] and db.get_songs()[0]["file_sha1"] == "AB12"
    assert sorted(db.query("00FF" * 5)) == [(1, 4), (1, 9), (2, 7)]
    assert db.get_song_by_id(a) == {"song_name": "a", "file_sha1": "AB12", "total_hashes": 3}
    other = memdb.PostgreSQLDatabase(database="unit")   # a second "connection" sees the same tables
    assert other.get_num_fingerprints() == 4
    db.save(str(tmp_path / "db.npz"))
    db.setup()                                       # song b never finished: removed with its rows
    assert db.get_num_fingerprints() == 3 and db.get_song_by_id(b) is None
    db.load(str(tmp_path / "db.npz"))
    assert db.get_num_fingerprints() == 4 and db.get_song_by_id(b)["song_name"] == "b"
    assert memdb.split_hash("ffffffffffffffff0001") == (2 ** 64 - 1, 1)
