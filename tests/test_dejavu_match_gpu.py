"""GPU parity: in-memory Dejavu index (lookup + offset vote) and the evaluation sums, through the drop-in
`Dejavu` / `FileRecognizer` / `testing.metrics` names, against golden vectors made by the reference's own
return_matches / align_matches / Recall / Precision / F1score (oracle/make_golden_dejavu_match.py)."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from oracle import dejavu_np as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def mods():
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu", "testing")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    import afp.dejavu.dejavu as dj
    import afp.dejavu.file_recognizer as fr
    import afp.dejavu.postgres_database as pg
    import testing.metrics as tm

    yield {"dj": dj, "fr": fr, "pg": pg, "tm": tm}
    sys.path.remove(DROPIN)
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu", "testing")]:
        del sys.modules[k]


def _golden():
    g = np.load(os.path.join(GOLD, "dejavu_match.npz"))
    qs = [list(zip(g[f"q{i}_hash"].tolist(), g[f"q{i}_offset"].tolist())) for i in range(int(g["n_queries"]))]
    return g, qs


def test_return_matches_and_align_equal_the_reference(mods):
    g, qs = _golden()
    settings = {"samplerate": 8000, "n_hop": 256}
    djv = mods["dj"].Dejavu({"database": {"database": "golden"}}, settings, "clear")
    for sid in range(1, 41):
        assert djv.db.insert_song(f"song{sid}", "AB" * 20, 300) == sid
    for sid in range(1, 41):
        sel = g["table_song"] == sid
        djv.db.insert_hashes(sid, zip(g["table_hash"][sel].tolist(), g["table_offset"][sel].tolist()))
        djv.db.set_song_fingerprinted(sid)
    for i, hashes in enumerate(qs):
        matches, dedup = djv.db.return_matches(hashes)
        assert np.array_equal(np.array(sorted(matches), np.int32).reshape(-1, 2), g[f"q{i}_matches"]), i
        assert np.array_equal(np.array(sorted(dedup.items()), np.int32).reshape(-1, 2), g[f"q{i}_dedup"]), i
        m2, dedup2, _ = djv.find_matches(hashes)
        assert len(m2) == len(matches) and sorted(m2) == sorted(matches) and dedup2 == dedup
        for source in (m2, matches):                       # pairs kept on the device, and a plain list re-uploaded
            res = djv.align_matches(source, dedup, len(hashes))
            want = g[f"q{i}_top"]
            if len(want) == 0:
                assert res == []
                continue
            r0 = res[0]
            assert (r0["song_id"], r0["offset"], r0["nb_matches_with_offset"], r0["hashes_matched_in_input"],
                    r0["input_total_hashes"]) == tuple(int(v) for v in want), i
            assert np.allclose([r0["input_confidence"], r0["input_confidence_2"], r0["fingerprinted_confidence"], r0["offset_seconds"]],
                               g[f"q{i}_conf"]), i
            assert r0["song_name"] == f"song{int(want[0])}".encode("utf8")


def test_vote_tie_breaks(mods):
    """Equal counts: the smaller song id wins, then the smaller offset difference - the first element of the reference's
    stable sorts (dejavu.py:330-345)."""
    from musicfpaugment_b200 import lib, runtime

    ix = lib.DejavuIndex(runtime.get_context(), [1], [0], [1], [0], 8)
    pairs = [(5, 10)] * 3 + [(3, -7)] * 3 + [(3, -9)] * 3 + [(7, 0)] * 2
    rng = np.random.default_rng(0)
    for _ in range(4):
        p = torch.tensor([pairs[i] for i in rng.permutation(len(pairs))], dtype=torch.int32).cuda()
        assert ix.align(p) == (3, -9, 3) == D.align_top([tuple(r) for r in p.cpu().tolist()])
    assert ix.align(torch.zeros(0, 2, dtype=torch.int32).cuda())[0] == -1
    ix.close()


def test_recognize_file_end_to_end(mods, tmp_path):
    """dejavu_exps.create_fp_database + compute_accuracy call sequence (testing/dejavu_exps.py:16-79) on synthetic
    query pickles: Dejavu.fingerprint_directory, FileRecognizer.recognize_file -> the indexed track is identified."""
    from musicfpaugment_b200 import synth

    tracks = synth.music_like(6, n_samples=8 * 8000, seed=77).numpy()
    paths = []
    for i, x in enumerate(tracks):
        p = tmp_path / f"track{i}.pkl"
        with open(p, "wb") as fh:
            pickle.dump(np.asarray(x, np.float32), fh)
        paths.append(str(p))
    settings = {"samplerate": 8000, "n_fft": 512, "n_hop": 256, "fan_value": 3, "amp_min": 50, "peak_neighborhood_size": 10}
    djv = mods["dj"].Dejavu({"database": {"database": "e2e"}}, settings, "clear")
    djv.fingerprint_directory(paths)
    assert djv.db.get_num_songs() == 6 and len(djv.songhashes_set) == 6
    n_fp = djv.db.get_num_fingerprints()
    djv.fingerprint_directory(paths)                       # already indexed: skipped (dejavu.py:206-208)
    assert djv.db.get_num_fingerprints() == n_fp
    rec = mods["fr"].FileRecognizer(djv)
    rng = np.random.default_rng(5)
    hits = 0
    for i, x in enumerate(tracks):
        q = tmp_path / f"query{i}.pkl"
        with open(q, "wb") as fh:                          # a 4 s excerpt with a little noise
            pickle.dump(np.asarray(x[16000:48000] + 0.01 * rng.standard_normal(32000), np.float32), fh)
        out = rec.recognize_file(str(q))
        assert set(out) == {"total_time", "fingerprint_time", "query_time", "align_time", "results", "match"}
        if out["match"] and out["results"][0]["song_name"].decode("utf-8") == f"track{i}":
            hits += 1
            assert out["results"][0]["offset"] * 256 == pytest.approx(16000, abs=512)
    assert hits == 6


def test_metrics_equal_the_reference(mods, mfpa_ctx):
    g = np.load(os.path.join(GOLD, "metrics.npz"))
    tm = mods["tm"]
    for i in range(int(g["n_cases"])):
        pred, gt = torch.from_numpy(g[f"pred{i}"]), torch.from_numpy(g[f"gt{i}"])
        got = [tm.Recall()(pred, gt), tm.Precision()(pred, gt), tm.F1score()(pred, gt)]
        assert np.allclose(got, g[f"res{i}"], rtol=0, atol=1e-12), (i, got)
    rng = np.random.default_rng(3)
    a, b = rng.random((1, 257, 251)), rng.random((1, 257, 251))
    assert tm.psnr(torch.from_numpy(a), torch.from_numpy(b)).item() == pytest.approx(D.psnr(a, b), rel=1e-6)   # a float32 0-d tensor
