"""CPU: the numpy oracle must reproduce the golden vectors that the REAL reference
produced in the build container (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import audfprint_np as O
from oracle import dejavu_np as D

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def g_afp():
    return np.load(os.path.join(GOLD, "audfprint.npz"))


def test_audfprint_peaks_landmarks_hashes(g_afp):
    for i in range(int(g_afp["n_cases"])):
        x = g_afp[f"x{i}"]
        pk, mask, spec = O.find_peaks(x)
        assert np.array_equal(np.asarray(pk, np.int32).reshape(-1, 2), g_afp[f"peaks{i}"]), i
        lm = O.peaks2landmarks(pk)
        assert np.array_equal(np.asarray(lm, np.int32).reshape(-1, 4), g_afp[f"landmarks{i}"]), i
        assert np.array_equal(O.landmarks2hashes(lm), g_afp[f"hashes{i}"]), i


@pytest.mark.parametrize("shifts", [1, 4])
def test_audfprint_wavfile2hashes(g_afp, shifts):
    for i in range(int(g_afp["n_cases"])):
        got = O.wave2hashes(g_afp[f"x{i}"], shifts)
        want = g_afp[f"wf2h_s{shifts}_{i}"]
        assert got.dtype == np.int32 and np.array_equal(got, want), i


def test_audfprint_spectrogram_and_filter_bit_exact(g_afp):
    for i in (4, 5):
        spec = O.normalise(O.stft_mag(g_afp[f"x{i}"]))
        assert np.array_equal(spec, g_afp[f"spec{i}"])
        assert np.array_equal(O.onset_filter(spec), g_afp[f"sgram{i}"])
        pk, mask = O.peaks_from_sgram(g_afp[f"sgram{i}"])
        assert np.array_equal(mask, g_afp[f"mask{i}"])


def _golden_table(g):
    ht = O.HashTable()
    idx = g["counts_nonzero_idx"]
    ht.counts[idx] = g["counts_nonzero"]
    ht.table[idx] = g["table_rows"]
    ht.hashesperid = g["hashesperid"]
    ht.names = [f"track{t:04d}" for t in range(int(g["n_tracks"]))]
    return ht


def test_match_hashes_golden():
    g = np.load(os.path.join(GOLD, "match.npz"))
    ht = _golden_table(g)
    assert np.array_equal(ht.get_hits(g["q0"]), g["hits0"])
    for i in range(int(g["n_queries"])):
        q = g[f"q{i}"]
        assert len(ht.get_hits(q)) == int(g["hits_n"][i])
        got = O.match_hashes(ht, q)
        want = g[f"res{i}"]
        assert got.shape == want.shape, i
        # ties in the filtered count may be ordered differently (unstable argsort)
        assert np.array_equal(got[:, 1], want[:, 1])
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))


def test_match_exact_count_time_range_hashesfor_golden():
    """exact_count / find_time_range / hashesfor restated (audfprint_match.py:130-233, 296-304) vs the reference
    Matcher's own output on the golden index (oracle/make_golden_match_exact.py)."""
    g = np.load(os.path.join(GOLD, "match.npz"))
    e = np.load(os.path.join(GOLD, "match_exact.npz"))
    ht = _golden_table(g)
    n = 0
    for i in range(int(g["n_queries"])):
        if f"exact{i}" not in e.files:
            continue
        q, want = g[f"q{i}"], e[f"exact{i}"]
        got, hf = O.match_hashes_exact(ht, q, find_time_range=True, hashesfor=0 if len(want) else None)
        assert got.shape == want.shape and sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist())), i
        if len(want):
            assert np.array_equal(hf, e[f"hashesfor{i}"]), i
            n += 1
        tr = O.approx_time_ranges(ht, q, O.match_hashes(ht, q))
        assert sorted(map(tuple, tr.tolist())) == sorted(map(tuple, e[f"approx_tr{i}"].tolist())), i
    assert n >= 20


def test_hash_table_store_matches_layout():
    """store(): bucket fill order, saturation counting, id/time packing (hash_table.py:70-116)."""
    ht = O.HashTable()
    rows = np.array([[5, 77], [6, 77], [16385, 77], [9, (1 << 20) + 3]], np.int32)
    ht.store("a", rows)
    ht.store("b", rows[:2])
    assert ht.counts[77] == 5 and ht.counts[3] == 1
    assert ht.table[77, :5].tolist() == [(1 << 14) + 5, (1 << 14) + 6, (1 << 14) + 1, (2 << 14) + 5, (2 << 14) + 6]
    assert ht.hashesperid.tolist() == [4, 2]
    hits = ht.get_hits(np.array([[2, 77]], np.int32))
    assert hits.tolist() == [[0, 3, 77, 2], [0, 4, 77, 2], [0, -1, 77, 2], [1, 3, 77, 2], [1, 4, 77, 2]]


def test_dejavu_peaks_golden():
    g = np.load(os.path.join(GOLD, "dejavu.npz"))
    for i in range(4):
        pk, mask = D.get_2d_peaks(g[f"arr{i}"], amp_min=float(g[f"amp_min{i}"]))
        assert np.array_equal(np.asarray(pk, np.int32).reshape(-1, 2), g[f"peaks{i}"]), i
        assert np.array_equal(mask.astype(np.uint8), g[f"mask{i}"]), i


def test_augment_chain_golden():
    """AugmentFP arithmetic vs the reference's transform classes (dumped parameters);
    1e-4 relative is the north-star tolerance, the oracle is ~1e-6."""
    from oracle import augment_np as A

    g = np.load(os.path.join(GOLD, "augment.npz"))
    keys = ("fc1", "ir", "noise", "snr_db", "gain_factor", "clip_p", "fc2", "fc3")
    for i in range(int(g["n_cases"])):
        prm = {k: (g[f"p{i}_{k}"] if g[f"p{i}_{k}"].ndim else float(g[f"p{i}_{k}"])) for k in keys if f"p{i}_{k}" in g}
        y, stages = A.augment_chain(g[f"x{i}"], prm, return_stages=True)
        ref = g[f"out{i}"]
        assert y.dtype == np.float32 and y.shape == ref.shape
        assert np.abs(y - ref).max() / np.abs(ref).max() < 5e-6, i
        if i == 0:
            for k in A.STAGES:
                r = g[f"stage0_{k}"]
                assert np.abs(stages[k] - r).max() / np.abs(r).max() < 5e-6, k


def test_fir_taps_properties():
    from oracle import augment_np as A

    for fc, half in ((150.0, 213), (30.0, 1066), (3999.0, 8), (3000.0, 10)):
        h = A.lowpass_taps(A.cutoff_fraction(fc, 8000))
        assert len(h) == 2 * half + 1 and abs(float(h.sum()) - 1) < 1e-5
        assert np.allclose(h, h[::-1], atol=1e-7)
    for bad in (-0.1, 0.6, 0.0):
        with pytest.raises(ValueError):
            A.lowpass_taps(bad)


def test_unet_oracle_reproduces_reference_golden():
    """oracle/unet_torch.py == the real reference UNet (tests/golden/unet.npz), bit for bit on the CPU
    for torch.manual_seed(0) weights; state_dict keys (= parameter blob order) identical."""
    import torch

    from oracle.unet_torch import seeded_unet

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet.npz"))
    torch.set_num_threads(1)
    net = seeded_unet(0, randomize_bn=False)
    assert list(net.state_dict().keys()) == [str(k) for k in g["keys"]]
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"])).numpy()
    np.testing.assert_allclose(y, g["y"], rtol=0, atol=1e-6)


def test_dejavu_specgram_restatement_matches_scipy_definition():
    """oracle/dejavu_np.specgram_psd (restating matplotlib.mlab.specgram, not installed) against
    scipy.signal.spectrogram, an independent implementation of the same one-sided density periodogram."""
    import scipy.signal

    from musicfpaugment_b200 import synth

    for i, T in enumerate((64000, 5000, 513)):
        x = synth.music_like(1, seed=40 + i).numpy()[0, :T].astype(np.float64)
        got = D.specgram_psd(x)
        _, _, want = scipy.signal.spectrogram(x, fs=8000.0, window=np.hanning(512), nperseg=512, noverlap=256, nfft=512,
                                              detrend=False, return_onesided=True, scaling="density", mode="psd")
        assert got.shape == want.shape == (257, (T - 256) // 256)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-18)


def test_julius_taps_restatement_against_scipy_firwin():
    """Independent pin of the restated julius.lowpass_filter taps (oracle/augment_np.lowpass_taps; julius 0.2.7 is a
    third-party dependency absent from the reference tree and the image): scipy.signal.firwin builds the same
    Hann-windowed sinc with unit DC gain from its own code (cut-off there is relative to Nyquist: 2 x julius')."""
    from scipy.signal import firwin

    from oracle import augment_np as A

    for fc_hz in (150.0, 30.0, 3999.0, 3000.0, 4.0, 0.9):
        cutoff = fc_hz / 8000.0
        taps = A.lowpass_taps(cutoff)
        half = int(8.0 / cutoff / 2)
        assert len(taps) == 2 * half + 1 and abs(float(taps.sum(dtype=np.float64)) - 1.0) < 1e-5
        ref = firwin(2 * half + 1, 2.0 * cutoff, window="hann")
        assert np.abs(taps - ref).max() <= 2e-6 * np.abs(ref).max() + 1e-9, fc_hz
