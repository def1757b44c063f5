"""GPU: the drop-in modules (reference import paths) end to end against the oracle."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from oracle import audfprint_np as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")
PRM = {"density": 20, "pks-per-frame": 5, "freq-sd": 30, "shifts": 1, "samplerate": 8000, "n_fft": 512, "n_hop": 256}


@pytest.fixture(scope="module")
def mods():
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    import afp.audfprint.audfprint_match as m
    import afp.audfprint.hash_table as ht
    import afp.audfprint.peak_extractor as pe
    import afp.dejavu.fingerprint as fp
    import augmentation as aug

    yield {"pe": pe, "ht": ht, "match": m, "fp": fp, "aug": aug}
    sys.path.remove(DROPIN)
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu")]:
        del sys.modules[k]


def _queries(n):
    from musicfpaugment_b200 import synth

    return synth.music_like(n, seed=808).numpy()


def test_find_peaks_and_landmarks(mods):
    """find_peaks through the drop-in: types and shapes of the reference's 3-tuple, spectrogram within 1e-4, and the
    north star's end-to-end gate (>= 99.9 %) on the peak sets, measured over 12 queries (one flipped peak of a single
    query is ~1 % of ITS set, so the gate is an aggregate)."""
    a = mods["pe"].Audfprint_peaks(PRM)
    inter = union = 0
    for x in _queries(12):
        pk, mask, spec = a.find_peaks(x)
        pk_o, mask_o, spec_o = O.find_peaks(x)
        assert mask.shape == mask_o.shape and mask.dtype == np.float32 and spec.shape == spec_o.shape and spec.dtype == np.float64
        assert np.abs(spec - spec_o).max() < 1e-4
        inter += len(set(pk) & set(pk_o))
        union += len(set(pk) | set(pk_o))
    assert inter >= 0.999 * union, (inter, union)
    assert a.peaks2landmarks(pk_o) == O.peaks2landmarks(pk_o)  # bit-exact given the same peaks


def test_wavfile2hashes_and_match_file(mods, tmp_path):
    pe, ht_mod, m_mod = mods["pe"], mods["ht"], mods["match"]
    X = _queries(6)
    index = pe.Audfprint_peaks(PRM)
    ht = ht_mod.HashTable()
    for i, x in enumerate(X):  # "index" the clean tracks through ingest()
        p = str(tmp_path / f"track{i}.pkl")
        with open(p, "wb") as f:
            pickle.dump(x, f)
        dur, n = index.ingest(ht, p)
        assert abs(dur - 8.0) < 1e-6 and n == len(O.wave2hashes(x))
    qa = pe.Audfprint_peaks(PRM)
    qa.shifts = 4  # audfprint_exps.py:167
    matcher = m_mod.Matcher()
    hits = inter = union = 0
    for i, x in enumerate(X):
        seg = x[8000:40000] + 0.01 * np.random.default_rng(i).standard_normal(32000).astype(np.float32)
        p = str(tmp_path / f"query{i}.pkl")
        with open(p, "wb") as f:
            pickle.dump(seg.astype(np.float32), f)
        h = qa.wavfile2hashes(p)
        want = O.wave2hashes(seg.astype(np.float32), 4)
        a, b = {tuple(r) for r in h.tolist()}, {tuple(r) for r in want.tolist()}
        inter += len(a & b)
        union += len(a | b)
        status, name, n = matcher.file_match_to_msgs(qa, ht, p)
        hits += int(status == "MATCH" and name.endswith(f"track{i}.pkl"))
    assert hits == len(X)
    assert inter >= 0.999 * union, (inter, union)   # the north star's end-to-end gate over the 6 x 4-shift queries
    # get_hits parity on the shim table
    oracle_ht = O.HashTable()
    oracle_ht.table, oracle_ht.counts, oracle_ht.hashesperid = ht.table, ht.counts, ht.hashesperid
    assert np.array_equal(ht.get_hits(want), oracle_ht.get_hits(want))
    res, _ = matcher.match_hashes(ht, want)
    ref = O.match_hashes(oracle_ht, want)
    assert res.shape == ref.shape and np.array_equal(res[:, :4], ref[:, :4])


def test_matcher_exact_count_time_range_hashesfor(mods):
    """Matcher(exact_count / find_time_range) and hashesfor on the golden index: rows and matching hashes equal the
    reference Matcher's own output (tests/golden/match_exact.npz, oracle/make_golden_match_exact.py)."""
    gold = os.path.join(ROOT, "tests", "golden")
    g, e = np.load(os.path.join(gold, "match.npz")), np.load(os.path.join(gold, "match_exact.npz"))
    ht = mods["ht"].HashTable()
    idx = g["counts_nonzero_idx"]
    ht.table[idx], ht.counts[idx] = g["table_rows"], g["counts_nonzero"]
    ht.hashesperid = g["hashesperid"].copy()
    ht.names = [f"track{t:04d}" for t in range(int(g["n_tracks"]))]
    rows = lambda a: sorted(map(tuple, np.asarray(a).tolist()))
    checked = 0
    for i in range(int(g["n_queries"])):
        if f"exact{i}" not in e.files:
            continue
        q, want = g[f"q{i}"], e[f"exact{i}"]
        m = mods["match"].Matcher()
        m.exact_count, m.find_time_range = True, True
        got, hf = m.match_hashes(ht, q, hashesfor=0 if len(want) else None)
        assert got.dtype == np.int32 and got.shape == want.shape and rows(got) == rows(want), i
        if len(want):
            assert np.array_equal(hf, e[f"hashesfor{i}"]), i
            checked += 1
        m2 = mods["match"].Matcher()
        m2.find_time_range = True
        assert rows(m2.match_hashes(ht, q)[0]) == rows(e[f"approx_tr{i}"]), i
    assert checked >= 20


def test_add_colored_noise_drop_in(mods):
    """augmentation.transformations.colored_noise.AddColoredNoise: same seeded draws as the reference (gate, SNRs,
    decays, noise periods) and the mix on the GPU within 2e-6 of the reference's output
    (tests/golden/colored_noise.npz, oracle/make_golden_colored_noise.py)."""
    from augmentation.transformations.colored_noise import AddColoredNoise

    g = np.load(os.path.join(ROOT, "tests", "golden", "colored_noise.npz"))
    t = AddColoredNoise(min_snr_in_db=3.0, max_snr_in_db=30.0, min_f_decay=-2.0, max_f_decay=2.0, p=0.7,
                        sample_rate=int(g["sample_rate"]))
    torch.manual_seed(int(g["seed"]))
    out = t(samples=torch.from_numpy(g["x"].copy()), sample_rate=int(g["sample_rate"]))
    tp = t.transform_parameters
    assert np.array_equal(tp["should_apply"].numpy(), g["should_apply"])
    assert np.array_equal(tp["snr_in_db"].numpy(), g["snr_in_db"]) and np.array_equal(tp["f_decay"].numpy(), g["f_decay"])
    got = out.samples.numpy()
    assert got.shape == g["out"].shape and out.sample_rate == int(g["sample_rate"])
    assert np.array_equal(got[~g["should_apply"]], g["x"][~g["should_apply"]])     # untouched examples pass through
    assert float(np.abs(got - g["out"]).max()) <= 2e-6                                # outputs are peak-normalised: |y| <= 1
    with pytest.raises(RuntimeError):
        t(samples=torch.zeros(4, 8000), sample_rate=8000)


def test_band_filters_drop_in(mods):
    """BandPassFilter / BandStopFilter: the seeded draws equal the reference's (tests/golden/band_params.npz, made by the
    reference's own class), the filtering equals the restated julius.bandpass_filter (oracle/augment_np.bandpass -
    julius is absent, that part is unpinned) within 1e-5 of the signal's peak; one example needs an 11 909-tap window
    (a 5.4 Hz low cut-off: the partitioned FFT path), the others 100-600 taps."""
    from augmentation.transformations.band_filters import BandPassFilter, BandStopFilter
    from oracle import augment_np as A

    g = np.load(os.path.join(ROOT, "tests", "golden", "band_params.npz"))
    B, sr, T = int(g["batch"]), int(g["sample_rate"]), 24000
    x = np.random.default_rng(5).standard_normal((B, 1, T)).astype(np.float32) * 0.3
    for cls in (BandPassFilter, BandStopFilter):
        t = cls(min_center_frequency=200, max_center_frequency=1900, min_bandwidth_fraction=0.5, max_bandwidth_fraction=1.99,
                p=0.8, sample_rate=sr)
        torch.manual_seed(int(g["seed"]))
        out = t(samples=torch.from_numpy(x.copy()), sample_rate=sr).samples.numpy()
        tp = t.transform_parameters
        gate = g["should_apply"]
        assert np.array_equal(tp["should_apply"].numpy(), gate)
        assert np.array_equal(tp["center_freq"].numpy(), g["center_freq"]) and np.array_equal(tp["bandwidth"].numpy(), g["bandwidth"])
        assert np.array_equal(out[~gate], x[~gate])
        worst = 0.0
        for k, i in enumerate(np.nonzero(gate)[0]):
            band = A.bandpass(x[i, 0], float(g["low"][k]), float(g["high"][k]))
            want = band if cls is BandPassFilter else x[i, 0] - band
            worst = max(worst, float(np.abs(out[i, 0] - want).max()))
        assert worst <= 1e-5 * float(np.abs(x).max()) * 4, worst
    with pytest.raises(ValueError):      # a band reaching past sr / 2: julius raises, so does the drop-in
        t = BandPassFilter(min_center_frequency=3900, max_center_frequency=4000, min_bandwidth_fraction=1.0,
                           max_bandwidth_fraction=1.5, p=1.0, sample_rate=sr)
        t(samples=torch.from_numpy(x[:2].copy()), sample_rate=sr)


def test_get_2d_peaks(mods):
    from oracle import dejavu_np as D

    g = np.load(os.path.join(ROOT, "tests", "golden", "dejavu.npz"))
    pk, mask = mods["fp"].get_2D_peaks(g["arr0"], plot=False, amp_min=50)
    assert [tuple(map(int, p)) for p in pk] == [tuple(r) for r in g["peaks0"].tolist()]
    assert mask.dtype == np.float64 and np.array_equal(mask.astype(np.uint8), g["mask0"])


def test_augmentfp_call_matches_oracle_on_dumped_parameters(mods):
    from oracle import augment_np as A

    aug = mods["aug"]
    g = torch.Generator().manual_seed(3)
    irs = [{"samples": torch.randn(1, n, generator=g) * torch.exp(-torch.arange(n) / 200.0), "sample_rate": 8000} for n in (500, 800)]
    bg = {"street": [{"samples": torch.randn(1, 90000, generator=g), "sample_rate": 8000}]}
    params = dict(aug.DEFAULT_PARAMETERS)
    params.update({k: 1.0 for k in params if k.startswith("proba_")})
    params["min_cutoff_freq1"] = 20.0
    a = aug.AugmentFP(bg, 8000, parameters=params, impulse_response_dir=irs)
    a.augmentation_pipeline.freeze_parameters(42)
    x = torch.from_numpy(_queries(1)[0][:32000])
    y = a(x.unsqueeze(0))
    assert y.shape == (1, 32000) and not y.is_cuda
    t = a.augmentation_pipeline.transforms
    prm = dict(fc1=float(t[0].transform_parameters["cutoff_freq"][0]), ir=t[1].transform_parameters["ir"][0, 0].cpu().numpy(),
               noise=t[2].transform_parameters["background"][0, 0].cpu().numpy(), snr_db=float(t[2].transform_parameters["snr_in_db"][0]),
               gain_factor=float(t[3].transform_parameters["gain_factors"].reshape(-1)[0]),
               clip_p=float(t[4].transform_parameters["percentile_threshold"].reshape(-1)[0]),
               fc2=float(t[5].transform_parameters["cutoff_freq"][0]), fc3=float(t[6].transform_parameters["cutoff_freq"][0]))
    ref = A.augment_chain(x.numpy(), prm)
    assert np.abs(y[0].numpy() - ref).max() / np.abs(ref).max() < 1e-4
    yb = a.batch_augment(torch.from_numpy(_queries(3)[:, None, :16000]))
    assert yb.shape == (3, 1, 16000) and torch.isfinite(yb).all()


def test_noise_rows_from_the_device_bank_equal_the_host_assembly(mods):
    """AddBackgroundNoise.random_background (background_noise.py:64-141): with the same `random` seed the
    device path (piece descriptors -> mfpa_noise_assemble on a bank of decoded files) gives the rows the
    host path builds with torch: single files longer and shorter than the query (several pieces per row),
    a mix-up pair, and both RMS normalisations."""
    import random

    aug = mods["aug"]
    g = torch.Generator().manual_seed(11)
    mk = lambda n, a=1.0: {"samples": a * torch.randn(1, n, generator=g), "sample_rate": 8000}
    pair = [mk(70000, 0.3), mk(90000, 2.0)]
    bg = {"street": [mk(90000), mk(5000, 0.1), mk(20000, 3.0)], "cafe": [pair, mk(64000)], "rain": [mk(1234)]}
    t = aug.AddBackgroundNoise(bg, min_snr_in_db=0.0, max_snr_in_db=10.0, p=1.0, sample_rate=8000)
    assert t.device_bank
    T, n = 64000, 24
    random.seed(5)
    dev = t.random_backgrounds_device(n, T)
    state_after_device = random.getstate()
    assert dev is not None and dev.is_cuda and dev.shape == (n, 1, T)
    random.seed(5)
    host = torch.stack([t.random_background(T) for _ in range(n)])
    assert random.getstate() == state_after_device          # the same number of draws
    err = (dev.cpu() - host).abs().max().item()
    assert err < 2e-6 * host.abs().max().item(), err
    rms = dev.square().mean(dim=-1).sqrt()
    assert torch.allclose(rms, torch.ones_like(rms), atol=1e-5)
    # a file at another sample rate is not bankable: the transform falls back to the host assembly
    t2 = aug.AddBackgroundNoise({"x": [{"samples": torch.randn(1, 200000, generator=g), "sample_rate": 16000}]}, p=1.0, sample_rate=8000)
    assert t2.random_backgrounds_device(2, 8000) is None


def test_find_peaks_with_unet_denoising(mods):
    """Audfprint_peaks(params, denoising=True, denoising_model="unet") (peak_extractor.py:263-269): the
    returned spectrogram is the network output; the picker run on it must agree with the oracle picker
    fed that same spectrogram, and wavfile2hashes' batched form must equal the per-item chain."""
    from oracle.unet_torch import seeded_unet

    pe = mods["pe"]
    net = seeded_unet(0)
    pe.Audfprint_peaks.set_unet_state_dict(net.state_dict())
    an = pe.Audfprint_peaks(dict(PRM), denoising=True, denoising_model="unet")
    X = _queries(2)
    for x in X:
        pk, mask, spec = an.find_peaks(x)
        assert spec.shape == (257, 251) and spec.dtype == np.float32 and mask.shape == (256, 251)
        # fp32 torch UNet on the oracle's normalised spectrogram: bf16 tolerance (3 % of the range)
        import torch

        sg = O.normalise(O.stft_mag(x)).astype(np.float32)
        with torch.no_grad():
            ref = net(torch.from_numpy(sg)[None, None])[0, 0].numpy()
        assert np.abs(spec - ref).max() <= 0.03 * (ref.max() - ref.min())
        opk, omask = O.peaks_from_sgram(O.onset_filter(spec.astype(np.float64)))
        agree = len(set(pk) & set(opk)) / max(1, len(set(pk) | set(opk)))
        assert agree >= 0.99, agree
    # batched hashes (shifts = 4) == per-shift find_peaks -> landmarks -> hashes -> unique/sort
    an.shifts = 4
    got = an.waves2hashes(X[:1], shifts=4)[0]
    rows = []
    for s in range(4):
        off = int(s / 4 * 256)
        pk, _, _ = an.find_peaks(X[0][off:])
        rows.append(O.landmarks2hashes(O.peaks2landmarks(pk)))
    want = O.unique_sorted_hashes(np.concatenate(rows))
    assert np.array_equal(got, want)
    with pytest.raises(NotImplementedError):
        pe.Audfprint_peaks(dict(PRM), denoising=True, denoising_model="demucs")


def test_dejavu_fingerprint_drop_in(mods):
    """afp.dejavu.fingerprint.fingerprint (fingerprint.py:34-91): hashes are generate_hashes of the peaks
    of the GPU log-spectrogram; with get_masks it also returns the mask and the normalised PSD."""
    from oracle import dejavu_np as D
    from oracle.unet_torch import seeded_unet

    fp = mods["fp"]
    x = _queries(1)[0]
    hashes, mask, spec = fp.fingerprint(x, get_masks=True)
    assert mask.shape == spec.shape == (257, 249)
    want_arr, want_psd = D.dejavu_fingerprint_arr(x)
    assert np.abs(spec - want_psd).max() < 1e-4
    want_pk, _ = D.get_2d_peaks(want_arr)
    got_pk = set(zip(*np.nonzero(mask)))
    assert len(got_pk & set(want_pk)) / max(1, len(got_pk | set(want_pk))) >= 0.99
    assert hashes == fp.generate_hashes([(f, t) for f, t in sorted(got_pk)], fan_value=3)
    assert fp.fingerprint(x) == hashes
    # denoised: arr2D = unet(arr2D) ** 2 before the log (:70-75)
    fp.set_unet_state_dict(seeded_unet(0).state_dict())
    h2, m2, s2 = fp.fingerprint(x, denoising=True, denoising_model="unet", get_masks=True)
    assert s2.shape == (257, 249) and (s2 >= 0).all() and isinstance(h2, list)
    with pytest.raises(NotImplementedError):
        fp.fingerprint(x, denoising=True, denoising_model="demucs")
