"""CPU: host-side logic of the drop-in modules (no kernels are launched here)."""
import os
import pickle
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")


@pytest.fixture()
def dropin_modules():
    """Import the drop-ins under their reference module names, then restore sys.modules."""
    from musicfpaugment_b200 import build

    build.build()
    from oracle import ref_loader

    if ref_loader.available():
        ref_loader.load()  # import the real reference first; its modules are parked while the drop-ins own the names
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("afp", "augmentation", "dejavu")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        import afp.audfprint.audfprint_match as m
        import afp.audfprint.hash_table as ht
        import afp.audfprint.peak_extractor as pe
        import afp.dejavu.fingerprint as fp
        import augmentation as aug

        yield {"pe": pe, "ht": ht, "match": m, "fp": fp, "aug": aug}
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_landmark_hash_round_trip(dropin_modules):
    from oracle import audfprint_np as O

    pe = dropin_modules["pe"]
    lms = [(3, 10, 40, 2), (3, 250, 220, 62), (9, 0, 30, 5), (9, 255, 225, 33)]
    h = pe.landmarks2hashes(lms)
    assert np.array_equal(h, O.landmarks2hashes(lms)) and h.dtype == np.int32
    assert pe.hashes2landmarks(h) == lms
    assert pe.landmarks2hashes([]).shape == (0, 2)
    v = np.array([1.0, 3.0, 3.0, 2.0, 5.0])
    assert np.array_equal(pe.locmax(v), O.locmax(v))


def test_analyzer_surface_and_errors(dropin_modules):
    pe = dropin_modules["pe"]
    prm = {"density": 20, "pks-per-frame": 5, "freq-sd": 30, "shifts": 1, "samplerate": 8000, "n_fft": 512, "n_hop": 256}
    a = pe.Audfprint_peaks(prm)
    assert (a.maxpairsperpeak, a.mindt, a.targetdt, a.targetdf) == (3, 2, 63, 31)
    p = a._afp()
    from oracle import audfprint_np as O

    assert p.a_dec == O.a_dec() and p.maxpks == 5
    assert a.find_peaks(np.zeros(0, np.float32))[0] == []
    with pytest.raises(NotImplementedError):  # Demucs is outside the hot path; the UNet is built
        pe.Audfprint_peaks(prm, denoising=True, denoising_model="demucs")
    assert pe.Audfprint_peaks(prm, denoising=True, denoising_model="unet").denoising
    with pytest.raises(ValueError):
        pe.Audfprint_peaks(dict(prm, n_fft=1024))


def test_hash_table_store_and_pickle_format(dropin_modules, tmp_path):
    from oracle import audfprint_np as O

    ht_mod = dropin_modules["ht"]
    r = np.random.default_rng(5)
    ht, ref = ht_mod.HashTable(), O.HashTable()
    for t in range(6):
        rows = np.stack([np.sort(r.integers(0, 900, 300)), r.integers(0, 8, 300) * 977], axis=1).astype(np.int32)
        random.seed(t)
        ht.store(f"t{t}", rows)
        random.seed(t)
        ref.store(f"t{t}", rows)
    assert ht.counts.max() > ht.depth  # overflowing buckets exercised the random overwrite
    assert np.array_equal(ht.table, ref.table) and np.array_equal(ht.counts, ref.counts)
    assert np.array_equal(ht.hashesperid, ref.hashesperid) and ht.names == ref.names
    path = str(tmp_path / "db.pklz")
    ht.save(path)
    back = ht_mod.HashTable(path)
    assert np.array_equal(back.table, ht.table) and back.names == ht.names and back.depth == 100 and back.hashbits == 20
    import gzip

    with gzip.open(path, "rb") as f:
        raw = pickle.load(f)
    assert type(raw).__module__ == "afp.audfprint.hash_table"
    assert not any(k.startswith("_") for k in ht.__getstate__())     # device state and version counter are not pickled


def test_hash_table_remove_retrieve_list_and_version(dropin_modules, capsys):
    """hash_table.py:277-326: remove() compacts the item's buckets and frees its name slot (re-used by the next new
    name), retrieve() returns its (time, hash) rows, list() prints the live items; every mutation bumps the version the
    device copy is keyed on."""
    ht = dropin_modules["ht"].HashTable()
    r = np.random.default_rng(9)
    rows = {n: np.stack([r.integers(0, 500, 40), r.integers(0, 5, 40) * 1000 + 3], axis=1).astype(np.int32) for n in "abc"}
    versions = [ht._version]
    for n in "abc":
        ht.store(n, rows[n])
        versions.append(ht._version)
    assert versions == sorted(set(versions))            # strictly increasing
    for n in "abc":
        got = ht.retrieve(n)
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, rows[n].tolist()))
    before = int(ht.totalhashes())
    ht.remove("b")
    assert ht._version > versions[-1] and ht.names == ["a", None, "c"] and ht.hashesperid[1] == 0
    assert int(ht.totalhashes()) == before - 40 and len(ht.retrieve("a")) == 40 and len(ht.retrieve("c")) == 40
    assert not ((ht.table >> ht.maxtimebits) == 2).any()
    filled = np.arange(ht.depth)[None, :] < ht.counts[:, None]
    assert (ht.table[filled] != 0).all() and (ht.table[~filled] == 0).all()     # buckets stay compact
    ht.store("d", rows["b"])
    assert ht.names == ["a", "d", "c"] and ht.hashesperid[1] == 40               # freed slot re-used (:262-266)
    lines = []
    ht.list(lines.append)
    assert lines == ["a (40 hashes)", "d (40 hashes)", "c (40 hashes)"]
    assert "Removed b ( 40 hashes)." in capsys.readouterr().out
    with pytest.raises(ValueError):
        ht.name_to_id("nobody")


def test_hash_table_files_cross_load_with_the_reference_class(dropin_modules, tmp_path):
    """f1 (on-disk format): a file the REFERENCE's HashTable.save wrote loads in the drop-in with the same contents,
    and a file the drop-in wrote loads in the reference's class (hash_table.py:118-198)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference checkout not mounted")
    ref_ht_mod = ref_loader.load().hash_table      # imported (and parked) before the drop-ins took the module names
    r = np.random.default_rng(12)
    rows = [np.stack([np.sort(r.integers(0, 900, 250)), r.integers(0, 6, 250) * 4099], axis=1).astype(np.int32) for _ in range(5)]

    def fill(ht):
        for t, rw in enumerate(rows):
            random.seed(100 + t)
            ht.store(f"track{t}", rw)
        return ht

    ref, mine = fill(ref_ht_mod.HashTable()), fill(dropin_modules["ht"].HashTable())
    assert ref.counts.max() > ref.depth
    assert np.array_equal(ref.table, mine.table) and np.array_equal(ref.counts, mine.counts) and ref.names == mine.names
    # the pickles name the class by module path: give each loader the class it expects under afp.audfprint.hash_table
    p_ref, p_mine = str(tmp_path / "ref.pklz"), str(tmp_path / "mine.pklz")
    current = sys.modules["afp.audfprint.hash_table"]
    sys.modules["afp.audfprint.hash_table"] = ref_ht_mod
    try:
        ref.save(p_ref)
        import gc

        gc.collect()     # the reference never closes its gzip stream (hash_table.py:123-125): flushed on collection
    finally:
        sys.modules["afp.audfprint.hash_table"] = current
    mine.save(p_mine)
    loaded = dropin_modules["ht"].HashTable(p_ref)                   # reference file -> drop-in
    assert np.array_equal(loaded.table, ref.table) and np.array_equal(loaded.counts, ref.counts)
    assert loaded.names == ref.names and np.array_equal(loaded.hashesperid, ref.hashesperid)
    assert (loaded.hashbits, loaded.depth, loaded.maxtimebits) == (ref.hashbits, ref.depth, ref.maxtimebits)
    sys.modules["afp.audfprint.hash_table"] = ref_ht_mod
    try:
        back = ref_ht_mod.HashTable(p_mine)                          # drop-in file -> reference
    finally:
        sys.modules["afp.audfprint.hash_table"] = current
    assert np.array_equal(back.table, mine.table) and np.array_equal(back.counts, mine.counts) and back.names == mine.names


def test_matcher_defaults(dropin_modules):
    m = dropin_modules["match"].Matcher()
    assert (m.window, m.threshcount, m.max_returns, m.search_depth, m.exact_count, m.max_alignments_per_id) == (2, 5, 1, 100, False, 100)
    p = m._params()
    assert (p.window, p.threshcount, p.search_depth, p.max_alignments_per_id) == (2, 5, 100, 100)


def test_dejavu_generate_hashes(dropin_modules):
    import hashlib

    fp = dropin_modules["fp"]
    peaks = [(10, 5), (20, 3), (30, 3), (40, 9)]
    hs = fp.generate_hashes(list(peaks), fan_value=3)
    assert hs[0] == (hashlib.sha1(b"20|30|0").hexdigest()[:20], 3)
    assert len(hs) == 5


def _in_memory_sources():
    g = torch.Generator().manual_seed(3)
    irs = [{"samples": torch.randn(1, n, generator=g), "sample_rate": 8000} for n in (500, 800, 300)]
    bg = {"street": [{"samples": torch.randn(1, 30000, generator=g), "sample_rate": 8000}],
          "cafe": [{"samples": torch.randn(1, 5000, generator=g), "sample_rate": 8000},
                   {"samples": torch.randn(1, 9000, generator=g), "sample_rate": 8000}]}
    return irs, bg


def test_augmentfp_parameter_draws(dropin_modules):
    aug = dropin_modules["aug"]
    irs, bg = _in_memory_sources()
    a = aug.AugmentFP(bg, 8000, impulse_response_dir=irs)
    pipe = a.augmentation_pipeline
    assert len(pipe.transforms) == 8
    pipe.freeze_parameters(42)
    arr, ir, noise = pipe.pack(64, 16000)
    on = lambda bit: (arr["apply"] & bit) != 0
    from musicfpaugment_b200 import lib

    assert on(lib.AUG_NORM).all() and 0.5 < on(lib.AUG_HPF1).mean() < 1.0
    assert ((arr["fc1_hz"][on(lib.AUG_HPF1)] >= 0) & (arr["fc1_hz"][on(lib.AUG_HPF1)] <= 150)).all()
    assert ((arr["fc2_hz"][on(lib.AUG_LPF)] >= 3000) & (arr["fc2_hz"][on(lib.AUG_LPF)] <= 3999)).all()
    assert ((arr["fc3_hz"][on(lib.AUG_HPF3)] >= 30) & (arr["fc3_hz"][on(lib.AUG_HPF3)] <= 150)).all()
    assert (np.abs(arr["snr_db"][on(lib.AUG_NOISE)]) <= 10).all()
    g = arr["gain_factor"][on(lib.AUG_GAIN)]
    assert ((g >= 10 ** (-5 / 20) - 1e-6) & (g <= 10 ** (5 / 20) + 1e-6)).all()
    assert ((arr["clip_p"][on(lib.AUG_CLIP)] >= 0) & (arr["clip_p"][on(lib.AUG_CLIP)] <= 0.01)).all()
    assert noise.shape == (64, 16000) and ir.shape[0] == 64
    rms = noise[torch.from_numpy(on(lib.AUG_NOISE))].square().mean(dim=1).sqrt()
    assert torch.allclose(rms, torch.ones_like(rms), atol=1e-4)  # double RMS normalisation (background_noise.py:139-141)
    # same seed -> same draws
    pipe.freeze_parameters(42)
    arr2, _, _ = pipe.pack(64, 16000)
    assert np.array_equal(arr, arr2)
    # per-transform dumps follow the App. C schema
    t = pipe.transforms
    assert t[0].transform_parameters["should_apply"].shape == (64,)
    assert t[3].transform_parameters["gain_factors"].dim() == 3 and t[4].transform_parameters["percentile_threshold"].dim() == 2
    with pytest.raises(RuntimeError):
        pipe(torch.zeros(4, 16000))


def test_noise_piece_descriptors_describe_random_background(dropin_modules):
    """The device path of AddBackgroundNoise draws mfpa_noise_piece rows with random_background's RNG calls
    (background_noise.py:64-141).  Assembled here in numpy (gather, mix-up mean, both RMS normalisations -
    what mfpa_noise_assemble does on the GPU), the descriptors must give the rows the host path builds, and
    consume exactly the same random draws."""
    import random

    aug = dropin_modules["aug"]
    g = torch.Generator().manual_seed(7)
    mk = lambda n, a=1.0: {"samples": a * torch.randn(1, n, generator=g), "sample_rate": 8000}
    bg = {"a": [mk(30000), mk(700, 0.2), mk(9000, 4.0)], "b": [[mk(26000, 0.5), mk(40000)], mk(25000)]}
    t = aug.AddBackgroundNoise(bg, p=1.0, sample_rate=8000)
    T, n = 24000, 16
    random.seed(3)
    rows = []
    for r in range(n):
        pcs = t._draw_pieces(r, T)
        assert pcs is not None and sum(p[4] for p in pcs) == T
        rows.append(pcs)
    state = random.getstate()
    bank = torch.cat(t._bank_new).numpy().astype(np.float64)
    random.seed(3)
    for r in range(n):
        host = t.random_background(T)[0].numpy()
        out = np.zeros(T)
        for a, b, q, dst, ln in rows[r]:
            assert q == r
            piece = bank[a: a + ln] if b < 0 else (bank[a: a + ln] + bank[b: b + ln]) / 2
            out[dst: dst + ln] = piece / (np.sqrt(np.mean(piece ** 2)) + 1e-8)
        out = out / (np.sqrt(np.mean(out ** 2)) + 1e-8)
        assert np.abs(out - host).max() < 1e-5 * np.abs(host).max(), r
    assert random.getstate() == state
    assert any(len(pcs) > 1 for pcs in rows) and any(p[1] >= 0 for pcs in rows for p in pcs)   # multi-piece rows, mix-up pairs


@pytest.mark.skipif(not os.path.isdir("/root/reference/augmentation"), reason="reference mount absent")
def test_parameter_draws_match_reference_rng_order(dropin_modules):
    """With the same seeds the drop-in draws exactly the reference's parameters for the
    pass filters, gain and clipping (transform.py:101-114 + each randomize_parameters)."""
    aug = dropin_modules["aug"]
    from oracle import ref_loader

    ns = ref_loader.load()
    x = torch.zeros(16, 1, 8000)
    pairs = [
        (aug.HighPassFilter(0.0, 150.0, p=0.8, sample_rate=8000), ns.pass_filters.HighPassFilter(0.0, 150.0, p=0.8, sample_rate=8000), "cutoff_freq"),
        (aug.LowPassFilter(3000.0, 3999.0, p=0.8, sample_rate=8000), ns.pass_filters.LowPassFilter(3000.0, 3999.0, p=0.8, sample_rate=8000), "cutoff_freq"),
        (aug.Gain(-5.0, 5.0, p=0.8), ns.gain.Gain(-5.0, 5.0, p=0.8), "gain_factors"),
        (aug.Clipping(0, 0.01, p=0.8), ns.clipping.Clipping(0, 0.01, p=0.8), "percentile_threshold"),
    ]
    for mine, ref, key in pairs:
        torch.manual_seed(7)
        mine.draw(16, 8000)
        torch.manual_seed(7)
        ref.transform_parameters = {"should_apply": ref.bernoulli_distribution.sample((16,)).to(torch.bool)}
        ref.randomize_parameters(x[ref.transform_parameters["should_apply"]])
        assert torch.equal(mine.transform_parameters["should_apply"], ref.transform_parameters["should_apply"])
        assert torch.equal(mine.transform_parameters[key], ref.transform_parameters[key]), key


def test_band_filter_and_colored_noise_draws_match_the_reference(dropin_modules):
    """BandPassFilter / AddColoredNoise host side: gate and parameter draws in the reference's RNG order
    (tests/golden/band_params.npz and colored_noise.npz were made by the reference's own classes)."""
    from augmentation.transformations.band_filters import BandPassFilter
    from augmentation.transformations.colored_noise import AddColoredNoise

    gold = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(gold, "band_params.npz"))
    t = BandPassFilter(min_center_frequency=200, max_center_frequency=1900, min_bandwidth_fraction=0.5, max_bandwidth_fraction=1.99,
                       p=0.8, sample_rate=int(g["sample_rate"]))
    torch.manual_seed(int(g["seed"]))
    gate = torch.distributions.Bernoulli(torch.tensor(0.8)).sample(sample_shape=(int(g["batch"]),)).to(torch.bool)
    assert np.array_equal(gate.numpy(), g["should_apply"])
    t.randomize_parameters(torch.zeros(int(gate.sum()), 1, 16))
    assert np.array_equal(t.transform_parameters["center_freq"].numpy(), g["center_freq"])
    assert np.array_equal(t.transform_parameters["bandwidth"].numpy(), g["bandwidth"])
    for bad in (dict(max_center_frequency=100), dict(min_bandwidth_fraction=0.0), dict(max_bandwidth_fraction=2.0)):
        with pytest.raises(ValueError):
            BandPassFilter(**bad)
    c = np.load(os.path.join(gold, "colored_noise.npz"))
    n = AddColoredNoise(p=0.7, sample_rate=int(c["sample_rate"]))
    torch.manual_seed(int(c["seed"]))
    gate = torch.distributions.Bernoulli(torch.tensor(0.7)).sample(sample_shape=(len(c["x"]),)).to(torch.bool)
    assert np.array_equal(gate.numpy(), c["should_apply"])
    n.randomize_parameters(torch.zeros(int(gate.sum()), 1, 16))
    assert np.array_equal(n.transform_parameters["snr_in_db"].numpy(), c["snr_in_db"])
    assert np.array_equal(n.transform_parameters["f_decay"].numpy(), c["f_decay"])
    with pytest.raises(ValueError):
        AddColoredNoise(min_snr_in_db=10.0, max_snr_in_db=5.0)
