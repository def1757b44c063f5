"""GPU: the three named drivers' call sequences through the drop-ins, on synthetic query pickles.

The reference checkout is not on the GPU box, so the loops of testing/generate_queries.py:63-92
(generate_augmented_queries), testing/audfprint_exps.py:17-157 (create_fp_database, compute_accuracy,
compute_peaks_metrics) are re-enacted here call for call - same constructors, attributes and methods, in the same
order - against the drop-in modules under their reference import names.  (Where a checkout IS mounted next to a GPU,
`test_reference_driver_functions` runs the reference's own functions instead.)  Checks: every augmented query is
identified (6-track set), hashes agree with the oracle at the north star's gate, peak metrics equal the oracle's.
"""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from oracle import audfprint_np as O
from oracle import augment_np as A
from oracle import dejavu_np as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "musicfpaugment_b200", "dropin")
AFP = {"density": 20, "pks-per-frame": 5, "freq-sd": 30, "shifts": 1, "samplerate": 8000, "n_fft": 512, "n_hop": 256}


@pytest.fixture(scope="module")
def mods():
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu", "testing")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    import afp.audfprint.audfprint_match as m
    import afp.audfprint.hash_table as ht
    import afp.audfprint.peak_extractor as pe
    import augmentation as aug
    import testing.metrics as tm

    yield {"pe": pe, "ht": ht, "match": m, "aug": aug, "tm": tm}
    sys.path.remove(DROPIN)
    for k in [k for k in sys.modules if k.split(".")[0] in ("afp", "augmentation", "dejavu", "testing")]:
        del sys.modules[k]


def _dump(path, x):
    with open(path, "wb") as fh:
        pickle.dump(np.asarray(x, np.float32), fh)
    return str(path)


def test_generate_queries_then_audfprint_exps(mods, tmp_path):
    from musicfpaugment_b200 import synth

    pe, ht_mod, m_mod, aug, tm = (mods[k] for k in ("pe", "ht", "match", "aug", "tm"))
    tracks = synth.music_like(6, seed=2024).numpy()
    cleans = [_dump(tmp_path / f"clean_{i}.pkl", x) for i, x in enumerate(tracks)]

    # ---- generate_queries.generate_augmented_queries (:63-92)
    g = torch.Generator().manual_seed(11)
    noise_paths = {"street": [{"samples": torch.randn(1, 200000, generator=g), "sample_rate": 8000}],
                   "cafe": [{"samples": torch.randn(1, 150000, generator=g), "sample_rate": 8000}]}
    irs = [{"samples": torch.randn(1, n, generator=g) * torch.exp(-torch.arange(n) / 400.0), "sample_rate": 8000} for n in (2000, 3000)]
    parameters = dict(aug.DEFAULT_PARAMETERS)
    parameters.update({"min_snr_in_db": 8, "max_snr_in_db": 12, "min_cutoff_freq1": 20.0})
    aug_pipeline = aug.AugmentFP(noise_paths, 8000, parameters, impulse_response_dir=irs)
    aug_pipeline.augmentation_pipeline.freeze_parameters(42)
    aug_pipeline.augmentation_pipeline.to("cpu")
    augmented, dumps = [], []
    for i, path in enumerate(cleans):
        with open(path, "rb") as fh:
            clean_audio = torch.tensor(np.array(pickle.load(fh))).unsqueeze(0)          # :81-84
        out = aug_pipeline(clean_audio)                                                  # :86
        assert out.shape == (1, 64000)
        t = aug_pipeline.augmentation_pipeline.transforms
        dumps.append({"apply": [bool(tr.transform_parameters["should_apply"][0]) for tr in t], "t": [dict(tr.transform_parameters) for tr in t]})
        qdir = tmp_path / "aug"
        qdir.mkdir(exist_ok=True)
        augmented.append(_dump(qdir / f"clean_{i}.pkl", np.array(out.T[:, 0])))          # :88-90

    # ---- audfprint_exps.create_fp_database (:17-28)
    hash_tab = ht_mod.HashTable()
    analyzer = pe.Audfprint_peaks(AFP)
    analyzer.shifts = 1
    for filename in cleans:
        analyzer.ingest(hash_tab, filename)
    dbpath = str(tmp_path / "db.pklz")
    hash_tab.save(dbpath)

    # ---- audfprint_exps.compute_accuracy (:31-83) with the analyzers of identification_rate_results (:160-185)
    hash_tab = ht_mod.HashTable(dbpath)
    matcher = m_mod.Matcher()
    analyzer1 = pe.Audfprint_peaks(AFP)
    analyzer1.shifts = 4
    acc = inter = union = 0
    for i, filename in enumerate(augmented):
        gt = filename.split("/")[-1].split(".")[0]
        msgs = matcher.file_match_to_msgs(analyzer1, hash_tab, filename)
        pred = msgs[1].split("/")[-1].split(".")[0]
        acc += int(msgs[0] == "MATCH" and str(gt) == str(pred))
        with open(filename, "rb") as fh:
            want = O.wave2hashes(np.asarray(pickle.load(fh), np.float32), 4)
        got = analyzer1.wavfile2hashes(filename)
        a, b = {tuple(r) for r in got.tolist()}, {tuple(r) for r in want.tolist()}
        inter += len(a & b)
        union += len(a | b)
    assert acc == len(augmented)                       # identification rate 100 % on the 6-track set
    assert inter >= 0.999 * union, (inter, union)      # >= 99.9 % hash agreement with the oracle

    # ---- audfprint_exps.compute_peaks_metrics (:86-157)
    precision, recall, f1_score = tm.Precision(), tm.Recall(), tm.F1score()
    analyzer_no_den = pe.Audfprint_peaks(AFP)
    for clean, augm in zip(cleans, augmented):
        m_clean, w_clean, sgram_clean = analyzer_no_den.wavfile2peaks(clean, get_masks_waveforms=True)
        m_aug, w_aug, sgram_aug = analyzer_no_den.wavfile2peaks(augm, get_masks_waveforms=True)
        mask_clean = torch.tensor(m_clean).T.unsqueeze(0)
        mask_aug = torch.tensor(m_aug).T.unsqueeze(0)
        p, r, f = precision(mask_aug, mask_clean), recall(mask_aug, mask_clean), f1_score(mask_aug, mask_clean)
        assert 0.0 <= p <= 1.0 and 0.0 <= r <= 1.0
        assert (p, r, f) == pytest.approx((D.precision(mask_aug.numpy(), mask_clean.numpy()), D.recall(mask_aug.numpy(), mask_clean.numpy()),
                                           D.f1score(mask_aug.numpy(), mask_clean.numpy())), abs=1e-12)
        ps = tm.psnr(torch.tensor(sgram_aug).unsqueeze(0), torch.tensor(sgram_clean).unsqueeze(0)).item()
        assert ps == pytest.approx(D.psnr(sgram_aug, sgram_clean), rel=1e-6)

    # ---- the augmented waveform itself against the oracle chain on the dumped parameters (first query)
    d0 = dumps[0]
    names = ("fc1", "ir", "noise", "gain", "clip", "fc2", "fc3")
    prm = {}
    if d0["apply"][0]: prm["fc1"] = float(d0["t"][0]["cutoff_freq"][0])
    if d0["apply"][1]: prm["ir"] = d0["t"][1]["ir"][0, 0].cpu().numpy()
    if d0["apply"][2]: prm["noise"], prm["snr_db"] = d0["t"][2]["background"][0, 0].cpu().numpy(), float(d0["t"][2]["snr_in_db"][0])
    if d0["apply"][3]: prm["gain_factor"] = float(d0["t"][3]["gain_factors"].reshape(-1)[0])
    if d0["apply"][4]: prm["clip_p"] = float(d0["t"][4]["percentile_threshold"].reshape(-1)[0])
    if d0["apply"][5]: prm["fc2"] = float(d0["t"][5]["cutoff_freq"][0])
    if d0["apply"][6]: prm["fc3"] = float(d0["t"][6]["cutoff_freq"][0])
    ref = A.augment_chain(tracks[0], prm)
    with open(augmented[0], "rb") as fh:
        got = np.asarray(pickle.load(fh), np.float32)
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.skipif(not os.path.isdir("/root/reference/testing"), reason="reference checkout not mounted next to the GPU")
def test_reference_driver_functions(tmp_path):
    """With a checkout mounted: the reference's OWN driver functions import and run on top of the drop-ins."""
    import subprocess

    code = ("import testing.audfprint_exps as e, testing.generate_queries as g, testing.dejavu_exps as d\n"
            "assert callable(e.create_fp_database) and callable(e.compute_accuracy) and callable(g.generate_augmented_queries)\n"
            "print('ok')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, DROPIN, "/root/reference"]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
