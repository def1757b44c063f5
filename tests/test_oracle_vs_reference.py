"""CPU, build container only: the numpy oracle against the LIVE reference (/root/reference imported through
oracle/ref_loader.py) on inputs that are NOT in the committed goldens - fresh seeds each run of the suite would be
non-deterministic, so the seeds below are fixed but differ from oracle/make_golden*.py's.  Skipped where the
reference is not mounted (the GPU box); the committed goldens carry the same pin there."""
import numpy as np
import pytest

from oracle import audfprint_np as O
from oracle import augment_np as A
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")


def test_wave2hashes_equals_reference_wavfile2hashes():
    from musicfpaugment_b200 import synth
    from oracle.make_golden import _wavfile2hashes

    ns = ref_loader.load()
    settings = dict(ns.parameters.afp_settings["audfprint"])
    for seed, n, shifts in ((9001, 64000, 4), (9002, 30011, 1), (9003, 5000, 4)):
        x = synth.music_like(1, n_samples=n, seed=seed).numpy()[0]
        an = ns.peak_extractor.Audfprint_peaks(dict(settings, shifts=shifts))
        want = _wavfile2hashes(ns, an, x)
        got = O.wave2hashes(x, shifts)
        assert np.array_equal(np.asarray(want, np.int32), got), (seed, len(want), len(got))


def test_match_hashes_equals_reference_matcher():
    from musicfpaugment_b200 import synth

    ns = ref_loader.load()
    table, counts, hpid, th = synth.hash_index(400, 300, seed=9100, depth=100)
    ref_ht = ns.hash_table.HashTable()
    ref_ht.table, ref_ht.counts, ref_ht.hashesperid = table.copy(), counts.copy(), hpid.copy()
    ref_ht.names = [f"t{i}" for i in range(400)]
    ora_ht = O.HashTable()
    ora_ht.table, ora_ht.counts, ora_ht.hashesperid = table, counts, hpid
    q, nq, _ = synth.planted_queries(th, 6, n_hashes=250, frac=0.35, seed=9101)
    m = ns.match.Matcher()
    for i in range(len(q)):
        hashes = q[i, : nq[i]]
        want, _ = m.match_hashes(ref_ht, hashes)
        got = O.match_hashes(ora_ht, hashes)
        assert want.shape == got.shape and np.array_equal(want[:, 1], got[:, 1])
        assert sorted(map(tuple, want[:, :4].tolist())) == sorted(map(tuple, got[:, :4].tolist()))
        assert np.array_equal(ref_ht.get_hits(hashes), ora_ht.get_hits(hashes))


def test_augment_chain_equals_reference_transforms():
    from musicfpaugment_b200 import synth
    from oracle.make_golden_augment import reference_chain

    ns = ref_loader.load()
    x = synth.music_like(2, n_samples=24000, seed=9200).numpy()
    ir = synth.impulse_responses(2, length=3000, seed=9201).numpy()
    nz = synth.rms_noise(2, n_samples=24000, seed=9202).numpy()
    pr = synth.augment_params(2, seed=9203)
    for i in range(2):
        prm = dict(fc1=float(pr["fc1"][i]), ir=ir[i], noise=nz[i], snr_db=float(pr["snr_db"][i]),
                   gain_factor=float(10 ** (pr["gain_db"][i] / 20)), clip_p=float(pr["clip_p"][i]), fc2=float(pr["fc2"][i]),
                   fc3=float(pr["fc3"][i]))
        want = reference_chain(ns, x[i], prm)["norm"]
        got = A.augment_chain(x[i], prm)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), i
