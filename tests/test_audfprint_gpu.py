"""GPU parity: CUDA audfprint path (through the C ABI) vs the numpy oracle.

Tolerances (BASELINE.json north_star): spectrograms within 1e-4 relative;
peak / landmark / hash sets bit-exact when fed the reference spectrogram;
>= 99.9 % hash agreement end to end.
"""
import numpy as np
import pytest
import torch

from oracle import audfprint_np as O

pytestmark = pytest.mark.gpu


def _lib():
    from musicfpaugment_b200 import lib

    return lib


def _queries(n_music=6, n_noise=2):
    from musicfpaugment_b200 import synth

    return np.concatenate([synth.music_like(n_music).numpy(), synth.white_noise(n_noise).numpy()])


def _rec_to_pklist(rec_row):
    out = []
    for c, r in enumerate(rec_row.tolist()):
        r &= (1 << 64) - 1
        for i in range(r & 0xFF):
            out.append((c, (r >> (8 * (i + 1))) & 0xFF))
    return out


def _params(lib):
    p = lib.afp_defaults()
    assert p.a_dec == O.a_dec()
    return p


def test_stft_mag_matches_oracle(mfpa_ctx):
    lib = _lib()
    X = _queries()
    mag, qmax = mfpa_ctx.stft_mag(torch.from_numpy(X).cuda(), shifts=1)
    mag = mag.cpu().numpy()
    for i, x in enumerate(X):
        ref = O.stft_mag(x)  # [257, N] float64
        got = mag[i, :, :257].T.astype(np.float64)
        assert got.shape == ref.shape
        err = np.abs(got - ref).max() / ref.max()
        assert err < 1e-4, err  # north_star tolerance
        assert err < 2e-6, err  # what the fp32 FFT actually delivers
        assert abs(float(qmax[i]) - ref.max()) / ref.max() < 2e-6


def test_stft_shifts_and_odd_lengths(mfpa_ctx):
    X = _queries(2, 0)[:, :12345]
    mag, qmax = mfpa_ctx.stft_mag(torch.from_numpy(X).cuda().contiguous(), shifts=4)
    mag = mag.cpu().numpy()
    for q in range(2):
        for s, off in enumerate(O.shift_offsets(4)):
            ref = O.stft_mag(X[q][off:])
            got = mag[q * 4 + s, : ref.shape[1], :257].T
            assert np.abs(got - ref).max() / ref.max() < 2e-6


def test_spec_from_mag_layout(mfpa_ctx):
    X = _queries(2, 0)
    mag, qmax = mfpa_ctx.stft_mag(torch.from_numpy(X).cuda(), shifts=1)
    spec = mfpa_ctx.spec_from_mag(mag, qmax, X.shape[1]).cpu().numpy()
    for i, x in enumerate(X):
        ref = O.normalise(O.stft_mag(x))
        assert spec[i].shape == ref.shape
        assert np.abs(spec[i] - ref).max() < 1e-5


@pytest.mark.parametrize("stage", [1, 0])
def test_peaks_bit_exact_from_reference_spectrogram(mfpa_ctx, stage):
    """stage 1: fed the filtered sgram (deterministic float64 ops only);
    stage 0: fed find_peaks' normalised magnitude spectrogram."""
    lib = _lib()
    X = list(_queries()) + [_queries(1, 0)[0][:2000], _queries(1, 0)[0][:12345]]
    p = _params(lib)
    for x in X:
        spec = O.normalise(O.stft_mag(x))
        sg = O.onset_filter(spec)
        want, _ = O.peaks_from_sgram(sg)
        feed = sg if stage == 1 else spec
        rec, npk = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(feed[None].copy()).cuda(), stage, p)
        got = _rec_to_pklist(rec[0].cpu().numpy())
        assert got == want
        assert int(npk[0]) == len(want)


def test_peaks_unnormalised_spectrogram(mfpa_ctx):
    lib = _lib()
    x = _queries(1, 0)[0]
    mag = O.stft_mag(x)
    want = O.peaks_from_mag(mag)[0]
    rec, _ = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(mag[None].copy()).cuda(), 0, _params(lib))
    assert _rec_to_pklist(rec[0].cpu().numpy()) == want


def test_peaks_list_and_mask(mfpa_ctx):
    lib = _lib()
    X = _queries(2, 1)
    specs = np.stack([O.normalise(O.stft_mag(x)) for x in X])
    rec, npk = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(specs).cuda(), 0, _params(lib))
    peaks, n2 = mfpa_ctx.peaks_list(rec)
    mask = mfpa_ctx.peaks_mask(rec).cpu().numpy()
    for i, x in enumerate(X):
        pk, m, _ = O.find_peaks(x)
        assert int(n2[i]) == len(pk) == int(npk[i])
        assert [tuple(r) for r in peaks[i, : len(pk)].cpu().numpy().tolist()] == pk
        assert np.array_equal(mask[i], m)


def test_landmark_hashes_bit_exact(mfpa_ctx):
    lib = _lib()
    X = _queries(3, 2)
    specs = np.stack([O.normalise(O.stft_mag(x)) for x in X])
    p = _params(lib)
    rec, _ = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(specs).cuda(), 0, p)
    h_ref_order, nh = mfpa_ctx.landmark_hashes(rec, p, sorted_rows=False)
    h_sorted, nh2 = mfpa_ctx.landmark_hashes(rec, p, sorted_rows=True)
    for i, x in enumerate(X):
        pk = O.find_peaks(x)[0]
        want = O.landmarks2hashes(O.peaks2landmarks(pk))
        assert int(nh[i]) == len(want) == int(nh2[i])
        assert np.array_equal(h_ref_order[i, : len(want)].cpu().numpy(), want)
        assert np.array_equal(h_sorted[i, : len(want)].cpu().numpy(), O.unique_sorted_hashes(want))


@pytest.mark.parametrize("n_samples", [200000, 300000])
def test_landmark_hashes_long_items(mfpa_ctx, n_samples):
    """Items of 782 frames (lane-per-peak list kernel) and 1172 frames (above its shared-memory bound: the
    lane-per-frame kernel) against the oracle's landmarks, on white noise (dense peaks: 5 per frame)."""
    lib = _lib()
    r = np.random.default_rng(n_samples)
    X = r.standard_normal((2, n_samples)).astype(np.float32)
    X[1] *= np.linspace(0.0, 1.0, n_samples, dtype=np.float32) ** 2
    specs = np.stack([O.normalise(O.stft_mag(x)) for x in X])
    p = _params(lib)
    rec, _ = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(specs).cuda(), 0, p)
    h_ref_order, nh = mfpa_ctx.landmark_hashes(rec, p, sorted_rows=False)
    for i, x in enumerate(X):
        want = O.landmarks2hashes(O.peaks2landmarks(O.find_peaks(x)[0]))
        assert int(nh[i]) == len(want) > 1000
        assert np.array_equal(h_ref_order[i, : len(want)].cpu().numpy(), want)


@pytest.mark.parametrize("shifts", [1, 4])
def test_fingerprint_end_to_end(mfpa_ctx, shifts):
    lib = _lib()
    X = _queries(6, 2)
    out, nh = mfpa_ctx.fingerprint(torch.from_numpy(X).cuda(), shifts, _params(lib))
    out, nh = out.cpu().numpy(), nh.cpu().numpy()
    agree = total = 0
    for i, x in enumerate(X):
        want = {tuple(r) for r in O.wave2hashes(x, shifts).tolist()}
        rows = out[i, : nh[i]]
        key = rows[:, 0].astype(np.int64) << 32 | rows[:, 1]
        assert np.all(np.diff(key) > 0), "rows must be unique and sorted by (time, hash)"
        got = {tuple(r) for r in rows.tolist()}
        agree += len(want & got)
        total += len(want | got)
    assert agree / total >= 0.999, (agree, total)


def test_fingerprint_host_matches_device(mfpa_ctx):
    lib = _lib()
    X = _queries(5, 2)
    p = _params(lib)
    for shifts in (1, 4):
        d_out, d_nh = mfpa_ctx.fingerprint(torch.from_numpy(X).cuda(), shifts, p)
        rows, offs = mfpa_ctx.fingerprint_host(X, shifts, p)
        assert np.array_equal(np.diff(offs), d_nh.cpu().numpy())
        for i in range(len(X)):
            assert np.array_equal(d_out[i, : int(d_nh[i])].cpu().numpy(), rows[offs[i]: offs[i + 1]])
    # tiny capacity -> retried with the exact size
    rows2, offs2 = mfpa_ctx.fingerprint_host(torch.from_numpy(X).pin_memory(), 1, p, rows=torch.empty(8, 2, dtype=torch.int32))
    r1, o1 = mfpa_ctx.fingerprint_host(X, 1, p)
    assert np.array_equal(rows2, r1) and np.array_equal(offs2, o1)


def test_silent_and_tiny_inputs(mfpa_ctx):
    lib = _lib()
    p = _params(lib)
    x = np.zeros((2, 4000), np.float32)
    x[1] = _queries(1, 0)[0][:4000]
    out, nh = mfpa_ctx.fingerprint(torch.from_numpy(x).cuda(), 1, p)
    assert int(nh[0]) == 0  # all-zero input -> no peaks (peak_extractor.py:272-280)
    assert int(nh[1]) == len(O.wave2hashes(x[1]))
    for T in (1, 2, 255, 256, 257, 700):
        xs = _queries(1, 0)[:, :T].copy()
        o, n = mfpa_ctx.fingerprint(torch.from_numpy(xs).cuda(), 1, p)
        want = O.wave2hashes(xs[0])
        assert int(n[0]) == len(want)
        assert np.array_equal(o[0, : len(want)].cpu().numpy(), want)


def test_bad_arguments_raise(mfpa_ctx):
    lib = _lib()
    p = _params(lib)
    p.maxpks = 9
    with pytest.raises(lib.MfpaError):
        mfpa_ctx.fingerprint(torch.zeros(1, 1000, device="cuda"), 1, p)
    with pytest.raises(lib.MfpaError):
        mfpa_ctx.fingerprint(torch.zeros(1, 1000, device="cuda"), 9, _params(lib))


def test_fingerprint_agreement_many_queries(mfpa_ctx):
    """>= 99.9 % of hashes agree with the oracle over a larger seeded sample (fast float path)."""
    from musicfpaugment_b200 import synth

    lib = _lib()
    X = np.concatenate([synth.music_like(40, seed=4321).numpy(), synth.white_noise(8, seed=99).numpy()])
    out, nh = mfpa_ctx.fingerprint(torch.from_numpy(X).cuda(), 1, _params(lib))
    out, nh = out.cpu().numpy(), nh.cpu().numpy()
    agree = total = exact = 0
    for i, x in enumerate(X):
        want = O.wave2hashes(x, 1)
        got = out[i, : nh[i]]
        exact += int(np.array_equal(want, got))
        a, b = {tuple(r) for r in want.tolist()}, {tuple(r) for r in got.tolist()}
        agree += len(a & b)
        total += len(a | b)
    assert agree / total >= 0.999, (agree, total, exact)


def test_scale_invariance(mfpa_ctx):
    """The path is scale invariant (sgram /= max) over 24 orders of magnitude of input level: a property test at the
    ends of float32's range (one query per level; the >= 99.9 % agreement gate is measured by
    test_fingerprint_agreement_many_queries, here a level may move one or two peaks of the single query)."""
    lib = _lib()
    x = _queries(1, 0)[:, :16000]
    for scale in (1e-12, 1e-4, 1.0, 1e12):
        xs = (x.astype(np.float64) * scale).astype(np.float32)
        out, nh = mfpa_ctx.fingerprint(torch.from_numpy(xs).cuda(), 1, _params(lib))
        want = O.wave2hashes(xs[0], 1)
        got = out[0, : int(nh[0])].cpu().numpy()
        a, b = {tuple(r) for r in want.tolist()}, {tuple(r) for r in got.tolist()}
        assert len(a & b) >= 0.98 * len(a | b), (scale, len(a), len(b))


def test_float32_picker_equals_float64_picker(mfpa_ctx):
    """The hot path's float32 picker and the float64 picker (MFPA_OPT_PEAKS_F64) fed the same float32
    magnitudes produce the same records on music-like and white-noise queries, all shift offsets."""
    from musicfpaugment_b200 import synth

    lib = _lib()
    p = _params(lib)
    X = torch.cat([synth.music_like(48, seed=31), synth.white_noise(16, seed=32), 1e-9 * synth.music_like(4, seed=33)]).cuda()
    X[5, 20000:] = 0.0  # digital silence: exercises the 1e-6 floor and the plateau tie-breaks
    T = X.shape[1]
    same = total = 0
    for shifts in (1, 4):
        mag, qmax = mfpa_ctx.stft_mag(X, shifts)
        rec32, n32 = mfpa_ctx.audfprint_peaks(mag, qmax, T, shifts, p)
        mfpa_ctx.set_option(lib.OPT_PEAKS_F64, 1)
        try:
            rec64, n64 = mfpa_ctx.audfprint_peaks(mag, qmax, T, shifts, p)
        finally:
            mfpa_ctx.set_option(lib.OPT_PEAKS_F64, 0)
        same += int((rec32 == rec64).sum())
        total += rec32.numel()
        assert int((n32 - n64).abs().sum()) <= 2
    assert same / total >= 0.9995, (same, total)


def test_pcm16_host_entry_equals_float_entry(mfpa_ctx):
    """mfpa_fingerprint_host_pcm16(x) == mfpa_fingerprint_host(float32(x) / 32768), bit for bit."""
    from musicfpaugment_b200 import lib, synth

    x = synth.music_like(5, seed=21).numpy()
    x16 = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
    x16[3] = 0  # a silent query
    p = lib.afp_defaults()
    for shifts in (1, 4):
        rows_f, offs_f = mfpa_ctx.fingerprint_host(x16.astype(np.float32) / np.float32(32768.0), shifts, p)
        rows_i, offs_i = mfpa_ctx.fingerprint_host(x16, shifts, p)
        assert np.array_equal(offs_f, offs_i) and np.array_equal(rows_f, rows_i)
        assert offs_i[-1] > 0


def test_picker_edge_cases_on_crafted_sgrams(mfpa_ctx):
    """SURVEY.md App. E cases on hand-made filtered spectrograms (stage 1 entry, bit-exact): plateaus
    (locmax keeps the LAST element, sort ties favour the higher bin), more than 5 candidates in a frame,
    a peak repeated in the same bin of the next frame (backward pass clears the later one), fewer than 10
    frames (initial threshold from min(10, n) columns), a single frame, and an all-equal spectrogram."""
    lib = _lib()
    p = _params(lib)
    rng = np.random.default_rng(99)
    cases = []
    sg = np.zeros((256, 40))
    sg[10:14, :] = 3.0                       # a 4-bin plateau in every frame
    sg[100, 5] = sg[100, 6] = 9.0            # same bin, consecutive frames
    sg[200:206, 20] = [1, 2, 2, 2, 1, 0]     # plateau inside a bump
    cases.append(sg)
    sg = rng.standard_normal((256, 30))
    sg[::8, 3] += 50.0                       # 32 strong candidates in one frame -> top 5 kept
    cases.append(sg)
    cases.append(rng.standard_normal((256, 4)))    # fewer than 10 frames
    cases.append(rng.standard_normal((256, 1)))    # a single frame
    cases.append(np.full((256, 12), 0.25))         # everything equal
    cases.append(np.round(rng.standard_normal((256, 64)) * 2) / 2)  # heavy ties
    for i, sg in enumerate(cases):
        want, _ = O.peaks_from_sgram(sg)
        rec, npk = mfpa_ctx.audfprint_peaks_from_spec(torch.from_numpy(np.ascontiguousarray(sg)[None]).cuda(), 1, p)
        got = _rec_to_pklist(rec[0].cpu().numpy())
        assert got == want, i
        h, nh = mfpa_ctx.landmark_hashes(rec, p, sorted_rows=False)
        wh = O.landmarks2hashes(O.peaks2landmarks(want))
        assert int(nh[0]) == len(wh) and np.array_equal(h[0, : len(wh)].cpu().numpy(), wh), i
