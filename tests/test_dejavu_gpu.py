"""GPU parity: Dejavu get_2D_peaks kernel vs golden vectors from the reference and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import dejavu_np as D

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_cases(mfpa_ctx):
    g = np.load(os.path.join(GOLD, "dejavu.npz"))
    for i in range(4):
        arr = g[f"arr{i}"]
        mask, peaks, n = mfpa_ctx.dejavu_peaks(torch.from_numpy(arr[None].copy()).cuda(), amp_min=float(g[f"amp_min{i}"]))
        assert np.array_equal(mask[0].cpu().numpy(), g[f"mask{i}"]), i
        assert np.array_equal(peaks[0, : int(n[0])].cpu().numpy(), g[f"peaks{i}"]), i


def test_batch_vs_oracle_full_size(mfpa_ctx):
    """[257, 249] log spectrograms (the size fingerprint() produces for 8 s queries), float64 and float32."""
    from musicfpaugment_b200 import synth
    from oracle import audfprint_np as O

    x = synth.music_like(3, seed=71).numpy()
    arrs = np.stack([D.log_spectrogram(O.normalise(O.stft_mag(xi)[:, :249] ** 2)) for xi in x])
    arrs[1] = np.round(arrs[1])           # plateaus
    arrs[2, 100:160, 50:120] = 0.0        # zero background block
    for dt in (np.float64, np.float32):
        a = arrs.astype(dt)
        mask, peaks, n = mfpa_ctx.dejavu_peaks(torch.from_numpy(a).cuda(), amp_min=20.0)
        for i in range(3):
            pk, m = D.get_2d_peaks(a[i], amp_min=20.0)
            assert np.array_equal(mask[i].cpu().numpy(), m.astype(np.uint8)), (dt, i)
            assert [tuple(r) for r in peaks[i, : int(n[i])].cpu().numpy().tolist()] == pk
            assert len(pk) > 10


def test_fingerprint_front_end_vs_oracle(mfpa_ctx):
    """fingerprint.py:60-79 on the GPU (float32) vs the float64 oracle: PSD/max within 1e-4 of the peak,
    log array within 1e-3 (10 ln of a 1e-6-floored value), and the peak sets agree >= 99 %."""
    from musicfpaugment_b200 import synth
    from oracle import dejavu_np as D

    X = synth.music_like(4, seed=50, device=torch.device("cuda"))
    psd = mfpa_ctx.dejavu_psd(X)
    arr = mfpa_ctx.dejavu_log(psd)
    mask, peaks, n = mfpa_ctx.dejavu_peaks(arr, cap=257 * 249)
    assert psd.shape == (4, 257, 249)
    for i in range(4):
        want_arr, want_psd = D.dejavu_fingerprint_arr(X[i].cpu().numpy())
        assert np.abs(psd[i].cpu().numpy() - want_psd).max() < 1e-4
        assert np.abs(arr[i].cpu().numpy() - want_arr).max() < 2e-3
        want_pk, _ = D.get_2d_peaks(want_arr)
        got_pk = set(map(tuple, peaks[i, : int(n[i])].cpu().numpy().tolist()))
        agree = len(got_pk & set(want_pk)) / max(1, len(got_pk | set(want_pk)))
        assert agree >= 0.99, agree
        # the peak finder itself is exact on the GPU's own array
        same_pk, _ = D.get_2d_peaks(arr[i].cpu().numpy())
        assert got_pk == set(same_pk)
    # short input: one segment; too short raises
    assert mfpa_ctx.dejavu_psd(X[:1, :700]).shape == (1, 257, 1)
    from musicfpaugment_b200 import lib

    with pytest.raises(lib.MfpaError):
        mfpa_ctx.dejavu_psd(X[:1, :400])
