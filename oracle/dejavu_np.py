"""TEST INFRASTRUCTURE ONLY — numpy restatement of Dejavu's 2-D peak finder.

Follows afp/dejavu/fingerprint.py:94-171 (constants afp/dejavu/variables.py:18-19,
testing/parameters.py:27-34).  Pinned by tests/golden/dejavu.npz.

`specgram_psd` restates the third-party front end of fingerprint() (fingerprint.py:60-66):
matplotlib.mlab.specgram (matplotlib 3.8.0, poetry.lock:2097-2098, NOT installed in the image, so the
reference's fingerprint() itself cannot run here).  Its published definition (mlab._spectral_helper,
mode "psd", detrend none, one-sided, scale_by_freq) is the Welch-segment periodogram that
scipy.signal.spectrogram(scaling="density", mode="psd", detrend=False) also implements;
tests/test_oracle_golden.py pins this restatement against scipy's.  Parity status of that one
function: pinned to the shared definition, unpinned against matplotlib proper.
"""
from __future__ import annotations

import numpy as np

PEAK_NEIGHBORHOOD_SIZE = 10
AMP_MIN = 50


def _sliding(a: np.ndarray, k: int, axis: int, fn) -> np.ndarray:
    """fn-reduce over windows of length k along `axis` of an already padded array."""
    n = a.shape[axis] - k + 1
    out = None
    for s in range(k):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(s, s + n)
        w = a[tuple(sl)]
        out = w.copy() if out is None else fn(out, w)
    return out


def maximum_filter_square(arr: np.ndarray, r: int) -> np.ndarray:
    """scipy.ndimage.maximum_filter(arr, footprint=ones((2r+1, 2r+1))) with its default
    boundary mode 'reflect' (edge sample repeated = numpy 'symmetric').  fingerprint.py:128.
    iterate_structure(generate_binary_structure(2, 2), 10) is the full 21 x 21 square (:118-125)."""
    p = np.pad(arr, r, mode="symmetric")
    k = 2 * r + 1
    return _sliding(_sliding(p, k, 0, np.maximum), k, 1, np.maximum)


def erode_square(mask: np.ndarray, r: int) -> np.ndarray:
    """binary_erosion(mask, structure=ones((2r+1, 2r+1)), border_value=1).  fingerprint.py:131-134."""
    p = np.pad(mask.astype(bool), r, mode="constant", constant_values=True)
    k = 2 * r + 1
    return _sliding(_sliding(p, k, 0, np.logical_and), k, 1, np.logical_and)


def get_2d_peaks(arr2d: np.ndarray, amp_min: float = AMP_MIN, r: int = PEAK_NEIGHBORHOOD_SIZE):
    """-> (list[(freq, time)] in row-major order, mask float64).  fingerprint.py:94-171."""
    arr2d = np.asarray(arr2d)
    local_max = maximum_filter_square(arr2d, r) == arr2d
    eroded_bg = erode_square(arr2d == 0, r)
    det = (local_max != eroded_bg) & (arr2d > amp_min)
    f, t = np.nonzero(det)
    return list(zip(f.tolist(), t.tolist())), det.astype(np.float64)


def log_spectrogram(spec: np.ndarray) -> np.ndarray:
    """10*ln(max(spec, max/1e6)) - mean.  fingerprint.py:77-79 (natural log)."""
    s = 10 * np.log(np.maximum(spec, np.max(spec) / 1e6))
    return s - np.mean(s)


def specgram_psd(x: np.ndarray, nfft: int = 512, fs: float = 8000.0, noverlap: int = 256) -> np.ndarray:
    """mlab.specgram(x, NFFT=nfft, Fs=fs, window=mlab.window_hanning, noverlap=noverlap)[0]
    (fingerprint.py:60-66) -> float64 [nfft/2+1, n_seg], n_seg = (len(x) - noverlap) // (nfft - noverlap).
    Segments start every nfft-noverlap samples, no padding; window np.hanning(nfft) (symmetric);
    |rfft|^2, interior bins doubled (one-sided), divided by fs and by sum(window^2)."""
    x = np.asarray(x, dtype=np.float64)
    step = nfft - noverlap
    n_seg = (len(x) - noverlap) // step
    if n_seg <= 0:
        return np.zeros((nfft // 2 + 1, 0))
    win = np.hanning(nfft)
    idx = np.arange(nfft)[:, None] + step * np.arange(n_seg)[None, :]
    seg = x[idx] * win[:, None]
    spec = np.fft.rfft(seg, axis=0)
    psd = (np.conj(spec) * spec).real
    psd[1:-1] *= 2.0
    return psd / fs / (win ** 2).sum()


def dejavu_fingerprint_arr(x: np.ndarray, denoise=None):
    """fingerprint() up to the array handed to get_2D_peaks (fingerprint.py:60-79): PSD, / max,
    [unet(.)**2], 10*ln(max(., max/1e6)) - mean.  Returns (arr2D, specgram)."""
    arr = specgram_psd(x)
    arr = arr / arr.max()
    if denoise is not None:
        arr = np.asarray(denoise(arr.astype(np.float32)), dtype=np.float32) ** 2
    return log_spectrogram(arr), arr
