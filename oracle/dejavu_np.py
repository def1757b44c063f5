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


# ---------------------------------------------------------------------------------------------------
# Matching: CommonDatabase.return_matches (afp/dejavu/postgres_database.py:180-229) and
# Dejavu.align_matches (afp/dejavu/dejavu.py:312-378), restated over a plain list of fingerprint rows.
# Pinned by tests/golden/dejavu_match.npz, which oracle/make_golden_dejavu_match.py produces by running the
# reference's own two methods against a dict-backed stand-in for the Postgres cursor.
# ---------------------------------------------------------------------------------------------------
def return_matches(rows, hashes):
    """rows: iterable of (hash hex str, song id, offset) = the fingerprints table; hashes: iterable of
    (hash, offset) of the query.  -> ([(song id, db offset - query offset), ...], {song id: matched rows}),
    each distinct query hash looked up once (:198-204) and every matching row paired with every offset the
    query saw that hash at (:224-226)."""
    table = {}
    for h, sid, off in rows:
        table.setdefault(h.upper(), []).append((int(sid), int(off)))
    mapper = {}
    for h, off in hashes:
        mapper.setdefault(h.upper(), []).append(int(off))
    results, dedup = [], {}
    for h, q_offsets in mapper.items():
        for sid, off in table.get(h, []):
            dedup[sid] = dedup.get(sid, 0) + 1
            for qo in q_offsets:
                results.append((sid, off - qo))
    return results, dedup


def align_top(matches):
    """The first row of align_matches' `songs_matches` (:330-345): per song the offset difference with the most
    votes (first maximum in ascending offset order), songs ordered by that count descending with Python's stable
    sort, so ties keep ascending song id.  -> (song id, offset difference, votes) or None."""
    counts = {}
    for sid, diff in matches:
        counts[(int(sid), int(diff))] = counts.get((int(sid), int(diff)), 0) + 1
    best = None
    for (sid, diff), c in sorted(counts.items()):
        if best is None or c > best[2]:
            best = (sid, diff, c)
    return best


# ---------------------------------------------------------------------------------------------------
# Evaluation (testing/metrics.py:10-192).  For every non-zero (b, i, j) of one mask the reference multiplies a 3x3
# window of the OTHER mask by a kernel that is zero off its centre.  In the interior and at the high edges the centre
# lands on (i, j); at i == 0 or j == 0 the window is cut on the low side while the kernel is cut on the HIGH side
# (`window[f : f + 2] * kernel[:2]`, :44-47, :72-75), so the centre lands on i + 1 resp. j + 1 - a quirk of the
# reference, reproduced.  Pinned by tests/golden/metrics.npz (reference classes on random masks).
# psnr restates torchmetrics.PeakSignalNoiseRatio with data_range=None (third-party, absent: parity unpinned):
# range = max(target, 0) - min(target, 0), 10 log10(range^2 / mse).
# ---------------------------------------------------------------------------------------------------
def _looked_at(other, idx):
    b, i, j = idx
    i2 = np.where((i == 0) & (other.shape[1] > 1), i + 1, i)
    j2 = np.where((j == 0) & (other.shape[2] > 1), j + 1, j)
    return other[b, i2, j2]


def recall(predicted, gt):
    predicted, gt = np.asarray(predicted, np.float64), np.asarray(gt, np.float64)
    idx = np.nonzero(gt)
    return 0.0 if len(idx[0]) == 0 else float(_looked_at(predicted, idx).sum() / len(idx[0]))


def precision(predicted, gt):
    predicted, gt = np.asarray(predicted, np.float64), np.asarray(gt, np.float64)
    idx = np.nonzero(predicted)
    return 0.0 if len(idx[0]) == 0 else float(_looked_at(gt, idx).sum() / len(idx[0]))


def f1score(predicted, gt):
    p, r = precision(predicted, gt), recall(predicted, gt)
    return 0.0 if np.isclose(p + r, 0.0) else float(2.0 * p * r / (p + r))


def psnr(pred, target):
    pred, target = np.asarray(pred, np.float64), np.asarray(target, np.float64)
    rng = max(target.max(), 0.0) - min(target.min(), 0.0)
    return float(10.0 * np.log10(rng ** 2 / np.mean((pred - target) ** 2)))
