"""TEST INFRASTRUCTURE ONLY — numpy restatement of Dejavu's 2-D peak finder.

Follows afp/dejavu/fingerprint.py:94-171 (constants afp/dejavu/variables.py:18-19,
testing/parameters.py:27-34).  Pinned by tests/golden/dejavu.npz.
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
