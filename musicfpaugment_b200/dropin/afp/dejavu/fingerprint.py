"""Drop-in for afp/dejavu/fingerprint.py: get_2D_peaks on the GPU, generate_hashes as in the
reference (host SHA-1 of short strings).  `fingerprint()` itself needs matplotlib's
mlab.specgram PSD, which is a "next" row of SURVEY.md §8(f) and raises for now."""
from __future__ import annotations

import hashlib
from operator import itemgetter
from typing import List, Tuple

import numpy as np

from .variables import FINGERPRINT_REDUCTION, MAX_HASH_TIME_DELTA, MIN_HASH_TIME_DELTA, PEAK_NEIGHBORHOOD_SIZE


def get_2D_peaks(arr2D, plot: bool = False, amp_min: float = 50):
    """-> (list[(freq, time)], mask float64).  fingerprint.py:94-171."""
    import torch

    from musicfpaugment_b200 import runtime

    arr = np.ascontiguousarray(arr2D)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    t = torch.from_numpy(arr[None]).cuda()
    mask, peaks, n = runtime.get_context().dejavu_peaks(t, amp_min=float(amp_min), neighborhood=PEAK_NEIGHBORHOOD_SIZE,
                                                      cap=int(arr.size))
    pk = peaks[0, : int(n[0])].cpu().numpy()
    return [(f, t_) for f, t_ in zip(pk[:, 0], pk[:, 1])], mask[0].cpu().numpy().astype(np.float64)


def generate_hashes(peaks: List[Tuple[int, int]], fan_value: int = 3):
    """fingerprint.py:174-213 (pairs each peak with the next fan_value-1 peaks in time order)."""
    peaks.sort(key=itemgetter(1))
    hashes = []
    for i in range(len(peaks)):
        for j in range(1, fan_value):
            if (i + j) < len(peaks):
                f1, f2 = peaks[i][0], peaks[i + j][0]
                t1, t2 = peaks[i][1], peaks[i + j][1]
                dt = t2 - t1
                if MIN_HASH_TIME_DELTA <= dt <= MAX_HASH_TIME_DELTA:
                    h = hashlib.sha1(f"{str(f1)}|{str(f2)}|{str(dt)}".encode("utf-8"))
                    hashes.append((h.hexdigest()[0:FINGERPRINT_REDUCTION], t1))
    return hashes


def fingerprint(*args, **kwargs):
    raise NotImplementedError("Dejavu's mlab.specgram front end is not built yet (SURVEY.md §8(f) item 2); "
                              "get_2D_peaks and generate_hashes are")
