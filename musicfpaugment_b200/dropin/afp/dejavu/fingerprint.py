"""Drop-in for afp/dejavu/fingerprint.py: fingerprint() = mlab.specgram PSD -> / max -> [UNet ** 2] ->
10 ln - mean -> get_2D_peaks, all on the GPU; generate_hashes as in the reference (host SHA-1 of
short strings)."""
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


_unet = None
_unet_state_dict = None
UNET_CHECKPOINT = "/workspace/src/training/checkpoints/unet_lr_0.001_BS_128/best_epoch.pt"  # fingerprint.py:29


def set_unet_state_dict(state_dict) -> None:
    """Weights for `denoising=True` (the reference loads a hard-coded checkpoint at import, :27-31)."""
    global _unet, _unet_state_dict
    _unet_state_dict = state_dict
    if _unet is not None:
        _unet.close()
        _unet = None


def _denoiser(ctx):
    global _unet
    if _unet is None:
        from musicfpaugment_b200 import lib

        sd = _unet_state_dict
        if sd is None:
            import os

            import torch

            sd = torch.load(os.environ.get("MFPA_UNET_CHECKPOINT", UNET_CHECKPOINT), map_location="cpu")["model_state_dict"]
        _unet = lib.UNetDenoiser(ctx, sd)
    return _unet


def fingerprint(channel_samples, Fs: float = 8000, wsize: int = 512, n_hop: int = 256, fan_value: int = 3,
                amp_min: int = 50, denoising: bool = False, denoising_model: str = "unet", get_masks="False"):
    """fingerprint.py:34-91.  Defaults are afp_settings["dejavu"] (testing/parameters.py:27-34)."""
    import torch

    from musicfpaugment_b200 import runtime

    if denoising:
        assert denoising_model in ["unet", "demucs"]
        if denoising_model == "demucs":
            raise NotImplementedError("denoising_model='demucs' is outside the B200 hot path (SURVEY.md section 8f)")
    if wsize != 512 or n_hop != 256:
        raise ValueError(f"the CUDA path is built for NFFT=512, noverlap=256 (testing/parameters.py:27-34), got {wsize}/{n_hop}")
    ctx = runtime.get_context()
    x = torch.as_tensor(np.asarray(channel_samples), dtype=torch.float32).reshape(1, -1).cuda()
    psd = ctx.dejavu_psd(x)                      # specgram / max (Fs only scales the PSD and cancels here)
    if denoising is True:
        psd = _denoiser(ctx).forward(psd)        # arr2D = unet(arr2D) ... ** 2 (:70-75)
        arr = ctx.dejavu_log(psd, square=True)
        specgram = (psd * psd)[0].cpu().numpy()
    else:
        arr = ctx.dejavu_log(psd)
        specgram = psd[0].cpu().numpy()
    mask, peaks, n = ctx.dejavu_peaks(arr, amp_min=float(amp_min), neighborhood=PEAK_NEIGHBORHOOD_SIZE, cap=int(arr[0].numel()))
    pk = peaks[0, : int(n[0])].cpu().numpy()
    local_maxima = [(f, t_) for f, t_ in zip(pk[:, 0], pk[:, 1])]
    hashes = generate_hashes(local_maxima, fan_value=fan_value)
    if get_masks is True:
        return hashes, mask[0].cpu().numpy().astype(np.float64), specgram
    return hashes
