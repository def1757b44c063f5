"""Drop-in for afp/audfprint/stft.py: `stft()` is only called by find_peaks, which here runs the
fused CUDA STFT-magnitude kernel; this module keeps the import path and offers the magnitude."""
from __future__ import annotations

import numpy as np


def stft_magnitude(signal, n_fft: int = 512, hop_length: int = 256):
    """|stft(signal)| as float32 [n_fft/2+1, frames] computed on the GPU (reference: stft.py:15-62 + abs)."""
    import torch

    from musicfpaugment_b200 import lib, runtime

    if n_fft != lib.N_FFT or hop_length != lib.HOP:
        raise ValueError(f"the CUDA path is built for n_fft={lib.N_FFT}, hop={lib.HOP}")
    x = torch.as_tensor(np.asarray(signal, dtype=np.float32)).reshape(1, -1).cuda()
    mag, _ = runtime.get_context().stft_mag(x)
    return mag[0, :, : lib.BINS].T.contiguous().cpu().numpy()


def stft(*args, **kwargs):
    raise NotImplementedError(
        "complex STFT output is not part of the B200 hot path; use stft_magnitude() "
        "(Audfprint_peaks.find_peaks calls the fused kernel directly)")
