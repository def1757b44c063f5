"""Drop-in for afp/audfprint/stft.py.

`stft()` keeps the reference's signature and result (complex128 [n_fft/2+1, frames], float64
arithmetic, stft.py:15-62) and runs on the GPU (`mfpa_stft_complex`).  find_peaks does not call it:
the analyzer uses the fused float32 magnitude kernel (`stft_magnitude`, `mfpa_stft_mag`).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def stft(signal, n_fft: int, hop_length: Optional[int] = None, window=None) -> np.ndarray:
    """stft.py:15-62: reflect-pad by n_fft // 2, frames of len(window) samples every `hop_length`,
    times the window, rfft of n_fft points, transposed."""
    import torch

    from musicfpaugment_b200 import runtime

    signal = np.asarray(signal)
    if signal.ndim != 1:
        raise ValueError("stft() takes a 1-D signal")
    if window is None:
        window = np.hanning(n_fft + 2)[1:-1]                     # :40-42
    window = np.asarray(window, dtype=np.float64)
    if hop_length is None:
        hop_length = len(window) // 2                            # :46-47
    if signal.shape[0] == 0:
        raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")   # np.pad, :51
    dev = torch.device("cuda", runtime.get_context().device)
    x = torch.as_tensor(signal.astype(np.float64), device=dev)
    w = torch.as_tensor(window, device=dev)
    return runtime.get_context().stft_complex(x, int(n_fft), int(hop_length), w).cpu().numpy()


def stft_magnitude(signal, n_fft: int = 512, hop_length: int = 256):
    """|stft(signal)| as float32 [n_fft/2+1, frames] from the fused kernel the analyzer uses
    (reference: stft.py:15-62 + abs, peak_extractor.py:257-261)."""
    import torch

    from musicfpaugment_b200 import lib, runtime

    if n_fft != lib.N_FFT or hop_length != lib.HOP:
        raise ValueError(f"the fused kernel is built for n_fft={lib.N_FFT}, hop={lib.HOP}; use stft() for other sizes")
    x = torch.as_tensor(np.asarray(signal, dtype=np.float32)).reshape(1, -1).cuda()
    mag, _ = runtime.get_context().stft_mag(x)
    return mag[0, :, : lib.BINS].T.contiguous().cpu().numpy()
