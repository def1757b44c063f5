"""TEST INFRASTRUCTURE ONLY — numpy restatement of the AugmentFP arithmetic.

Follows the `apply_transform` bodies under augmentation/transformations/ of the
reference (cited per function) for one query at a time (the reference drivers
call the chain with B = 1, testing/generate_queries.py:81-86).  Parameter
sampling and file IO are NOT restated: parity is on dumped parameters
(SURVEY.md §8c, App. C).

Parity status
  * everything except the FIR taps: pinned by tests/golden/augment.npz, produced by
    the reference's own transform classes (oracle/make_golden_augment.py);
  * `julius.lowpass_filter` (julius 0.2.7, pyproject.toml:42 / poetry.lock:1503-1504,
    NOT vendored in /root/reference, not installed, no network): restated from its
    published algorithm (windowed-sinc, zeros = 8, Hann window, replicate padding,
    conv1d) — **parity unpinned** for that one function.  The golden vectors use the
    same restatement (oracle/ref_loader.py::_julius_lowpass_filter, torch float32).
    The taps are pinned independently against scipy.signal.firwin (tests/test_oracle_golden.py);
  * `julius.bandpass_filter` (`bandpass`, `bandstop` below; same package, same situation): restated,
    **parity unpinned**; the parameter draws of the band filters ARE pinned (tests/golden/band_params.npz).

Stage boundaries are float32 like the reference's tensors; inside a stage the
oracle accumulates in float64 (differences to torch's float32 kernels are ~1e-6,
two orders below the 1e-4 tolerance of the north star).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------ FIR (julius restated)
def lowpass_taps(cutoff: float, zeros: float = 8.0, width: float | None = None) -> np.ndarray:
    """julius.lowpass.LowPassFilters.__init__ for one cutoff (fraction of the sample rate):
    half = int(zeros / cutoff / 2); hann(2*half+1, periodic=False) * 2c * sinc(2c*pi*t), sum -> 1.
    float32 like torch (window, argument and sinc are float32 tensors there).  `width`: the cut-off that sets the
    window (a bank of several cut-offs takes half from the lowest positive one)."""
    if cutoff < 0:
        raise ValueError("Minimum cutoff must be larger than zero.")
    if cutoff > 0.5:
        raise ValueError("A cutoff above 0.5 does not make sense.")
    if cutoff == 0:
        raise ValueError("min() arg is an empty sequence")
    half = int(zeros / (cutoff if width is None else width) / 2)
    n = 2 * half + 1
    k = np.arange(n, dtype=np.float64)
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))).astype(F32) if n > 1 else np.ones(1, F32)
    t = np.arange(-half, half + 1).astype(F32)
    arg = (t * F32(2 * cutoff * np.pi)).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(arg == 0, F32(1), np.sin(arg, dtype=F32) / arg).astype(F32)
    h = (F32(2 * cutoff) * win * sinc).astype(F32)
    return (h / h.sum(dtype=F32)).astype(F32)


def lowpass(x: np.ndarray, cutoff: float, width: float | None = None) -> np.ndarray:
    """julius.lowpass_filter(x, cutoff, fft=False): replicate-pad `half` each side, conv1d.
    Call site: augmentation/transformations/pass_filters.py:100-102."""
    h = lowpass_taps(cutoff, width=width).astype(np.float64)
    half = (len(h) - 1) // 2
    xp = np.pad(np.asarray(x, np.float64), half, mode="edge")
    if len(h) > 256:  # long filters: FFT convolution (same numbers to ~1e-12)
        n = len(xp) + len(h) - 1
        nfft = 1 << (n - 1).bit_length()
        full = np.fft.irfft(np.fft.rfft(xp, nfft) * np.fft.rfft(h, nfft), nfft)[:n]
        y = full[len(h) - 1: len(h) - 1 + len(x)]
    else:
        y = np.convolve(xp, h, mode="valid")  # h is symmetric: correlation == convolution
    return y.astype(F32)


def highpass(x: np.ndarray, cutoff: float) -> np.ndarray:
    """HighPassFilter.apply_transform: x - lowpass(x).  pass_filters.py:144-155."""
    x = np.asarray(x, F32)
    return (x - lowpass(x, cutoff)).astype(F32)


def bandpass(x: np.ndarray, cutoff_low: float, cutoff_high: float) -> np.ndarray:
    """julius.bandpass_filter(x, cutoff_low, cutoff_high, fft=False) - julius 0.2.7 bands.py / lowpass.py, restated
    (julius is not in this image: PARITY UNPINNED for this function): LowPassFilters([low, high]) builds both filters
    on the window of the lowest positive cut-off, the band is high-passed-low = lows[1] - lows[0]; a low cut-off of 0
    makes lows[0] zero, a high cut-off of 0.5 makes lows[1] the identity.
    Call site: augmentation/transformations/band_filters.py:126-135."""
    if cutoff_low > cutoff_high:
        raise ValueError("Lower cutoff must be smaller than higher cutoff.")
    x = np.asarray(x, F32)
    width = min(c for c in (cutoff_low, cutoff_high) if c > 0)
    low = lowpass(x, cutoff_low, width) if cutoff_low > 0 else np.zeros_like(x)
    high = lowpass(x, cutoff_high, width)
    return (high - low).astype(F32)


def bandstop(x: np.ndarray, cutoff_low: float, cutoff_high: float) -> np.ndarray:
    """BandStopFilter.apply_transform (band_filters.py:187-199): x - bandpass(x)."""
    x = np.asarray(x, F32)
    return (x - bandpass(x, cutoff_low, cutoff_high)).astype(F32)


def cutoff_fraction(cutoff_hz, sample_rate: int) -> float:
    """pass_filters.py:94-96 — float32 cutoff tensor / int sample rate, then .item()."""
    return float(F32(cutoff_hz) / F32(sample_rate))


# ------------------------------------------------------------------ impulse response
def next_fast_len(size: int) -> int:
    """impulse_response.py:170-201 — next 5-smooth number."""
    n = size
    while True:
        r = n
        for p in (2, 3, 5):
            while r % p == 0:
                r //= p
        if r == 1:
            return n
        n += 1


def apply_ir(x: np.ndarray, ir: np.ndarray) -> np.ndarray:
    """ApplyImpulseResponse.apply_transform (impulse_response.py:73-116) with convolve()
    (:119-164): full FFT convolution, divide by the peak of the FULL result, keep first T."""
    x = np.asarray(x, F32)
    ir = np.asarray(ir, F32)
    n = len(x) + len(ir) - 1
    nfft = next_fast_len(n)
    full = np.fft.irfft(np.fft.rfft(x.astype(np.float64), nfft) * np.fft.rfft(ir.astype(np.float64), nfft), nfft)[:n]
    full = full.astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        full = full / np.max(np.abs(full))
    return full[: len(x)].astype(F32)


# ------------------------------------------------------------------ noise / gain / clip / norm
def add_noise(x: np.ndarray, noise: np.ndarray, snr_db: float) -> np.ndarray:
    """AddBackgroundNoise.apply_transform (background_noise.py:183-213), calculate_rms (utils.py:23-29)."""
    x = np.asarray(x, F32)
    rms = F32(np.sqrt(np.mean(np.square(x.astype(np.float64)))))
    scale = F32(rms / F32(10.0 ** (F32(snr_db) / F32(20.0))))
    y = (x + scale * np.asarray(noise, F32)).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (y / np.max(np.abs(y))).astype(F32)


def gain(x: np.ndarray, gain_factor: float) -> np.ndarray:
    """Gain.apply_transform (gain.py:62-70); gain_factor = 10**(dB/20) (utils.py:32-33)."""
    return (np.asarray(x, F32) * F32(gain_factor)).astype(F32)


def quantile_f32(x: np.ndarray, q: float) -> np.float32:
    """torch.quantile(x, q) with linear interpolation on float32 data:
    rank = q*(n-1) (float32), lo = floor(rank), sorted[lo] + (sorted[lo+1]-sorted[lo])*(rank-lo)."""
    s = np.sort(np.asarray(x, F32).ravel())
    n = len(s)
    rank = F32(q) * F32(n - 1)
    lo = int(np.floor(rank))
    hi = min(lo + 1, n - 1)
    w = F32(rank - F32(lo))
    return F32(s[lo] + (s[hi] - s[lo]) * w)


def clip(x: np.ndarray, percentile_threshold: float) -> np.ndarray:
    """Clipping.apply_transform (clipping.py:67-101), B = 1: clamp to the p/2 and 1-p/2 quantiles."""
    x = np.asarray(x, F32)
    lo_q = F32(percentile_threshold) / F32(2)
    lo = quantile_f32(x, lo_q)
    hi = quantile_f32(x, F32(1) - lo_q)
    return np.clip(x, lo, hi).astype(F32)


def peak_normalise(x: np.ndarray) -> np.ndarray:
    """PeakNormalization (peak_normalization.py:38-67): divide by max|x| when it is > 0."""
    x = np.asarray(x, F32)
    m = np.max(np.abs(x))
    return (x / m).astype(F32) if m > 0 else x


# ------------------------------------------------------------------ the chain
STAGES = ("hpf1", "ir", "noise", "gain", "clip", "lpf", "hpf3", "norm")


def augment_chain(x: np.ndarray, prm: dict, sample_rate: int = 8000, return_stages: bool = False):
    """AugmentFP.__call__ for one query with dumped parameters (augmentation/__init__.py:46-97).
    prm keys (any may be absent = transform not applied for this query):
      fc1 [Hz], ir [L] float32, noise [T] float32 + snr_db, gain_factor, clip_p, fc2 [Hz], fc3 [Hz];
    the final PeakNormalization always runs (p = 1, :92)."""
    y = np.asarray(x, F32)
    stages = {}
    if prm.get("fc1") is not None:
        y = highpass(y, cutoff_fraction(prm["fc1"], sample_rate))
    stages["hpf1"] = y
    if prm.get("ir") is not None:
        y = apply_ir(y, prm["ir"])
    stages["ir"] = y
    if prm.get("noise") is not None:
        y = add_noise(y, prm["noise"], prm["snr_db"])
    stages["noise"] = y
    if prm.get("gain_factor") is not None:
        y = gain(y, prm["gain_factor"])
    stages["gain"] = y
    if prm.get("clip_p") is not None:
        y = clip(y, prm["clip_p"])
    stages["clip"] = y
    if prm.get("fc2") is not None:
        y = lowpass(y, cutoff_fraction(prm["fc2"], sample_rate))
    stages["lpf"] = y
    if prm.get("fc3") is not None:
        y = highpass(y, cutoff_fraction(prm["fc3"], sample_rate))
    stages["hpf3"] = y
    y = peak_normalise(y)
    stages["norm"] = y
    return (y, stages) if return_stages else y
