"""Drop-in for augmentation/transformations/band_filters.py: `BandPassFilter` and `BandStopFilter` with the reference's
constructors, parameter draws and call surface (band_filters.py:15-199, transform.py:24-165).

Host side: the Bernoulli gate and the centre-frequency (uniform in mel) / bandwidth draws, with the reference's torch
RNG calls in its order.  Device side: `julius.bandpass_filter(x, low, high)` = the two filters of
`LowPassFilters([low, high])` - both on the window of the low cut-off - subtracted, each run by
`mfpa_lowpass_filters` (the low-pass stage of the chain: FFT overlap-save for the long windows a 1 Hz cut-off needs).
julius itself is not needed.
"""
from __future__ import annotations

import random
import warnings
from typing import Any, Dict, Optional

import numpy as np
import torch

from musicfpaugment_b200 import runtime

from .colored_noise import ObjectDict


def convert_frequencies_to_mels(f: torch.Tensor) -> torch.Tensor:
    return 2595.0 * torch.log10(1.0 + f / 700.0)


def convert_mels_to_frequencies(m: torch.Tensor) -> torch.Tensor:
    return 700.0 * (10 ** (m / 2595.0) - 1.0)


class BandPassFilter:
    supported_modes = {"per_batch", "per_example"}
    supports_multichannel = True
    requires_sample_rate = True

    def __init__(self, min_center_frequency: int = 200, max_center_frequency: int = 4000, min_bandwidth_fraction: float = 0.5,
                 max_bandwidth_fraction: float = 1.99, p: float = 0.5, sample_rate: Optional[int] = None) -> None:
        assert 0.0 <= p <= 1.0
        self.p, self.sample_rate = p, sample_rate
        self.transform_parameters: Dict[Any, Any] = {}
        self.are_parameters_frozen = False
        self.min_center_frequency, self.max_center_frequency = min_center_frequency, max_center_frequency
        self.min_bandwidth_fraction, self.max_bandwidth_fraction = min_bandwidth_fraction, max_bandwidth_fraction
        if max_center_frequency < min_center_frequency:
            raise ValueError(f"max_center_frequency ({max_center_frequency}) should be larger than "
                             f"min_center_frequency ({min_center_frequency}).")
        if min_bandwidth_fraction <= 0.0:
            raise ValueError("min_bandwidth_fraction must be a positive number")
        if max_bandwidth_fraction < min_bandwidth_fraction:
            raise ValueError(f"max_bandwidth_fraction ({max_bandwidth_fraction}) should be larger than "
                             f"min_bandwidth_fraction ({min_bandwidth_fraction}).")
        if max_bandwidth_fraction >= 2.0:
            raise ValueError(f"max_bandwidth_fraction ({max_bandwidth_fraction}) should be smaller than 2.0,"
                             f"since otherwise low_cut_frequency of the band can be smaller than 0 Hz.")

    def freeze_parameters(self, seed: int = 0) -> None:
        self.are_parameters_frozen = True
        random.seed(seed)
        torch.manual_seed(seed)

    def unfreeze_parameters(self) -> None:
        self.are_parameters_frozen = False

    def randomize_parameters(self, samples: torch.Tensor) -> None:
        """Centre frequency uniform in mel between the two bounds, bandwidth fraction uniform (band_filters.py:76-115)."""
        n = samples.shape[0]
        mel = lambda hz: convert_frequencies_to_mels(torch.tensor(hz, dtype=torch.float32))
        centre = torch.distributions.Uniform(low=mel(self.min_center_frequency), high=mel(self.max_center_frequency),
                                             validate_args=True)
        self.transform_parameters["center_freq"] = convert_mels_to_frequencies(centre.sample(sample_shape=(n,)))
        width = torch.distributions.Uniform(low=torch.tensor(self.min_bandwidth_fraction, dtype=torch.float32),
                                            high=torch.tensor(self.max_bandwidth_fraction, dtype=torch.float32))
        self.transform_parameters["bandwidth"] = width.sample(sample_shape=(n,))

    def _band(self, samples: torch.Tensor, sample_rate: int) -> torch.Tensor:
        """julius.bandpass_filter per example: [n, C, T] -> [n, C, T] on the GPU."""
        tp = self.transform_parameters
        low = tp["center_freq"] * (1 - 0.5 * tp["bandwidth"]) / sample_rate
        high = tp["center_freq"] * (1 + 0.5 * tp["bandwidth"]) / sample_rate
        n, channels, num_samples = samples.shape
        lo = np.repeat(np.array([v.item() for v in low], dtype=np.float64), channels)
        hi = np.repeat(np.array([v.item() for v in high], dtype=np.float64), channels)
        if (lo > hi).any():
            raise ValueError("Lower cutoff must be smaller than higher cutoff.")
        if (lo < 0).any():
            raise ValueError("Minimum cutoff must be larger than zero.")
        if (hi > 0.5).any():
            raise ValueError("A cutoff above 0.5 does not make sense.")
        ctx = runtime.get_context()
        x = samples.reshape(n * channels, num_samples).to(device="cuda", dtype=torch.float32).contiguous()
        band = ctx.lowpass_filters(x, hi, lo) - ctx.lowpass_filters(x, lo, lo)     # lows[1] - lows[0], one window
        return band.reshape(n, channels, num_samples)

    def apply_transform(self, samples: torch.Tensor, sample_rate: int) -> ObjectDict:
        return ObjectDict(samples=self._band(samples, sample_rate).to(samples.device), sample_rate=sample_rate)

    def forward(self, samples: torch.Tensor, sample_rate: int) -> Any:
        if not isinstance(samples, torch.Tensor) or len(samples.shape) != 3:
            raise RuntimeError("torch-audiomentations expects three-dimensional input tensors, with dimension ordering like "
                               "[batch_size, num_channels, num_samples]. If your audio is mono, you can use a shape like "
                               "[batch_size, 1, num_samples].")
        batch_size, num_channels, num_samples = samples.shape
        if batch_size * num_channels * num_samples == 0:
            warnings.warn("An empty samples tensor was passed to {}".format(self.__class__.__name__))
            return ObjectDict(samples=samples, sample_rate=sample_rate)
        gate = torch.distributions.Bernoulli(torch.tensor(float(self.p))).sample(sample_shape=(batch_size,)).to(torch.bool)
        self.transform_parameters = {"should_apply": gate}
        if gate.any():
            out = samples.clone()
            selected = out[gate]
            self.randomize_parameters(samples=selected)
            perturbed = self.apply_transform(samples=selected, sample_rate=sample_rate)
            out[gate] = perturbed.samples
            return ObjectDict(samples=out, sample_rate=perturbed.sample_rate)
        return ObjectDict(samples=samples, sample_rate=sample_rate)

    __call__ = forward


class BandStopFilter(BandPassFilter):
    """x - bandpass(x) (band_filters.py:159-199)."""

    def apply_transform(self, samples: torch.Tensor, sample_rate: int) -> ObjectDict:
        band = self._band(samples, sample_rate).to(samples.device)
        return ObjectDict(samples=samples - band, sample_rate=sample_rate)
