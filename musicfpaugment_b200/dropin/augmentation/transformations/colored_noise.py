"""Drop-in for augmentation/transformations/colored_noise.py: `AddColoredNoise` with the reference's constructor,
`randomize_parameters`, `transform_parameters` and call surface (colored_noise.py:41-171, transform.py:24-165).

Host side: the Bernoulli gate, the SNR / spectral-decay draws and the one-second noise period of each selected
example, with the reference's torch RNG calls in the reference's order (so a seeded run draws the same numbers).
Device side: the mix - RMS of the signal, noise scaled to the SNR, peak normalisation - is the noise stage of
`mfpa_augment` (the stage AugmentFP's AddBackgroundNoise runs; same arithmetic, background_noise.py:183-213).

`augmentation/transformations/` has no __init__.py on either side, so the other modules of the reference's directory
stay importable next to this one.
"""
from __future__ import annotations

import random
import warnings
from math import ceil
from typing import Any, Dict, Optional

import numpy as np
import torch

from musicfpaugment_b200 import lib, runtime


class ObjectDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _gen_noise(f_decay, num_samples: int, device: str = "cpu", sample_rate: Optional[int] = None) -> torch.Tensor:
    """One second of unit-RMS noise with a 1/f**f_decay amplitude profile, repeated to num_samples
    (colored_noise.py:12-38).  Drawn on the host generator, like a reference run on CPU tensors."""
    if sample_rate is None:
        sample_rate = 44100
    white = torch.normal(0.0, 1.0, (sample_rate,))
    spec = torch.fft.rfft(white)
    spec *= 1 / (torch.linspace(1, (sample_rate / 2) ** 0.5, spec.shape[0]) ** f_decay)
    period = torch.fft.irfft(spec).unsqueeze(0)
    period = (period / (period.square().mean(dim=-1, keepdim=True).sqrt() + 1e-8)).squeeze()   # Audio.rms_normalize
    return torch.cat([period] * int(ceil(num_samples / sample_rate)))[:num_samples]


class AddColoredNoise:
    supported_modes = {"per_batch", "per_example", "per_channel"}
    supports_multichannel = True
    requires_sample_rate = True

    def __init__(self, min_snr_in_db: float = 3.0, max_snr_in_db: float = 30.0, min_f_decay: float = -2.0,
                 max_f_decay: float = 2.0, p: float = 0.5, sample_rate: Optional[int] = None):
        assert 0.0 <= p <= 1.0
        self.p, self.sample_rate = p, sample_rate
        self.transform_parameters: Dict[Any, Any] = {}
        self.are_parameters_frozen = False
        self.min_snr_in_db, self.max_snr_in_db = min_snr_in_db, max_snr_in_db
        if self.min_snr_in_db > self.max_snr_in_db:
            raise ValueError("min_snr_in_db must not be greater than max_snr_in_db")
        self.min_f_decay, self.max_f_decay = min_f_decay, max_f_decay
        if self.min_f_decay > self.max_f_decay:
            raise ValueError("min_f_decay must not be greater than max_f_decay")

    def freeze_parameters(self, seed: int = 0) -> None:
        self.are_parameters_frozen = True
        random.seed(seed)
        torch.manual_seed(seed)

    def unfreeze_parameters(self) -> None:
        self.are_parameters_frozen = False

    def randomize_parameters(self, samples: torch.Tensor) -> None:
        n = samples.shape[0]
        for name, lo, hi in (("snr_in_db", self.min_snr_in_db, self.max_snr_in_db), ("f_decay", self.min_f_decay, self.max_f_decay)):
            dist = torch.distributions.Uniform(torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32),
                                               validate_args=True)
            self.transform_parameters[name] = dist.sample(sample_shape=(n,))

    def apply_transform(self, samples: torch.Tensor, sample_rate: int) -> ObjectDict:
        n, channels, num_samples = samples.shape
        if channels != 1:
            # the reference's own peak normalisation (`peak_values.expand(-1, T)`, colored_noise.py:139-142) only
            # holds for one channel
            raise RuntimeError("AddColoredNoise: mono input [batch, 1, samples] only")
        noise = torch.stack([_gen_noise(self.transform_parameters["f_decay"][i], num_samples, "cpu", self.sample_rate)
                             for i in range(n)])
        prm = np.zeros(n, dtype=lib.AUG_DTYPE)
        prm["apply"] = lib.AUG_NOISE
        prm["snr_db"] = self.transform_parameters["snr_in_db"].numpy()
        prm["gain_factor"] = 1.0
        ctx = runtime.get_context()
        x = samples[:, 0, :].to(device="cuda", dtype=torch.float32).contiguous()
        out = ctx.augment(x, prm, noise=noise.to(device="cuda", dtype=torch.float32).contiguous(), sample_rate=sample_rate)
        return ObjectDict(samples=out.unsqueeze(1).to(samples.device), sample_rate=sample_rate)

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
