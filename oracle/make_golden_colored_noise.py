"""TEST INFRASTRUCTURE ONLY — tests/golden/colored_noise.npz from the REAL reference.

Run in the build container:  python -m oracle.make_golden_colored_noise

`AddColoredNoise` (augmentation/transformations/colored_noise.py:41-171) called through BaseWaveformTransform.forward
(augmentation/transform.py:60-141) on a seeded batch of 6 mono signals, p = 0.7: the gate, the drawn SNRs and decays
and the output samples.  The module needs nothing that is absent here (torch only).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SEED, SR, B, T = 4321, 8000, 6, 20000


def signals():
    r = np.random.default_rng(99)
    t = np.arange(T) / SR
    x = np.stack([np.sin(2 * np.pi * (200.0 + 150.0 * i) * t) * (0.2 + 0.1 * i) + 0.05 * r.standard_normal(T) for i in range(B)])
    return x.astype(np.float32)[:, None, :]


def main():
    ref_loader.load()
    from augmentation.transformations.colored_noise import AddColoredNoise

    x = signals()
    t = AddColoredNoise(min_snr_in_db=3.0, max_snr_in_db=30.0, min_f_decay=-2.0, max_f_decay=2.0, p=0.7, sample_rate=SR)
    torch.manual_seed(SEED)
    out = t(samples=torch.from_numpy(x.copy()), sample_rate=SR)
    tp = t.transform_parameters
    np.savez_compressed(os.path.join(GOLD, "colored_noise.npz"), x=x, out=out.samples.numpy(), seed=SEED, sample_rate=SR,
                        should_apply=tp["should_apply"].numpy(), snr_in_db=tp["snr_in_db"].numpy(), f_decay=tp["f_decay"].numpy())
    print("colored_noise.npz:", out.samples.shape, "applied", tp["should_apply"].tolist(), "snr", tp["snr_in_db"].tolist(),
          "decay", tp["f_decay"].tolist())


if __name__ == "__main__":
    main()
