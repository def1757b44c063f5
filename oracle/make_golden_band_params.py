"""TEST INFRASTRUCTURE ONLY — tests/golden/band_params.npz from the REAL reference.

Run in the build container:  python -m oracle.make_golden_band_params

The seeded parameter draws of the reference's `BandPassFilter` (augmentation/transformations/band_filters.py:76-115):
gate, centre frequencies, bandwidth fractions, and the two cut-off fractions `apply_transform` derives from them
(:122-133).  The filtering itself calls julius.bandpass_filter (julius is not in this image): that part is restated in
oracle/augment_np.bandpass and stays PARITY UNPINNED.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SEED, SR, B = 2468, 8000, 12


def main():
    ref_loader.load()
    from augmentation.transformations.band_filters import BandPassFilter

    t = BandPassFilter(min_center_frequency=200, max_center_frequency=1900, min_bandwidth_fraction=0.5, max_bandwidth_fraction=1.99,
                       p=0.8, sample_rate=SR)
    torch.manual_seed(SEED)
    gate = t.bernoulli_distribution.sample(sample_shape=(B,)).to(torch.bool)     # transform.py:101-105
    t.randomize_parameters(torch.zeros(int(gate.sum()), 1, 16))
    tp = t.transform_parameters
    low = tp["center_freq"] * (1 - 0.5 * tp["bandwidth"]) / SR
    high = tp["center_freq"] * (1 + 0.5 * tp["bandwidth"]) / SR
    np.savez_compressed(os.path.join(GOLD, "band_params.npz"), seed=SEED, sample_rate=SR, batch=B, should_apply=gate.numpy(),
                        center_freq=tp["center_freq"].numpy(), bandwidth=tp["bandwidth"].numpy(),
                        low=np.array([v.item() for v in low]), high=np.array([v.item() for v in high]))
    print("band_params.npz: applied", int(gate.sum()), "centre", tp["center_freq"].tolist()[:4], "low", [v.item() for v in low][:4])


if __name__ == "__main__":
    main()
