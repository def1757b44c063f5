"""TEST INFRASTRUCTURE ONLY — tests/golden/clip_pooled.npz: the reference's Gain + Clipping on a BATCH of 3.

`Clipping.apply_transform` (augmentation/transformations/clipping.py:67-101) calls torch.quantile with a vector of
q and no dim, so every row is clipped to quantiles over the POOLED rows.  The golden is the
reference classes' own output for B = 3 with one row left out by the Bernoulli gate, replayed like
BaseWaveformTransform.forward does for the selected sub-batch (transform.py:107-123).

Run in the build container:  python -m oracle.make_golden_clip_pooled
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    import torch

    from musicfpaugment_b200 import synth
    from oracle import ref_loader
    from oracle.make_golden import _meta

    ns = ref_loader.load()
    x = synth.music_like(4, n_samples=12000, seed=909).numpy() * np.array([[1.0], [0.3], [0.7], [0.5]], np.float32)
    gains = np.array([1.4, 0.8, 1.0, 0.6], np.float32)
    clip_p = np.array([0.004, 0.009, 0.0, 0.001], np.float32)
    selected = np.array([True, True, False, True])          # row 2: Clipping's gate did not fire
    s = torch.from_numpy(x).reshape(4, 1, -1).clone()
    g = ns.gain.Gain(min_gain_in_db=-5, max_gain_in_db=5, p=1)
    g.transform_parameters = {"gain_factors": torch.from_numpy(gains).reshape(4, 1, 1)}
    s = g.apply_transform(s.clone(), 8000).samples
    c = ns.clipping.Clipping(min_percentile_threshold=0.0, max_percentile_threshold=0.01, p=1)
    sel = torch.from_numpy(selected)
    c.transform_parameters = {"percentile_threshold": torch.from_numpy(clip_p[selected]).reshape(-1, 1)}
    out = s.clone()
    out[sel] = c.apply_transform(s[sel].clone(), 8000).samples
    np.savez_compressed(os.path.join(GOLD, "clip_pooled.npz"), meta=_meta(), x=x, gain_factor=gains, clip_p=clip_p,
                        selected=selected, out=out.numpy()[:, 0, :])
    print("wrote clip_pooled.npz")


if __name__ == "__main__":
    main()
