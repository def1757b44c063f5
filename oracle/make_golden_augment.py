"""TEST INFRASTRUCTURE ONLY — golden vectors for the AugmentFP chain.

Instantiates the reference's transform classes (no file IO: the IR / noise
tensors and every sampled parameter are written into `transform_parameters`
directly, the App. C dump schema) and replays `apply_transform` stage by
stage exactly like BaseWaveformTransform.forward does for a selected
sub-batch (augmentation/transform.py:107-123).
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def reference_chain(ns, x, prm, sr=8000):
    """Run the reference transforms on one query x [T] with dumped params; returns stage outputs."""
    import torch

    HP, LP = ns.pass_filters.HighPassFilter, ns.pass_filters.LowPassFilter
    s = torch.from_numpy(np.asarray(x, np.float32)).reshape(1, 1, -1).clone()
    out = {}

    def run(t, params):
        nonlocal s
        t.transform_parameters = params
        s = t.apply_transform(s.clone(), None if isinstance(t, (HP, LP)) else sr).samples

    if prm.get("fc1") is not None:
        run(HP(min_cutoff_freq=0.0, max_cutoff_freq=150.0, p=1, sample_rate=sr),
            {"cutoff_freq": torch.tensor([prm["fc1"]], dtype=torch.float32)})
    out["hpf1"] = s.numpy().ravel().copy()
    if prm.get("ir") is not None:
        t = ns.impulse_response.ApplyImpulseResponse.__new__(ns.impulse_response.ApplyImpulseResponse)
        torch.nn.Module.__init__(t)
        t.convolve_mode, t.compensate_for_propagation_delay = "full", False
        run(t, {"ir": torch.from_numpy(np.asarray(prm["ir"], np.float32)).reshape(1, 1, -1)})
    out["ir"] = s.numpy().ravel().copy()
    if prm.get("noise") is not None:
        t = ns.background_noise.AddBackgroundNoise.__new__(ns.background_noise.AddBackgroundNoise)
        torch.nn.Module.__init__(t)
        run(t, {"background": torch.from_numpy(np.asarray(prm["noise"], np.float32)).reshape(1, -1),
                "snr_in_db": torch.tensor([prm["snr_db"]], dtype=torch.float32)})
    out["noise"] = s.numpy().ravel().copy()
    if prm.get("gain_factor") is not None:
        run(ns.gain.Gain(min_gain_in_db=-5, max_gain_in_db=5, p=1),
            {"gain_factors": torch.tensor([prm["gain_factor"]], dtype=torch.float32).reshape(1, 1, 1)})
    out["gain"] = s.numpy().ravel().copy()
    if prm.get("clip_p") is not None:
        run(ns.clipping.Clipping(min_percentile_threshold=0.0, max_percentile_threshold=0.01, p=1),
            {"percentile_threshold": torch.tensor([[prm["clip_p"]]], dtype=torch.float32)})
    out["clip"] = s.numpy().ravel().copy()
    if prm.get("fc2") is not None:
        run(LP(min_cutoff_freq=3000.0, max_cutoff_freq=3999.0, p=1, sample_rate=sr),
            {"cutoff_freq": torch.tensor([prm["fc2"]], dtype=torch.float32)})
    out["lpf"] = s.numpy().ravel().copy()
    if prm.get("fc3") is not None:
        run(HP(min_cutoff_freq=30.0, max_cutoff_freq=150.0, p=1, sample_rate=sr),
            {"cutoff_freq": torch.tensor([prm["fc3"]], dtype=torch.float32)})
    out["hpf3"] = s.numpy().ravel().copy()
    t = ns.peak_normalization.PeakNormalization(p=1)
    t.transform_parameters = {}
    t.randomize_parameters(s)
    s = t.apply_transform(s.clone(), sr).samples
    out["norm"] = s.numpy().ravel().copy()
    return out


def cases():
    """Seeded (x, params) cases: full chain at two sizes plus partial chains."""
    from musicfpaugment_b200 import synth

    x = synth.music_like(3, seed=555).numpy()
    ir = synth.impulse_responses(3, length=8000, seed=2000).numpy()
    nz = synth.rms_noise(3, seed=3000).numpy()
    pr = synth.augment_params(3, seed=4000)
    out = []
    # 0: full chain, 2 s query, 0.25 s IR
    out.append((x[0][:16000], dict(fc1=float(pr["fc1"][0]), ir=ir[0][:2000], noise=nz[0][:16000], snr_db=float(pr["snr_db"][0]),
                                   gain_factor=float(10 ** (pr["gain_db"][0] / 20)), clip_p=float(pr["clip_p"][0]),
                                   fc2=float(pr["fc2"][0]), fc3=float(pr["fc3"][0]))))
    # 1: full chain, 8 s query, 1 s IR (BASELINE.json config 3 shape)
    out.append((x[1], dict(fc1=float(pr["fc1"][1]), ir=ir[1], noise=nz[1], snr_db=float(pr["snr_db"][1]),
                           gain_factor=float(10 ** (pr["gain_db"][1] / 20)), clip_p=float(pr["clip_p"][1]),
                           fc2=float(pr["fc2"][1]), fc3=float(pr["fc3"][1]))))
    # 2: noise only at -10 dB ("bn_m10", testing/parameters.py:37-56)
    out.append((x[2][:12000], dict(noise=nz[2][:12000], snr_db=-10.0)))
    # 3: reverb only, IR longer than the 0.5 s signal
    out.append((x[2][:4000], dict(ir=ir[2][:6000])))
    # 4: recording device only: very low loudspeaker cut-off (FIR longer than the signal), clip, LPF, HPF
    out.append((x[0][:8000], dict(fc1=3.0, clip_p=0.01, fc2=3000.0, fc3=30.0)))
    return out


def make_augment(ns):
    out = {"n_cases": 0}
    from oracle.make_golden import _meta

    out["meta"] = _meta()
    for i, (x, prm) in enumerate(cases()):
        ref = reference_chain(ns, x, prm)
        out[f"x{i}"] = np.asarray(x, np.float32)
        for k, v in prm.items():
            out[f"p{i}_{k}"] = np.asarray(v, np.float32)
        out[f"out{i}"] = ref["norm"]
        if i == 0:
            for k, v in ref.items():
                out[f"stage0_{k}"] = v
        out["n_cases"] = i + 1
    np.savez_compressed(os.path.join(GOLD, "augment.npz"), **out)
    print("augment.npz:", out["n_cases"], "cases")
