"""TEST INFRASTRUCTURE ONLY — import the real reference with stub modules.

Works only where ``/root/reference`` is mounted (the build container).  Used
by ``oracle/make_golden.py`` to produce ``tests/golden`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the mount is absent).

Recipe: SURVEY.md Appendix D.  The reference dies at import on GPUtil /
tensorflow / matplotlib / julius / librosa (none installed) and on hard-coded
checkpoint paths (afp/audfprint/peak_extractor.py:24-37,
afp/dejavu/fingerprint.py:27-31), so those are stubbed *before* import.
"""
from __future__ import annotations

import math
import os
import sys
import types

REF = os.environ.get("MFPA_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "afp", "audfprint"))


# --------------------------------------------------------------------------
# julius 0.2.7 `lowpass_filter` restated (third-party, un-vendored; see
# oracle/augment_np.py for the numpy twin and the "parity unpinned" note).
# Call site: augmentation/transformations/pass_filters.py:100-102.
# --------------------------------------------------------------------------
def _julius_lowpass_filter(x, cutoff: float, zeros: float = 8, fft=None, stride: int = 1, pad: bool = True):
    import torch
    import torch.nn.functional as F

    if cutoff < 0:
        raise ValueError("Minimum cutoff must be larger than zero.")
    if cutoff > 0.5:
        raise ValueError("A cutoff above 0.5 does not make sense.")
    if cutoff == 0:
        # julius builds an empty filter list for cutoff 0 and `min()` of it raises.
        raise ValueError("min() arg is an empty sequence")
    half_size = int(zeros / cutoff / 2)
    window = torch.hann_window(2 * half_size + 1, periodic=False)
    time = torch.arange(-half_size, half_size + 1)
    arg = 2 * cutoff * math.pi * time
    sinc = torch.where(arg == 0, torch.ones_like(arg, dtype=torch.float32), torch.sin(arg) / arg)
    filt = 2 * cutoff * window * sinc
    filt = filt / filt.sum()
    shape = list(x.shape)
    inp = x.reshape(-1, 1, shape[-1])
    if pad:
        inp = F.pad(inp, (half_size, half_size), mode="replicate")
    out = F.conv1d(inp, filt[None, None, :], stride=stride)
    shape[-1] = out.shape[-1]
    return out.reshape(shape)


def _install_stubs() -> None:
    import torch

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "tensorflow" not in sys.modules:
        tf = mod("tensorflow", __version__="stub")
        tf.config = types.SimpleNamespace(set_visible_devices=lambda *a, **k: None)
        tf.random = types.SimpleNamespace(set_seed=lambda *a, **k: None)
    if "GPUtil" not in sys.modules:
        mod("GPUtil", getAvailable=lambda *a, **k: [0])
    if "librosa" not in sys.modules:
        mod("librosa")
    if "julius" not in sys.modules:
        mod("julius", lowpass_filter=_julius_lowpass_filter)
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mlab = mod("matplotlib.mlab", window_hanning=None)
        plt = mod("matplotlib.pyplot")
        mpl.mlab, mpl.pyplot = mlab, plt
    for name in ("GPUtil",):
        pass

    # Hard-coded checkpoints → seeded random init (same key names).
    _orig_load = torch.load

    def _fake_load(path, *a, **k):
        if isinstance(path, str) and path.startswith("/workspace/"):
            from training.model import Demucs
            from training.unet import UNet

            g = torch.random.get_rng_state()
            torch.manual_seed(0)
            sd = (UNet(1, 1, rate=0.05) if "unet" in path else Demucs()).state_dict()
            torch.random.set_rng_state(g)
            return {"model_state_dict": sd}
        return _orig_load(path, *a, **k)

    torch.load = _fake_load


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path modules imported."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference not mounted at {REF}")
    _install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import io
    import contextlib

    import training.utils as tu

    tu.set_gpus = lambda *a, **k: "cpu"
    with contextlib.redirect_stdout(io.StringIO()):
        import afp.audfprint.stft as r_stft
        import afp.audfprint.peak_extractor as r_pe
        import afp.audfprint.hash_table as r_ht
        import afp.audfprint.audfprint_match as r_match
        import afp.dejavu as r_dj
        import afp.dejavu.variables as r_djv

        sys.modules["dejavu"] = r_dj
        sys.modules["dejavu.variables"] = r_djv
        import afp.dejavu.fingerprint as r_djfp
        import augmentation as r_aug
        import augmentation.transformations.pass_filters as r_pf
        import augmentation.transformations.impulse_response as r_ir
        import augmentation.transformations.background_noise as r_bn
        import augmentation.transformations.gain as r_gain
        import augmentation.transformations.clipping as r_clip
        import augmentation.transformations.peak_normalization as r_pn
        import testing.parameters as r_params
        import training.unet as r_unet
    ns = types.SimpleNamespace(
        stft=r_stft, peak_extractor=r_pe, hash_table=r_ht, match=r_match,
        dejavu_fingerprint=r_djfp, augmentation=r_aug, pass_filters=r_pf,
        impulse_response=r_ir, background_noise=r_bn, gain=r_gain,
        clipping=r_clip, peak_normalization=r_pn, parameters=r_params,
        unet=r_unet, julius_lowpass_filter=_julius_lowpass_filter,
    )
    _loaded["ns"] = ns
    return ns
