"""Process-wide libmfpa context used by the drop-in modules (one per visible GPU)."""
from __future__ import annotations

import numpy as np

_ctx = {}


def spread_table(npoints: int = 256, width: float = 30.0) -> np.ndarray:
    """The reference's cached Gaussian profile (afp/audfprint/peak_extractor.py:163-165),
    computed with numpy so the table the kernels read is bit-identical to it."""
    return np.exp(-0.5 * ((np.arange(-npoints, npoints + 1) / width) ** 2))


def get_context(device: int | None = None):
    import torch

    from . import lib

    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _ctx:
        _ctx[device] = lib.Context(device, spread_table=spread_table())
    return _ctx[device]


def a_dec(density: float, n_hop: int) -> float:
    """peak_extractor.py:295, evaluated in numpy float64 like the reference."""
    return float(1 - 0.01 * (density * np.sqrt(n_hop / 352.8) / 35))
