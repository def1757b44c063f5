"""TEST INFRASTRUCTURE ONLY — tests/golden/unet.npz from the REAL reference UNet
(/root/reference/training/unet.py via oracle/ref_loader.py).

Run in the build container:  python -m oracle.make_golden_unet
The checkpoint of the paper is not distributed, so the pin uses torch.manual_seed(0)
initial weights (the reference constructor draws them) and a seeded 48x40 input.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import torch

    from oracle import ref_loader

    ns = ref_loader.load()
    UNet = sys.modules["training.unet"].UNet
    torch.manual_seed(0)
    net = UNet(1, 1, rate=0.05).eval()
    keys = [k for k in net.state_dict().keys()]
    rng = np.random.default_rng(77)
    x = rng.random((2, 1, 48, 40), dtype=np.float32)
    torch.set_num_threads(1)
    with torch.no_grad():
        y = net(torch.from_numpy(x)).numpy()
    meta = json.dumps({"torch": torch.__version__, "numpy": np.__version__, "generator": "oracle/make_golden_unet.py",
                       "seed": 0, "n_params": int(sum(p.numel() for p in net.parameters()))})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "unet.npz"), x=x, y=y, keys=np.array(keys), meta=meta)
    print("wrote unet.npz", y.shape, float(np.abs(y).max()), len(keys))


if __name__ == "__main__":
    main()
