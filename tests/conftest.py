import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """libmfpa.so is a build artefact (git-ignored): make sure the one in the tree matches the sources
    before any test imports the binding.  build() is a no-op when its source stamp is current; nvcc
    cross-compiles, so this works on the CPU-only box too."""
    from musicfpaugment_b200 import build

    try:
        build.build()
    except FileNotFoundError as e:  # no nvcc on this machine: the tests that need the library will say so themselves
        print(f"conftest: could not build libmfpa.so: {e}", file=sys.stderr)
    # any other failure (a compile error) propagates: a stale library must never stand in for the sources


@pytest.fixture(scope="session")
def mfpa_ctx():
    """One libmfpa context on cuda:0 with the numpy-computed spreading table."""
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from musicfpaugment_b200 import lib
    from oracle import audfprint_np as O

    ctx = lib.Context(0, spread_table=O.gaussian_table(256, 30.0))
    yield ctx
    ctx.close()
