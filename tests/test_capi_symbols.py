"""CPU: libmfpa.so builds, loads, and exports every symbol include/mfpa.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from musicfpaugment_b200 import build

    return build.build()


def _declared():
    names = set()
    for f in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mfpa_[a-z0-9_]+)\s*\(", src))
    return names


def test_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    declared = _declared()
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_python_binding_covers_header(libpath):
    from musicfpaugment_b200 import lib

    assert set(lib.ALL_SIGNATURES) == _declared()
    assert lib.raw().mfpa_abi_version() == 1


def test_pure_host_entry_points(libpath):
    from musicfpaugment_b200 import lib
    from oracle import audfprint_np as O

    assert lib.num_frames(64000) == 251 == O.num_frames(64000)
    assert [lib.shift_offset(s, 4) for s in range(4)] == O.shift_offsets(4)
    assert [lib.shift_offset(s, 3) for s in range(3)] == O.shift_offsets(3)
    p = lib.afp_defaults()
    assert p.a_dec == O.a_dec() and p.maxpks == 5 and p.fanout == 3 and p.targetdt == 63


def test_no_gpu_fails_loudly(libpath):
    import torch

    from musicfpaugment_b200 import lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.MfpaError):
        lib.Context(0)
