"""The FFT-convolution core of augment.cu, run on the CPU (no GPU needed).

csrc/fftconv_core.cuh is __host__ __device__: tests/host/fftconv_check.cu executes the passes of one
256-thread block (pack, three FFT passes, split/multiply, inverse, block geometry) thread by thread and
compares with a float64 direct convolution under the reference's boundary rules
(augmentation/transformations/pass_filters.py:84-155, impulse_response.py:119-164).
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fftconv_core_on_host(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "fftconv_check")
    inc = "/usr/local/cuda/include"
    r = subprocess.run([cxx, "-O2", "-std=c++17", "-x", "c++", "-I", inc, "-o", exe,
                        os.path.join(ROOT, "tests", "host", "fftconv_check.cu"), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ok" in r.stdout
