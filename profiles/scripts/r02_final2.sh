python -m pytest tests/test_dejavu_match_gpu.py tests/test_drivers_gpu.py tests/test_dejavu_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -2
