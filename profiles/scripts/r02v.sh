set -u
python -m pytest tests/test_dropin_gpu.py tests/test_augment_gpu.py -m gpu -q -x 2>&1 | tail -12
