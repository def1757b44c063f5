#!/bin/bash
# Run on the GPU box via gpurun: tests, bench, ncu launch list, ncu full capture of the hot kernels.
# Usage: bash profiles/run_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
KREGEX='regex:stft_mag_kernel|peaks_kernel|peaks_fast_kernel|landmark_kernel|merge_shifts_kernel'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 40 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "$KREGEX" -s 3 -c 3 \
    -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT > $OUT/ls_$TAG.txt
