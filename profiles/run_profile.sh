#!/bin/bash
# Run on the GPU box via gpurun: tests, bench, ncu launch list, ncu full capture of the hot kernels.
# Usage: bash profiles/run_profile.sh <tag> [skip-tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
fi
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
OURS='stft_mag_kernel|peaks_kernel|peaks_fast_kernel|landmark_kernel|merge_shifts_kernel|offsets_scan_kernel|compact_rows_kernel'
AUG='fftconv_kernel|mix_kernel|clip_select_kernel|clip_lpf_kernel|norm_kernel'
MAT='match_counts|match_select_kernel|match_collect_kernel|match_align_kernel'
# launch list of one short bench run: every kernel of ours (headline + e2e + full chain + match)
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$OURS|$AUG|$MAT" -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
# full capture: the three headline kernels (after the 3 warm-up steps) ...
ncu --set full --clock-control none --import-source on -k "regex:stft_mag_kernel|peaks_fast_kernel|landmark_kernel" -s 9 -c 3 \
    -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_full_$TAG.log 2>&1
# ... the augmentation kernels (second call of the full-chain leg) ...
ncu --set full --clock-control none --import-source on -k "regex:$AUG" -s 8 -c 8 \
    -o $OUT/prof_aug_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_full_aug_$TAG.log 2>&1
# ... and the matching kernels
ncu --set full --clock-control none --import-source on -k "regex:$MAT" -s 4 -c 4 \
    -o $OUT/prof_match_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also match > $OUT/ncu_full_match_$TAG.log 2>&1
ls -la $OUT > $OUT/ls_$TAG.txt
