#!/bin/bash
# Run on the GPU box via gpurun (ONE GPU): tests, smoke, both bench arms, ncu launch list and ncu full captures of the
# chain (the headline) and of the matcher.  The .ncu-rep files are summarised ON THE BOX (profiles/summarize.py,
# profiles/sass_by_line.py) and deleted unless KEEP_REP=1 - gpurun_out/ only travels back when it stays under 64 MiB.
# Every ncu run sets MFPA_NO_PULL=1 (Nsight Compute hangs on kernels that read mapped host memory; the pull kernel
# only exists on the host-pipeline path, which is not profiled) and sits under `timeout`.
# Usage: bash profiles/run_profile.sh <tag> [skip-tests]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
CHAIN='fftconv_kernel|filter_spectrum_kernel|mix_kernel|clip_sample_kernel|clip_finish_kernel|clip_lpf_kernel|norm_kernel|stft_mag_kernel|peaks_fast_kernel|landmark_list_kernel|landmark_kernel|merge_shifts_kernel|offsets_scan_kernel|compact_rows_kernel|noise_assemble'
MAT='match_fused_kernel|match_align_kernel'
export MFPA_NO_PULL=1
finish() {  # $1 = report stem, $2 = launch list
  python profiles/summarize.py $OUT/$1.ncu-rep $2 $OUT/$1_summary.txt > /dev/null 2>&1
  if [ "${KEEP_REP:-0}" != "1" ]; then rm -f $OUT/$1.ncu-rep; fi
}
# launch list of the bench command (device-timed part only): every kernel of the chain, per launch
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$CHAIN|$MAT" -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also match > $OUT/ncu_bench_$TAG.log 2>&1
# full capture of one step of the chain (after the warm-up steps' launches)
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$CHAIN" -s 39 -c 13 \
    -o $OUT/prof_chain_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_full_chain_$TAG.log 2>&1
python profiles/sass_by_line.py $OUT/prof_chain_$TAG.ncu-rep "fftconv_kernel" augment 40 > $OUT/conv_lines_$TAG.txt 2>&1
finish prof_chain_$TAG $OUT/launches_$TAG.csv
# the single-shard matcher
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$MAT" -s 9 -c 3 \
    -o $OUT/prof_match_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also match > $OUT/ncu_full_match_$TAG.log 2>&1
finish prof_match_$TAG $OUT/launches_$TAG.csv
ls -la $OUT > $OUT/ls_$TAG.txt
