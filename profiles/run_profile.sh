#!/bin/bash
# Run on the GPU box via gpurun: tests, bench, ncu launch lists, ncu full captures of the hot kernels.
# The .ncu-rep files are summarised ON THE BOX (profiles/summarize.py) and deleted unless KEEP_REP=1,
# because gpurun_out/ only travels back when it stays under 64 MiB.
# Usage: bash profiles/run_profile.sh <tag> [skip-tests]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "skip-tests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
fi
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
OURS='stft_mag_kernel|peaks_kernel|peaks_fast_kernel|landmark_kernel|landmark_list_kernel|merge_shifts_kernel|offsets_scan_kernel|compact_rows_kernel'
AUG='fftconv_kernel|mix_kernel|clip_sample_kernel|clip_finish_kernel|clip_lpf_kernel|norm_kernel|filter_spectrum_kernel'
MAT='match_fused_kernel|match_counts|match_select_kernel|match_collect_kernel|match_align_kernel'
UNET='conv_gemm_kernel|conv_halo_kernel|conv_in_kernel|maxpool_kernel'
finish() {  # $1 = report stem, $2 = launch list
  python profiles/summarize.py $OUT/$1.ncu-rep $2 $OUT/$1_summary.txt > /dev/null 2>&1
  if [ "${KEEP_REP:-0}" != "1" ]; then rm -f $OUT/$1.ncu-rep; fi
}
# launch list of one short bench run: every kernel of ours (headline + e2e + full chain + match)
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$OURS|$AUG|$MAT" -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain,match > $OUT/ncu_bench_$TAG.log 2>&1
# full capture: the three headline kernels (after the 3 warm-up steps) ...
ncu --set full --clock-control none --import-source on -k "regex:stft_mag_kernel|peaks_fast_kernel|landmark_list_kernel" -s 9 -c 3 \
    -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_full_$TAG.log 2>&1
finish prof_$TAG $OUT/launches_$TAG.csv
# ... the augmentation kernels (second call of the full-chain leg) ...
ncu --set full --clock-control none --import-source on -k "regex:$AUG" -s 10 -c 10 \
    -o $OUT/prof_aug_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_full_aug_$TAG.log 2>&1
finish prof_aug_$TAG $OUT/launches_$TAG.csv
# ... the matching kernels ...
ncu --set full --clock-control none --import-source on -k "regex:$MAT" -s 6 -c 2 \
    -o $OUT/prof_match_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also match > $OUT/ncu_full_match_$TAG.log 2>&1
finish prof_match_$TAG $OUT/launches_$TAG.csv
if [ "${SKIP_UNET:-0}" = "1" ]; then ls -la $OUT > $OUT/ls_$TAG.txt; exit 0; fi
# ... and the UNet denoiser: per-layer launch list (profiles/unet_layers.py turns it into TFLOP/s per layer)
# plus a full capture of the tcgen05 GEMM launches of one forward pass
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$UNET" -c 60 --csv \
    --log-file $OUT/launches_unet_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also unet --unet-queries 32 > $OUT/ncu_unet_$TAG.log 2>&1
python profiles/unet_layers.py $OUT/launches_unet_$TAG.csv 32 > $OUT/unet_layers_$TAG.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_kernel|conv_halo_kernel" -s 22 -c 22 \
    -o $OUT/prof_unet_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also unet --unet-queries 32 > $OUT/ncu_full_unet_$TAG.log 2>&1
finish prof_unet_$TAG $OUT/launches_unet_$TAG.csv
ls -la $OUT > $OUT/ls_$TAG.txt
