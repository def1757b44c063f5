set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_augment_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -15 > $OUT/pytest_r02b.log
AUG='fftconv_kernel|mix_kernel|clip_sample_kernel|clip_finish_kernel|clip_lpf_kernel|norm_kernel|filter_spectrum_kernel|stft_mag_kernel|peaks_fast_kernel|landmark'
for occ in 3 2; do
MFPA_CONV_OCC=$occ python bench.py --steps 5 --warmup 3 --also chain --no-cpu-baseline > $OUT/bench_r02b_occ$occ.json 2> $OUT/bench_r02b_occ$occ.err
MFPA_CONV_OCC=$occ ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$AUG" -c 200 --csv --log-file $OUT/launches_r02b_occ$occ.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_bench_r02b.log 2>&1
done
ncu --set full --clock-control none --import-source on -k "regex:fftconv_kernel|mix_kernel|clip_lpf_kernel|filter_spectrum_kernel" -s 8 -c 8 -o $OUT/prof_aug_r02b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_full_aug_r02b.log 2>&1
python profiles/summarize.py $OUT/prof_aug_r02b.ncu-rep $OUT/launches_r02b_occ3.csv $OUT/prof_aug_r02b_summary.txt > /dev/null 2>&1
cat $OUT/pytest_r02b.log
for occ in 3 2; do python - <<PY
import json
d=json.loads(open("$OUT/bench_r02b_occ$occ.json").read()); print($occ, d['full_chain']['ms_per_step'], d['full_chain']['value'])
PY
done
