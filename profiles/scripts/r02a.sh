set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_augment_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -15 > $OUT/pytest_r02a.log
python bench.py --steps 5 --warmup 3 --also chain --no-cpu-baseline > $OUT/bench_r02a.json 2> $OUT/bench_r02a.err
AUG='fftconv_kernel|mix_kernel|clip_sample_kernel|clip_finish_kernel|clip_lpf_kernel|norm_kernel|filter_spectrum_kernel|stft_mag_kernel|peaks_fast_kernel|landmark'
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$AUG" -c 200 --csv --log-file $OUT/launches_r02a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_bench_r02a.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:fftconv_kernel|mix_kernel|clip_lpf_kernel|filter_spectrum_kernel" -s 8 -c 8 -o $OUT/prof_aug_r02a -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also chain > $OUT/ncu_full_aug_r02a.log 2>&1
python profiles/summarize.py $OUT/prof_aug_r02a.ncu-rep $OUT/launches_r02a.csv $OUT/prof_aug_r02a_summary.txt > /dev/null 2>&1
ls -la $OUT/prof_aug_r02a.ncu-rep
cat $OUT/pytest_r02a.log; cat $OUT/bench_r02a.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['stage_ms'], d.get('full_chain'))"
