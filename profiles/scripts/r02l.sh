set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_gpu.py tests/test_augment_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02l.json 2> $OUT/bench_r02l.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02l.json").read())
print(round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), round(d["e2e_pcm16"]["ms_per_step"],2))
PY
export MFPA_NO_PULL=1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fftconv_kernel|clip_lpf_kernel" -s 3 -c 3 -o $OUT/prof_conv_r02l -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_conv_r02l.log 2>&1
tail -2 $OUT/ncu_conv_r02l.log
python profiles/sass_by_line.py $OUT/prof_conv_r02l.ncu-rep "fftconv_kernel" augment 60 > $OUT/conv_lines_r02l.txt 2>&1
python profiles/summarize.py $OUT/prof_conv_r02l.ncu-rep profiles/r02b_launches.csv $OUT/prof_conv_r02l_summary.txt > /dev/null 2>&1
rm -f $OUT/prof_conv_r02l.ncu-rep
head -75 $OUT/conv_lines_r02l.txt
