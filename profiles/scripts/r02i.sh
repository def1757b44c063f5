set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_gpu.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
export MFPA_NO_PULL=1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fftconv_kernel" -s 3 -c 2 -o $OUT/prof_conv_r02i -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_conv_r02i.log 2>&1
tail -3 $OUT/ncu_conv_r02i.log
python profiles/sass_by_line.py $OUT/prof_conv_r02i.ncu-rep "fftconv_kernel<0" augment 45 > $OUT/conv_lines_r02i.txt 2>&1
python profiles/summarize.py $OUT/prof_conv_r02i.ncu-rep profiles/r02b_launches.csv $OUT/prof_conv_r02i_summary.txt > /dev/null 2>&1
rm -f $OUT/prof_conv_r02i.ncu-rep
head -60 $OUT/conv_lines_r02i.txt
