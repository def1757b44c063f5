set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_peer_gpu.py tests/test_match_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -5
python profiles/scripts/owner_profile.py
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:match_owner_kernel" -s 4 -c 1 -o $OUT/prof_owner -f python profiles/scripts/owner_profile.py > $OUT/ncu_owner.log 2>&1
tail -2 $OUT/ncu_owner.log
python profiles/sass_by_line.py $OUT/prof_owner.ncu-rep "match_owner_kernel" match 30 > $OUT/owner_lines.txt 2>&1
python profiles/summarize.py $OUT/prof_owner.ncu-rep profiles/r02b_launches.csv $OUT/prof_owner_summary.txt > /dev/null 2>&1
rm -f $OUT/prof_owner.ncu-rep
head -34 $OUT/owner_lines.txt | cut -c1-100
grep -v "^#" $OUT/prof_owner_summary.txt | head -28
