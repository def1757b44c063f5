set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_peer_gpu.py tests/test_match_gpu.py -m gpu -q -x 2>&1 | tail -5
python profiles/scripts/owner_profile.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/owner_launches.csv python profiles/scripts/owner_profile.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open("$OUT/owner_launches.csv") if l.startswith('"'))]
hdr=rows[0]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[k][:60]].append(float(r[v].replace(",","")))
for n,t in d.items(): print(n, len(t), "last ms", t[-1]/1e6)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:match_owner_kernel|match_emit_kernel" -s 4 -c 2 -o $OUT/prof_owner -f python profiles/scripts/owner_profile.py > $OUT/ncu_owner.log 2>&1
tail -2 $OUT/ncu_owner.log
python profiles/sass_by_line.py $OUT/prof_owner.ncu-rep "match_owner_kernel" match 40 > $OUT/owner_lines.txt 2>&1
python profiles/summarize.py $OUT/prof_owner.ncu-rep $OUT/owner_launches.csv $OUT/prof_owner_summary.txt > /dev/null 2>&1
rm -f $OUT/prof_owner.ncu-rep
head -60 $OUT/owner_lines.txt
grep -v "^#" $OUT/prof_owner_summary.txt | head -60
