set -u
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 5 --warmup 3 --also fingerprint > $OUT/bench_r02c.json 2> $OUT/bench_r02c.err
tail -5 $OUT/bench_r02c.err
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $OUT/pytest_r02c.log
cat $OUT/pytest_r02c.log
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02c.json").read())
for k in ("value","ms_per_step","stage_ms","roofline","e2e","e2e_pcm16","h2d_ceiling","cpu_baseline","parity","fingerprint_only"): print(k, d.get(k))
PY
