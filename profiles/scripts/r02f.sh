set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q -x 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02f.json 2> $OUT/bench_r02f.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02f.json").read())
for k in ("value","e2e","e2e_pcm16","h2d_ceiling"): print(k, d.get(k))
PY
