set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_gpu.py -m gpu -q -x 2>&1 | tail -8
MFPA_DEBUG_PIPE=1 python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02d.json 2> $OUT/bench_r02d.err
grep "pipe chunk" $OUT/bench_r02d.err | tail -24
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02d.json").read())
for k in ("value","e2e","e2e_pcm16","h2d_ceiling"): print(k, d.get(k))
PY
