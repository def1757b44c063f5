set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_peer_gpu.py tests/test_match_gpu.py tests/test_dropin_gpu.py tests/test_drivers_gpu.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --also match --no-cpu-baseline > $OUT/bench_r02r.json 2> $OUT/bench_r02r.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02r.json").read())
print(round(d["ms_per_step"],3), d.get("match"))
PY
