set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_audfprint_gpu.py tests/test_dejavu_gpu.py -m gpu -q -x 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --also fingerprint > $OUT/bench_r02y3.json 2> $OUT/bench_r02y.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02y3.json").read())
print(round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, d["parity"]["hash_agreement"], d["fingerprint_only"]["stage_ms"])
PY
