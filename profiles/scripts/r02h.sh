set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_augment_gpu.py tests/test_dropin_gpu.py tests/test_drivers_gpu.py -m gpu -q 2>&1 | tail -4
for occ in 3 2; do
MFPA_CONV_OCC=$occ python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02h_occ$occ.json 2> $OUT/bench_r02h.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02h_occ$occ.json").read())
print($occ, round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), round(d["e2e_pcm16"]["ms_per_step"],2))
PY
done
