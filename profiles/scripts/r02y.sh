set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_augment_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -2
for gx in 0 4; do
if [ $gx != 0 ]; then export MFPA_LPF_GX=$gx; fi
python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02y_$gx.json 2> $OUT/bench_r02y.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02y_$gx.json").read())
print("gx", $gx, round(d["ms_per_step"],3), "clip_lpf", round(d["stage_ms"]["clip_lpf"],3))
PY
done
