set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -12
for mb in 256 384; do
MFPA_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02g_$mb.json 2> $OUT/bench_r02g.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02g_$mb.json").read())
print($mb, "e2e", round(d["e2e"]["ms_per_step"],2), "pcm16", round(d["e2e_pcm16"]["ms_per_step"],2), "ceiling", round(d["h2d_ceiling"]["ms_per_step"],2))
PY
done
