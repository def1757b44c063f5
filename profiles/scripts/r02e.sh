set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -4
for mb in 256 512 768 1280; do
MFPA_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --also none --no-cpu-baseline > $OUT/bench_r02e_$mb.json 2> $OUT/bench_r02e.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02e_$mb.json").read())
print($mb, "e2e", round(d["e2e"]["ms_per_step"],2), "pcm16", round(d["e2e_pcm16"]["ms_per_step"],2), "ceiling", round(d["h2d_ceiling"]["ms_per_step"],2))
PY
done
python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline --queries 1048 > $OUT/bench_r02e_q1048.json 2>> $OUT/bench_r02e.err
python bench.py --steps 5 --warmup 3 --also none --no-cpu-baseline --queries 2500 > $OUT/bench_r02e_q2500.json 2>> $OUT/bench_r02e.err
python - <<PY
import json
for q in (1048, 2500):
    d=json.loads(open("$OUT/bench_r02e_q%d.json"%q).read())
    print(q, d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms"].items()})
PY
