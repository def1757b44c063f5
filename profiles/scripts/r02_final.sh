set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > $OUT/bench_r02_final.json 2> $OUT/bench_r02_final.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02_final.json").read())
print(round(d["value"]), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()})
print("e2e", round(d["e2e"]["value"]), "pcm", round(d["e2e_pcm16"]["value"]), "parity", d["parity"]["hash_agreement"], "cpu", round(d["cpu_baseline"]["value"]), "match", round(d["match"]["ms_per_step"],3), "fp", round(d["fingerprint_only"]["value"]), "unet", round(d["unet"]["value"]))
PY
