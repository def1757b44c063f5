set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_match_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --also match --no-cpu-baseline > $OUT/bench_r02k_n2.json 2> $OUT/bench_r02k_n2.err; echo rc=$?
tail -5 $OUT/bench_r02k_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02k_n2.json").read())
for k in ("value","ms_per_step","e2e","e2e_pcm16","h2d_ceiling","match"): print(k, d.get(k))
PY
