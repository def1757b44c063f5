set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_match_peer_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --also match --no-cpu-baseline > $OUT/bench_r02x_n2.json 2> $OUT/bench_r02x_n2.err; echo rc=$?
tail -3 $OUT/bench_r02x_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02x_n2.json").read())
for k in ("value","ms_per_step","e2e","e2e_pcm16","h2d_ceiling","match"): print(k, d.get(k))
PY
MFPA_PHASE_SUBS=10000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 profiles/scripts/match_phases.py 2>/dev/null | tee $OUT/match_phases_x_n2.txt
