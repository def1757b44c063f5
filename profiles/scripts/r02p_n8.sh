set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --also match --no-cpu-baseline > $OUT/bench_r02w_n$N.json 2> $OUT/bench_r02w_n$N.err; echo rc=$?
tail -3 $OUT/bench_r02w_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_r02w_n$N.json").read())
for k in ("value","ms_per_step","e2e","e2e_pcm16","h2d_ceiling","match"): print(k, d.get(k))
PY
MFPA_PHASE_SUBS=10000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 profiles/scripts/match_phases.py > $OUT/match_phases_w_n$N.txt 2> $OUT/match_phases_n$N.err; echo rc=$?
cat $OUT/match_phases_w_n$N.txt
