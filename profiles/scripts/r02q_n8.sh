set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
for v in on off; do
  if [ $v = off ]; then export MFPA_PEER_NO_FENCE=1; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 profiles/scripts/peer_stress.py 2>$OUT/peer_stress_$v.err | tee $OUT/peer_stress_n${N}_$v.txt
  tail -2 $OUT/peer_stress_$v.err
done
unset MFPA_PEER_NO_FENCE
MFPA_PHASE_SUBS=10000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 profiles/scripts/match_phases.py 2>/dev/null | tee $OUT/match_phases_n${N}_fence.txt
