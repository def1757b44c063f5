set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
timeout 300 python -m pytest tests/test_match_peer_gpu.py tests/test_match_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 profiles/scripts/match_phases.py > $OUT/match_phases_n$N.txt 2> $OUT/match_phases_n$N.err; echo rc=$?
tail -5 $OUT/match_phases_n$N.err
cat $OUT/match_phases_n$N.txt
