set -u
OUT=gpurun_out; mkdir -p $OUT
echo "--- ncu on a trivial torch program"
timeout 120 ncu --metrics gpu__time_duration.sum -c 2 python -c "import torch; a=torch.ones(1000,device='cuda'); print((a+a).sum().item())" 2>&1 | tail -5
echo "--- ncu, bench, default"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fftconv_kernel" -c 4 python -X faulthandler bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_dbg1.log 2>&1; echo rc=$?
tail -12 $OUT/ncu_dbg1.log
echo "--- ncu, bench, MFPA_NO_PULL"
MFPA_NO_PULL=1 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fftconv_kernel" -c 4 python -X faulthandler bench.py --steps 2 --warmup 3 --no-cpu-baseline --also none > $OUT/ncu_dbg2.log 2>&1; echo rc=$?
tail -12 $OUT/ncu_dbg2.log
