set -u
OUT=gpurun_out; mkdir -p $OUT
# memcheck: matcher (incl. the peer-memory kernels against same-GPU buffers), picker, augmentation chain, drop-ins
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_augment_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -k "not unet" > $OUT/sanitizer_r02_aug.txt 2>&1; echo rc=$?
grep -n "ERROR SUMMARY\|passed\|failed" $OUT/sanitizer_r02_aug.txt | head -5
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_match_gpu.py tests/test_match_peer_gpu.py -m gpu -q -x -k "sparse or emit_peer or golden" > $OUT/sanitizer_r02_race.txt 2>&1; echo rc=$?
grep -n "RACECHECK SUMMARY\|passed\|failed\|hazard" $OUT/sanitizer_r02_race.txt | head -8
