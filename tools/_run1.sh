mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/all_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/all_pytest.log
for o in "chunks=0" "chunks=2" "chunks=3" "chunks=4" "chunks=6" "segments=0"; do timeout 200 python tools/e2e_probe.py --opts $o --reps 8 2>&1 | grep "^calls+stats " ; done
timeout 200 python tools/e2e_probe.py --reps 3 --timeline > gpurun_out/e2e_timeline.log 2>&1
