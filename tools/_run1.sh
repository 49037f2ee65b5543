mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "segmented" > gpurun_out/seg_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/seg_pytest.log
timeout 200 python tools/perf_probe.py --seg 1 > gpurun_out/seg_probe1.log 2>&1; tail -6 gpurun_out/seg_probe1.log
