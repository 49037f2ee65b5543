mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/all_pytest.log
