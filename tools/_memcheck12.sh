#!/bin/bash
# compute-sanitizer memcheck over the ingestion-layout tests (12-bit rows with odd bin counts, 16-bit) and the sample-chunk pipeline
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2l_memcheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "12_bit or 16_bit" > gpurun_out/r2l_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -2 gpurun_out/r2l_memcheck_pytest.log; tail -3 gpurun_out/r2l_memcheck.log
