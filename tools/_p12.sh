#!/bin/bash
# host call: parity tests of the ingestion layouts and the sample-chunk pipeline, then the call timed with 12-bit and 16-bit counts
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "12_bit or 16_bit or sample_chunk or exomecount or full_size or capacity" > gpurun_out/p12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p12_pytest.log
timeout 60 python tools/e2e_probe.py --reps 12 --p12 > gpurun_out/p12_probe.log 2>&1
timeout 60 python tools/e2e_probe.py --calls-only --reps 12 >> gpurun_out/p12_probe.log 2>&1
cat gpurun_out/p12_probe.log
