#!/bin/bash
# 12-bit ingestion layout: parity tests, then the host call timed with 16-bit and 12-bit counts
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "12_bit or 16_bit or sample_chunk" > gpurun_out/p12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p12_pytest.log
timeout 60 python tools/e2e_probe.py --calls-only --reps 12 > gpurun_out/p12_probe.log 2>&1
timeout 100 python tools/e2e_probe.py --calls-only --reps 12 --p12 --opts ";chunks=4,head=50;chunks=5,head=50;chunks=5,head=35;chunks=6,head=50;chunks=5,head=50,reserve=48;chunks=5,head=70;chunks=6,head=35;chunks=3" >> gpurun_out/p12_probe.log 2>&1
timeout 60 python tools/e2e_probe.py --calls-only --reps 12 --states 3 --p12 --opts ";chunks=5,head=50" >> gpurun_out/p12_probe.log 2>&1
timeout 60 python tools/e2e_probe.py --calls-only --reps 3 --p12 --timeline --opts "chunks=5,head=50" > gpurun_out/p12_timeline.log 2>&1
cat gpurun_out/p12_probe.log
