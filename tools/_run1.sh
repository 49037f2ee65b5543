mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_refset.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/refset_probe.py --samples 256 2>&1 | grep -i "gram\|standard"
timeout 300 python tools/refset_probe.py --samples 2000 --rows 250 2>&1 | grep -i "gram\|standard"
