mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/all_pytest.log
for f in 256 512 768 1024 1280 1536 1792; do timeout 300 python tools/perf_probe.py --first $f --gen 256 --reps 2 2>&1 | grep "^segments\|emission+viterbi" | cut -c1-330; done
