#!/bin/bash
# One GPU-box pass: parity tests, smoke, the default bench line, then a compute-sanitizer memcheck over the small
# parity cases.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-check}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 150 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print(j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], j["roofline_other"]["frac"])
    print(json.dumps(j["aux"].get("small_panel")))
except Exception as e:
    print("bench parse failed", e)
PY
if [ "${2:-}" = "memcheck" ]; then
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${tag}_memcheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "kat1 or kat3 or ragged or special_values or equal_length or small_panel_shape" \
    > gpurun_out/${tag}_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/${tag}_memcheck_pytest.log; grep -c "Invalid\|out of bounds" gpurun_out/${tag}_memcheck.log; tail -3 gpurun_out/${tag}_memcheck.log
fi
