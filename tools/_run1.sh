mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/all_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
j = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"], "roof", j["roofline"]["kernel"], j["roofline"]["frac"], j["roofline_other"]["kernel"], j["roofline_other"]["frac"])
print(j["kernel_ms_per_step"]); print(j["segmented_sweep"]); print(j["parity"]); print(j["clocks"])
print({k: (v.get("ms_per_step"), v.get("value")) for k, v in j["e2e"].items() if isinstance(v, dict) and "ms_per_step" in v})
print(json.dumps(j["aux"].get("large_cohort_one_gpu"))[:600]); print(json.dumps(j["aux"].get("small_panel"))[:900])
PY
