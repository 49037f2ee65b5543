#!/bin/bash
# Final single-GPU pass after the 12-bit ingestion layout: full GPU test suite, smoke, the cohort / 3-state / small-panel bench
# lines, e2e timeline, ncu launch list of the default bench command.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-final}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 python bench.py --states 3 --no-aux --no-large > gpurun_out/${tag}_bench_s3.json 2>> gpurun_out/bench.err; echo "bench s3 rc=$? $(( $(date +%s) - t0 ))s"
timeout 120 python bench.py --workload small_panel > gpurun_out/${tag}_bench_panel.json 2>> gpurun_out/bench.err; echo "panel rc=$? $(( $(date +%s) - t0 ))s"
timeout 60 python tools/e2e_probe.py --calls-only --reps 5 --p12 --timeline > gpurun_out/${tag}_e2e_timeline.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-aux --no-large --no-cpu --no-ll --no-parity > /dev/null 2>> gpurun_out/bench.err; echo "ncu list rc=$? $(( $(date +%s) - t0 ))s"
python - <<PY
import json
for name in ("bench", "bench_s3", "bench_panel"):
    try:
        j = json.loads(open("gpurun_out/${tag}_%s.json" % name).read().strip().splitlines()[-1])
        e = j.get("e2e") or {}
        r = j.get("roofline") or {}
        print(name, "value %.4g" % j["value"], "ms %.4f" % j["ms_per_step"], "e2e %.4g" % (e.get("value") or 0), "e2e ms", e.get("ms_per_step"),
              "u16", (e.get("uint16_counts") or {}).get("ms_per_step"), "frac", r.get("frac"), (j.get("roofline_other") or {}).get("frac"), "parity", j.get("parity"))
    except Exception as ex:
        print(name, "parse failed", ex)
PY
tail -5 gpurun_out/bench.err
