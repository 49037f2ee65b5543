#!/bin/bash
# Final single-GPU pass of a round: smoke, the four bench lines and the reference arm, the ncu launch list of the default
# bench command.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-final}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 python bench.py --states 3 --no-aux --no-large > gpurun_out/${tag}_bench_s3.json 2>> gpurun_out/bench.err; echo "bench s3 rc=$? $(( $(date +%s) - t0 ))s"
timeout 120 python bench.py --workload small_panel > gpurun_out/${tag}_bench_panel.json 2>> gpurun_out/bench.err; echo "panel rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 python bench.py --workload refset > gpurun_out/${tag}_bench_refset.json 2>> gpurun_out/bench.err; echo "refset rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-aux --no-large --no-cpu --no-ll --no-parity > /dev/null 2>> gpurun_out/bench.err; echo "ncu list rc=$? $(( $(date +%s) - t0 ))s"
python - <<PY
import json
for name in ("bench", "bench_s3", "bench_panel", "bench_refset", "bench_reference"):
    try:
        j = json.loads(open("gpurun_out/${tag}_%s.json" % name).read().strip().splitlines()[-1])
        e = j.get("e2e") or {}
        r = j.get("roofline") or {}
        print(name, "value %.4g" % j["value"], "ms %.4f" % j["ms_per_step"], "e2e %.4g" % (e.get("value") or 0), "e2e ms", e.get("ms_per_step"), "frac", r.get("frac"), (j.get("roofline_other") or {}).get("frac"))
    except Exception as ex:
        print(name, "parse failed", ex)
PY
tail -5 gpurun_out/bench.err
if [ "${2:-}" = "ncu" ]; then
# one --set full capture of each kernel of the device-resident step (first launch after the warm-up launches)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"emission_table|viterbi_tpc|viterbi_tilemap|viterbi_trace|viterbi_expand" --launch-skip 10 --launch-count 5 \
    -o gpurun_out/${tag}_full -f python tools/perf_probe.py --reps 1 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu full rc=$? $(( $(date +%s) - t0 ))s"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${tag}_memcheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "table_and_direct or panel_lattice or sample_chunk or kat1 or ragged" \
    > gpurun_out/${tag}_memcheck_pytest.log 2>&1; echo "memcheck rc=$? $(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/${tag}_memcheck_pytest.log; tail -2 gpurun_out/${tag}_memcheck.log
fi
