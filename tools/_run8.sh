mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-aux --no-ll > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload refset --steps 5 --warmup 3 > gpurun_out/r2h_bench_refset_n$N.json 2>> gpurun_out/bench_n$N.err; echo "refset N=$N rc=$?"
python - <<PY
import json
for f in ["gpurun_out/r2h_bench_n$N.json", "gpurun_out/r2h_bench_refset_n$N.json"]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["n_gpus"], j["value"], j["ms_per_step"], (j.get("e2e") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "parse failed", e)
PY
