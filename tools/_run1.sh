mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
timeout 300 python bench.py --workload small_panel > gpurun_out/r2h_bench_panel.json 2>> gpurun_out/bench.err; echo "panel rc=$?"
timeout 300 python bench.py --workload refset --steps 5 > gpurun_out/r2h_bench_refset.json 2>> gpurun_out/bench.err; echo "refset rc=$?"
timeout 300 python bench.py --states 3 --no-aux --no-ll > gpurun_out/r2h_bench_s3.json 2>> gpurun_out/bench.err; echo "s3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-aux --no-parity > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:viterbi_tpc_kernel.*\(bool\)1|emission_table_kernel' -s 2 -c 2 -o gpurun_out/r2h_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-aux --no-parity > /dev/null 2>&1; echo "ncu full rc=$?"
