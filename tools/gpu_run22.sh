mkdir -p gpurun_out
timeout 150 python bench.py --workload c1 --size-mb 16 --steps 1 --warmup 1 > gpurun_out/bench_c1_small.json 2> gpurun_out/bench_c1_small.err; echo rc=$?; wc -l gpurun_out/bench_c1_small.json; head -c 700 gpurun_out/bench_c1_small.json; tail -3 gpurun_out/bench_c1_small.err
