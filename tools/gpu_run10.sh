set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_tagscan -c 1 -o gpurun_out/k1_v4_full python tools/k1_probe.py 256 > gpurun_out/ncu_k1_v4.log 2>&1; tail -2 gpurun_out/ncu_k1_v4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c2_64.csv python bench.py --size-mb 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1; tail -2 gpurun_out/ncu_launch2.log
