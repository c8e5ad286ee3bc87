set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/perf_probe.py 256 > gpurun_out/perf_probe.log 2>&1; cat gpurun_out/perf_probe.log
timeout 300 python tools/lzma_probe.py 512 2048 > gpurun_out/lzma_probe.log 2>&1; cat gpurun_out/lzma_probe.log
timeout 600 python bench.py --workload c2n --steps 1 --warmup 3 > gpurun_out/bench_c2n.log 2>&1; cat gpurun_out/bench_c2n.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2n_256.csv python bench.py --workload c2n --size-mb 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_tagscan -c 2 -o gpurun_out/k1_full python tools/prof_small.py 64 > gpurun_out/ncu_k1.log 2>&1; tail -3 gpurun_out/ncu_k1.log
