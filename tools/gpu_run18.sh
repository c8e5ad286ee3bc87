set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu3.log; tail -2 gpurun_out/pytest_gpu3.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload c2n --size-mb 100 --steps 1 --warmup 1 > gpurun_out/bench_n2_c2n.json 2> gpurun_out/bench_n2_c2n.err; tail -c 1800 gpurun_out/bench_n2_c2n.json; tail -5 gpurun_out/bench_n2_c2n.err
