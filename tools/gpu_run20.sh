set -x
mkdir -p gpurun_out
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_backend.py -x -q --tb=short 2>&1 | tail -25 > gpurun_out/pytest_backend_rep$i.log; tail -3 gpurun_out/pytest_backend_rep$i.log; done
