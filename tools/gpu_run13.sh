set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 300 python tools/prof_small.py 256 > gpurun_out/k2_debug5.log 2>&1; tail -3 gpurun_out/k2_debug5.log
timeout 400 python -m pytest tests/test_gpu_rzip.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_rzip4.log; tail -3 gpurun_out/pytest_rzip4.log
