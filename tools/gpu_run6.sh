set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rzip.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_rzip2.log; tail -8 gpurun_out/pytest_rzip2.log
LRZGPU_DEBUG=1 timeout 300 python tools/prof_small.py 64 > gpurun_out/k2_debug3.log 2>&1; cat gpurun_out/k2_debug3.log
