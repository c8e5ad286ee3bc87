set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 300 python tools/prof_small.py 64 > gpurun_out/k2_debug.log 2>&1; cat gpurun_out/k2_debug.log
timeout 900 python -m pytest tests/test_gpu_backend.py -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_backend.log; tail -30 gpurun_out/pytest_backend.log
LRZGPU_DEBUG=1 timeout 600 python tools/lzma_probe.py 2048 10240 > gpurun_out/lzma_probe2.log 2>&1; cat gpurun_out/lzma_probe2.log
