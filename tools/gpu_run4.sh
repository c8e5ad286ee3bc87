set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backend.py -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_backend2.log; tail -30 gpurun_out/pytest_backend2.log
LRZGPU_DEBUG=1 timeout 600 python tools/lzma_probe.py 2048 10240 > gpurun_out/lzma_probe3.log 2>&1; cat gpurun_out/lzma_probe3.log
