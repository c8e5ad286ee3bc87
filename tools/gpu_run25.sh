mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_backend.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/pytest_backend_final.log; tail -4 gpurun_out/pytest_backend_final.log
