mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_backend.py -x -q --tb=short -k "3-4194304 or 1-262144 or kw3" 2>&1 | tail -15 > gpurun_out/pytest_fast.log; tail -6 gpurun_out/pytest_fast.log
