mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_rzip.py -x -q --tb=short -k "two_phase" 2>&1 | tail -12 > gpurun_out/pytest_twophase.log; tail -5 gpurun_out/pytest_twophase.log
