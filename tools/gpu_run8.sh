set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 120 python tools/prof_small.py 8 > gpurun_out/k2_xcheck2.log 2>&1; tail -5 gpurun_out/k2_xcheck2.log
grep -q "^[0-9]" gpurun_out/k2_xcheck2.log && (timeout 400 python -m pytest tests/test_gpu_rzip.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_rzip3.log; tail -8 gpurun_out/pytest_rzip3.log)
