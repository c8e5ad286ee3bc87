set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 120 python tools/prof_small.py 8 > gpurun_out/k2_xcheck.log 2>&1; tail -5 gpurun_out/k2_xcheck.log
