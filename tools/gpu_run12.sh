set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 400 python tools/prof_small.py 256 > gpurun_out/k2_xcheck3.log 2>&1; tail -5 gpurun_out/k2_xcheck3.log
