set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 200 python tools/prof_small.py 64 > gpurun_out/k2_confstat.log 2>&1; tail -5 gpurun_out/k2_confstat.log
