set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_commit --launch-skip 5 -c 1 -o gpurun_out/k2_full python tools/prof_small.py 128 > gpurun_out/ncu_k2.log 2>&1; tail -3 gpurun_out/ncu_k2.log
