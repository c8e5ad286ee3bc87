set -x
mkdir -p gpurun_out
timeout 400 python tools/flaky_probe.py 12 > gpurun_out/flaky.log 2>&1; tail -20 gpurun_out/flaky.log | cut -c1-600
