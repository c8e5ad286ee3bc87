set -x
mkdir -p gpurun_out
# 1. cross-check build: cooperative evaluator vs single-lane evaluator on 256 MiB (wide modes reached)
touch lrzip_next_b200/csrc/k2_commit.cu; make -s -j8 -C lrzip_next_b200/csrc EXTRA=-DK2_CROSSCHECK > /dev/null 2>&1
LRZGPU_DEBUG=1 timeout 300 python tools/prof_small.py 256 > gpurun_out/k2_xcheck4.log 2>&1; tail -2 gpurun_out/k2_xcheck4.log | cut -c1-300
# 2. product build
touch lrzip_next_b200/csrc/k2_commit.cu; make -s -j8 -C lrzip_next_b200/csrc > /dev/null 2>&1
LRZGPU_DEBUG=1 timeout 300 python tools/prof_small.py 256 > gpurun_out/k2_debug6.log 2>&1; tail -2 gpurun_out/k2_debug6.log | cut -c1-700
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu2.log; tail -3 gpurun_out/pytest_gpu2.log
