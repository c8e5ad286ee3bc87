set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
LRZGPU_DEBUG=1 timeout 1500 python bench.py --verify > gpurun_out/bench_c2_v3.json 2> gpurun_out/bench_c2_v3.err; tail -c 1500 gpurun_out/bench_c2_v3.json; grep "commit:" gpurun_out/bench_c2_v3.err | tail -1 | cut -c1-600; grep "lzma wave" gpurun_out/bench_c2_v3.err | tail -1
