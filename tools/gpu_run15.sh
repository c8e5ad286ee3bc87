set -x
mkdir -p gpurun_out
LRZGPU_DEBUG=1 timeout 1500 python bench.py --verify > gpurun_out/bench_c2_v2.json 2> gpurun_out/bench_c2_v2.err; tail -c 3500 gpurun_out/bench_c2_v2.json; grep "commit:" gpurun_out/bench_c2_v2.err | tail -2 | cut -c1-600; grep "lzma wave" gpurun_out/bench_c2_v2.err | tail -2
