set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lzma_block_kernel -c 1 -o gpurun_out/lzma_parser_full python tools/lzma_probe.py 256 > gpurun_out/ncu_lzma.log 2>&1; tail -3 gpurun_out/ncu_lzma.log
