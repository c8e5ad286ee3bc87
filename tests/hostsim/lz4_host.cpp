// tests/hostsim/lz4_host.cpp -- TEST INFRASTRUCTURE ONLY: host build of the product's LZ4 size
// emulation (lrzip_next_b200/csrc/lz4_size.cuh) for checks against the system liblz4.so.1.
#include <stdlib.h>
#include "../../lrzip_next_b200/csrc/lz4_size.cuh"

extern "C" int hostsim_lz4_size(const uint8_t *src, int n, int cap)
{
	uint32_t *t = (uint32_t *)calloc(4096, 4);
	const int r = lrz::lz4s::compress_size(src, n, cap, t);
	free(t);
	return r;
}

extern "C" int hostsim_lz4_gate(const uint8_t *src, int64_t n, int threshold)
{
	uint32_t *t = (uint32_t *)calloc(4096, 4);
	const int r = lrz::lz4s::gate(src, n, threshold, t);
	free(t);
	return r;
}
