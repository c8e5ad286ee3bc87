// tests/hostsim/filters_simt.cpp -- TEST INFRASTRUCTURE ONLY.
// The block-filter kernels (lrzip_next_b200/csrc/filters.cu: one thread per word / bundle, one per block for the serial
// converters, Delta in two passes over tiles) launched by the product's own filter_blocks_launch through the SIMT
// emulator of simt.h, over a stretch of several stream blocks -- so that the CPU tests can compare them with the
// reference's converters applied block by block.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LRZ_SIMT_HOST 1
#include "../../lrzip_next_b200/csrc/filters.cu"

// converts s[0, n) in place as stream blocks of `bs` bytes (the last one may be short); enc = 0: the decode side
extern "C" int simt_filter_blocks(int filter, int delta, uint8_t *s, int64_t n, int64_t bs, int enc)
{
	std::vector<uint8_t> side(lrz::filter_side_bytes(filter, n, bs) + 64);
	return lrz::filter_blocks_launch(filter, delta, s, 0, n, bs, side.data(), nullptr, nullptr, enc != 0);
}
