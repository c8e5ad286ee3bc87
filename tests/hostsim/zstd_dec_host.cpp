// tests/hostsim/zstd_dec_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the product's zstd frame decoder (lrzip_next_b200/csrc/zstd_dec.cuh) so that the CPU-only container can
// check it against frames made by the system's libzstd at several levels (tests/test_host_logic.py).  The product
// library runs the same source on the GPU, one thread per frame.
#include <stdlib.h>

#include "../../lrzip_next_b200/csrc/zstd_dec.cuh"

extern "C" int64_t hostsim_zstd_decode(const uint8_t *src, int64_t len, uint8_t *out, int64_t cap)
{
	lrz::zd::Work *w = (lrz::zd::Work *)malloc(sizeof(lrz::zd::Work));
	if (!w)
		return -100;
	const int64_t rc = lrz::zd::decode_frame(src, len, out, cap, w);
	free(w);
	return rc;
}
