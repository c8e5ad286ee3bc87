// tests/hostsim/lzma_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the product's LZMA block encoder (lrzip_next_b200/csrc/lzma_enc.cuh), so that the
// CPU-only container can check it bit-for-bit against the reference's own LzmaCompress
// (oracle/_ref/liblzmaref.so).  The product library runs the same source on the GPU and never
// contains this host build.
#include <stdlib.h>
#include <string.h>

#include "../../lrzip_next_b200/csrc/lzma_enc.cuh"

using namespace lrz::lzma;

extern "C" int hostsim_lzma_encode(const uint8_t *src, int64_t n, int level, uint32_t dict, uint32_t fb, uint8_t *out,
				   int64_t cap, int64_t *out_len)
{
	Config c;
	if (!make_config(level, dict, fb, (uint64_t)n, c))
		return -1;
	Enc *e = (Enc *)malloc(sizeof(Enc));
	uint32_t *h2 = (uint32_t *)calloc(kHash2Size, 4), *h3 = (uint32_t *)calloc(kHash3Size, 4);
	uint32_t *h4 = (uint32_t *)calloc(c.hash4Entries, 4), *son = (uint32_t *)malloc(c.sonEntries * 4);
	if (!e || !h2 || !h3 || !h4 || !son)
		return -2;
	enc_init(e, c, src, (uint32_t)n, out, (uint64_t)cap, h2, h3, h4, son);
	const uint64_t len = enc_run(e);
	const int ovf = e->overflow;
	free(e);
	free(h2);
	free(h3);
	free(h4);
	free(son);
	*out_len = (int64_t)len;
	return ovf ? 1 : 0;
}
