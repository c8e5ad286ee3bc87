// tests/hostsim/lzma_mf_simt.cpp -- TEST INFRASTRUCTURE ONLY.
// K7a, the data-parallel LZMA match finder (lrzip_next_b200/csrc/lzma_mf.cu: radix-sort kernels, predecessor kernels,
// the bucket-per-thread tree walk / the hash-chain kernel), with its own host orchestration (mf_prepare_block,
// mf_walk_launch) running every launch through the SIMT emulator of simt.h; the block is then encoded by the host build
// of the product's encoder (lzma_enc.cuh) over those match lists, so that the CPU-only container can compare the
// payload with the reference's LzmaCompress.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LRZ_SIMT_HOST 1
#include "../../lrzip_next_b200/csrc/lzma_mf.cu"
#include "../../lrzip_next_b200/csrc/lzma_enc.cuh"

using namespace lrz::lzma;

extern "C" int simt_lzma_encode_mf(const uint8_t *src, int64_t n, int level, uint32_t dict, uint32_t fb, uint8_t *out,
				   int64_t cap, int64_t *out_len, int64_t *pool_words)
{
	Config c;
	if (!make_config(level, dict, fb, (uint64_t)n, c))
		return -1;
	if (mf_init_tables())
		return -3;
	const uint32_t minAvail = c.fastMode ? 5 : 4;
	MfBlock B;
	memset(&B, 0, sizeof(B));
	std::vector<uint8_t> padded((size_t)n + 64, 0);
	memcpy(padded.data(), src, (size_t)n);
	B.src = padded.data();
	B.P.n = (uint32_t)n;
	B.P.fb = c.fb;
	B.P.mc = c.mc;
	B.P.hashMask = c.hashMask;
	B.P.bigHash = c.bigHash;
	B.P.historySize = c.historySize;
	B.P.cyclicSize = c.cyclicSize;
	B.P.hc5 = c.fastMode;
	B.count = (uint32_t)n >= minAvail ? (uint32_t)n - (minAvail - 1) : 0;
	std::vector<uint32_t> son(2 * ((size_t)n + 2), 0xDDDDDDDDu), c2((size_t)B.count + 1), c3((size_t)B.count + 1),
		sorted((size_t)B.count + 1);
	std::vector<uint64_t> rec((size_t)n + 1, 0);
	const uint64_t poolCap = 48ull * (uint64_t)n + 65536;
	std::vector<uint32_t> pool((size_t)poolCap + 1, 0);
	unsigned long long cursor = 0;
	int overflow = 0;
	B.son = son.data();
	B.c2 = c2.data();
	B.c3 = c3.data();
	B.sorted = sorted.data();
	B.rec = rec.data();
	B.pool = pool.data();
	B.poolCap = poolCap;
	B.cursor = &cursor;
	B.overflow = &overflow;
	std::vector<uint8_t> scratch(mf_sort_scratch_bytes(B.count ? B.count : 1) + 64);
	if (mf_prepare_block(B, scratch.data(), nullptr, nullptr))
		return -4;
	const uint64_t seg[2] = { 0, B.count };
	if (mf_walk_launch(&B, 1, seg, B.count, c.fastMode != 0, nullptr, nullptr))
		return -5;
	if (overflow)
		return -6;
	if (pool_words)
		*pool_words = (int64_t)cursor;
	Enc *e = (Enc *)malloc(sizeof(Enc));
	if (!e)
		return -2;
	memset(e, 0xA5, sizeof(Enc));
	enc_init(e, c, src, (uint32_t)n, out, (uint64_t)cap, nullptr, nullptr, nullptr, nullptr);
	e->preRec = rec.data();
	e->prePool = pool.data();
	const uint64_t len = enc_run(e);
	const int ovf = e->overflow;
	free(e);
	*out_len = (int64_t)len;
	return ovf ? 1 : 0;
}
