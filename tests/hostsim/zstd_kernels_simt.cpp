// tests/hostsim/zstd_kernels_simt.cpp -- TEST INFRASTRUCTURE ONLY.
// K6 on the emulated device: the match lists come from K7a's kernels (lzma_mf.cu, through its own orchestration), then
// backend.cu's zstd_encode_kernel (one warp per 128 KiB zstd block) and zstd_assemble_kernel (frame header, block
// offsets, headers and payloads, 256 threads) build the frame of one stream block, with the product's launch shapes.
// The CPU tests hand the frame to the system's libzstd and to the product's own decoder.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LRZ_SIMT_HOST 1
#include "../../lrzip_next_b200/csrc/lzma_mf.cu"
#include "../../lrzip_next_b200/csrc/backend.cu"

using namespace lrz;
using namespace lrz::lzma;

// returns the frame length, 0 when the block stays stored (frame not smaller: `why` says so), < 0 on failure
extern "C" int64_t simt_zstd_frame(const uint8_t *src, int64_t n, int level, uint32_t dict, uint8_t *out, int64_t cap, int *why,
				   int64_t *blocks_compressed)
{
	Config c;
	const uint32_t fb = 64;
	if (!make_config(level < 5 ? 5 : level, dict, fb, (uint64_t)n, c))
		return -1;
	if (mf_init_tables())
		return -3;
	MfBlock B;
	memset(&B, 0, sizeof(B));
	std::vector<uint8_t> padded((size_t)n + 4096, 0);
	memcpy(padded.data(), src, (size_t)n);
	B.src = padded.data();
	B.P.n = (uint32_t)n;
	B.P.fb = c.fb;
	B.P.mc = c.mc;
	B.P.hashMask = c.hashMask;
	B.P.bigHash = c.bigHash;
	B.P.historySize = c.historySize;
	B.P.cyclicSize = c.cyclicSize;
	B.P.hc5 = 0;
	B.count = n >= 4 ? (uint32_t)n - 3 : 0;
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
	if (B.count) {
		std::vector<uint8_t> scratch(mf_sort_scratch_bytes(B.count) + 64);
		if (mf_prepare_block(B, scratch.data(), nullptr, nullptr))
			return -4;
		const uint64_t seg[2] = { 0, B.count };
		if (mf_walk_launch(&B, 1, seg, B.count, false, nullptr, nullptr))
			return -5;
	}
	zs::Tables T;
	zs::build_tables(T);
	const uint32_t nzb = (uint32_t)((n + zs::kBlockMax - 1) / zs::kBlockMax);
	std::vector<zs::Seq> zseq((size_t)nzb * kZsSeqPerBlock);
	std::vector<uint8_t> zlit((size_t)nzb * zs::kBlockMax + 64), zstage((size_t)nzb * kZsStagePerBlock + 64);
	std::vector<uint32_t> zsize(2 * (size_t)nzb + 2, 0xEEEEEEEEu);
	LzmaJob j;
	memset(&j, 0, sizeof(j));
	j.src = padded.data();
	j.n = (uint32_t)n;
	j.out = out;
	j.outCap = (uint64_t)cap;
	j.rec = rec.data();
	j.pool = pool.data();
	j.cfg = c;
	j.mf_overflow = &overflow;
	j.pool_cap = poolCap;
	j.zseq = zseq.data();
	j.zlit = zlit.data();
	j.zstage = zstage.data();
	j.zsize = zsize.data();
	j.nzb = nzb;
	j.gate_result = nullptr;
	if (!simt::run_grid(dim3(nzb, 1), 32, [&]() { zstd_encode_kernel(&j, &T, fb); }, 512 << 10))
		return -6;
	if (!simt::run_grid(dim3(1), 256, [&]() { zstd_assemble_kernel(&j); }))
		return -7;
	*why = j.skipped;
	if (blocks_compressed) {
		*blocks_compressed = 0;
		for (uint32_t x = 0; x < nzb; x++)
			*blocks_compressed += (zsize[x] >> 28) == 2;
	}
	return j.skipped ? 0 : (int64_t)j.outLen;
}
