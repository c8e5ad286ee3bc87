// tests/hostsim/zstd_host.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the product's zstd block encoder (lrzip_next_b200/csrc/zstd_enc.cuh) over match lists computed
// the way lzma_mf.cu computes them, so that the CPU-only container can check the frames with the system's
// libzstd decoder (tests/test_host_logic.py).  The product library runs the same source on the GPU.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../lrzip_next_b200/csrc/lzma_enc.cuh"
#include "../../lrzip_next_b200/csrc/lzma_mf.cuh"
#include "../../lrzip_next_b200/csrc/zstd_enc.cuh"

using namespace lrz::lzma;

extern "C" int64_t hostsim_zstd_compress(const uint8_t *src, int64_t n, int level, uint32_t dict, uint8_t *out, int64_t cap,
					  int64_t *blocks_compressed)
{
	Config c;
	const uint32_t fb = 64;
	if (!make_config(level < 5 ? 5 : level, dict, fb, (uint64_t)n, c))
		return -1;
	uint32_t crc[256];
	for (uint32_t i = 0; i < 256; i++)
		crc[i] = mf_crc_entry(i);
	MfParams P = { (uint32_t)n, c.fb, c.mc, c.hashMask, c.bigHash, c.historySize, c.cyclicSize, 0 };
	const uint32_t count = n >= 4 ? (uint32_t)n - 3 : 0;
	std::vector<uint32_t> c2(count), c3(count), order(count), son(2 * ((size_t)n + 2));
	std::vector<uint64_t> rec((size_t)n, 0);
	std::vector<uint32_t> pool;
	auto prev_by = [&](auto keyf, std::vector<uint32_t> &dst) {
		std::iota(order.begin(), order.end(), 0u);
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keyf(a) < keyf(b); });
		for (uint32_t s = 0; s < count; s++)
			dst[order[s]] = (s > 0 && keyf(order[s - 1]) == keyf(order[s])) ? order[s - 1] + 1 : 0;
	};
	prev_by([&](uint32_t i) { return mf_hash2(crc, src + i); }, c2);
	prev_by([&](uint32_t i) { return mf_hash3(crc, src + i); }, c3);
	uint32_t d[2 * 273 + 8];
	auto h4 = [&](uint32_t i) { return mf_hash4(crc, src + i, P.hashMask, P.bigHash); };
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h4(a) < h4(b); });
	for (uint32_t s = 0; s < count;) {
		const uint32_t hv = h4(order[s]);
		uint32_t prev = 0;
		for (; s < count && h4(order[s]) == hv; s++) {
			const uint32_t i = order[s], pos = i + 1;
			const uint32_t nbt = mf_bt_insert(src, P, son.data(), pos, prev, d + 4);
			const uint32_t nd = mf_mix(src, P, pos, c2[i], c3[i], d, nbt);
			rec[i] = ((uint64_t)pool.size() << kMfCountBits) | nd;
			pool.insert(pool.end(), d, d + nd);
			prev = pos;
		}
	}
	pool.push_back(0);

	lrz::zs::Tables T;
	lrz::zs::build_tables(T);
	uint8_t hdr[16];
	int64_t o = lrz::zs::frame_header((uint64_t)n, hdr);
	if (o > cap)
		return -2;
	memcpy(out, hdr, (size_t)o);
	std::vector<lrz::zs::Seq> seq(lrz::zs::kBlockMax / 3 + 2);
	std::vector<uint8_t> lit(lrz::zs::kBlockMax + 16), stage(lrz::zs::kBlockMax + 64);
	int64_t ncomp = 0;
	const uint32_t nb = (uint32_t)((n + lrz::zs::kBlockMax - 1) / lrz::zs::kBlockMax);
	for (uint32_t b = 0; b < nb; b++) {
		const uint32_t lo = b * lrz::zs::kBlockMax, hi = (uint32_t)std::min<int64_t>(n, (int64_t)lo + lrz::zs::kBlockMax);
		const uint32_t size = hi - lo, last = b == nb - 1;
		const uint32_t cs = lrz::zs::encode_block(T, src, (uint32_t)n, lo, hi, rec.data(), pool.data(), fb, seq.data(), lit.data(),
							  stage.data(), (uint32_t)stage.size());
		bool rle = size > 0;
		for (uint32_t i = lo + 1; i < hi && rle; i++)
			rle = src[i] == src[lo];
		uint32_t type, bsz, payload;
		if (rle) {
			type = 1;
			bsz = size;
			payload = 1;
		} else if (cs) {
			type = 2;
			bsz = cs;
			payload = cs;
			ncomp++;
		} else {
			type = 0;
			bsz = size;
			payload = size;
		}
		if (o + 3 + payload > cap)
			return -2;
		const uint32_t h = last | (type << 1) | (bsz << 3);
		out[o++] = (uint8_t)h;
		out[o++] = (uint8_t)(h >> 8);
		out[o++] = (uint8_t)(h >> 16);
		if (type == 1)
			out[o++] = src[lo];
		else if (type == 2) {
			memcpy(out + o, stage.data(), cs);
			o += cs;
		} else {
			memcpy(out + o, src + lo, size);
			o += size;
		}
	}
	if (blocks_compressed)
		*blocks_compressed = ncomp;
	return o;
}
