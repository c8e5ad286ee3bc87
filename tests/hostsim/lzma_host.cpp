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
	memset(e, 0xA5, sizeof(Enc)); // the device keeps this state in shared memory, which starts with stale contents
	for (uint32_t i = 0; i < kNumOpts; i++) { // ... e.g. cells a previous block left behind as "rep0, length 5"
		e->opt[i].len = 5;
		e->opt[i].dist = 0;
	}
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

// The same encoder over the match lists of the data-parallel pre-pass (lzma_mf.cuh), computed here the
// way lzma_mf.cu does it -- positions grouped by hash, buckets processed one after another in HASH order
// (not in position order) -- to check on the CPU that bucket-wise insertion reproduces the serial finder.
#include <algorithm>
#include <chrono>
#include <numeric>
#include <vector>

#include "../../lrzip_next_b200/csrc/lzma_mf.cuh"

extern "C" {
double hostsim_last_parse_seconds = 0; // parser + range coder alone (development aid)
}

extern "C" int hostsim_lzma_encode_pre(const uint8_t *src, int64_t n, int level, uint32_t dict, uint32_t fb, uint8_t *out,
				       int64_t cap, int64_t *out_len, int64_t *pool_words)
{
	Config c;
	if (!make_config(level, dict, fb, (uint64_t)n, c))
		return -1;
	uint32_t crc[256];
	for (uint32_t i = 0; i < 256; i++)
		crc[i] = mf_crc_entry(i);
	MfParams P = { (uint32_t)n, c.fb, c.mc, c.hashMask, c.bigHash, c.historySize, c.cyclicSize, c.fastMode };
	const uint32_t minAvail = c.fastMode ? 5 : 4; // positions with fewer bytes left are neither searched nor inserted
	const uint32_t count = n >= minAvail ? (uint32_t)n - (minAvail - 1) : 0;
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
	if (c.fastMode) {
		// hash chains: link[pos] = previous position with the same 5-byte hash; then every position on its own,
		// here in DESCENDING order to make the point that nothing is carried between positions
		std::vector<uint32_t> prev5(count);
		prev_by([&](uint32_t i) { return mf_hash5(crc, src + i, P.hashMask); }, prev5);
		std::vector<uint32_t> link((size_t)n + 2, 0);
		for (uint32_t i = 0; i < count; i++)
			link[i + 1] = prev5[i];
		std::vector<std::vector<uint32_t>> lists(count);
		for (uint32_t k = count; k-- > 0;) {
			const uint32_t nd = mf_hc5_matches(src, P, link.data(), k + 1, c2[k], c3[k], d);
			lists[k].assign(d, d + nd);
		}
		for (uint32_t i = 0; i < count; i++) {
			rec[i] = ((uint64_t)pool.size() << kMfCountBits) | lists[i].size();
			pool.insert(pool.end(), lists[i].begin(), lists[i].end());
		}
	} else {
		auto h4 = [&](uint32_t i) { return mf_hash4(crc, src + i, P.hashMask, P.bigHash); };
		std::iota(order.begin(), order.end(), 0u);
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h4(a) < h4(b); });
		for (uint32_t s = 0; s < count;) {
			const uint32_t hv = h4(order[s]);
			uint32_t prev = 0;
			for (; s < count && h4(order[s]) == hv; s++) {
				const uint32_t i = order[s], pos = i + 1;
				// the device's pre-pass (mf_first_kernel): the common prefix with the bucket's previous position,
				// worked out independently of the tree, handed to the insertion as its first comparison
				uint32_t first = kMfFirstNone;
				if (prev) {
					const uint32_t cmCheck = pos <= P.cyclicSize ? 0 : pos - P.cyclicSize;
					if (cmCheck < prev) {
						const uint32_t avail = P.n - i, lenLimit = avail < P.fb ? avail : P.fb;
						first = mf_extend(src + (prev - 1), src + i, 0, lenLimit);
					}
				}
				const uint32_t nbt = mf_bt_insert(src, P, son.data(), pos, prev, d + 4, first);
				const uint32_t nd = mf_mix(src, P, pos, c2[i], c3[i], d, nbt);
				rec[i] = ((uint64_t)pool.size() << kMfCountBits) | nd;
				pool.insert(pool.end(), d, d + nd);
				prev = pos;
			}
		}
	}
	if (pool_words)
		*pool_words = (int64_t)pool.size();
	pool.push_back(0);
	Enc *e = (Enc *)malloc(sizeof(Enc));
	if (!e)
		return -2;
	memset(e, 0xA5, sizeof(Enc));
	enc_init(e, c, src, (uint32_t)n, out, (uint64_t)cap, nullptr, nullptr, nullptr, nullptr);
	e->preRec = rec.data();
	e->prePool = pool.data();
	const auto t0 = std::chrono::steady_clock::now();
	const uint64_t len = enc_run(e);
	hostsim_last_parse_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	const int ovf = e->overflow;
	free(e);
	*out_len = (int64_t)len;
	return ovf ? 1 : 0;
}
