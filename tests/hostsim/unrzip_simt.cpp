// tests/hostsim/unrzip_simt.cpp -- TEST INFRASTRUCTURE ONLY.
// The decode path's kernels (lrzip_next_b200/csrc/unrzip.cu: stream-0 parser, literal scatter, in-order match replay,
// the LZMA and zstd block decoders) compiled as C++ and run under the SIMT emulator of simt.h with the launch shapes the
// product uses, so that the CPU-only container can feed them the oracle's streams and the reference's LZMA payloads.
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LRZ_SIMT_HOST 1
#include "../../lrzip_next_b200/csrc/unrzip.cu"

using namespace lrz;

// runzip_chunk (src/runzip.c:261-370) on the emulated device: returns the summary's status (0 ok), fills out[0..chunk_size)
extern "C" int simt_unrzip_chunk(const uint8_t *s0, int64_t s0_len, const uint8_t *s1, int64_t s1_len, int cb, int64_t chunk_size,
				 uint8_t *out, int64_t *out_len, uint32_t *crc)
{
	const int64_t cap = s0_len / 3 + 2;
	std::vector<DecLit> lits((size_t)cap);
	std::vector<DecMatch> matches((size_t)cap);
	DecSummary sum;
	memset(&sum, 0, sizeof(sum));
	if (!simt::run_grid(1, 32, [&]() { s0_parse_kernel(s0, s0_len, cb, chunk_size, lits.data(), matches.data(), cap, &sum); }))
		return -100;
	if (sum.status)
		return sum.status;
	if (sum.lit_len != s1_len || sum.out_len != chunk_size)
		return -50;
	if (s1_len > 0 && sum.n_lit > 0)
		if (!simt::run_grid(3, 256, [&]() { lit_scatter_kernel(s1, s1_len, lits.data(), sum.n_lit, out); }))
			return -100;
	if (sum.n_match > 0)
		if (!simt::run_grid(1, 1024, [&]() { match_replay_kernel(matches.data(), sum.n_match, out); }))
			return -100;
	*out_len = sum.out_len;
	*crc = sum.crc;
	return 0;
}

// which: 0 = lzma_dec_kernel (raw LZMA stream, lc3 lp0 pb2, as lzma_decompress_buf hands it over), 1 = zstd_dec_kernel
extern "C" int64_t simt_block_decode(int which, const uint8_t *src, int64_t c_len, uint8_t *out, int64_t u_len)
{
	LzmaDecJob job;
	memset(&job, 0, sizeof(job));
	job.src = src;
	job.c_len = c_len;
	job.out = out;
	job.u_len = u_len;
	bool ok;
	if (which == 0) {
		std::vector<uint16_t> probs((size_t)kNumProbs);
		ok = simt::run_grid(1, 32, [&]() { lzma_dec_kernel(&job, 1, probs.data()); });
	} else {
		std::vector<zd::Work> work(1);
		ok = simt::run_grid(1, 32, [&]() { zstd_dec_kernel(&job, 1, work.data()); }, 256 << 10);
	}
	if (!ok)
		return -100;
	return job.status ? (int64_t)job.status : job.produced;
}
