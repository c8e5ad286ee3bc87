// tests/hostsim/hostsim_common.h -- TEST INFRASTRUCTURE ONLY: scalar stand-ins for the K1 tag scan and the K4 emit
// kernels, shared by the CPU checks of the commit logic (hostsim.cpp: scalar primitives; k2_simt.cpp: the batched
// commit kernel itself under the SIMT emulator).
#pragma once
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <zlib.h>

#include "../../lrzip_next_b200/csrc/k2_commit.cuh"
#include "../../lrzip_next_b200/csrc/lrz_host.h"

using namespace lrz;

static void scalar_k1(const uint8_t *buf, int64_t n, int64_t lo, int64_t hi, int64_t mask, const int64_t *hi_tab,
		      std::vector<Cand> &cand, std::vector<uint32_t> &tc)
{
	const int64_t first_tile = lo / kTile, ntiles = (hi - 1) / kTile - first_tile + 1, end = n - kMinMatch;
	cand.assign((size_t)ntiles * kTile, Cand{ 0, 0 });
	tc.assign((size_t)ntiles, 0);
	for (int64_t p = lo < 1 ? 1 : lo; p < hi && p <= end; p++) {
		int64_t t = 0;
		for (int i = 0; i < kMinMatch; i++)
			t ^= hi_tab[buf[p + i]];
		if ((t & mask) != mask)
			continue;
		const int64_t tile = p / kTile - first_tile;
		cand[(size_t)(tile * kTile + tc[(size_t)tile]++)] = Cand{ p, t };
	}
}

// records -> stream 0 / stream 1 (malloc'ed)
static void scalar_k4(const uint8_t *buf, int64_t n, int cb, const ScanState &st, const std::vector<MatchRec> &recs,
		      uint8_t **s0_out, uint8_t **s1_out)
{
	uint8_t *s0 = (uint8_t *)malloc((size_t)st.s0_len + 1), *s1 = (uint8_t *)malloc((size_t)st.s1_len + 1);
	for (int64_t i = 0; i < st.n_rec; i++) {
		const MatchRec &r = recs[(size_t)i];
		uint8_t *w = s0 + r.s0_off;
		int64_t left = r.lit_len;
		while (left > 0) {
			const int64_t l = left > 0xFFFF ? 0xFFFF : left;
			*w++ = 0;
			*w++ = (uint8_t)l;
			*w++ = (uint8_t)(l >> 8);
			left -= l;
		}
		memcpy(s1 + r.s1_off, buf + r.p - r.lit_len, (size_t)r.lit_len);
		left = r.len;
		while (left > 0) {
			const int64_t l = left > 0xFFFF ? 0xFFFF : left;
			*w++ = 1;
			*w++ = (uint8_t)l;
			*w++ = (uint8_t)(l >> 8);
			put_le(w, r.p - r.ofs, cb);
			w += cb;
			left -= l;
		}
		if (i == st.n_rec - 1) {
			uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
			for (int64_t d = 0; d < n; d += 1 << 30)
				crc = (uint32_t)crc32(crc, buf + d, (uInt)(n - d > (1 << 30) ? (1 << 30) : n - d));
			w[0] = w[1] = w[2] = 0;
			w[3] = (uint8_t)(crc >> 24);
			w[4] = (uint8_t)(crc >> 16);
			w[5] = (uint8_t)(crc >> 8);
			w[6] = (uint8_t)crc;
		}
	}
	*s0_out = s0;
	*s1_out = s1;
}
