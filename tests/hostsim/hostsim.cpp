// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Builds the product's own commit control logic (lrzip_next_b200/csrc/k2_commit.cuh) with the
// scalar primitives, driven by a scalar stand-in for the K1 tag scan and the K4 emit kernels, so the
// order-sensitive logic (candidate stepping, segment resume, unrolled insert recursion, record
// offsets) can be checked against the oracle on a machine without a GPU.  The product library never
// contains or calls this code.
#include "hostsim_common.h"

extern "C" int hostsim_rzip_chunk(const uint8_t *buf, int64_t n, int rzip_level, int cb, int64_t *victim_round,
				  int64_t seg, uint8_t **s0_out, int64_t *s0_len, uint8_t **s1_out, int64_t *s1_len,
				  int64_t *stats /* [8]: inserts, lookups, hits, misses, evictions, sweeps, hash_count, min_mask */)
{
	int64_t hi_tab[256];
	make_hash_index(hi_tab);
	ScanState st;
	const int64_t rec_cap = n / kMinMatch + 8;
	k2_init_state(&st, n, rzip_level, cb, *victim_round, rec_cap);
	std::vector<HEntry> tab((size_t)1 << st.hash_bits, HEntry{ 0, 0 });
	std::vector<MatchRec> recs((size_t)rec_cap);
	seg = (seg + kTile - 1) / kTile * kTile;
	const int64_t nseg = (n + seg - 1) / seg;
	int64_t mask_lag[2] = { st.min_mask, st.min_mask }; // K1(i) sees the mask as of the end of K2(i-2)
	for (int64_t i = 0; i < nseg; i++) {
		const int64_t lo = i * seg, hi = lo + seg < n ? lo + seg : n;
		std::vector<Cand> cand;
		std::vector<uint32_t> tc;
		scalar_k1(buf, n, lo, hi, mask_lag[i & 1], hi_tab, cand, tc);
		ScalarPrim prim;
		prim.buf = buf;
		prim.tab = tab.data();
		prim.hmask = ((int64_t)1 << st.hash_bits) - 1;
		prim.cand = cand.data();
		prim.tile_count = tc.data();
		prim.first_tile = lo / kTile;
		prim.num_tiles = (hi - 1) / kTile - lo / kTile + 1;
		prim.seg_hi = hi;
		prim.tile = 0;
		prim.idx = 0;
		k2_commit_segment(prim, &st, recs.data(), i == nseg - 1);
		mask_lag[i & 1] = st.min_mask;
	}
	if (st.status != kStatusChunkDone)
		return -1;
	*victim_round = st.victim_round;
	uint8_t *s0, *s1;
	scalar_k4(buf, n, cb, st, recs, &s0, &s1);
	*s0_out = s0;
	*s0_len = st.s0_len;
	*s1_out = s1;
	*s1_len = st.s1_len;
	if (stats) {
		stats[0] = st.st_inserts;
		stats[1] = st.st_lookups;
		stats[2] = st.st_tag_hits;
		stats[3] = st.st_tag_misses;
		stats[4] = st.st_evictions;
		stats[5] = st.st_sweeps;
		stats[6] = st.hash_count;
		stats[7] = st.min_mask;
	}
	return 0;
}

extern "C" void hostsim_free(void *p) { free(p); }
