// tests/hostsim/k2_simt.cpp -- TEST INFRASTRUCTURE ONLY.
// The product's batched commit kernel (lrzip_next_b200/csrc/k2_commit.cu: commit warp + 7 evaluator warps, named
// barriers, warp collectives) compiled as C++ and run under the SIMT emulator of simt.h, next to the scalar form of
// the same control logic (k2_commit.cuh with ScalarPrim, itself pinned against the oracle).  `table_bits` overrides
// the hash-table size of the rzip level so that every phase of a long scan -- table full, sweeps, gate tightening,
// the narrow / wide evaluation modes -- is reached within a megabyte of input.
#include "hostsim_common.h"

#define LRZ_SIMT_HOST 1
#include "../../lrzip_next_b200/csrc/k2_commit.cu"
#include "../../lrzip_next_b200/csrc/k4_emit.cu"
#include "../../lrzip_next_b200/csrc/k1_tagscan.cu"

// K1 on the emulated device: the tag scan of positions [lo, hi) with the product's launch arithmetic (k1_launch), a
// grid of three CTAs; the TMA bulk copies are memcpys that complete at once (k1_tagscan.cu, LRZ_SIMT_HOST)
static bool simt_k1(const uint8_t *buf, int64_t n, int64_t lo, int64_t hi, int64_t mask, std::vector<Cand> &cand,
		    std::vector<uint32_t> &tc)
{
	static bool tables = false;
	if (!tables) {
		if (k1_init_tables())
			return false;
		tables = true;
	}
	const int64_t first_tile = lo / kTile, num_tiles = (hi - 1) / kTile - first_tile + 1;
	const int64_t first_step = lo / K1_STEP, num_steps = (hi - 1) / K1_STEP - first_step + 1;
	cand.assign((size_t)num_tiles * kTile, Cand{ -1, -1 });
	tc.assign((size_t)num_tiles, 0xFFFFFFFFu);
	const unsigned grid = (unsigned)(num_steps < 3 ? num_steps : 3);
	return simt::run_grid(grid, K1_THREADS, [&]() {
		k1_tagscan_kernel(buf, n, lo, hi, mask, nullptr, 0, cand.data(), tc.data(), first_tile, num_tiles, first_step,
				  num_steps);
	}, 64 << 10, K1_SMEM);
}

// K4 on the emulated device: chunk CRC, stream-0 headers, stream-1 literal gather (k4_emit.cu) from the commit
// kernel's records, with small grids of the product's block shapes
static bool simt_k4(const uint8_t *buf, int64_t n, int cb, const ScanState &st, const std::vector<MatchRec> &recs,
		    uint8_t **s0_out, uint8_t **s1_out)
{
	static bool tables = false;
	if (!tables) {
		if (k4_init_tables())
			return false;
		tables = true;
	}
	uint8_t *s0 = (uint8_t *)malloc((size_t)st.s0_len + 64), *s1 = (uint8_t *)malloc((size_t)st.s1_len + 64);
	memset(s0, 0xEE, (size_t)st.s0_len + 64);
	memset(s1, 0xEE, (size_t)st.s1_len + 64);
	uint32_t crc = 0;
	bool ok = true;
	if (n > 0)
		ok = simt::run_grid(3, CRC_THREADS, [&]() { crc32_kernel(buf, n, &crc); }, 64 << 10, CRC_SMEM);
	const int64_t n_rec = st.n_rec;
	if (ok && n_rec > 0) {
		int64_t blocks = (n_rec + 7) / 8;
		if (blocks > 4)
			blocks = 4;
		ok = simt::run_grid((unsigned)blocks, 256, [&]() { k4_headers_kernel(recs.data(), n_rec, cb, &crc, s0); });
	}
	if (ok && st.s1_len > 0 && n_rec > 0) {
		int64_t tiles = (st.s1_len + LIT_TILE - 1) / LIT_TILE, grid = tiles < 5 ? tiles : 5;
		ok = simt::run_grid((unsigned)grid, LIT_THREADS, [&]() { k4_literals_kernel(buf, recs.data(), n_rec, 0, st.s1_len, s1); });
	}
	*s0_out = s0;
	*s1_out = s1;
	return ok;
}

// mode 0: scalar primitives, one candidate after the other; mode 1: k2_commit_kernel under the emulator; mode 2: as 1,
// and the streams are made by the K4 kernels under the emulator as well (instead of the scalar stand-in); mode 3: as 2,
// and the candidates come from the K1 kernel under the emulator: every kernel of the rzip stage
extern "C" int simt_rzip_chunk(const uint8_t *data, int64_t n, int rzip_level, int cb, int64_t *victim_round, int64_t seg,
			       int table_bits, int flags, int mode, uint8_t **s0_out, int64_t *s0_len, uint8_t **s1_out,
			       int64_t *s1_len, int64_t *stats /* [10]: [8] as hostsim_rzip_chunk, hash of the final table, records */,
			       int64_t *dbg /* [16] or null */)
{
	int64_t hi_tab[256];
	make_hash_index(hi_tab);
	// the layout the kernels see in HBM: 256 zero bytes in front, kInputPad behind (vector over-reads)
	std::vector<uint8_t> padded((size_t)n + 256 + kInputPad + 64, 0);
	uint8_t *buf = padded.data() + 256;
	buf += (16 - ((uintptr_t)buf & 15)) & 15;
	memcpy(buf, data, (size_t)n);
	ScanState st;
	const int64_t rec_cap = n / kMinMatch + 8;
	k2_init_state(&st, n, rzip_level, cb, *victim_round, rec_cap);
	if (table_bits) {
		st.hash_bits = table_bits;
		st.hash_limit = ((int64_t)1 << table_bits) / 3 * 2;
	}
	st.flags = flags;
	std::vector<HEntry> tab((size_t)1 << st.hash_bits, HEntry{ 0, 0 });
	std::vector<MatchRec> recs((size_t)rec_cap);
	seg = (seg + kTile - 1) / kTile * kTile;
	const int64_t nseg = (n + seg - 1) / seg;
	int64_t mask_lag[2] = { st.min_mask, st.min_mask }; // K1(i) sees the mask as of the end of K2(i-2)
	for (int64_t i = 0; i < nseg; i++) {
		const int64_t lo = i * seg, hi = lo + seg < n ? lo + seg : n;
		std::vector<Cand> cand;
		std::vector<uint32_t> tc;
		if (mode == 3) {
			if (!simt_k1(buf, n, lo, hi, mask_lag[i & 1], cand, tc))
				return -4;
		} else
			scalar_k1(buf, n, lo, hi, mask_lag[i & 1], hi_tab, cand, tc);
		const int64_t first_tile = lo / kTile, num_tiles = (hi - 1) / kTile - lo / kTile + 1;
		cand.resize(cand.size() + 1024, Cand{ 0, 0 }); // the list prefetch looks past the last tile
		if (mode == 0) {
			ScalarPrim prim;
			prim.buf = buf;
			prim.tab = tab.data();
			prim.hmask = ((int64_t)1 << st.hash_bits) - 1;
			prim.cand = cand.data();
			prim.tile_count = tc.data();
			prim.first_tile = first_tile;
			prim.num_tiles = num_tiles;
			prim.seg_hi = hi;
			prim.tile = 0;
			prim.idx = 0;
			k2_commit_segment(prim, &st, recs.data(), i == nseg - 1);
		} else {
			const bool last = i == nseg - 1;
			const bool ok = simt::run_block(K2_THREADS, 0, [&]() {
				k2_commit_kernel(buf, &st, tab.data(), cand.data(), tc.data(), first_tile, num_tiles, hi, recs.data(),
						 last ? 1 : 0, 0, 0);
			});
			if (!ok)
				return -2;
		}
		mask_lag[i & 1] = st.min_mask;
	}
	if (dbg)
		memcpy(dbg, st.dbg, sizeof(st.dbg));
	if (const char *path = getenv("K2_SIMT_TABLE_DUMP")) { // debugging aid: the final hash table
		if (FILE *f = fopen(path, "wb")) {
			fwrite(tab.data(), sizeof(HEntry), tab.size(), f);
			fclose(f);
		}
	}
	if (st.status != kStatusChunkDone)
		return st.status == -9 ? -9 : -1;
	*victim_round = st.victim_round;
	uint8_t *s0, *s1;
	if (mode >= 2) {
		if (!simt_k4(buf, n, cb, st, recs, &s0, &s1))
			return -3;
	} else
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
		uint64_t hsh = 1469598103934665603ull; // the final hash table itself: a misplaced entry shows here first
		for (const HEntry &e : tab) {
			hsh = (hsh ^ (uint64_t)e.offset) * 1099511628211ull;
			hsh = (hsh ^ (uint64_t)e.tag) * 1099511628211ull;
		}
		stats[8] = (int64_t)hsh;
		stats[9] = st.n_rec;
	}
	return 0;
}

// All-values speculation of the cross-window counter (lrzgpu_chunk_begin_all): the kernel as a grid of `nvar` CTAs, CTA v
// starting from victim_round = v with its own state, table and record array (blockIdx.x strides).  Per variant:
// out[v * 6 ..] = status, outgoing victim_round, records, stream-0 length, stream-1 length, hash of the final table.
extern "C" int simt_rzip_chunk_variants(const uint8_t *data, int64_t n, int rzip_level, int cb, int nvar, int64_t seg,
					int table_bits, int flags, int mode, int64_t *out)
{
	int64_t hi_tab[256];
	make_hash_index(hi_tab);
	std::vector<uint8_t> padded((size_t)n + 256 + kInputPad + 64, 0);
	uint8_t *buf = padded.data() + 256;
	buf += (16 - ((uintptr_t)buf & 15)) & 15;
	memcpy(buf, data, (size_t)n);
	const int64_t rec_cap = n / kMinMatch + 8;
	std::vector<ScanState> st((size_t)nvar);
	for (int v = 0; v < nvar; v++) {
		k2_init_state(&st[(size_t)v], n, rzip_level, cb, v, rec_cap);
		if (table_bits) {
			st[(size_t)v].hash_bits = table_bits;
			st[(size_t)v].hash_limit = ((int64_t)1 << table_bits) / 3 * 2;
		}
		st[(size_t)v].flags = flags;
	}
	const int64_t tab_stride = (int64_t)1 << st[0].hash_bits, loosest = st[0].min_mask; // the level's initial gate
	std::vector<HEntry> tab((size_t)(tab_stride * nvar), HEntry{ 0, 0 });
	std::vector<MatchRec> recs((size_t)(rec_cap * nvar));
	seg = (seg + kTile - 1) / kTile * kTile;
	const int64_t nseg = (n + seg - 1) / seg;
	for (int64_t i = 0; i < nseg; i++) {
		const int64_t lo = i * seg, hi = lo + seg < n ? lo + seg : n;
		std::vector<Cand> cand;
		std::vector<uint32_t> tc;
		scalar_k1(buf, n, lo, hi, loosest, hi_tab, cand, tc); // one list for all variants, made with the loosest gate
		const int64_t first_tile = lo / kTile, num_tiles = (hi - 1) / kTile - lo / kTile + 1;
		cand.resize(cand.size() + 1024, Cand{ 0, 0 });
		for (int v = 0; v < nvar; v++) {
			if (mode == 0) {
				ScalarPrim prim;
				prim.buf = buf;
				prim.tab = tab.data() + v * tab_stride;
				prim.hmask = tab_stride - 1;
				prim.cand = cand.data();
				prim.tile_count = tc.data();
				prim.first_tile = first_tile;
				prim.num_tiles = num_tiles;
				prim.seg_hi = hi;
				prim.tile = 0;
				prim.idx = 0;
				k2_commit_segment(prim, &st[(size_t)v], recs.data() + v * rec_cap, i == nseg - 1);
			} else {
				const bool last = i == nseg - 1;
				if (!simt::run_block(K2_THREADS, (unsigned)v, [&]() {
					    k2_commit_kernel(buf, st.data(), tab.data(), cand.data(), tc.data(), first_tile, num_tiles, hi,
							     recs.data(), last ? 1 : 0, tab_stride, rec_cap);
				    }))
					return -2;
			}
		}
	}
	for (int v = 0; v < nvar; v++) {
		const ScanState &s = st[(size_t)v];
		uint64_t hsh = 1469598103934665603ull;
		for (int64_t k = 0; k < tab_stride; k++) {
			const HEntry &e = tab[(size_t)(v * tab_stride + k)];
			hsh = (hsh ^ (uint64_t)e.offset) * 1099511628211ull;
			hsh = (hsh ^ (uint64_t)e.tag) * 1099511628211ull;
		}
		int64_t *o = out + v * 6;
		o[0] = s.status;
		o[1] = s.victim_round;
		o[2] = s.n_rec;
		o[3] = s.s0_len;
		o[4] = s.s1_len;
		o[5] = (int64_t)hsh;
	}
	return 0;
}

// K1 alone: candidates of positions [lo, hi) under `mask`, compacted in position order.  Returns their number (or < 0);
// pos / tag must hold hi - lo entries.
extern "C" int64_t simt_tag_scan(const uint8_t *data, int64_t n, int64_t lo, int64_t hi, int64_t mask, int64_t *pos, int64_t *tag)
{
	std::vector<uint8_t> padded((size_t)n + 256 + kInputPad + 64, 0);
	uint8_t *buf = padded.data() + 256;
	buf += (16 - ((uintptr_t)buf & 15)) & 15;
	memcpy(buf, data, (size_t)n);
	std::vector<Cand> cand;
	std::vector<uint32_t> tc;
	if (hi <= lo)
		return 0;
	if (!simt_k1(buf, n, lo, hi, mask, cand, tc))
		return -1;
	int64_t k = 0;
	for (size_t t = 0; t < tc.size(); t++) {
		if (tc[t] > (uint32_t)kTile)
			return -2; // a tile count the kernel did not write
		for (uint32_t i = 0; i < tc[t]; i++) {
			pos[k] = cand[t * kTile + i].pos;
			tag[k] = cand[t * kTile + i].tag;
			k++;
		}
	}
	return k;
}

extern "C" void simt_free(void *p) { free(p); }
