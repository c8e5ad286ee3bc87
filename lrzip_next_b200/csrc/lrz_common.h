// lrz_common.h -- constants and plain structs shared by host code and sm_100a kernels.
//
// Domain vocabulary follows lrzip-next: a *chunk* is one rzip window; *stream 0* carries match /
// literal headers, *stream 1* the literal bytes; streams are cut into *blocks* of `bufsize`.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LRZ_HD __host__ __device__ __forceinline__
#else
#define LRZ_HD inline
#endif

namespace lrz {

constexpr int kMinMatch = 31;        // MINIMUM_MATCH (reference src/include/lrzip_private.h)
constexpr int kGreatMatch = 1024;    // GREAT_MATCH
constexpr int kTile = 512;           // positions per candidate tile (one warp of a K1 step); candidate regions are tile-strided
constexpr int kInputPad = 8192;      // zero padding after the chunk in HBM (tile / vector over-reads)

// One candidate position that passed the tag mask: same 16-byte layout as the reference's
// struct hash_entry {i64 offset; tag t;} (src/rzip.c:61-64).
struct Cand {
	int64_t pos;
	int64_t tag;
};

struct HEntry {
	int64_t offset;
	int64_t tag;
};

// One emitted long-range match (before 0xFFFF splitting) preceded by its literal run
// [p - lit_len, p).  s0_off / s1_off are the byte offsets of this record's headers in stream 0
// and of its literal bytes in stream 1 (running sums kept by the serial commit kernel, so the
// emit kernels need no scan).  The chunk's closing record has len == 0 (tail literal +
// terminator + CRC).
struct MatchRec {
	int64_t p;      // start of the match in the chunk (already moved back by the reverse extension)
	int64_t ofs;    // start of the earlier copy
	int64_t len;
	int64_t lit_len;
	int64_t s0_off;
	int64_t s1_off;
};

// levels[] of src/rzip.c:67-82
struct RzipLevel {
	unsigned mb_used, initial_freq, max_chain_len;
};
static const RzipLevel kLevels[10] = {
	{1, 4, 1}, {2, 4, 2}, {4, 4, 2}, {8, 4, 2}, {16, 4, 3},
	{32, 4, 4}, {32, 2, 6}, {64, 1, 16}, {64, 1, 32}, {64, 1, 128},
};

// hash_index[] (src/rzip.c:765-771): (random() << 16) ^ random() from the never-seeded glibc
// random() == TYPE_3 additive feedback generator, seed 1.  Pure arithmetic, evaluated on the host
// once and copied to constant / shared memory.
inline void make_hash_index(int64_t hi[256])
{
	int32_t r[344 + 512 + 8];
	r[0] = 1;
	for (int i = 1; i < 31; i++) {
		int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
		if (v < 0)
			v += 2147483647LL;
		r[i] = (int32_t)v;
	}
	for (int i = 31; i < 34; i++)
		r[i] = r[i - 31];
	for (int i = 34; i < 344 + 512; i++)
		r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
	for (int i = 0; i < 256; i++) {
		int64_t a = (int64_t)(((uint32_t)r[344 + 2 * i]) >> 1);
		int64_t b = (int64_t)(((uint32_t)r[344 + 2 * i + 1]) >> 1);
		hi[i] = (a << 16) ^ b;
	}
}

// Persistent per-chunk scan state (struct rzip_state of the reference, src/include/
// lrzip_private.h:441-469, plus the locals of hash_search src/rzip.c:586-597).  Lives in HBM;
// the commit kernel loads it at the start of every segment and stores it back at the end.
struct ScanState {
	int64_t n;            // chunk size
	int64_t end;          // n - kMinMatch
	int64_t hash_count, hash_limit;
	int64_t min_mask;     // minimum_tag_mask (lookup gate)
	int64_t tag_mask;     // insert gate
	int64_t clean_ptr;    // tag_clean_ptr
	int64_t victim_round; // the reference's function-static counter (src/rzip.c:308)
	int64_t last_match;
	int64_t cur_p, cur_ofs, cur_len;
	int64_t scan_pos;     // the loop variable p: positions <= scan_pos are done
	int64_t n_rec;        // match records emitted so far
	int64_t rec_cap;
	int64_t s0_len, s1_len; // running stream lengths
	int32_t hash_bits, max_chain;
	int32_t chunk_bytes;
	int32_t status;       // see enum below
	// statistics (src/rzip.c:1238-1246) and bookkeeping
	int64_t st_inserts, st_tag_hits, st_tag_misses, st_lookups;
	int64_t st_matches, st_match_bytes, st_literals, st_literal_bytes;
	int64_t st_evictions, st_sweeps, st_displacements;
	// commit-kernel bookkeeping: 0 batch rounds, 1 lanes committed in batches, 2 serial steps (pending match),
	// 3 serial (needs serial logic), 4 serial (sweep wrap / window), 5 single-lane re-evaluations, 6 gate cuts,
	// 7 resumed rounds (no evaluation), 8 clock cycles in lane evaluation, 9 cycles in serial steps and match tails,
	// 10 total cycles, 11 cycles in queue refill, 12 sweep scan + classify, 13 validation, 14 commit loop, 15 match tails
	int64_t dbg[16];
	int32_t flags;        // development switches of the commit kernel (LRZGPU_K2_FLAGS): 1 no twin evaluation, 2 exact
	int32_t pad_;         // validation of every lane (no bit filter), 4 no resumed rounds, 8 no soft stoppers (matches
			      // and pending matches through the serial step)
};

enum { kStatusRunning = 0, kStatusChunkDone = 2, kStatusRecOverflow = -1 };

LRZ_HD int64_t pieces_of(int64_t len) { return (len + 0xFFFE) / 0xFFFF; } // 0xFFFF-byte pieces (src/rzip.c:211-225)

} // namespace lrz
