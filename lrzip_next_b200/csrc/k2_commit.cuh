// k2_commit.cuh -- K2: the exact, order-preserving part of rzip's hash_search (src/rzip.c:586-762).
//
// The reference walks every byte position with one thread.  Here the per-position work that is a
// pure function of the bytes (the 31-byte tag and the tag-mask gate) has already been done for the
// whole segment by K1; what remains is the state machine over the *candidate* positions, which must
// see exactly the reference's order of table lookups, inserts, sweeps and match decisions:
//
//   for each candidate (p, t) with p > scan_pos and (t & minimum_tag_mask) == minimum_tag_mask
//       find_best_match                    src/rzip.c:495-534   (probe until an empty slot)
//       insert_hash (+ clean_one_from_hash) src/rzip.c:304-383  when (t & tag_mask) == tag_mask
//       keep the longest match; emit it on GREAT_MATCH or 31 positions later  src/rzip.c:673-688
//       after an emit, jump scan_pos to the end of the match
//
// The control flow below is written once, against a small "primitive" policy P that supplies the
// data-parallel pieces (candidate fetch, probe-window classification, byte compares, sweep scan).
// k2_commit.cu instantiates it with warp-cooperative primitives (one warp owns the chunk's table);
// tests build it with the scalar primitives at the bottom of this header to check the control logic
// on a CPU.  Every value returned by a primitive is warp-uniform.
#pragma once
#include "lrz_common.h"

namespace lrz {

enum ProbeKind { kProbeEmpty = 0, kProbeDue = 1, kProbeDisplace = 2, kProbeChain = 3 };

struct ProbeResult {
	int64_t slot;
	int kind;
	HEntry occ; // the occupant of `slot` (valid for kProbeDisplace)
};

LRZ_HD int tz_ones(int64_t t) // ffsll(~t): 1 + number of trailing one bits (src/rzip.c:295-300)
{
#if defined(__CUDA_ARCH__)
	return __ffsll(~(long long)t);
#endif
	uint64_t v = ~(uint64_t)t;
	int n = 1;
	if (!v)
		return 0;
	while (!(v & 1)) {
		v >>= 1;
		n++;
	}
	return n;
}

// Registers-resident copy of ScanState fields the loop mutates.
struct CommitRegs {
	int64_t hash_count, min_mask, tag_mask, clean_ptr, victim_round, last_match;
	int64_t cur_p, cur_ofs, cur_len, p;
	int64_t n_rec, s0_len, s1_len;
};

template <class P>
LRZ_HD void k2_emit_record(P &prim, ScanState *st, CommitRegs &r, MatchRec *recs, int64_t mp, int64_t mofs,
			   int64_t mlen, int cb, bool closing)
{
	MatchRec rec;
	rec.p = mp;
	rec.ofs = mofs;
	rec.len = mlen;
	rec.lit_len = mp - r.last_match;
	rec.s0_off = r.s0_len;
	rec.s1_off = r.s1_len;
	prim.store_rec(recs + r.n_rec, rec);
	r.n_rec++;
	const int64_t lp = pieces_of(rec.lit_len), mpieces = pieces_of(mlen);
	r.s0_len += 3 * lp + (3 + cb) * mpieces;
	r.s1_len += rec.lit_len;
	if (prim.leader()) {
		st->st_literals += lp;
		st->st_literal_bytes += rec.lit_len;
		st->st_matches += mpieces;
		st->st_match_bytes += mlen;
	}
	if (closing) {
		r.s0_len += 3 + 4; // terminator literal (0,0) + CRC  (src/rzip.c:759-760)
		if (prim.leader())
			st->st_literals += 1;
	}
}

// src/rzip.c:304-353 insert_hash, with the recursion unrolled: all probing of a displacement chain
// happens on the unmodified table (the reference recurses *before* it overwrites the slot), then the
// writes are applied deepest-first.
template <class P>
LRZ_HD void k2_insert(P &prim, ScanState *st, CommitRegs &r, int64_t t, int64_t off, int max_chain)
{
	const int64_t better = (r.min_mask << 1) | 1;
	int64_t slots[48], tags[48], offs[48]; // bitness strictly decreases along a displacement chain: depth <= 47
	int depth = 0;
	for (;;) {
		ProbeResult pr;
		prim.probe(t, better, r.victim_round, max_chain, pr);
		if (pr.kind == kProbeDisplace) {
			slots[depth] = pr.slot;
			tags[depth] = t;
			offs[depth] = off;
			depth++;
			t = pr.occ.tag;
			off = pr.occ.offset;
			if (prim.leader())
				st->st_displacements++;
			continue;
		}
		if (pr.kind == kProbeDue)
			r.hash_count--;
		else if (pr.kind == kProbeChain) {
			r.hash_count--;
			if (++r.victim_round == max_chain)
				r.victim_round = 0;
			if (prim.leader())
				st->st_evictions++;
		}
		prim.store_entry(pr.slot, t, off); // the deepest insert lands first
		break;
	}
	while (depth--)
		prim.store_entry(slots[depth], tags[depth], offs[depth]);
}

// src/rzip.c:357-383 clean_one_from_hash
template <class P>
LRZ_HD int64_t k2_clean_one(P &prim, ScanState *st, CommitRegs &r, int hash_bits)
{
	const int64_t size = (int64_t)1 << hash_bits;
	for (;;) {
		const int64_t better = (r.min_mask << 1) | 1;
		int64_t found;
		if (prim.clean_scan(r.clean_ptr, size, better, found)) {
			prim.clear_entry(found);
			r.clean_ptr = found; // the reference leaves tag_clean_ptr on the deleted slot
			r.hash_count--;
			return better;
		}
		r.min_mask = better;
		r.clean_ptr = 0;
		if (prim.leader())
			st->st_sweeps++;
	}
}

struct CommitConst {
	int64_t end, n, hash_limit, rec_cap;
	int max_chain, hash_bits, cb;
};

struct CommitCounters {
	int64_t lookups = 0, inserts = 0, hits = 0, misses = 0;
};

LRZ_HD void k2_load_regs(const ScanState *st, CommitRegs &r, CommitConst &c)
{
	r.hash_count = st->hash_count;
	r.min_mask = st->min_mask;
	r.tag_mask = st->tag_mask;
	r.clean_ptr = st->clean_ptr;
	r.victim_round = st->victim_round;
	r.last_match = st->last_match;
	r.cur_p = st->cur_p;
	r.cur_ofs = st->cur_ofs;
	r.cur_len = st->cur_len;
	r.p = st->scan_pos;
	r.n_rec = st->n_rec;
	r.s0_len = st->s0_len;
	r.s1_len = st->s1_len;
	c.end = st->end;
	c.n = st->n;
	c.hash_limit = st->hash_limit;
	c.rec_cap = st->rec_cap;
	c.max_chain = st->max_chain;
	c.hash_bits = st->hash_bits;
	c.cb = st->chunk_bytes;
}

LRZ_HD void k2_store_regs(ScanState *st, const CommitRegs &r, const CommitCounters &n, int status)
{
	st->hash_count = r.hash_count;
	st->min_mask = r.min_mask;
	st->tag_mask = r.tag_mask;
	st->clean_ptr = r.clean_ptr;
	st->victim_round = r.victim_round;
	st->last_match = r.last_match;
	st->cur_p = r.cur_p;
	st->cur_ofs = r.cur_ofs;
	st->cur_len = r.cur_len;
	st->scan_pos = r.p;
	st->n_rec = r.n_rec;
	st->s0_len = r.s0_len;
	st->s1_len = r.s1_len;
	st->status = status;
	st->st_lookups += n.lookups;
	st->st_inserts += n.inserts;
	st->st_tag_hits += n.hits;
	st->st_tag_misses += n.misses;
}

// The second half of the loop body at candidate p, once its lookup (best match mlen / offset / reverse, 0 = none) and its
// insert are done: keep the longer match, emit on GREAT_MATCH or 31 positions past the pending match's start
// (src/rzip.c:673-688).  Touches no table slot.  Returns k2_step's `again`.
template <class P>
LRZ_HD bool k2_step_tail(P &prim, ScanState *st, CommitRegs &r, const CommitConst &c, MatchRec *recs, int64_t p,
			 int64_t mlen, int64_t offset, int64_t reverse, int &status)
{
	bool again = false;
	if (mlen > r.cur_len) {
		r.cur_p = p - reverse;
		r.cur_len = mlen;
		r.cur_ofs = offset;
	}
	if ((r.cur_len >= kGreatMatch || p >= r.cur_p + kMinMatch) && r.cur_len >= kMinMatch) {
		if (r.n_rec + 2 > c.rec_cap) {
			status = kStatusRecOverflow;
			return false;
		}
		k2_emit_record(prim, st, r, recs, r.cur_p, r.cur_ofs, r.cur_len, c.cb, false);
		r.last_match = r.cur_p + r.cur_len;
		r.cur_p = r.p = r.last_match;
		r.cur_len = 0;
		again = r.last_match < p;
		prim.publish(r.p, r.min_mask);
	}
	return again;
}

// One iteration of the reference's loop body at candidate (p, t): src/rzip.c:660-688.
// Returns true when p must be examined again: a match is emitted at the first candidate p that lies
// 31 positions past its start; when the match ends before p the reference's loop variable moves
// BACKWARDS to the end of the match (src/rzip.c:685) and walks up to p again, so p itself is looked up
// and inserted a second time (no other candidate can lie in between).
template <class P>
LRZ_HD bool k2_step(P &prim, ScanState *st, CommitRegs &r, const CommitConst &c, CommitCounters &n, MatchRec *recs,
		    int64_t p, int64_t t, int &status)
{
	int64_t mlen = 0, offset = 0, reverse = 0;
	r.p = p;
	prim.publish(p, r.min_mask);
	n.lookups++;
	prim.lookup(t, p, c.end, r.last_match, mlen, offset, reverse, n.hits, n.misses);
	if ((t & r.tag_mask) == r.tag_mask) {
		n.inserts++;
		r.hash_count++;
		k2_insert(prim, st, r, t, p, c.max_chain);
		if (r.hash_count > c.hash_limit)
			r.tag_mask = k2_clean_one(prim, st, r, c.hash_bits);
	}
	return k2_step_tail(prim, st, r, c, recs, p, mlen, offset, reverse, status);
}

// src/rzip.c:710-711 tail literal, :759-760 terminator + CRC
template <class P>
LRZ_HD void k2_close_chunk(P &prim, ScanState *st, CommitRegs &r, const CommitConst &c, MatchRec *recs, int &status)
{
	k2_emit_record(prim, st, r, recs, c.n, 0, 0, c.cb, true);
	r.last_match = c.n;
	status = kStatusChunkDone;
}

// Process the candidates of one segment, strictly one after the other.  `last_segment` closes the
// chunk when the candidates run out.
template <class P>
LRZ_HD void k2_commit_segment(P &prim, ScanState *st, MatchRec *recs, bool last_segment)
{
	CommitRegs r;
	CommitConst c;
	CommitCounters n;
	k2_load_regs(st, r, c);
	int status = st->status;
	if (status == kStatusRunning) {
		int64_t p = 0, t = 0;
		bool again = false;
		for (;;) {
			if (again) {
				again = false;
				if ((t & r.min_mask) != r.min_mask) { // the gate may have tightened meanwhile
					r.p = p;
					continue;
				}
			} else if (!prim.next(r.p, r.min_mask, p, t))
				break;
			again = k2_step(prim, st, r, c, n, recs, p, t, status);
			if (status != kStatusRunning)
				break;
		}
		if (status == kStatusRunning && last_segment)
			k2_close_chunk(prim, st, r, c, recs, status);
	}
	if (prim.leader())
		k2_store_regs(st, r, n, status);
}

// Initial state for a chunk (src/rzip.c:590-626).
inline void k2_init_state(ScanState *st, int64_t n, int rzip_level, int chunk_bytes, int64_t victim_round,
			  int64_t rec_cap)
{
	*st = ScanState();
	const RzipLevel &lv = kLevels[rzip_level];
	const int64_t hashsize = (int64_t)lv.mb_used * (1048576 / 16);
	int bits = 0;
	while (((int64_t)1 << bits) < hashsize)
		bits++;
	st->n = n;
	st->end = n - kMinMatch;
	st->hash_bits = bits;
	st->hash_limit = ((int64_t)1 << bits) / 3 * 2;
	st->min_mask = st->tag_mask = ((int64_t)1 << lv.initial_freq) - 1;
	st->max_chain = (int32_t)lv.max_chain_len;
	st->victim_round = victim_round;
	st->chunk_bytes = chunk_bytes;
	st->rec_cap = rec_cap;
	st->status = kStatusRunning;
	st->flags = 0;
}

// ---------------------------------------------------------------------------------------------
// Scalar primitives: the single-lane meaning of every primitive.  Used by the CPU check of the
// control logic (tests/hostsim) -- never by the product library, whose only path is the warp
// version in k2_commit.cu.
struct ScalarPrim {
	const uint8_t *buf;
	HEntry *tab;
	int64_t hmask;
	const Cand *cand;            // tile-strided candidates of the segment
	const uint32_t *tile_count;
	int64_t first_tile, num_tiles, seg_hi;
	int64_t tile, idx;

	bool leader() const { return true; }
	void publish(int64_t, int64_t) {}
	void store_rec(MatchRec *dst, const MatchRec &r) { *dst = r; }
	void store_entry(int64_t slot, int64_t t, int64_t off)
	{
		tab[slot].tag = t;
		tab[slot].offset = off;
#ifdef LRZ_TRACE
		if (dbg) fprintf(dbg, "W %lld %lld %lld\n", (long long)slot, (long long)t, (long long)off);
#endif
	}
	void clear_entry(int64_t slot) { tab[slot].tag = 0; tab[slot].offset = 0; }

	bool next(int64_t after, int64_t min_mask, int64_t &pos, int64_t &tag)
	{
		for (; tile < num_tiles; tile++, idx = 0) {
			const Cand *c = cand + tile * (int64_t)kTile;
			const int64_t cnt = tile_count[tile];
			for (; idx < cnt; idx++) {
				if (c[idx].pos > after && (c[idx].tag & min_mask) == min_mask) {
					if (c[idx].pos >= seg_hi)
						return false;
					pos = c[idx].pos;
					tag = c[idx].tag;
					idx++;
					return true;
				}
			}
		}
		return false;
	}

	// src/rzip.c:431-461 single_match_len
	int64_t match_len(int64_t p0, int64_t op, int64_t end, int64_t last_match, int64_t &rev) const
	{
		int64_t p = p0, len;
		if (op >= p0)
			return 0;
		while (p < end && buf[p] == buf[op]) {
			p++;
			op++;
		}
		len = p - p0;
		p = p0;
		op -= len;
		const int64_t lo = last_match > 0 ? last_match : 0;
		while (p > lo && op > 0 && buf[op - 1] == buf[p - 1]) {
			op--;
			p--;
		}
		rev = p0 - p;
		len += rev;
		return len < kMinMatch ? 0 : len;
	}

	void lookup(int64_t t, int64_t p, int64_t end, int64_t last_match, int64_t &mlen, int64_t &offset,
		    int64_t &reverse, int64_t &hits, int64_t &misses)
	{
		int64_t h = t & hmask;
		mlen = 0;
		reverse = 0;
		while (tab[h].offset | tab[h].tag) {
			if (tab[h].tag == t) {
				int64_t rev = 0;
				const int64_t l = match_len(p, tab[h].offset, end, last_match, rev);
				if (l) {
					if (l > mlen) {
						mlen = l;
						offset = tab[h].offset - rev;
						reverse = rev;
					}
					hits++;
				} else
					misses++;
			}
			h = (h + 1) & hmask;
		}
	}

	void probe(int64_t t, int64_t better, int64_t victim_round, int max_chain, ProbeResult &pr)
	{
		int64_t h = t & hmask, victim_h = 0;
		int round = 0;
		for (;;) {
			const HEntry he = tab[h];
			if (!(he.offset | he.tag)) {
				pr.slot = h;
				pr.kind = kProbeEmpty;
				return;
			}
			if ((he.tag & better) != better) {
				pr.slot = h;
				pr.kind = kProbeDue;
				return;
			}
			if (tz_ones(he.tag) < tz_ones(t)) {
				pr.slot = h;
				pr.kind = kProbeDisplace;
				pr.occ = he;
				return;
			}
			if (he.tag == t) {
				if (round == victim_round)
					victim_h = h;
				if (++round == max_chain) {
					pr.slot = victim_h;
					pr.kind = kProbeChain;
					return;
				}
			}
			h = (h + 1) & hmask;
		}
	}

	bool clean_scan(int64_t from, int64_t size, int64_t better, int64_t &found)
	{
		for (int64_t i = from; i < size; i++) {
			if (!(tab[i].offset | tab[i].tag))
				continue;
			if ((tab[i].tag & better) != better) {
				found = i;
				return true;
			}
		}
		return false;
	}
};

} // namespace lrz
