// k2_commit.cu -- warp-cooperative primitives for the serial commit stage (see k2_commit.cuh) and
// the kernel that runs them: ONE CTA owns a chunk's hash table (64 MiB at rzip level 7, resident in
// the 126 MB L2) and replays the candidates K1 produced, in position order.
//
// Data-parallel pieces, each one L2/L1 round trip wide instead of one per slot / per byte:
//   * probe windows: 32 consecutive 16-byte slots per load (512 B, 4 lines); empty / equal-tag /
//     due-for-cleaning / lesser-bitness classification by __ballot_sync, first hit by __ffs
//     (find_best_match src/rzip.c:511-531, insert_hash :313-349, clean_one_from_hash :363-378)
//   * match extension: 512 bytes per step forwards and backwards, 16 B per lane, first mismatch by
//     ballot + ffs/clz (single_match_len src/rzip.c:441-454)
//   * candidate fetch: 32 {pos,tag} records per load, mask filter by ballot
//
// The commit warp is latency bound: every candidate costs a chain of dependent loads (probe window,
// then the bytes at the matching offsets).  Candidates are therefore evaluated in batches of up to 32 by
// ALL 8 warps of the CTA (group_eval_t: 8 lanes per candidate, several probe windows in flight), against
// the table as it stands; the commit warp validates in order that no candidate read a slot an earlier one
// of the batch writes, commits the conflict-free prefix and re-evaluates / serialises the rest, so
// exactness rests on the serial step (k2_commit.cuh) and on the ordered validation alone.
#include "k2_commit.cuh"
#if defined(LRZ_SIMT_HOST) // tests/hostsim: this file compiled for the CPU, CUDA threads emulated as fibers (simt.h)
#include "simt.h"
#define K2_PREFETCH_L1(p) ((void)(p))
#else
#include "kernels.h"
#define K2_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#endif

namespace lrz {

static constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void load16u(const uint8_t *p, uint64_t &lo, uint64_t &hi)
{
	const uintptr_t a = (uintptr_t)p;
	const uint64_t *q = (const uint64_t *)(a & ~(uintptr_t)7);
	const unsigned sh = (unsigned)(a & 7) * 8;
	const uint64_t w0 = __ldg(q), w1 = __ldg(q + 1);
	if (sh) {
		const uint64_t w2 = __ldg(q + 2);
		lo = (w0 >> sh) | (w1 << (64 - sh));
		hi = (w1 >> sh) | (w2 << (64 - sh));
	} else {
		lo = w0;
		hi = w1;
	}
}

// Certain "no match of >= 31 bytes between p0 and the earlier offset op" from 16 bytes either side of both
// positions (single_match_len src/rzip.c:431-461 returns 0).  The rolling tag is an XOR over the window, so it
// is blind to byte order and to pairs of equal bytes: on text nearly every lookup meets equal-tag entries
// whose bytes differ at once.  This test lets one lane dismiss one such entry with four independent loads;
// false means "a match, or agreement too long to tell" and the full compare decides.
__device__ __forceinline__ bool quick_no_match(const uint8_t *__restrict__ buf, int64_t p0, int64_t op, int64_t end,
					       int64_t last_match)
{
	if (op >= p0)
		return true;
	uint64_t a0, a1, b0, b1, c0, c1, d0, d1;
	load16u(buf + p0, a0, a1);
	load16u(buf + op, b0, b1);
	load16u(buf + p0 - 16, c0, c1);
	load16u(buf + op - 16, d0, d1);
	const uint64_t x0 = a0 ^ b0, x1 = a1 ^ b1;
	const int cf = x0 ? ((__ffsll((long long)x0) - 1) >> 3) : (x1 ? 8 + ((__ffsll((long long)x1) - 1) >> 3) : 16);
	const int64_t fwd_cap = (end - p0 < kMinMatch) ? end - p0 : kMinMatch;
	if (cf >= 16 && fwd_cap > 16)
		return false;
	const int64_t fwd = cf < fwd_cap ? cf : fwd_cap;
	const int64_t need = kMinMatch - fwd;
	const int64_t lo = last_match > 0 ? last_match : 0;
	int64_t rev_cap = need;
	if (p0 - lo < rev_cap)
		rev_cap = p0 - lo;
	if (op < rev_cap)
		rev_cap = op;
	const uint64_t y0 = c0 ^ d0, y1 = c1 ^ d1;
	const int cr = y1 ? (__clzll((long long)y1) >> 3) : (y0 ? 8 + (__clzll((long long)y0) >> 3) : 16);
	if (cr >= 16 && rev_cap > 16)
		return false;
	const int64_t rev = cr < rev_cap ? cr : rev_cap;
	return rev < need;
}

__device__ __forceinline__ HEntry ld_entry(const HEntry *p)
{
	const longlong2 v = *reinterpret_cast<const longlong2 *>(p);
	HEntry e;
	e.offset = v.x;
	e.tag = v.y;
	return e;
}

__device__ __forceinline__ int nth_set_bit(uint32_t m, int k)
{
	while (k--)
		m &= m - 1;
	return __ffs(m) - 1;
}

struct Progress { // shared memory, written by the commit warp, polled by the helpers
	volatile long long pos;
	volatile long long min_mask;
	volatile int done;
};

// Table slots written since the last batch evaluation (batch commits and serial steps alike): what a lane of that
// evaluation must not have read if its result is to be used in a later round (see "resumed rounds" below).
static constexpr int K2_WLOG = 192;

struct WarpPrim {
	Progress *prog;
	unsigned *wlog = nullptr; // shared memory, K2_WLOG entries
	int *wlog_n = nullptr;    // entries written; > K2_WLOG = overflow
	const uint8_t *buf;
	HEntry *tab;
	int64_t hmask;
	const Cand *cand;
	const uint32_t *tile_count;
	int64_t first_tile, num_tiles, seg_hi;
	int lane;
	// candidate cursor
	int64_t tile;
	uint32_t idx, cnt;
	bool cnt_valid;
	int64_t bpos, btag;
	uint32_t bmask;

	__device__ __forceinline__ bool leader() const { return lane == 0; }
	__device__ __forceinline__ void publish(int64_t p, int64_t mask)
	{
		if (lane == 0) {
			prog->pos = p;
			prog->min_mask = mask;
		}
	}
	__device__ __forceinline__ void store_rec(MatchRec *dst, const MatchRec &r)
	{
		if (lane == 0)
			*dst = r;
	}
	__device__ __forceinline__ void store_entry(int64_t slot, int64_t t, int64_t off)
	{
		if (lane == 0) {
			*reinterpret_cast<longlong2 *>(tab + slot) = make_longlong2(off, t);
			if (wlog) {
				const int k = *wlog_n;
				if (k < K2_WLOG)
					wlog[k] = (unsigned)slot;
				*wlog_n = k + 1;
			}
		}
		__syncwarp();
	}
	__device__ __forceinline__ void clear_entry(int64_t slot) { store_entry(slot, 0, 0); }

	__device__ bool next(int64_t after, int64_t min_mask, int64_t &pos, int64_t &tag)
	{
		for (;;) {
			if (bmask) {
				const bool ok = ((bmask >> lane) & 1) && bpos > after && (btag & min_mask) == min_mask;
				const uint32_t m = __ballot_sync(FULL, ok);
				if (m) {
					const int l = __ffs(m) - 1;
					pos = __shfl_sync(FULL, bpos, l);
					tag = __shfl_sync(FULL, btag, l);
					bmask &= (l == 31) ? 0u : (FULL << (l + 1));
					return pos < seg_hi;
				}
				bmask = 0;
			}
			const int64_t want = (after + 1) / kTile - first_tile;
			if (want > tile) {
				tile = want;
				idx = 0;
				cnt_valid = false;
			}
			for (;;) {
				if (tile >= num_tiles)
					return false;
				if (!cnt_valid) {
					cnt = __ldg(tile_count + tile);
					cnt_valid = true;
				}
				if (idx < cnt)
					break;
				tile++;
				idx = 0;
				cnt_valid = false;
			}
			const uint32_t i = idx + lane;
			const bool v = i < cnt;
			if (v) {
				const longlong2 c = *reinterpret_cast<const longlong2 *>(cand + tile * (int64_t)kTile + i);
				bpos = c.x;
				btag = c.y;
			}
			bmask = __ballot_sync(FULL, v);
			idx += 32;
		}
	}

	// Next 32 records of the K1 list (one per lane, `valid` false past the tile's count); false when
	// the segment's list is exhausted.  Tiles that lie wholly at or before `after` are skipped.
	__device__ bool load_window(int64_t after, int64_t &pos, int64_t &tag, bool &valid)
	{
		const int64_t want = (after + 1) / kTile - first_tile;
		if (want > tile) {
			tile = want;
			idx = 0;
			cnt_valid = false;
		}
		for (;;) {
			if (tile >= num_tiles)
				return false;
			if (!cnt_valid) {
				cnt = __ldg(tile_count + tile);
				cnt_valid = true;
			}
			if (idx < cnt)
				break;
			tile++;
			idx = 0;
			cnt_valid = false;
		}
		const uint32_t i = idx + lane;
		valid = i < cnt;
		if (valid) {
			const longlong2 c = *reinterpret_cast<const longlong2 *>(cand + tile * (int64_t)kTile + i);
			pos = c.x;
			tag = c.y;
		}
		{
			// The list is consumed front to back and, at a tight gate, a refill scans dozens of windows for a handful
			// of keepers: pull the 8 KB behind this window towards L1 (a line per lane, twice) so that those windows
			// do not each wait for L2 / HBM.
			const Cand *pa = cand + tile * (int64_t)kTile + idx + 32 + lane * 8;
			const Cand *lim = cand + num_tiles * (int64_t)kTile;
			if (pa < lim)
				K2_PREFETCH_L1(pa);
			if (pa + 256 < lim)
				K2_PREFETCH_L1(pa + 256);
		}
		idx += 32;
		return true;
	}

	// src/rzip.c:431-461 single_match_len, 512 bytes per step
	__device__ int64_t match_len(int64_t p0, int64_t op, int64_t end, int64_t last_match, int64_t &rev)
	{
		rev = 0;
		if (op >= p0)
			return 0;
		const int64_t maxf = end - p0;
		int64_t f = 0;
		while (f < maxf) {
			const int64_t o = f + lane * 16;
			int64_t neq = 0;
			if (o < maxf) {
				uint64_t a0, a1, b0, b1;
				load16u(buf + p0 + o, a0, a1);
				load16u(buf + op + o, b0, b1);
				const uint64_t x0 = a0 ^ b0, x1 = a1 ^ b1;
				neq = x0 ? ((__ffsll((long long)x0) - 1) >> 3) : (x1 ? 8 + ((__ffsll((long long)x1) - 1) >> 3) : 16);
				if (neq > maxf - o)
					neq = maxf - o;
			}
			const uint32_t m = __ballot_sync(FULL, neq < 16);
			if (m) {
				const int l = __ffs(m) - 1;
				f += l * 16 + __shfl_sync(FULL, neq, l);
				break;
			}
			f += 512;
		}
		const int64_t lo = last_match > 0 ? last_match : 0;
		int64_t maxb = p0 - lo;
		if (op < maxb)
			maxb = op;
		int64_t b = 0;
		while (b < maxb) {
			const int64_t o = b + lane * 16;
			int64_t cntb = 0;
			if (o < maxb) {
				uint64_t a0, a1, b0, b1;
				load16u(buf + p0 - o - 16, a0, a1);
				load16u(buf + op - o - 16, b0, b1);
				const uint64_t x0 = a0 ^ b0, x1 = a1 ^ b1;
				cntb = x1 ? (__clzll((long long)x1) >> 3) : (x0 ? 8 + (__clzll((long long)x0) >> 3) : 16);
				if (cntb > maxb - o)
					cntb = maxb - o;
			}
			const uint32_t m = __ballot_sync(FULL, cntb < 16);
			if (m) {
				const int l = __ffs(m) - 1;
				b += l * 16 + __shfl_sync(FULL, cntb, l);
				break;
			}
			b += 512;
		}
		rev = b;
		const int64_t len = f + b;
		return len < kMinMatch ? 0 : len;
	}

	__device__ void lookup(int64_t t, int64_t p, int64_t end, int64_t last_match, int64_t &mlen, int64_t &offset,
			       int64_t &reverse, int64_t &hits, int64_t &misses)
	{
		int64_t h = t & hmask;
		mlen = 0;
		reverse = 0;
		for (;;) {
			const HEntry e = ld_entry(tab + ((h + lane) & hmask));
			const bool emp = !(e.offset | e.tag);
			const uint32_t em = __ballot_sync(FULL, emp);
			const uint32_t valid = em ? ((1u << (__ffs(em) - 1)) - 1) : FULL;
			uint32_t eq = __ballot_sync(FULL, e.tag == t) & valid;
			if (eq) { // every lane dismisses its own equal-tag entry if its bytes differ at once
				const bool mine = (eq >> lane) & 1;
				const bool no = mine && quick_no_match(buf, p, e.offset, end, last_match);
				const uint32_t nom = __ballot_sync(FULL, no);
				misses += __popc(nom);
				eq &= ~nom;
			}
			while (eq) {
				const int l = __ffs(eq) - 1;
				eq &= eq - 1;
				const int64_t off = __shfl_sync(FULL, e.offset, l);
				int64_t rev;
				const int64_t len = match_len(p, off, end, last_match, rev);
				if (len) {
					if (len > mlen) {
						mlen = len;
						offset = off - rev;
						reverse = rev;
					}
					hits++;
				} else
					misses++;
			}
			if (em)
				break;
			h += 32;
		}
	}

	__device__ void probe(int64_t t, int64_t better, int64_t victim_round, int max_chain, ProbeResult &pr)
	{
		int64_t h = t & hmask, victim = 0;
		int round = 0;
		const int my_ones = tz_ones(t);
		for (;;) {
			const HEntry e = ld_entry(tab + ((h + lane) & hmask));
			const bool emp = !(e.offset | e.tag);
			const bool due = !emp && (e.tag & better) != better;
			const bool lesser = !emp && !due && tz_ones(e.tag) < my_ones;
			const uint32_t em = __ballot_sync(FULL, emp), dm = __ballot_sync(FULL, due);
			const uint32_t stopm = em | dm | __ballot_sync(FULL, lesser);
			const uint32_t valid = stopm ? ((1u << (__ffs(stopm) - 1)) - 1) : FULL;
			const uint32_t eqm = __ballot_sync(FULL, !emp && e.tag == t) & valid;
			const int c = __popc(eqm);
			if (c) {
				const int take = (round + c >= max_chain) ? (max_chain - round) : c;
				if (victim_round >= round && victim_round < round + take)
					victim = (h + nth_set_bit(eqm, (int)(victim_round - round))) & hmask;
				if (round + c >= max_chain) {
					pr.slot = victim;
					pr.kind = kProbeChain;
					return;
				}
				round += c;
			}
			if (stopm) {
				const int l = __ffs(stopm) - 1;
				pr.slot = (h + l) & hmask;
				pr.kind = ((em >> l) & 1) ? kProbeEmpty : (((dm >> l) & 1) ? kProbeDue : kProbeDisplace);
				pr.occ.offset = __shfl_sync(FULL, e.offset, l);
				pr.occ.tag = __shfl_sync(FULL, e.tag, l);
				return;
			}
			h += 32;
		}
	}

	__device__ bool clean_scan(int64_t from, int64_t size, int64_t better, int64_t &found)
	{
		for (int64_t i = from; i < size; i += 32) {
			const int64_t k = i + lane;
			bool q = false;
			if (k < size) {
				const HEntry e = ld_entry(tab + k);
				q = (e.offset | e.tag) && (e.tag & better) != better;
			}
			const uint32_t m = __ballot_sync(FULL, q);
			if (m) {
				found = i + __ffs(m) - 1;
				return true;
			}
		}
		return false;
	}
};

static constexpr int K2_THREADS = 256;   // commit warp + 7 evaluator warps
static constexpr int K2_MAXW = 4;        // insert writes one lane may carry (displacement depth 3)
static constexpr unsigned K2_MAXWALK = 2048; // slots one candidate may walk before it is handed to the serial step

struct LaneEval {
	unsigned wslot[K2_MAXW + 1]; // insert writes, then (optionally) the sweep deletion
	long long wtag[K2_MAXW], woff[K2_MAXW];
	unsigned rlo[K2_MAXW], rlen[K2_MAXW]; // probe ranges read (start slot, length), modulo the table size
	int nw, nr, net, ins, miss;
	int chain; // the insert met max_chain_len equal-tag entries: the victim slot is chosen at commit time (eslot[])
	int twin;  // evaluated on the table as the earlier same-tag candidates of the batch leave it (see group_eval_t)
	unsigned twmask; // those candidates (bit = evaluation slot): their inserts are not conflicts for this lane
	unsigned surv; // equal-tag entries (bit i = i-th met on the walk, offsets in eoff[]) that may give a match of >= 31
		       // bytes: the commit warp measures them (match tail); everything else about the lane is batch work
	bool cx;
};

static constexpr int K2_MAXEQ = 16; // equal-tag entries one candidate may meet in its chain
static constexpr int K2_TWMAX = 7;  // same-tag candidates earlier in the batch that one candidate may build on
static constexpr int K2_RESUME_MIN = 4; // a round without evaluation is worth its validation from this many lanes on

struct FastShared {
	long long qpos[64], qtag[64]; // queue of upcoming candidates that pass the current gate
	unsigned dslot[32];           // sweep deletions of the current batch, in order
	LaneEval ev[32];              // evaluation results of the batch, written by the 8-lane groups
	long long eoff[32][K2_MAXEQ];     // per candidate: offsets of the equal-tag entries met on the walk
	unsigned eslot[32][K2_MAXEQ];     // per candidate: the slots of those entries, in walk order (chain-cap victims)
	unsigned wmask[4096];             // validation filter: bit l of wmask[slot >> 10] = lane l of the batch writes there
	unsigned wlog[K2_WLOG];           // slots written since the last evaluation (WarpPrim::wlog)
	int wlog_n;
	// batch evaluation command, written by the commit warp before barrier 1 (see k2_eval_worker)
	long long cmd_tag_mask, cmd_better, cmd_end, cmd_last_match;
	int cmd_nb, cmd_max_chain, cmd_exit, cmd_mode, cmd_flags;
	Progress prog;
};

// Named barriers of the commit CTA: all 8 warps meet at K2_BAR_GO when the commit warp has queued a batch
// (or wants the workers to leave), and at K2_BAR_DONE when every warp has evaluated its four candidates.
static constexpr int K2_BAR_GO = 1, K2_BAR_DONE = 2;
#if defined(LRZ_SIMT_HOST)
__device__ __forceinline__ void k2_bar(int id) { simt::bar_sync(id, 256); }
#else
__device__ __forceinline__ void k2_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(256) : "memory"); }
#endif

// ---- batched commit ---------------------------------------------------------------------------------
// The reference's loop is serial, but consecutive candidates almost never interact: a candidate reads
// its probe chain [home, first empty slot] and writes one or a few slots.  The commit warp therefore
// evaluates up to 32 queued candidates at once, ONE PER LANE, against the table as it stands at the
// start of the batch, and then validates in order that no candidate read a slot written (inserted into
// or swept) by an earlier candidate of the same batch.  The longest conflict-free prefix is committed
// with exactly the effects the serial loop would have had.  A candidate that may match, or that triggers the
// emission of the pending match, is committed with that prefix too -- its lookup and insert are batch work -- and
// then gets the second half of the loop body from the commit warp (the "match tail", see soft stoppers below).
// The first candidate that needs anything beyond these cases -- more equal-tag entries than the walk records, a
// displacement chain deeper than 3, a sweep that wraps (mask promotion), a write inside the stretch of table the
// sweep is about to visit -- is handed to the serial k2_step().  Chain-cap evictions (victim_round) are ranked
// inside the batch.  The batch therefore never decides anything the serial code would decide differently;
// tests/hostsim runs this very kernel on the CPU (simt.h) against the scalar form of the logic.

__device__ __forceinline__ void prefetch_l1(const void *p) { K2_PREFETCH_L1(p); }

#ifdef K2_CROSSCHECK
// Development aid: the original one-lane evaluator, kept as an in-kernel cross-check of group_eval_t.
// Could the equal-tag entry at `op` give a match of >= 31 bytes at p0?  (single_match_len, bounded.)
__device__ __forceinline__ bool could_match(const uint8_t *__restrict__ buf, int64_t p0, int64_t op, int64_t end,
					    int64_t last_match)
{
	if (op >= p0)
		return false;
	int fwd = 0;
	while (fwd < kMinMatch && p0 + fwd < end && __ldg(buf + p0 + fwd) == __ldg(buf + op + fwd))
		fwd++;
	if (fwd >= kMinMatch)
		return true;
	const int need = kMinMatch - fwd;
	const int64_t lo = last_match > 0 ? last_match : 0;
	int rev = 0;
	while (rev < need && p0 - rev > lo && op - rev > 0 && __ldg(buf + op - rev - 1) == __ldg(buf + p0 - rev - 1))
		rev++;
	return rev >= need;
}

static constexpr int K2_WIDE = 8; // slots fetched per step of a lane's private probe walk (independent loads)

__device__ void lane_eval(const uint8_t *__restrict__ buf, const HEntry *tab, unsigned hmask, int64_t p, int64_t t,
			  int64_t tag_mask, int64_t better, int max_chain, int64_t end, int64_t last_match, int pf_lines,
			  LaneEval &L)
{
	L.nw = L.nr = L.net = L.ins = L.miss = 0;
	L.cx = false;
	const bool do_insert = (t & tag_mask) == tag_mask;
	const unsigned h = (unsigned)t & hmask;
	// inserted tags have all gate bits set, so homes cluster every 2^bits slots and a probe chain is about
	// 2/3 * 2^bits slots long: pull the expected span of the chain into L1 before walking it
	for (int l = 1; l <= pf_lines; l++)
		prefetch_l1(tab + ((h + 8u * (unsigned)l) & hmask));
	const int my_ones = tz_ones(t);
	bool stop = !do_insert;
	int kind = -1, round = 0;
	unsigned sslot = 0, s = 0;
	HEntry occ;
	occ.offset = occ.tag = 0;
	int64_t eq_off[K2_MAXEQ]; // offsets of the equal-tag entries met on the way, compared after the walk
	int neq = 0;
	prefetch_l1(buf + p);
	for (bool done = false; !done;) {
		if (s >= K2_MAXWALK) {
			L.cx = true;
			return;
		}
		HEntry e[K2_WIDE];
		prefetch_l1(tab + ((h + s + 8u * (unsigned)(pf_lines + 1)) & hmask));
#pragma unroll
		for (int k = 0; k < K2_WIDE; k++)
			e[k] = ld_entry(tab + ((h + s + k) & hmask));
#pragma unroll
		for (int k = 0; k < K2_WIDE; k++)
			if (e[k].tag == t && e[k].offset > 0 && e[k].offset < p)
				prefetch_l1(buf + e[k].offset);
#pragma unroll
		for (int k = 0; k < K2_WIDE; k++) {
			if (done)
				break;
			const unsigned slot = (h + s) & hmask;
			if (!(e[k].offset | e[k].tag)) {
				if (!stop) {
					kind = kProbeEmpty;
					sslot = slot;
				}
				done = true;
				break;
			}
			if (!stop) {
				if ((e[k].tag & better) != better) {
					kind = kProbeDue;
					sslot = slot;
					stop = true;
				} else if (tz_ones(e[k].tag) < my_ones) {
					kind = kProbeDisplace;
					occ = e[k];
					sslot = slot;
					stop = true;
				} else if (e[k].tag == t && ++round == max_chain) {
					L.cx = true; // chain cap: victim_round logic stays serial
					return;
				}
			}
			if (e[k].tag == t) {
				if (neq == K2_MAXEQ) {
					L.cx = true;
					return;
				}
				eq_off[neq++] = e[k].offset;
			}
			s++;
		}
	}
	// all lanes compare their q-th equal-tag entry at the same time, so the misses overlap
	for (int q = 0; q < neq; q++) {
		if (could_match(buf, p, eq_off[q], end, last_match)) {
			L.cx = true;
			return;
		}
		L.miss++;
	}
	L.rlo[0] = h;
	L.rlen[0] = s + 1;
	L.nr = 1;
	if (!do_insert)
		return;
	L.ins = 1;
	int64_t ct = t, coff = p;
	for (;;) {
		L.wslot[L.nw] = sslot;
		L.wtag[L.nw] = ct;
		L.woff[L.nw] = coff;
		L.nw++;
		if (kind != kProbeDisplace) {
			L.net = (kind == kProbeEmpty) ? 1 : 0;
			return;
		}
		if (L.nw == K2_MAXW) {
			L.cx = true;
			return;
		}
		// re-home the displaced occupant (the reference recurses before it overwrites the slot)
		ct = occ.tag;
		coff = occ.offset;
		const unsigned h2 = (unsigned)ct & hmask;
		const int ones2 = tz_ones(ct);
		round = 0;
		kind = -1;
		unsigned s2 = 0;
		for (int l = 1; l <= pf_lines; l++)
			prefetch_l1(tab + ((h2 + 8u * (unsigned)l) & hmask));
		for (bool done = false; !done;) {
			if (s2 >= K2_MAXWALK) {
				L.cx = true;
				return;
			}
			HEntry e[K2_WIDE];
			prefetch_l1(tab + ((h2 + s2 + 8u * (unsigned)(pf_lines + 1)) & hmask));
#pragma unroll
			for (int k = 0; k < K2_WIDE; k++)
				e[k] = ld_entry(tab + ((h2 + s2 + k) & hmask));
#pragma unroll
			for (int k = 0; k < K2_WIDE; k++) {
				if (done)
					break;
				sslot = (h2 + s2) & hmask;
				if (!(e[k].offset | e[k].tag)) {
					kind = kProbeEmpty;
					done = true;
					break;
				}
				if ((e[k].tag & better) != better) {
					kind = kProbeDue;
					done = true;
					break;
				}
				if (tz_ones(e[k].tag) < ones2) {
					kind = kProbeDisplace;
					occ = e[k];
					done = true;
					break;
				}
				if (e[k].tag == ct && ++round == max_chain) {
					L.cx = true;
					return;
				}
				s2++;
			}
		}
		L.rlo[L.nr] = h2;
		L.rlen[L.nr] = s2 + 1;
		L.nr++;
	}
}
#endif // K2_CROSSCHECK

// ---- cooperative evaluation -------------------------------------------------------------------------
// lane_eval above walks a candidate's probe chain with ONE lane: a long serial instruction stream, and the
// warp pays the slowest of its 32 lanes.  group_eval gives each candidate 8 lanes instead: one probe step is
// one 128-byte load of 8 consecutive slots, classified by ballot (empty / due / lesser-bitness / equal tag),
// so the per-candidate critical path is a handful of warp instructions per step.  Four candidates are
// evaluated per call (one per 8-lane group); results go to sh->ev[candidate].  The meaning of the result is
// exactly lane_eval's; `cx` (hand the candidate to the serial step) may only ever be set MORE often.

// could_match() with the group's 8 lanes: 4 bytes per lane forwards, then backwards.
__device__ __forceinline__ bool group_could_match(const uint8_t *__restrict__ buf, int64_t p0, int64_t op, int64_t end,
						  int64_t last_match, bool act, bool part, unsigned gshift, int gl)
{
	// act: the group has a pair to test (group-uniform); part: this lane is one of the 8 that compare bytes
	act = act && op < p0;
	part = part && act;
	int c = 0;
	bool stopf = false;
	if (part) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int i = 4 * gl + k;
			if (stopf)
				break;
			if (i >= kMinMatch || p0 + i >= end || __ldg(buf + p0 + i) != __ldg(buf + op + i))
				stopf = true;
			else
				c++;
		}
	}
	unsigned m = (__ballot_sync(FULL, stopf) >> gshift) & 0xffu;
	int fl = m ? __ffs(m) - 1 : 7;
	const int fwd = 4 * fl + __shfl_sync(FULL, c, (int)(gshift + fl));
	// no early exit: the other groups of the warp still need every lane at the ballots below
	const bool fwd_ok = fwd >= kMinMatch;
	const int need = kMinMatch - fwd;
	const int64_t lo = last_match > 0 ? last_match : 0;
	c = 0;
	stopf = false;
	if (part && !fwd_ok) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int j = 4 * gl + k;
			if (stopf)
				break;
			if (j >= need || p0 - j <= lo || op - j <= 0 || __ldg(buf + op - j - 1) != __ldg(buf + p0 - j - 1))
				stopf = true;
			else
				c++;
		}
	}
	m = (__ballot_sync(FULL, stopf) >> gshift) & 0xffu;
	fl = m ? __ffs(m) - 1 : 7;
	const int rev = 4 * fl + __shfl_sync(FULL, c, (int)(gshift + fl));
	if (fwd_ok)
		return act;
	return act && rev >= need;
}

__device__ __forceinline__ unsigned low_mask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// Evaluation modes: the probe chain of a candidate is about 2/3 * 2^bits slots long (bits = number of gate
// bits: inserted tags have them all set, so homes fall on every 2^bits-th slot and the entries between two
// homes pile up behind the first).  Short chains: 8 lanes per candidate, one 8-slot window per step.  Long
// chains (tight gates, i.e. large inputs): the whole warp per candidate and P windows of 32 slots loaded AT
// ONCE per step, so that a chain of hundreds of slots costs one L2 round trip instead of one per window.
enum { K2_MODE_NARROW = 0, K2_MODE_WIDE2 = 1, K2_MODE_WIDE8 = 2 };
__device__ __forceinline__ int k2_mode_for(int64_t tag_mask)
{
	const int bits = __popcll(tag_mask);
	return bits <= 4 ? K2_MODE_NARROW : (bits <= 6 ? K2_MODE_WIDE2 : K2_MODE_WIDE8);
}

// cand_idx: index into sh->qpos / sh->qtag / sh->ev of this lane's GROUP's candidate, or -1 (group idle).
// G lanes per candidate (8 or 32), P windows of G slots in flight per step.
template <int G, int P>
__device__ void group_eval_t(const uint8_t *__restrict__ buf, const HEntry *tab, unsigned hmask, FastShared *sh, int cand_idx,
			     int64_t tag_mask, int64_t better, int max_chain, int64_t end, int64_t last_match, int lane, int warp)
{
	constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
	const int g = lane / G, gl = lane % G;
	long long *eq_list = sh->eoff[cand_idx >= 0 ? cand_idx : 0];
	const unsigned gshift = (unsigned)(g * G);
	const unsigned ltg = (1u << gl) - 1;
	const bool active = cand_idx >= 0;
	int64_t p = 0, t = 0;
	if (active) {
		p = sh->qpos[cand_idx];
		t = sh->qtag[cand_idx];
	}
	LaneEval *R = &sh->ev[active ? cand_idx : 0];
	const bool do_insert = (t & tag_mask) == tag_mask;
	int nw = 0, nr = 0, net = 0, miss = 0;
	bool cx = false;

	// ---- the candidates earlier in the batch that carry the same tag (they walk the same chain and insert into it)
	unsigned pm = 0;
	if (!(sh->cmd_flags & 1)) {
		if (active && do_insert)
			for (int i = gl; i < cand_idx; i += G)
				if (sh->qtag[i] == t)
					pm |= 1u << i;
		for (int o = 1; o < G; o <<= 1)
			pm |= __shfl_xor_sync(FULL, pm, o);
	}
	const int npred = __popc(pm);
	// with one such predecessor that replaces a due entry, this candidate's insert stops at the NEXT stopping slot
	const bool want2 = npred == 1 && (t & better) == better;

	// ---- walk 1: the lookup chain [home, first empty slot], and on the way the insert target
	const unsigned h = (unsigned)t & hmask;
	const int my_ones = tz_ones(t);
	unsigned s = 0, sslot = 0, sslot2 = 0;
	bool stop = !do_insert, done = !active, stop2 = false;
	int kind = -1, round = 0, neq = 0, kind2 = -1, round2 = 0;
	int64_t occ_off = 0, occ_tag = 0, occ2_off = 0, occ2_tag = 0;
	while (__any_sync(FULL, !done)) {
		HEntry ew[P];
#pragma unroll
		for (int i = 0; i < P; i++) {
			ew[i].offset = ew[i].tag = 0;
			if (!done)
				ew[i] = ld_entry(tab + ((h + s + (unsigned)(i * G + gl)) & hmask));
		}
#pragma unroll
		for (int i = 0; i < P; i++) {
			const HEntry e = ew[i];
			const bool emp = !(e.offset | e.tag);
			const bool due = !emp && (e.tag & better) != better;
			const bool les = !emp && !due && tz_ones(e.tag) < my_ones;
			const bool eq = !emp && e.tag == t;
			if (eq && !done && e.offset > 0 && e.offset < p)
				prefetch_l1(buf + e.offset);
			// a window with nothing to act on (no empty slot, no equal tag, no insert target still wanted)
			// only moves the cursor: the common case inside a long chain
			if (!__any_sync(FULL, !done && (emp || eq || ((!stop || (want2 && kind == kProbeDue && !stop2)) && (due || les))))) {
				if (!done) {
					s += G;
					if (s >= K2_MAXWALK)
						cx = done = true;
				}
				continue;
			}
			const unsigned em = (__ballot_sync(FULL, emp) >> gshift) & GM;
			const unsigned dm = (__ballot_sync(FULL, due) >> gshift) & GM;
			const unsigned lm = (__ballot_sync(FULL, les) >> gshift) & GM;
			const unsigned qm = (__ballot_sync(FULL, eq) >> gshift) & GM;
			const int fe = em ? __ffs(em) - 1 : G;
			const unsigned valid = low_mask(fe);
			const unsigned sc = (dm | lm) & valid;
			const int fs = sc ? __ffs(sc) - 1 : G;
			const int src = (int)gshift + (fs % G);
			const int64_t oo = __shfl_sync(FULL, e.offset, src), ot = __shfl_sync(FULL, e.tag, src);
			bool just = false; // the first stopping slot lies in this window
			if (!done) {
				if (!stop) {
					const unsigned before = valid & low_mask(fs);
					round += __popc(qm & before);
					if (round >= max_chain) {
						// chain cap (src/rzip.c:332-343): the max_chain_len-th equal-tag entry comes before any
						// empty / due / lesser slot.  The insert replaces the victim_round-th of them; which one
						// is only known at commit time (rank among the evicting candidates of the batch).
						stop = true;
						kind = kProbeChain;
						if (max_chain > K2_MAXEQ)
							cx = true;
					} else if (sc) {
						stop = true;
						just = true;
						sslot = (h + s + (unsigned)fs) & hmask;
						kind = ((dm >> fs) & 1) ? kProbeDue : kProbeDisplace;
						occ_off = oo;
						occ_tag = ot;
					} else if (em) {
						kind = kProbeEmpty;
						sslot = (h + s + (unsigned)fe) & hmask;
					}
				}
			}
			if (__any_sync(FULL, !done && want2 && kind == kProbeDue && !stop2)) { // the second stopping slot
				const bool need2 = !done && want2 && kind == kProbeDue && !stop2;
				const unsigned from = just ? ~low_mask(fs + 1) : 0xffffffffu;
				const unsigned sc2 = sc & from;
				const int f2 = sc2 ? __ffs(sc2) - 1 : G;
				const int src2 = (int)gshift + (f2 % G);
				const int64_t oo2 = __shfl_sync(FULL, e.offset, src2), ot2 = __shfl_sync(FULL, e.tag, src2);
				if (need2) {
					round2 += __popc(qm & valid & from & low_mask(f2));
					if (sc2) {
						stop2 = true;
						sslot2 = (h + s + (unsigned)f2) & hmask;
						kind2 = ((dm >> f2) & 1) ? kProbeDue : kProbeDisplace;
						occ2_off = oo2;
						occ2_tag = ot2;
					} else if (em) {
						stop2 = true;
						kind2 = kProbeEmpty;
						sslot2 = (h + s + (unsigned)fe) & hmask;
					}
				}
			}
			if (!done) {
				const unsigned eqv = qm & valid;
				if (neq + __popc(eqv) > K2_MAXEQ)
					cx = true;
				else {
					if ((eqv >> gl) & 1) {
						eq_list[neq + __popc(eqv & ltg)] = e.offset;
						sh->eslot[cand_idx][neq + __popc(eqv & ltg)] = (h + s + (unsigned)gl) & hmask;
					}
					neq += __popc(eqv);
				}
				if (em) {
					s += (unsigned)fe;
					done = true;
				} else {
					s += G;
					if (s >= K2_MAXWALK)
						cx = true;
				}
				if (cx)
					done = true;
			}
		}
	}
	__syncwarp();
	// ---- same-tag predecessors ("twins").  The rolling tag is an XOR over the window, so positions p and p + 1 have the
	// SAME tag whenever buf[p] == buf[p + 31] (about one position in 16 on text), repeated phrases give equal tags further
	// apart, and runs give several in a row: a later one of such candidates reads exactly the chain the earlier ones write
	// to, which used to end the batch at it.  All of them walk the same chain from the same table, so this group knows what
	// its predecessors do and evaluates its own candidate on the table as they leave it:
	//   they append at the chain end E, E + 1, ... (those slots are empty): this candidate meets that many more equal-tag
	//     entries (tag misses unless the windows really match) and appends behind them; with ONE predecessor it may
	//     instead replace that entry when it is already due for cleaning, or evict when it completes the chain;
	//   they evict from a full chain: the chain keeps its slots, this candidate evicts the next victim;
	//   the (one) predecessor replaces the first due entry of the chain: this candidate's insert goes to the next slot
	//     that stops an insert (found on the same walk), or replaces the predecessor's entry when the tag itself is due.
	// Anything else is left to the validation (a conflict).  The commit warp skips the predecessors' inserts when it checks
	// this lane's reads (L.twmask); the other writes of those lanes, and every other lane's, are checked as usual.
	bool tw = active && !cx && npred >= 1 && npred <= K2_TWMAX;
	bool tw_evicts = false; // a predecessor may have removed one of the equal-tag entries this walk recorded
	if (__any_sync(FULL, tw)) {
		int ip = -1; // lane k of the group looks at the k-th predecessor
		if (tw && gl < npred) {
			unsigned m = pm;
			for (int k = 0; k < gl; k++)
				m &= m - 1;
			ip = __ffs(m) - 1;
		}
		const bool maym = ip >= 0 && !quick_no_match(buf, p, sh->qpos[ip >= 0 ? ip : 0], end, last_match);
		const bool anym = ((__ballot_sync(FULL, maym) >> gshift) & GM) != 0;
		const bool twE = tw && kind == kProbeEmpty;
		HEntry nx;
		nx.offset = nx.tag = 0;
		if (twE && gl < npred)
			nx = ld_entry(tab + ((sslot + 1u + (unsigned)gl) & hmask));
		const bool occupied = ((__ballot_sync(FULL, twE && gl < npred && (nx.offset | nx.tag) != 0) >> gshift) & GM) != 0;
		if (tw) {
			const bool tdue = (t & better) != better;
			if (anym)
				cx = true; // two of the windows may really match: the serial step decides
			else if (kind == kProbeEmpty) {
				if (occupied)
					tw = false;
				else if (tdue) {
					// while the table is still filling (insert gate == lookup gate) the predecessor's entry itself is
					// "due for cleaning anyway" (src/rzip.c:316-319): this candidate replaces it instead of walking on
					if (npred == 1) {
						miss += 1;
						kind = kProbeDue;
						s += 1;
					} else
						tw = false;
				} else if (round + npred >= max_chain) {
					if (npred == 1 && neq < K2_MAXEQ && max_chain <= K2_MAXEQ) {
						if (gl == 0)
							sh->eslot[cand_idx][neq] = sslot;
						kind = kProbeChain;
						miss += 1;
						s += 1;
						tw_evicts = true;
					} else
						tw = false;
				} else {
					miss += npred;
					sslot = (sslot + (unsigned)npred) & hmask;
					s += (unsigned)npred; // the read range grows by the slots behind the old chain end
				}
			} else if (kind == kProbeChain) {
				tw_evicts = true;
				if (npred >= max_chain)
					tw = false;
			} else if (kind == kProbeDue && npred == 1) {
				if (tdue) { // replaces the predecessor's entry, which sits where the due entry was
					if (occ_tag != t)
						miss += 1;
					tw_evicts = true;
				} else if (stop2 && round + 1 + round2 < max_chain) {
					miss += 1;
					sslot = sslot2;
					kind = kind2;
					occ_off = occ2_off;
					occ_tag = occ2_tag;
				} else
					tw = false;
			} else
				tw = false;
		}
		tw = tw && !cx;
	}
	__syncwarp();
	// ---- equal-tag entries: would any of them give a match?  (then the serial step must decide)
	// One lane per entry dismisses the ones whose bytes differ at once (all loads in flight together);
	// the rare survivors get the group-wide compare.
	unsigned surv = 0;
	const bool soft_on = !(sh->cmd_flags & 8);
	for (int qb = 0; __any_sync(FULL, active && !cx && qb < neq); qb += G) {
		const int q = qb + gl;
		const bool have = active && !cx && q < neq;
		const bool no = have && quick_no_match(buf, p, eq_list[q], end, last_match);
		const unsigned nom = (__ballot_sync(FULL, no) >> gshift) & GM;
		unsigned left = (__ballot_sync(FULL, have) >> gshift) & GM & ~nom;
		if (active && !cx)
			miss += __popc(nom);
		while (__any_sync(FULL, left != 0)) {
			const bool act = left != 0 && !cx;
			const int b = left ? __ffs(left) - 1 : 0;
			left &= left - 1;
			const int64_t op = act ? eq_list[qb + b] : 0;
			const bool cm = group_could_match(buf, p, op, end, last_match, act, gl < 8, gshift, gl & 7);
			if (act) {
				if (cm) {
					if (soft_on)
						surv |= 1u << (qb + b);
					else
						cx = true;
				} else
					miss++;
			}
		}
	}
	// a twin looked at the chain as it was BEFORE its predecessor's insert: if that insert evicts an entry, it may be one
	// of the survivors.  All misses are the same count either way; anything else is for the serial step.
	if (tw && tw_evicts && surv)
		cx = true;
	if (active && gl == 0) {
		R->rlo[0] = h;
		R->rlen[0] = s + 1;
	}
	nr = 1;
	// ---- the insert and the re-homing of displaced occupants (all probing on the unmodified table)
	bool chain = active && !cx && do_insert;
	const int ins = chain ? 1 : 0;
	int64_t wt = t, wo = p;
	while (__any_sync(FULL, chain)) {
		unsigned h2 = 0, s2 = 0;
		int ones2 = 0;
		bool wdone = true;
		if (chain) {
			if (gl == 0) {
				R->wslot[nw] = sslot;
				R->wtag[nw] = wt;
				R->woff[nw] = wo;
			}
			nw++;
			if (kind != kProbeDisplace) {
				net = (kind == kProbeEmpty) ? 1 : 0;
				chain = false;
			} else if (nw == K2_MAXW) {
				cx = true;
				chain = false;
			} else { // re-home the displaced occupant (the reference recurses before it overwrites the slot)
				wt = occ_tag;
				wo = occ_off;
				h2 = (unsigned)wt & hmask;
				ones2 = tz_ones(wt);
				round = 0;
				kind = -1;
				wdone = false;
			}
		}
		while (__any_sync(FULL, !wdone)) {
			HEntry ew[P];
#pragma unroll
			for (int i = 0; i < P; i++) {
				ew[i].offset = ew[i].tag = 0;
				if (!wdone)
					ew[i] = ld_entry(tab + ((h2 + s2 + (unsigned)(i * G + gl)) & hmask));
			}
#pragma unroll
			for (int i = 0; i < P; i++) {
				const HEntry e = ew[i];
				const bool emp = !(e.offset | e.tag);
				const bool due = !emp && (e.tag & better) != better;
				const bool les = !emp && !due && tz_ones(e.tag) < ones2;
				const bool eq = !emp && e.tag == wt;
				if (!__any_sync(FULL, !wdone && (emp || due || les || eq))) {
					if (!wdone) {
						s2 += G;
						if (s2 >= K2_MAXWALK) {
							cx = true;
							wdone = true;
							chain = false;
						}
					}
					continue;
				}
				const unsigned em = (__ballot_sync(FULL, emp) >> gshift) & GM;
				const unsigned dm = (__ballot_sync(FULL, due) >> gshift) & GM;
				const unsigned lm = (__ballot_sync(FULL, les) >> gshift) & GM;
				const unsigned qm = (__ballot_sync(FULL, eq) >> gshift) & GM;
				const unsigned xm = em | dm | lm;
				const int fx = xm ? __ffs(xm) - 1 : G;
				const int src = (int)gshift + (fx % G);
				const int64_t oo = __shfl_sync(FULL, e.offset, src), ot = __shfl_sync(FULL, e.tag, src);
				if (!wdone) {
					round += __popc(qm & low_mask(fx));
					if (round >= max_chain) {
						cx = true;
						wdone = true;
						chain = false;
					} else if (xm) {
						sslot = (h2 + s2 + (unsigned)fx) & hmask;
						kind = ((em >> fx) & 1) ? kProbeEmpty : (((dm >> fx) & 1) ? kProbeDue : kProbeDisplace);
						occ_off = oo;
						occ_tag = ot;
						s2 += (unsigned)fx;
						wdone = true;
						if (gl == 0) {
							R->rlo[nr] = h2;
							R->rlen[nr] = s2 + 1;
						}
						nr++;
					} else {
						s2 += G;
						if (s2 >= K2_MAXWALK) {
							cx = true;
							wdone = true;
							chain = false;
						}
					}
				}
			}
		}
	}
	if (active && gl == 0) {
		R->nw = nw;
		R->nr = nr;
		R->net = cx ? 0 : net;
		R->ins = ins;
		R->miss = miss;
		R->chain = (!cx && ins && nw == 1 && kind == kProbeChain) ? 1 : 0;
		R->twin = (tw && !cx) ? 1 : 0;
		R->twmask = (tw && !cx) ? pm : 0u;
		R->surv = cx ? 0u : surv;
		R->cx = cx;
	}
	__syncwarp();
}

// This warp's share of a batch of nb candidates (warp 0..7): four candidates at once, 8 lanes each, so that
// their table walks and their far-byte compares overlap; tighter gates get more windows in flight per step.
__device__ __noinline__ void k2_eval_share(const uint8_t *__restrict__ buf, const HEntry *tab, unsigned hmask, FastShared *sh, int nb, int mode,
			      int64_t tag_mask, int64_t better, int max_chain, int64_t end, int64_t last_match, int lane, int warp)
{
	const int ci = warp * 4 + (lane >> 3);
	if (warp * 4 >= nb)
		return;
	const int c = ci < nb ? ci : -1;
	if (mode == K2_MODE_NARROW)
		group_eval_t<8, 1>(buf, tab, hmask, sh, c, tag_mask, better, max_chain, end, last_match, lane, warp);
	else if (mode == K2_MODE_WIDE2)
		group_eval_t<8, 4>(buf, tab, hmask, sh, c, tag_mask, better, max_chain, end, last_match, lane, warp);
	else
		group_eval_t<8, 8>(buf, tab, hmask, sh, c, tag_mask, better, max_chain, end, last_match, lane, warp);
}

// Warps 1..7 of the commit CTA: evaluate candidates 4*warp .. 4*warp+3 of every batch the commit warp queues.
__device__ void k2_eval_worker(const uint8_t *__restrict__ buf, const HEntry *tab, unsigned hmask, FastShared *sh, int warp,
			       int lane)
{
	for (;;) {
		k2_bar(K2_BAR_GO);
		if (sh->cmd_exit)
			return;
		k2_eval_share(buf, tab, hmask, sh, sh->cmd_nb, sh->cmd_mode, sh->cmd_tag_mask, sh->cmd_better, sh->cmd_max_chain,
			      sh->cmd_end, sh->cmd_last_match, lane, warp);
		k2_bar(K2_BAR_DONE);
	}
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(FULL, v, o);
	return v;
}

__device__ void k2_commit_segment_batched(WarpPrim &prim, FastShared *sh, ScanState *st, MatchRec *recs, bool last_segment)
{
	CommitRegs r;
	CommitConst c;
	CommitCounters n;
	k2_load_regs(st, r, c);
	int status = st->status;
	const int lane = prim.lane;
	const unsigned lt = (1u << lane) - 1;
	const unsigned hmask = (unsigned)prim.hmask;
	const int64_t tsize = prim.hmask + 1;
	int qn = 0;
	bool list_done = false, again = false;
	int64_t again_p = 0, again_t = 0;
	int64_t n_disp = 0, n_evict = 0;
	int64_t dbg[16] = { 0 };
	const long long clk_start = clock64();
	// ---- resumed rounds.  A batch usually ends at a candidate that needs the serial step (a match, mostly), long before
	// its 32 evaluations are used up, and evaluating is two thirds of a round.  The results of the lanes behind that
	// candidate stay in sh->ev: they are still what an evaluation on the current table would return as long as no
	// slot they read has been written since -- by the lanes committed from that evaluation or by the serial steps in
	// between, all of which go through the write log -- and the gates are the same.  (last_match only grows, which can
	// only turn "could match" into "cannot": a lane evaluated as match-free stays match-free, and lanes that could
	// match go to the serial step anyway.)  The next round then skips the evaluation: queue entry i takes the result in
	// ev slot ev_base + i, is checked against the log on top of the usual in-batch validation, and a lane that fails
	// ends the round and sends everything from it on into a fresh evaluation.
	int ev_n = 0, ev_base = 0; // evaluated slots in sh->ev; queue entry 0 <-> slot ev_base
	int64_t ev_tag_mask = 0, ev_min_mask = 0;
	const bool resume_on = !(sh->cmd_flags & 4);
	// ---- soft stoppers.  A candidate that may have a match -- equal-tag entries the evaluation could not dismiss -- or
	// that triggers the emission of the pending match (31 positions past its start, src/rzip.c:678) still does nothing
	// to the table but its own lookup + insert, which the batch commits like any other lane's.  What is left is the
	// second half of the loop body: measure the surviving entries (match_len, with the current last_match), keep the
	// longest, emit.  The commit warp does that right after committing the lane ("match tail") and ends the round there,
	// because an emission moves the scan past candidates of the batch; the lanes behind it are taken up by a resumed
	// round.  While a match is pending the candidates inside its 31 positions that cannot match are plain lanes.
	const bool soft_on = !(sh->cmd_flags & 8);
	prim.wlog = sh->wlog;
	prim.wlog_n = &sh->wlog_n;
	if (lane == 0)
		sh->wlog_n = 0;
	__syncwarp();

	// drop queue entries the scan has moved past or that fail the (possibly tightened) gate
	auto filter_queue = [&]() {
		__syncwarp();
		long long p0 = 0, t0 = 0, p1 = 0, t1 = 0;
		bool k0 = false, k1 = false;
		if (lane < qn) {
			p0 = sh->qpos[lane];
			t0 = sh->qtag[lane];
			k0 = p0 > r.p && (t0 & r.min_mask) == r.min_mask;
		}
		if (lane + 32 < qn) {
			p1 = sh->qpos[lane + 32];
			t1 = sh->qtag[lane + 32];
			k1 = p1 > r.p && (t1 & r.min_mask) == r.min_mask;
		}
		const unsigned b0 = __ballot_sync(FULL, k0), b1 = __ballot_sync(FULL, k1);
		__syncwarp();
		if (k0) {
			const int d = __popc(b0 & lt);
			sh->qpos[d] = p0;
			sh->qtag[d] = t0;
		}
		if (k1) {
			const int d = __popc(b0) + __popc(b1 & lt);
			sh->qpos[d] = p1;
			sh->qtag[d] = t1;
		}
		{ // evaluated entries keep their slots only if what was dropped is a prefix of the queue
			const int qn_old = qn, removed = qn_old - (__popc(b0) + __popc(b1));
			const unsigned long long keep = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
			const unsigned long long all = qn_old >= 64 ? ~0ull : ((1ull << qn_old) - 1);
			const unsigned long long gone = removed >= 64 ? ~0ull : ((1ull << removed) - 1);
			if (keep == (all & ~gone))
				ev_base += removed;
			else
				ev_n = 0;
		}
		qn = __popc(b0) + __popc(b1);
		__syncwarp();
	};
	auto pop_front = [&](int k) {
		__syncwarp();
		long long p0 = 0, t0 = 0, p1 = 0, t1 = 0;
		const int i0 = lane + k, i1 = lane + 32 + k;
		if (i0 < qn) {
			p0 = sh->qpos[i0];
			t0 = sh->qtag[i0];
		}
		if (i1 < qn) {
			p1 = sh->qpos[i1];
			t1 = sh->qtag[i1];
		}
		__syncwarp();
		if (i0 < qn) {
			sh->qpos[lane] = p0;
			sh->qtag[lane] = t0;
		}
		if (i1 < qn) {
			sh->qpos[lane + 32] = p1;
			sh->qtag[lane + 32] = t1;
		}
		qn -= k;
		ev_base += k;
		__syncwarp();
	};

	while (status == kStatusRunning) {
		if (again) { // re-examine the candidate a match emission jumped back over (see k2_step)
			again = false;
			if ((again_t & r.min_mask) == r.min_mask) {
				again = k2_step(prim, st, r, c, n, recs, again_p, again_t, status);
				filter_queue();
			} else
				r.p = again_p;
			continue;
		}
		const long long cr0 = clock64();
		while (qn < 32 && !list_done) { // refill from the K1 list
			int64_t wp = 0, wt = 0;
			bool wv = false;
			if (!prim.load_window(r.p, wp, wt, wv)) {
				list_done = true;
				break;
			}
			const bool keep = wv && wp > r.p && wp < prim.seg_hi && (wt & r.min_mask) == r.min_mask;
			const unsigned bm = __ballot_sync(FULL, keep);
			if (keep) {
				const int d = qn + __popc(bm & lt);
				sh->qpos[d] = wp;
				sh->qtag[d] = wt;
			}
			qn += __popc(bm);
			__syncwarp();
		}
		dbg[11] += clock64() - cr0;
		if (qn == 0)
			break;
		if (r.cur_len > 0 && !soft_on) { // (development switch 8) a match is pending: strictly serial until it is emitted
			const int64_t p = sh->qpos[0], t = sh->qtag[0];
			pop_front(1);
			const long long c0 = clock64();
			again = k2_step(prim, st, r, c, n, recs, p, t, status);
			dbg[9] += clock64() - c0;
			dbg[2]++;
			again_p = p;
			again_t = t;
			filter_queue();
			continue;
		}

		// ---- evaluate up to 32 candidates, one per lane, on the table as it stands -- or take up the unused results
		// of the last evaluation
		const int wlog_now = __shfl_sync(FULL, sh->wlog_n, 0); // one read for the warp: lane 0 resets it further down
		const bool resumed = resume_on && ev_n - ev_base >= K2_RESUME_MIN && ev_n - ev_base <= qn &&
				     r.tag_mask == ev_tag_mask && r.min_mask == ev_min_mask && wlog_now <= K2_WLOG;
		const int nb = resumed ? ev_n - ev_base : (qn < 32 ? qn : 32);
		const int eb = resumed ? ev_base : 0; // ev / eslot slot of lane 0
		const int64_t better = (r.min_mask << 1) | 1;
		int pf_lines;
		{
			const int bits = __popcll(r.tag_mask);
			const int64_t span = bits >= 9 ? 384 : (((int64_t)2 << bits) / 3 + 8); // slots
			pf_lines = (int)((span + 7) >> 3);
			if (pf_lines > 24)
				pf_lines = 24;
		}
		const int mode = k2_mode_for(r.tag_mask);
		LaneEval L;
		L.nw = L.nr = L.net = L.ins = L.miss = L.chain = L.twin = 0;
		L.surv = L.twmask = 0;
		L.cx = false;
		int64_t myp = 0, myt = 0;
		const long long ce0 = clock64();
		if (resumed) {
			dbg[7]++;
			if (lane < nb) {
				L = sh->ev[eb + lane];
				myp = sh->qpos[lane];
			}
			__syncwarp();
		} else {
		dbg[0]++;
		if (lane < nb) { // pull each candidate's home line(s) and window bytes towards L1 before the passes
			myp = sh->qpos[lane];
			myt = sh->qtag[lane];
			const unsigned hh = (unsigned)myt & hmask;
			prefetch_l1(prim.buf + myp);
			if (mode == K2_MODE_NARROW)
				for (int l = 0; l <= pf_lines; l++)
					prefetch_l1(prim.tab + ((hh + 8u * (unsigned)l) & hmask));
		}
		if (lane == 0) { // all 8 warps of the CTA evaluate four candidates each
			sh->wlog_n = 0;
			sh->cmd_tag_mask = r.tag_mask;
			sh->cmd_better = better;
			sh->cmd_end = c.end;
			sh->cmd_last_match = r.last_match;
			sh->cmd_nb = nb;
			sh->cmd_max_chain = c.max_chain;
			sh->cmd_mode = mode;
		}
		k2_bar(K2_BAR_GO);
		k2_eval_share(prim.buf, prim.tab, hmask, sh, nb, mode, r.tag_mask, better, c.max_chain, c.end, r.last_match, lane, 0);
		k2_bar(K2_BAR_DONE);
		if (lane < nb)
			L = sh->ev[lane];
		__syncwarp();
		ev_n = nb;
		ev_base = 0;
		ev_tag_mask = r.tag_mask;
		ev_min_mask = r.min_mask;
#ifdef K2_CROSSCHECK
		{ // development aid: the single-lane evaluator must agree with the cooperative one
			bool bad = false;
			LaneEval L2;
			L2.nw = L2.nr = L2.net = L2.ins = L2.miss = 0;
			L2.cx = false;
			if (lane < nb) {
				lane_eval(prim.buf, prim.tab, hmask, myp, myt, r.tag_mask, better, c.max_chain, c.end, r.last_match, 0, L2);
				if (L2.cx)
					L2.net = 0;
				if (L.cx != L2.cx)
					bad = !L.cx && !L.chain; // the cooperative form may only be stricter (chain caps: decided at commit)
				else if (!L.cx) {
					bad = L.nw != L2.nw || L.nr != L2.nr || L.net != L2.net || L.ins != L2.ins || L.miss != L2.miss;
					for (int w = 0; w < L.nw && !bad; w++)
						bad = L.wslot[w] != L2.wslot[w] || L.wtag[w] != L2.wtag[w] || L.woff[w] != L2.woff[w];
					for (int q = 0; q < L.nr && !bad; q++)
						bad = L.rlo[q] != L2.rlo[q] || L.rlen[q] != L2.rlen[q];
				}
			}
			const unsigned bm = __ballot_sync(FULL, bad);
			if (bm) {
				if (lane == __ffs(bm) - 1) {
					st->dbg[0] = myp;
					st->dbg[1] = myt;
					st->dbg[2] = ((int64_t)L.cx << 32) | (unsigned)L2.cx;
					st->dbg[3] = ((int64_t)L.nw << 32) | (unsigned)L2.nw;
					st->dbg[4] = ((int64_t)L.nr << 32) | (unsigned)L2.nr;
					st->dbg[5] = ((int64_t)L.net << 32) | (unsigned)L2.net;
					st->dbg[6] = ((int64_t)L.wslot[0] << 32) | L2.wslot[0];
					st->dbg[7] = ((int64_t)L.wslot[1] << 32) | L2.wslot[1];
					st->dbg[8] = ((int64_t)L.rlo[0] << 32) | L2.rlo[0];
					st->dbg[9] = ((int64_t)L.rlen[0] << 32) | L2.rlen[0];
					st->dbg[10] = ((int64_t)L.rlen[1] << 32) | L2.rlen[1];
					st->dbg[11] = ((int64_t)L.miss << 32) | (unsigned)L2.miss;
					st->dbg[12] = L.wtag[0];
					st->dbg[13] = L2.wtag[0];
					st->dbg[14] = r.tag_mask;
					st->dbg[15] = ((int64_t)lane << 32) | (unsigned)nb;
					st->status = -9;
				}
				status = -9;
				break;
			}
		}
#endif
		} // fresh evaluation
		dbg[8] += clock64() - ce0;
		if (L.cx)
			L.net = 0;
		// chain-cap evictions (src/rzip.c:332-343) inside the batch: every evicting candidate advances the
		// reference's round-robin counter by one, so the k-th evicting lane of the batch sees
		// victim_round + k and replaces that one of the max_chain_len equal-tag entries its walk recorded.
		// Lanes commit as a prefix, so every evicting lane before a committing lane commits as well.
		const bool evict = lane < nb && !L.cx && L.chain != 0;
		const unsigned evm = __ballot_sync(FULL, evict);
		if (evict) {
			const int vi = (int)((r.victim_round + __popc(evm & lt)) % c.max_chain);
			L.wslot[0] = sh->eslot[eb + lane][vi];
		}
		const long long cs0 = clock64();
		// sweep deletions (clean_one_from_hash): the k-th insert that overfills the table removes the
		// k-th entry, in table order from tag_clean_ptr, that lacks the next-stricter mask
		// (only the lanes up to the first candidate that ends the round can commit: the sweep is scanned for them alone)
		const unsigned pre_stop = __ballot_sync(FULL, lane < nb && (L.cx || (soft_on && (L.surv != 0 ||
							  (r.cur_len > 0 && myp >= r.cur_p + kMinMatch)))));
		const int nv0 = pre_stop ? __ffs(pre_stop) : nb;
		const unsigned netm = __ballot_sync(FULL, L.net != 0);
		bool cl = lane < nv0 && L.net && (r.hash_count + __popc(netm & (lt | (1u << lane))) > c.hash_limit);
		const unsigned clm = __ballot_sync(FULL, cl);
		const int ncl = __popc(clm), crank = __popc(clm & lt);
		int found = 0;
		unsigned dmax = 0;
		if (ncl) {
			// four windows of 32 slots in flight per step: crossing a stretch without due entries (the
			// regions of tags that already satisfy the stricter mask) costs one L2 round trip per 128 slots
			for (int64_t cp = r.clean_ptr; found < ncl && cp < tsize; cp += 128) {
				HEntry e4[4];
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int64_t k = cp + u * 32 + lane;
					e4[u].offset = e4[u].tag = 0;
					if (k < tsize)
						e4[u] = ld_entry(prim.tab + k);
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int64_t k = cp + u * 32 + lane;
					const bool q = (e4[u].offset | e4[u].tag) && (e4[u].tag & better) != better;
					const unsigned bm = __ballot_sync(FULL, q);
					if (q) {
						const int rk = found + __popc(bm & lt);
						if (rk < ncl)
							sh->dslot[rk] = (unsigned)k;
					}
					found += __popc(bm);
				}
			}
			if (found > ncl)
				found = ncl;
			__syncwarp();
			if (found)
				dmax = sh->dslot[found - 1];
			{ // the sweep moves on from here in the next rounds: have the following 512 slots on their way
				const int64_t k0 = (found ? (int64_t)dmax : r.clean_ptr) + 128 + lane * 8;
				if (k0 < tsize)
					prefetch_l1(prim.tab + k0);
				if (k0 + 256 < tsize)
					prefetch_l1(prim.tab + k0 + 256);
			}
		}
		// flags of one lane after an evaluation: `stopper` = must go through the serial step
		bool stopper = false, soft = false;
		unsigned del = 0;
		int nwt = 0;
		auto classify = [&]() {
			stopper = L.cx;
			soft = soft_on && lane < nb && !L.cx && (L.surv != 0 || (r.cur_len > 0 && myp >= r.cur_p + kMinMatch));
			del = 0;
			if (cl) {
				if (crank >= found)
					stopper = true; // the sweep has to wrap first (mask promotion)
				else
					del = sh->dslot[crank];
			}
			if (found && !stopper) // a write inside the stretch the sweep visits in this batch changes the sweep
				for (int w = 0; w < L.nw; w++)
					if ((int64_t)L.wslot[w] >= r.clean_ptr && L.wslot[w] <= dmax)
						stopper = true;
			nwt = L.nw;
			if (cl && !stopper)
				L.wslot[nwt++] = del;
		};
		classify();
		dbg[12] += clock64() - cs0;
		const long long cv0 = clock64();
		// ---- ordered validation: cmask bit j = this lane read a slot that lane j (< lane) writes.
		// Only the lanes up to the first stopper can commit in this round, so only they are checked: their
		// writes are gathered (in lane order) into a shared list that every reader lane scans once.
		// ---- ordered validation: a lane may commit only if it read no slot that an earlier lane of the batch
		// writes.  Only lanes up to the first stopper can commit, and the batch ends at the first conflict, so all
		// that is needed is the FIRST lane with a conflict.  Filter: writers mark the 1024-slot stretch of each
		// write in a shared bit table, readers look up the stretches their probe ranges touch; the few lanes that
		// meet an earlier writer's mark are then checked exactly, in lane order, by the whole warp at once (every
		// earlier lane tests its own writes against the candidate's ranges).
		unsigned cmask = 0;
		{
			const unsigned stop0 = __ballot_sync(FULL, lane < nb && (stopper || soft));
			const int nv = stop0 ? __ffs(stop0) : nb;
			const int myn = lane < nv ? nwt : 0;
			const int mynr = lane < nv ? L.nr : 0;
			for (int w = 0; w < myn; w++)
				atomicOr(&sh->wmask[L.wslot[w] >> 10], 1u << lane);
			__syncwarp();
			unsigned seen = 0;
			for (int q = 0; q < mynr; q++) {
				const unsigned b0 = L.rlo[q] >> 10, b1 = ((L.rlo[q] + L.rlen[q] - 1) & hmask) >> 10;
				const unsigned nbk = (unsigned)(tsize >> 10);
				for (unsigned bk = b0;; bk = (bk + 1 == nbk) ? 0 : bk + 1) {
					seen |= sh->wmask[bk];
					if (bk == b1)
						break;
				}
			}
			unsigned F = __ballot_sync(FULL, (seen & lt) != 0 || ((sh->cmd_flags & 2) && lane < nv && lane > 0));
			__syncwarp();
			for (int w = 0; w < myn; w++)
				sh->wmask[L.wslot[w] >> 10] = 0;
			const unsigned w0 = myn > 0 ? L.wslot[0] : 0, w1 = myn > 1 ? L.wslot[1] : 0;
			int first_conf = 32;
			while (F) {
				const int f = __ffs(F) - 1;
				F &= F - 1;
				const int nrf = __shfl_sync(FULL, mynr, f);
				const unsigned twf = __shfl_sync(FULL, L.twmask, f); // same-tag predecessors, by evaluation slot
				bool hit = false;
#pragma unroll
				for (int q = 0; q < K2_MAXW; q++) {
					if (q >= nrf)
						break;
					const unsigned lo = __shfl_sync(FULL, L.rlo[q], f), len = __shfl_sync(FULL, L.rlen[q], f);
					if (lane < f) {
						// a twin was evaluated on the table as its predecessors' inserts leave it
						if (myn > 0 && !((twf >> (eb + lane)) & 1) && ((w0 - lo) & hmask) < len)
							hit = true;
						if (myn > 1 && ((w1 - lo) & hmask) < len)
							hit = true;
						for (int w = 2; w < myn; w++)
							if (((L.wslot[w] - lo) & hmask) < len)
								hit = true;
					}
				}
				if (__ballot_sync(FULL, hit)) {
					first_conf = f;
					break;
				}
			}
			if (resumed) { // reads against everything written since the evaluation
				const int nlog = sh->wlog_n;
				bool hitw = false;
				if (lane < nv) {
					if (L.twmask & low_mask(eb))
						hitw = true; // a predecessor it was evaluated behind is no longer part of the batch
					for (int q = 0; q < L.nr && !hitw; q++) {
						const unsigned lo = L.rlo[q], len = L.rlen[q];
						for (int w = 0; w < nlog; w++)
							if (((sh->wlog[w] - lo) & hmask) < len) {
								hitw = true;
								break;
							}
					}
				}
				const unsigned hw = __ballot_sync(FULL, hitw);
				if (hw && __ffs(hw) - 1 < first_conf)
					first_conf = __ffs(hw) - 1;
			}
			if (lane == first_conf)
				cmask = 1;
			__syncwarp();
		}
		dbg[13] += clock64() - cv0;
		const long long cc0 = clock64();

		// ---- commit the longest prefix of lanes that neither need the serial step nor read a slot an earlier lane writes
		int base = 0;
		bool serial_next = false, tail = false;
		const int64_t tag_mask0 = r.tag_mask;
		for (;;) {
			const unsigned stopm = __ballot_sync(FULL, lane >= base && lane < nb && (stopper || cmask != 0));
			const unsigned softm = __ballot_sync(FULL, lane >= base && lane < nb && soft && !stopper && cmask == 0);
			int k = stopm ? (__ffs(stopm) - 1) : nb;
			const int sl = softm ? (__ffs(softm) - 1) : 32;
			if (sl < k) // the first soft stopper commits with the lanes before it, then gets its match tail
				k = sl + 1;
			bool gate_cut = false;
			if (r.tag_mask != better) { // the first sweep deletion of a phase tightens the insert gate:
				const unsigned lo_m = (base >= 32) ? 0u : (FULL << base); // later lanes used the old gate
				const unsigned km0 = (k >= 32) ? FULL : ((1u << k) - 1);
				const unsigned first_cl = clm & lo_m & km0;
				if (first_cl) {
					const int f = __ffs(first_cl) - 1;
					if (f + 1 < k) {
						k = f + 1;
						gate_cut = true;
					}
				}
			}
			const bool mine = lane >= base && lane < k;
			// Two lanes of a batch never write the same slot -- except a twin, which may replace the entry its
			// predecessor has just inserted (when that entry is due for cleaning, or is the victim of the eviction
			// it completes): the twin's store has to land second.  A twin's predecessor is never a twin itself.
			if (mine && !L.twin) {
				for (int w = 0; w < L.nw; w++)
					*reinterpret_cast<longlong2 *>(prim.tab + L.wslot[w]) = make_longlong2(L.woff[w], L.wtag[w]);
				if (cl)
					*reinterpret_cast<longlong2 *>(prim.tab + del) = make_longlong2(0, 0);
			}
			__syncwarp();
			if (mine && L.twin) {
				for (int w = 0; w < L.nw; w++)
					*reinterpret_cast<longlong2 *>(prim.tab + L.wslot[w]) = make_longlong2(L.woff[w], L.wtag[w]);
				if (cl)
					*reinterpret_cast<longlong2 *>(prim.tab + del) = make_longlong2(0, 0);
			}
			if (mine) { // everything a committed lane wrote goes into the log (nwt: inserts + its sweep deletion)
				const int at = atomicAdd(&sh->wlog_n, nwt);
				if (at + nwt <= K2_WLOG)
					for (int w = 0; w < nwt; w++)
						sh->wlog[at + w] = L.wslot[w];
			}
			__syncwarp();
			if (k > base) {
				const unsigned km = ((k >= 32) ? FULL : ((1u << k) - 1)) & ((base >= 32) ? 0u : (FULL << base));
				const int s_ins = warp_sum(mine ? L.ins : 0), s_miss = warp_sum(mine ? L.miss : 0);
				const int s_disp = warp_sum(mine && L.nw > 1 ? L.nw - 1 : 0);
				const unsigned ck = clm & km;
				n.lookups += k - base;
				n.inserts += s_ins;
				n.misses += s_miss;
				n_disp += s_disp;
				r.hash_count += __popc(netm & km) - __popc(ck);
				const int nev = __popc(evm & km);
				if (nev) {
					r.victim_round = (r.victim_round + nev) % c.max_chain;
					n_evict += nev;
				}
				if (ck) {
					r.clean_ptr = sh->dslot[__popc(clm & ((k >= 32) ? FULL : ((1u << k) - 1))) - 1];
					r.tag_mask = better;
				}
				r.p = sh->qpos[k - 1];
			}
			dbg[1] += k - base;
			base = k;
			if (gate_cut)
				dbg[6]++;
			if (k == sl + 1) { // (a gate cut in front of it would have moved k)
				tail = true;
				break;
			}
			if (k >= nb || gate_cut || r.tag_mask != tag_mask0)
				break;
			if (__shfl_sync(FULL, (int)stopper, k)) {
				serial_next = true;
				dbg[__shfl_sync(FULL, (int)L.cx, k) ? 3 : 4]++;
				break;
			}
			// lane k only conflicts with committed lanes (it read a slot one of them wrote): it and everything
			// after it go into the next full-width batch, evaluated on the updated table
			dbg[5]++;
			ev_n = 0;
			break;
		}
		int64_t tail_p = 0, tail_t = 0;
		if (tail) {
			tail_p = sh->qpos[base - 1];
			tail_t = sh->qtag[base - 1];
		}
		if (base > 0) {
			prim.publish(r.p, r.min_mask);
			pop_front(base);
		}
		dbg[14] += clock64() - cc0;
		if (tail) { // match tail of the lane committed last: find_best_match's compares, then src/rzip.c:673-688
			const long long c0 = clock64();
			const int tl = base - 1;
			unsigned sv = __shfl_sync(FULL, L.surv, tl);
			int64_t mlen = 0, offset = 0, reverse = 0;
			while (sv) {
				const int b = __ffs(sv) - 1;
				sv &= sv - 1;
				const int64_t off = sh->eoff[eb + tl][b];
				int64_t rev;
				const int64_t len = prim.match_len(tail_p, off, c.end, r.last_match, rev);
				if (len) {
					if (len > mlen) {
						mlen = len;
						offset = off - rev;
						reverse = rev;
					}
					n.hits++;
				} else
					n.misses++;
			}
			again = k2_step_tail(prim, st, r, c, recs, tail_p, mlen, offset, reverse, status);
			again_p = tail_p;
			again_t = tail_t;
			dbg[9] += clock64() - c0;
			dbg[15]++;
			filter_queue();
		}
		if (serial_next) { // serial step for the candidate that needs it
			const int64_t p = sh->qpos[0], t = sh->qtag[0];
			pop_front(1);
			const long long c0 = clock64();
			again = k2_step(prim, st, r, c, n, recs, p, t, status);
			dbg[9] += clock64() - c0;
			again_p = p;
			again_t = t;
			filter_queue();
		}
	}
	if (status == kStatusRunning && last_segment)
		k2_close_chunk(prim, st, r, c, recs, status);
	if (lane == 0 && status != -9) {
		dbg[10] = clock64() - clk_start;
		for (int i = 0; i < 16; i++)
			st->dbg[i] += dbg[i];
		st->st_displacements += n_disp;
		st->st_evictions += n_evict;
		k2_store_regs(st, r, n, status);
	}
}

__global__ void __launch_bounds__(K2_THREADS, 1)
k2_commit_kernel(const uint8_t *__restrict__ buf, ScanState *st, HEntry *tab, const Cand *cand,
		 const uint32_t *tile_count, int64_t first_tile, int64_t num_tiles, int64_t seg_hi, MatchRec *recs,
		 int last_segment, int64_t tab_stride, int64_t rec_stride)
{
	__shared__ FastShared sh;
	// one CTA per variant of the window (all-values speculation of victim_round); a plain run has one
	st += blockIdx.x;
	tab += (int64_t)blockIdx.x * tab_stride;
	recs += (int64_t)blockIdx.x * rec_stride;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < 4096; i += K2_THREADS)
		sh.wmask[i] = 0;
	if (threadIdx.x == 0) {
		sh.prog.pos = st->scan_pos;
		sh.prog.min_mask = st->min_mask;
		sh.prog.done = (st->status != kStatusRunning) ? 1 : 0;
		sh.cmd_exit = 0;
		sh.cmd_nb = 0;
		sh.cmd_flags = st->flags;
	}
	__syncthreads();
	const int64_t hmask = ((int64_t)1 << st->hash_bits) - 1;
	if (warp != 0) {
		k2_eval_worker(buf, tab, (unsigned)hmask, &sh, warp, lane);
		return;
	}
	WarpPrim prim;
	prim.prog = &sh.prog;
	prim.buf = buf;
	prim.tab = tab;
	prim.hmask = hmask;
	prim.cand = cand;
	prim.tile_count = tile_count;
	prim.first_tile = first_tile;
	prim.num_tiles = num_tiles;
	prim.seg_hi = seg_hi;
	prim.lane = lane;
	prim.tile = 0;
	prim.idx = 0;
	prim.cnt = 0;
	prim.cnt_valid = false;
	prim.bpos = prim.btag = 0;
	prim.bmask = 0;
	k2_commit_segment_batched(prim, &sh, st, recs, last_segment != 0);
	__syncwarp();
	if (lane == 0) {
		sh.prog.done = 1;
		sh.cmd_exit = 1;
	}
	__syncwarp();
	k2_bar(K2_BAR_GO); // releases the workers
}

#if !defined(LRZ_SIMT_HOST)
static constexpr int kK2ReserveSmem = 64 * 1024;

int k2_launch(const uint8_t *d_buf, ScanState *d_state, HEntry *d_tab, const Cand *d_cand,
	      const uint32_t *d_tile_count, int64_t pos_lo, int64_t pos_hi, MatchRec *d_recs, bool last_segment,
	      int nvar, int64_t tab_stride, int64_t rec_stride, cudaStream_t stream)
{
	int64_t first_tile = 0, num_tiles = 0;
	if (pos_hi > pos_lo) {
		first_tile = pos_lo / kTile;
		num_tiles = (pos_hi - 1) / kTile - first_tile + 1;
	}
	// The commit CTA lives on L1: probe chains, candidate windows and the list are prefetched into it ahead of
	// their use.  A block encoder (K7b: ~200 KB of shared memory) landing on the same SM would shrink that L1 to a
	// fifth, so the commit CTA claims enough shared memory that no encoder CTA fits beside it.
	k2_commit_kernel<<<nvar, K2_THREADS, kK2ReserveSmem, stream>>>(d_buf, d_state, d_tab, d_cand, d_tile_count, first_tile, num_tiles,
							  pos_hi, d_recs, last_segment ? 1 : 0, tab_stride, rec_stride);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Load this file's kernels now (CUDA loads a kernel's code at its first launch, and that load waits for every kernel
// that is running -- block encoders run for tens of seconds).
int k2_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, k2_commit_kernel) == cudaSuccess;
	ok = ok && cudaFuncSetAttribute(k2_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2ReserveSmem) == cudaSuccess;
	return ok ? 0 : -1;
}
#endif // !LRZ_SIMT_HOST

} // namespace lrz
