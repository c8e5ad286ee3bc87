// k2_commit.cu -- warp-cooperative primitives for the serial commit stage (see k2_commit.cuh) and
// the kernel that runs them: ONE warp owns a chunk's hash table (64 MiB at rzip level 7, resident in
// the 126 MB L2) and replays the candidates K1 produced, in position order.
//
// Data-parallel pieces, each one L2/L1 round trip wide instead of one per slot / per byte:
//   * probe windows: 32 consecutive 16-byte slots per load (512 B, 4 lines); empty / equal-tag /
//     due-for-cleaning / lesser-bitness classification by __ballot_sync, first hit by __ffs
//     (find_best_match src/rzip.c:511-531, insert_hash :313-349, clean_one_from_hash :363-378)
//   * match extension: 512 bytes per step forwards and backwards, 16 B per lane, first mismatch by
//     ballot + ffs/clz (single_match_len src/rzip.c:441-454)
//   * candidate fetch: 32 {pos,tag} records per load, mask filter by ballot, table lines of the
//     upcoming candidates prefetched into L1
#include "k2_commit.cuh"
#include "kernels.h"

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

struct WarpPrim {
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
	__device__ __forceinline__ void store_rec(MatchRec *dst, const MatchRec &r)
	{
		if (lane == 0)
			*dst = r;
	}
	__device__ __forceinline__ void store_entry(int64_t slot, int64_t t, int64_t off)
	{
		if (lane == 0)
			*reinterpret_cast<longlong2 *>(tab + slot) = make_longlong2(off, t);
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
				const longlong2 c = __ldcs(reinterpret_cast<const longlong2 *>(cand + tile * (int64_t)kTile + i));
				bpos = c.x;
				btag = c.y;
				if ((btag & min_mask) == min_mask && bpos > after) {
					const HEntry *w = tab + (btag & hmask);
					asm volatile("prefetch.global.L1 [%0];" ::"l"(w));
					asm volatile("prefetch.global.L1 [%0];" ::"l"(w + 8));
				}
			}
			bmask = __ballot_sync(FULL, v);
			idx += 32;
		}
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

__global__ void __launch_bounds__(32, 1)
k2_commit_kernel(const uint8_t *__restrict__ buf, ScanState *st, HEntry *tab, const Cand *cand,
		 const uint32_t *tile_count, int64_t first_tile, int64_t num_tiles, int64_t seg_hi, MatchRec *recs,
		 int last_segment)
{
	WarpPrim prim;
	prim.buf = buf;
	prim.tab = tab;
	prim.hmask = ((int64_t)1 << st->hash_bits) - 1;
	prim.cand = cand;
	prim.tile_count = tile_count;
	prim.first_tile = first_tile;
	prim.num_tiles = num_tiles;
	prim.seg_hi = seg_hi;
	prim.lane = threadIdx.x;
	prim.tile = 0;
	prim.idx = 0;
	prim.cnt = 0;
	prim.cnt_valid = false;
	prim.bpos = prim.btag = 0;
	prim.bmask = 0;
	k2_commit_segment(prim, st, recs, last_segment != 0);
}

int k2_launch(const uint8_t *d_buf, ScanState *d_state, HEntry *d_tab, const Cand *d_cand,
	      const uint32_t *d_tile_count, int64_t pos_lo, int64_t pos_hi, MatchRec *d_recs, bool last_segment,
	      cudaStream_t stream)
{
	int64_t first_tile = 0, num_tiles = 0;
	if (pos_hi > pos_lo) {
		first_tile = pos_lo / kTile;
		num_tiles = (pos_hi - 1) / kTile - first_tile + 1;
	}
	k2_commit_kernel<<<1, 32, 0, stream>>>(d_buf, d_state, d_tab, d_cand, d_tile_count, first_tile, num_tiles,
					       pos_hi, d_recs, last_segment ? 1 : 0);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

} // namespace lrz
