// lzma_enc.cuh -- LZMA block encoder for the lrzip-next LZMA backend, one encoder instance per
// stream block (src/stream.c:429-494 lzma_compress_buf -> src/lzma/C/LzmaLib.c:12 LzmaCompress).
//
// Bit-exactness target: the raw LZMA stream LzmaCompress() writes for (level, dictSize, lc3 lp0 pb2,
// fb, numThreads = 2), i.e. the vendored 7-Zip SDK 24.07 encoder in "optimal" mode (levels 5-9)
// over its two-thread match finder, and in "fast" mode (levels 1-4: GetOptimumFast, LzmaEnc.c:1970-2098,
// over the single-threaded hc5 hash-chain finder, lzma_mf.cuh).  What is restated here, piece by piece:
//
//   match finder    binary tree over 4-byte hashes with cut value mc        LzFind.c:962-1029
//                   driven per position like the BT thread does             LzFindMt.c:571-729
//                   2/3-byte hash candidates mixed in front of the tree's   LzFindMt.c:1093-1131,
//                   pairs only when nearer than the first tree match          1274-1317, 1340-1350
//                   (LzFindOpt.c's long-match shortcut is equivalent to the plain tree step)
//   range coder     32-bit range, 11-bit adaptive probabilities, shift 5    LzmaEnc.c:685-826
//   price tables    4-bit fixed-point prices, refreshed every 64 matches /  LzmaEnc.c:830-1065,
//                   64 rep lengths                                            2202-2319, 2645-2657
//   optimal parser  price-based DP over up to 2048 positions, 4 reps,       LzmaEnc.c:1219-1968
//                   12-state machine, LIT/REP/MATCH + "x : LIT : REP0" trials
//   block loop      first byte literal, symbol coding, flush of 5 bytes     LzmaEnc.c:2383-2680
//
// The same source is compiled for the device (product: one encoder per CUDA block, see
// backend_lzma.cu) and for the host (tests/hostsim only: lets the CPU-only container check the
// restatement against the reference's own LzmaCompress before a GPU is involved).
#pragma once
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define LZ_FN __host__ __device__
#define LZ_INL __host__ __device__ __forceinline__
#else
#define LZ_FN
#define LZ_INL inline
#endif

namespace lrz {
namespace lzma {

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative execution model of the encoder (device): ALL 32 lanes of the block's warp run the
// encoder's scalar code in lockstep on identical values ("replicated": every lane computes and stores the
// same thing, so no broadcast is needed and a lane always sees its own stores), and the loops whose
// iterations are independent -- price-table rows, literal-price bits, the cells one position updates in the
// optimal-parse table, byte compares, the candidates of one position -- are split across lanes:
//     LZ_PFOR(i, n) { body(i) }   lane l runs i = l, l + 32, ...;  results go to memory (shared);
//     lz_sync()                   makes them visible to the whole warp and re-converges it.
// On the host the warp has one lane (LZ_W == 1), so the very same source degenerates to the serial
// algorithm -- which is what tests/hostsim checks against the reference's LzmaCompress.
#if defined(__CUDA_ARCH__)
#define LZ_W 32u
LZ_INL uint32_t lz_lane() { return threadIdx.x & 31u; }
LZ_INL void lz_sync() { __syncwarp(); }
LZ_INL uint32_t lz_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
LZ_INL uint32_t lz_ffs(uint32_t m) { return (uint32_t)__ffs((int)m); }
LZ_INL void lz_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
LZ_INL uint32_t lz_ld_acquire(const uint32_t *p) // pairs with the helper warp's fence + tag store
{
	uint32_t v;
	asm volatile("ld.acquire.cta.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
LZ_INL uint32_t lz_shfl(uint32_t v, uint32_t src) { return __shfl_sync(0xffffffffu, v, (int)src); }
LZ_INL uint32_t lz_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
#else
#define LZ_W 1u
LZ_INL uint32_t lz_lane() { return 0; }
LZ_INL void lz_sync() {}
LZ_INL uint32_t lz_ballot(bool p) { return p ? 1u : 0u; }
LZ_INL uint32_t lz_ffs(uint32_t m) { return m ? (uint32_t)__builtin_ffs((int)m) : 0; }
LZ_INL void lz_prefetch(const void *) {}
#endif
#define LZ_PFOR(i, n) for (uint32_t i = lz_lane(); i < (uint32_t)(n); i += LZ_W)
// Development aid (-DLZ_PROF): cycles of the encoder warp per section, printed at the end of the block.
#if defined(LZ_PROF) && defined(__CUDA_ARCH__)
#define LZ_T(i) do { const long long t_ = clock64(); e->prof[i] += (uint64_t)(t_ - e->profT); e->profN[i]++; e->profT = t_; } while (0)
#else
#define LZ_T(i) do { } while (0)
#endif

// First index in [from, limit] at which a and b differ (limit if they agree up to there); from <= limit.
LZ_FN inline uint32_t lz_extend(const uint8_t *a, const uint8_t *b, uint32_t from, uint32_t limit)
{
	for (;;) {
		const uint32_t i = from + lz_lane();
		const uint32_t m = lz_ballot(i >= limit || a[i] != b[i]);
		if (m)
			return from + lz_ffs(m) - 1;
		from += LZ_W;
	}
}

constexpr uint32_t kNumReps = 4;
constexpr uint32_t kNumOpts = 1u << 11;
constexpr uint32_t kNumStates = 12;
constexpr uint32_t kMatchMin = 2;
constexpr uint32_t kMatchMax = 273;
constexpr uint32_t kNumPosStatesMax = 16;
constexpr uint32_t kLenLow = 8, kLenHigh = 256, kLenTotal = kLenLow * 2 + kLenHigh;
constexpr uint32_t kNumLenToPos = 4;
constexpr uint32_t kNumAlignBits = 4, kAlignSize = 16, kAlignMask = 15;
constexpr uint32_t kStartPosModel = 4, kEndPosModel = 14, kNumFullDist = 1u << (kEndPosModel >> 1); // 128
constexpr uint32_t kDistTableMax = 64;
constexpr uint32_t kBitModelTotal = 1u << 11;
constexpr uint32_t kProbInit = kBitModelTotal >> 1;
constexpr uint32_t kMoveBits = 5;
constexpr uint32_t kMoveReducingBits = 4;
constexpr uint32_t kPriceShift = 4;
constexpr uint32_t kInfinity = 1u << 30;
constexpr uint32_t kTopValue = 1u << 24;
constexpr uint32_t kMarkLit = 0xFFFFFFFFu;
constexpr int kRepLenCount = 64;
constexpr uint32_t kHash2Size = 1u << 10, kHash3Size = 1u << 16;

typedef uint16_t Prob;

// ---- helpers of the device build (one CTA per block: the encoder warp plus helper warps, backend.cu) ------------
// Staged match lists: a look-ahead warp copies the lists of the next few positions from HBM into a small ring in
// shared memory and, for every (len, dist) pair, already does the far-byte comparison of the MATCH : LIT : REP0
// trial.  An entry is a pure cache: the encoder uses it when its tag says it holds the position it wants and
// falls back to HBM otherwise, so the output never depends on the helper.
constexpr uint32_t kLkSlots = 16;                 // ring entries (positions); the helper runs at most 6 ahead of the parser,
                                                  // whose second warp is up to two positions behind the first
constexpr uint32_t kLkMaxList = 128;              // uint32 of one staged list (pairs of len, dist - 1)
constexpr uint32_t kLkByLen = 1 + kLkMaxList + kLkMaxList / 2; // offset of the distance-by-length table (indexed by length)
constexpr uint32_t kLkWords = kLkByLen + kMatchMax + 1; // count | longest << 16, list, per pair (twoBytesEqual << 31 | end), table
// Range-coder queue: the encoder warp updates the probabilities (the next prices depend on them) and queues
// (probability, bit) for a coder thread that does the range arithmetic, carries and byte output on its own.
constexpr uint32_t kRcQ = 4096;                   // queue entries (power of two)
constexpr uint32_t kRcDirect = 0x80000000u;       // entry: direct bits, nbits << 26 | value
constexpr uint32_t kRcFlush = 0xFFFFFFFFu;        // entry: end of the block

struct LenProbs {
	Prob low[kNumPosStatesMax << 4]; // per posState: 16 probs (choice bits live in low[0], low[8])
	Prob high[kLenHigh];
};

struct LenPrices {
	uint32_t tableSize;
	uint32_t prices[kNumPosStatesMax][kLenTotal];
};

struct alignas(16) Opt {
	uint32_t price;
	uint16_t state;
	uint16_t extra; // 0 normal, 1 LIT : MATCH, >1 MATCH(extra-1) : LIT : REP0(len)
	uint32_t len;
	uint32_t dist;
	uint32_t reps[kNumReps];
};

// Hand-over of one position from the parser's first warp to its second (device product path, see opt_step_a / _b).
struct StepPkt {
	uint32_t cmd; // 1: a position, 0: leave
	uint32_t cur, pos, position, naf, hdr, state, posState;
	uint32_t reps[kNumReps];
	uint32_t matchPrice, repMatchPrice, normalMatchPrice, litPrice, nextIsLit, curByte, matchByte, last;
};

// Everything one block encoder owns besides the match-finder arrays; lives in HBM (device) / heap (host).
struct Enc {
	// ---- configuration
	const uint8_t *src;
	uint32_t n;
	uint32_t fb, mc, historySize, cyclicSize, hashMask, bigHash, distTableSize;
	uint32_t pbMask, lpMask, lc, fastMode;
	// ---- match finder (positions are 1-based: byte i has pos i + 1; 0 = empty)
	uint32_t *hash2, *hash3, *hash4, *son;
	// precomputed match lists (lzma_mf.cu): when preRec is set the serial finder above is not used
	const uint64_t *preRec;
	const uint32_t *prePool;
	// the tree walk may still be running (device pipeline): records [0, preWait) carry a "final" bit to wait for, a
	// record may point past the pool when the walk ran out of it (mfOverflow is set then and the block is redone)
	uint32_t preWait;
	uint64_t prePoolCap;
	const int *mfOverflow;
	// lz4 compressibility gate running beside the encoder (device): 0 undecided, 1 compressible, 2 leave the
	// block stored -- the encoder gives up as soon as it reads 2 (null: no gate)
	const int *gateState;
	int aborted;
	// staged match lists / range-coder queue (null on the host and in tests of the plain path)
	int lkOn, rcOn;         // the look-ahead warp / the coder thread are running (device product path only)
	int lkSlot;             // entry the current position's list was taken from, or -1
	uint32_t rcTail;        // entries queued so far (the encoder warp's private count)
	uint32_t pos;       // position the match finder will hand out next
	uint32_t cycPos;
	uint32_t crc[256];
	// ---- range coder
	uint64_t low, cacheSize;
	uint32_t range, cache;
	uint8_t *out;
	uint64_t outPos, outCap;
	int overflow;
	// ---- coder state
	uint32_t state, reps[kNumReps];
	uint32_t optCur, optEnd, longestMatchLen, numPairs, numAvail, additionalOffset, backRes, matchPriceCount;
	int repLenCounter;
	uint32_t probPrices[kBitModelTotal >> kMoveReducingBits];
	uint32_t matches[kMatchMax * 2 + 2];
	uint32_t btTmp[kMatchMax * 2 + 2]; // tree pairs of the current position before the 2/3-byte hash mix
	uint32_t alignPrices[kAlignSize];
	uint32_t posSlotPrices[kNumLenToPos][kDistTableMax];
	uint32_t distPrices[kNumLenToPos][kNumFullDist];
	Prob posAlign[kAlignSize];
	Prob isRep[kNumStates], isRepG0[kNumStates], isRepG1[kNumStates], isRepG2[kNumStates];
	Prob isMatch[kNumStates][kNumPosStatesMax], isRep0Long[kNumStates][kNumPosStatesMax];
	Prob posSlot[kNumLenToPos][64];
	Prob posEnc[kNumFullDist];
	LenProbs lenProbs, repLenProbs;
	LenPrices lenPrices, repLenPrices;
	Prob lit[0x300 << 4]; // lc + lp <= 4 supported (lrzip-next uses lc3 lp0)
	Opt opt[kNumOpts];
	// scratch the lanes of the warp exchange results through (see LZ_PFOR)
	uint32_t tmpDist[kNumFullDist];
	uint32_t xLit[8];
	uint32_t xRepLen[kNumReps], xRepLen2[kNumReps];
	uint32_t xPairLen2[kMatchMax + 2], xPairPrice[kMatchMax + 2];
#if defined(LZ_PROF)
	uint64_t prof[24], profN[24];
	long long profT;
#endif
	// The helpers' structures live inside the encoder so that the encoder warp reaches them with plain shared-memory
	// loads (through a pointer kept in the struct they become generic loads, which cost three times as much).
	int splitOn;        // a second parser warp takes the rep / match half of every staged position (backend.cu)
	uint32_t lastB;     // `last` as the second warp leaves it
	alignas(16) StepPkt pkt;
	alignas(16) uint32_t pubReps[kNumReps]; // the reps of the cell the parser is at: the look-ahead warp prefetches behind them
	alignas(8) uint32_t lkHead[kLkSlots][2]; // {position (1-based, as e->pos) the entry holds, count | longest length << 16}
	uint32_t lkRing[kLkSlots * kLkWords];
	uint32_t rcQueue[kRcQ];
	uint32_t rcTailPub, rcHeadPub; // published counts (encoder -> coder, coder -> encoder)
	int rcDone;
};

// ---------------------------------------------------------------------------------------------------
// small helpers
LZ_INL uint32_t top_bit(uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return 31u - (uint32_t)__clz((int)v);
#else
	return 31u - (uint32_t)__builtin_clz(v);
#endif
}

LZ_INL uint32_t pos_slot(uint32_t d) // GetPosSlot (LzmaEnc.c:167-246): 2*log2 + next bit
{
	if (d < 2)
		return d;
	const uint32_t b = top_bit(d);
	return 2 * b + ((d >> (b - 1)) & 1);
}

// a[i] for a four-element array kept in registers (a run-time index would push the array to local memory)
LZ_INL uint32_t pick4(const uint32_t *a, uint32_t i) { return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3])); }
LZ_INL bool is_lit_state(uint32_t s) { return s < 7; }
LZ_INL uint32_t st_lit(uint32_t s) { return s < 4 ? 0 : (s < 10 ? s - 3 : s - 6); }   // kLiteralNextStates
LZ_INL uint32_t st_match(uint32_t s) { return s < 7 ? 7 : 10; }                         // kMatchNextStates
LZ_INL uint32_t st_rep(uint32_t s) { return s < 7 ? 8 : 11; }                           // kRepNextStates
LZ_INL uint32_t st_shortrep(uint32_t s) { return s < 7 ? 9 : 11; }                      // kShortRepNextStates

LZ_INL uint32_t price_bit(const Enc *e, uint32_t prob, uint32_t bit)
{
	return e->probPrices[(prob ^ ((0u - bit) & (kBitModelTotal - 1))) >> kMoveReducingBits];
}
LZ_INL uint32_t price0(const Enc *e, uint32_t prob) { return e->probPrices[prob >> kMoveReducingBits]; }
LZ_INL uint32_t price1(const Enc *e, uint32_t prob) { return e->probPrices[(prob ^ (kBitModelTotal - 1)) >> kMoveReducingBits]; }

// ---------------------------------------------------------------------------------------------------
// range coder (LzmaEnc.c:685-760)
LZ_INL void rc_put(Enc *e, uint8_t b)
{
	if (e->outPos < e->outCap)
		e->out[e->outPos] = b;
	else
		e->overflow = 1;
	e->outPos++;
}

LZ_FN inline void rc_shift_low(Enc *e)
{
	const uint32_t low = (uint32_t)e->low;
	const uint32_t high = (uint32_t)(e->low >> 32);
	e->low = (uint32_t)(low << 8);
	if (low < 0xFF000000u || high != 0) {
		rc_put(e, (uint8_t)(e->cache + high));
		e->cache = low >> 24;
		if (e->cacheSize == 0)
			return;
		const uint8_t fill = (uint8_t)(high + 0xFF);
		do
			rc_put(e, fill);
		while (--e->cacheSize);
		return;
	}
	e->cacheSize++;
}

LZ_INL void rc_norm(Enc *e)
{
	if (e->range < kTopValue) {
		e->range <<= 8;
		rc_shift_low(e);
	}
}

#if defined(__CUDA_ARCH__)
LZ_INL void rcq_push(Enc *e, uint32_t v)
{
	e->rcQueue[e->rcTail & (kRcQ - 1)] = v; // every lane stores the same word
	e->rcTail++;
}
// make the entries queued so far visible to the coder thread; called once per symbol
LZ_INL void rcq_publish(Enc *e)
{
	__threadfence_block();
	*(volatile uint32_t *)&e->rcTailPub = e->rcTail;
}
// room for one more symbol (a symbol queues fewer than 64 entries)
LZ_INL void rcq_reserve(Enc *e)
{
	while (e->rcTail + 64 - *(volatile uint32_t *)&e->rcHeadPub > kRcQ)
		__nanosleep(100);
}
#endif

// Device, queue mode: the binary decisions of ONE symbol are collected one per lane (which probability, which
// bit: both follow from the symbol alone, never from a probability's value, and no probability is used twice
// within a symbol), then rcw_commit lets every lane update its probability and queue its (probability, bit) entry
// at once.  A symbol has at most 23 decisions (match: 2 + 10 length + 6 slot + 1 direct + 4 align).
struct RcW {
	uint32_t off; // the probability, as a byte offset inside the encoder (0: direct bits)
	uint32_t v, cnt;
	bool on;
};

LZ_INL void rcw_commit(Enc *e, RcW &w)
{
#if defined(__CUDA_ARCH__)
	if (!w.on)
		return;
	const uint32_t tail = e->rcTail;
	if (lz_lane() < w.cnt) {
		uint32_t ent = w.v;
		if (w.off) {
			Prob *pp = reinterpret_cast<Prob *>(reinterpret_cast<char *>(e) + w.off);
			const uint32_t q = *pp;
			*pp = w.v ? (Prob)(q - (q >> kMoveBits)) : (Prob)(q + ((kBitModelTotal - q) >> kMoveBits));
			ent = (q << 1) | w.v;
		}
		e->rcQueue[(tail + lz_lane()) & (kRcQ - 1)] = ent;
	}
	e->rcTail = tail + w.cnt;
	w.cnt = 0;
	lz_sync();
	rcq_publish(e);
#endif
}

LZ_INL void rc_bit(Enc *e, RcW &w, Prob *prob, uint32_t bit)
{
#if defined(__CUDA_ARCH__)
	if (w.on) { // decision number w.cnt of this symbol: lane w.cnt keeps it (rcw_commit)
		if (lz_lane() == w.cnt) {
			w.off = (uint32_t)(reinterpret_cast<const char *>(prob) - reinterpret_cast<const char *>(e));
			w.v = bit;
		}
		w.cnt++;
		return;
	}
#endif
	const uint32_t p = *prob;
	const uint32_t bound = (e->range >> 11) * p;
	if (bit == 0) {
		e->range = bound;
		*prob = (Prob)(p + ((kBitModelTotal - p) >> kMoveBits));
	} else {
		e->low += bound;
		e->range -= bound;
		*prob = (Prob)(p - (p >> kMoveBits));
	}
	rc_norm(e);
}

// The coder thread's side of rc_bit: the probability is the value the encoder saw, its update is already done.
LZ_INL void rc_bit_value(Enc *e, uint32_t p, uint32_t bit)
{
	const uint32_t bound = (e->range >> 11) * p;
	if (bit == 0)
		e->range = bound;
	else {
		e->low += bound;
		e->range -= bound;
	}
	rc_norm(e);
}

LZ_INL void rc_direct(Enc *e, RcW &w, uint32_t value, uint32_t nbits) // most significant bit first
{
#if defined(__CUDA_ARCH__)
	if (w.on) {
		if (nbits) {
			if (lz_lane() == w.cnt) {
				w.off = 0;
				w.v = kRcDirect | (nbits << 26) | value; // nbits <= 26, value < 2^26
			}
			w.cnt++;
		}
		return;
	}
#endif
	while (nbits--) {
		e->range >>= 1;
		if ((value >> nbits) & 1)
			e->low += e->range;
		rc_norm(e);
	}
}

LZ_INL void lit_encode(Enc *e, RcW &w, Prob *probs, uint32_t sym)
{
	sym |= 0x100;
	do {
		rc_bit(e, w, probs + (sym >> 8), (sym >> 7) & 1);
		sym <<= 1;
	} while (sym < 0x10000);
}

LZ_INL void lit_encode_matched(Enc *e, RcW &w, Prob *probs, uint32_t sym, uint32_t matchByte)
{
	uint32_t offs = 0x100;
	sym |= 0x100;
	do {
		matchByte <<= 1;
		Prob *prob = probs + (offs + (matchByte & offs) + (sym >> 8));
		const uint32_t bit = (sym >> 7) & 1;
		sym <<= 1;
		offs &= ~(matchByte ^ sym);
		rc_bit(e, w, prob, bit);
	} while (sym < 0x10000);
}

LZ_INL void rc_reverse(Enc *e, RcW &w, Prob *probs, uint32_t nbits, uint32_t sym)
{
	uint32_t m = 1;
	do {
		const uint32_t bit = sym & 1;
		sym >>= 1;
		rc_bit(e, w, probs + m, bit);
		m = (m << 1) | bit;
	} while (--nbits);
}

// LenEnc_Encode (LzmaEnc.c:928-960)
LZ_INL void len_encode(Enc *e, RcW &w, LenProbs *lp, uint32_t sym, uint32_t posState)
{
	Prob *probs = lp->low;
	if (sym >= kLenLow) {
		rc_bit(e, w, probs, 1);
		probs += kLenLow;
		if (sym >= kLenLow * 2) {
			rc_bit(e, w, probs, 1);
			lit_encode(e, w, lp->high, sym - kLenLow * 2);
			return;
		}
		sym -= kLenLow;
	}
	rc_bit(e, w, probs, 0);
	probs += posState << 4;
	uint32_t bit = sym >> 2;
	rc_bit(e, w, probs + 1, bit);
	uint32_t m = 2 + bit;
	bit = (sym >> 1) & 1;
	rc_bit(e, w, probs + m, bit);
	m = (m << 1) + bit;
	rc_bit(e, w, probs + m, sym & 1);
}

// ---------------------------------------------------------------------------------------------------
// prices
LZ_FN inline void init_prob_prices(uint32_t *pp) // LzmaEnc_InitPriceTables (LzmaEnc.c:830-852)
{
	for (uint32_t i = 0; i < (kBitModelTotal >> kMoveReducingBits); i++) {
		uint32_t w = (i << kMoveReducingBits) + (1u << (kMoveReducingBits - 1));
		uint32_t bits = 0;
		for (uint32_t j = 0; j < kPriceShift; j++) {
			w = w * w;
			bits <<= 1;
			while (w >= (1u << 16)) {
				w >>= 1;
				bits++;
			}
		}
		pp[i] = (11u << kPriceShift) - 15 - bits;
	}
}

// Serial forms (used inside LZ_PFOR bodies, where a lane prices a literal on its own).
LZ_FN inline uint32_t lit_price_1(const Enc *e, const Prob *probs, uint32_t sym)
{
	uint32_t price = 0;
	sym |= 0x100;
	do {
		const uint32_t bit = sym & 1;
		sym >>= 1;
		price += price_bit(e, probs[sym], bit);
	} while (sym >= 2);
	return price;
}

LZ_FN inline uint32_t lit_price_matched_1(const Enc *e, const Prob *probs, uint32_t sym, uint32_t matchByte)
{
	uint32_t price = 0, offs = 0x100;
	sym |= 0x100;
	do {
		matchByte <<= 1;
		price += price_bit(e, probs[offs + (matchByte & offs) + (sym >> 8)], (sym >> 7) & 1);
		sym <<= 1;
		offs &= ~(matchByte ^ sym);
	} while (sym < 0x10000);
	return price;
}

// Warp forms: the 8 binary decisions of a literal are priced by 8 lanes at once.  Decision k (most
// significant bit first) uses the node reached by the k bits above it, which is known from the symbol.
LZ_FN inline uint32_t lit_price(Enc *e, const Prob *probs, uint32_t sym)
{
	lz_sync();
	LZ_PFOR(k, 8) {
		const uint32_t node = (0x100u | sym) >> (8 - k);
		e->xLit[k] = price_bit(e, probs[node], (sym >> (7 - k)) & 1);
	}
	lz_sync();
	uint32_t price = 0;
	for (uint32_t k = 0; k < 8; k++)
		price += e->xLit[k];
	return price;
}

LZ_FN inline uint32_t lit_price_matched(Enc *e, const Prob *probs, uint32_t sym, uint32_t matchByte)
{
	lz_sync();
	LZ_PFOR(k, 8) {
		// offs stays 0x100 while the bits above agree with the match byte's (LitEnc_MatchedEncode, LzmaEnc.c:800-826)
		const uint32_t offs = (((matchByte ^ sym) >> (8 - k)) == 0) ? 0x100u : 0u;
		const uint32_t node = (0x100u | sym) >> (8 - k);
		const uint32_t mb = ((matchByte >> (7 - k)) & 1) << 8;
		e->xLit[k] = price_bit(e, probs[offs + (mb & offs) + node], (sym >> (7 - k)) & 1);
	}
	lz_sync();
	uint32_t price = 0;
	for (uint32_t k = 0; k < 8; k++)
		price += e->xLit[k];
	return price;
}

LZ_INL const Prob *lit_probs(const Enc *e, uint32_t pos, uint32_t prevByte)
{
	return e->lit + 3u * ((((pos << 8) + prevByte) & e->lpMask) << e->lc);
}

LZ_FN inline void set_prices_3(const Enc *e, const Prob *probs, uint32_t start, uint32_t *prices)
{
	for (uint32_t i = 0; i < 8; i += 2) {
		uint32_t price = start;
		price += price_bit(e, probs[1], i >> 2);
		price += price_bit(e, probs[2 + (i >> 2)], (i >> 1) & 1);
		const uint32_t prob = probs[4 + (i >> 1)];
		prices[i] = price + price0(e, prob);
		prices[i + 1] = price + price1(e, prob);
	}
}

// LenPriceEnc_UpdateTables (LzmaEnc.c:979-1065)
LZ_FN inline void len_update_prices(const Enc *e, LenPrices *lp, uint32_t numPosStates, const LenProbs *enc)
{
	uint32_t b;
	lz_sync();
	{
		const uint32_t prob = enc->low[0];
		b = price1(e, prob);
		const uint32_t a = price0(e, prob);
		const uint32_t c = b + price0(e, enc->low[kLenLow]);
		LZ_PFOR(q, numPosStates * 2) { // one lane per (posState, low | mid) group of 8 prices
			const uint32_t ps = q >> 1;
			uint32_t *prices = lp->prices[ps];
			const Prob *probs = enc->low + (ps << 4);
			if (q & 1)
				set_prices_3(e, probs + kLenLow, c, prices + kLenLow);
			else
				set_prices_3(e, probs, a, prices);
		}
	}
	const uint32_t ts = lp->tableSize;
	if (ts > kLenLow * 2) {
		const Prob *probs = enc->high;
		uint32_t *prices = lp->prices[0] + kLenLow * 2;
		const uint32_t cnt = (ts - (kLenLow * 2 - 1)) >> 1;
		b += price1(e, enc->low[kLenLow]);
		LZ_PFOR(i, cnt) {
			uint32_t sym = i + (1u << 7);
			uint32_t price = b;
			do {
				const uint32_t bit = sym & 1;
				sym >>= 1;
				price += price_bit(e, probs[sym], bit);
			} while (sym >= 2);
			const uint32_t prob = probs[i + (1u << 7)];
			prices[i * 2] = price + price0(e, prob);
			prices[i * 2 + 1] = price + price1(e, prob);
		}
		lz_sync();
		const uint32_t num = ts - kLenLow * 2;
		LZ_PFOR(q, (numPosStates - 1) * num) {
			const uint32_t ps = 1 + q / num, k = q % num;
			lp->prices[ps][kLenLow * 2 + k] = lp->prices[0][kLenLow * 2 + k];
		}
	}
	lz_sync();
}

LZ_FN inline void fill_align_prices(Enc *e) // LzmaEnc.c:2202-2222
{
	const Prob *probs = e->posAlign;
	lz_sync();
	LZ_PFOR(i, kAlignSize / 2) {
		uint32_t price = 0, sym = i, m = 1, bit;
		for (int k = 0; k < 3; k++) {
			bit = sym & 1;
			sym >>= 1;
			price += price_bit(e, probs[m], bit);
			m = (m << 1) + bit;
		}
		const uint32_t prob = probs[m];
		e->alignPrices[i] = price + price0(e, prob);
		e->alignPrices[i + 8] = price + price1(e, prob);
	}
	lz_sync();
}

LZ_FN inline void fill_distance_prices(Enc *e) // LzmaEnc.c:2225-2319
{
	uint32_t *temp = e->tmpDist;
	e->matchPriceCount = 0;
	lz_sync();
	LZ_PFOR(q, kNumFullDist / 2 - kStartPosModel / 2) {
		const uint32_t i = q + kStartPosModel / 2;
		const uint32_t slot = pos_slot(i);
		uint32_t footer = (slot >> 1) - 1;
		uint32_t base = (2 | (slot & 1)) << footer;
		const Prob *probs = e->posEnc + (size_t)base * 2;
		uint32_t price = 0, m = 1, sym = i;
		const uint32_t offset = 1u << footer;
		base += i;
		while (footer) {
			const uint32_t bit = sym & 1;
			sym >>= 1;
			price += price_bit(e, probs[m], bit);
			m = (m << 1) + bit;
			footer--;
		}
		const uint32_t prob = probs[m];
		temp[base] = price + price0(e, prob);
		temp[base + offset] = price + price1(e, prob);
	}
	const uint32_t half = (e->distTableSize + 1) >> 1;
	LZ_PFOR(q, kNumLenToPos * half) { // slot prices; slots >= kEndPosModel also carry their direct bits
		const uint32_t lps = q / half, slot = q % half;
		uint32_t *sp = e->posSlotPrices[lps];
		const Prob *probs = e->posSlot[lps];
		uint32_t sym = slot + (1u << 5), price = 0;
		for (int k = 0; k < 5; k++) {
			const uint32_t bit = sym & 1;
			sym >>= 1;
			price += price_bit(e, probs[sym], bit);
		}
		const uint32_t prob = probs[slot + (1u << 5)];
		uint32_t delta = 0;
		if (slot >= kEndPosModel / 2)
			delta = (((kEndPosModel / 2 - 1) - kNumAlignBits) + (slot - kEndPosModel / 2)) << kPriceShift;
		sp[slot * 2] = price + price0(e, prob) + delta;
		sp[slot * 2 + 1] = price + price1(e, prob) + delta;
	}
	lz_sync();
	LZ_PFOR(q, kNumLenToPos * kNumFullDist) {
		const uint32_t lps = q / kNumFullDist, i = q % kNumFullDist;
		const uint32_t *sp = e->posSlotPrices[lps];
		e->distPrices[lps][i] = i < 4 ? sp[i] : sp[pos_slot(i)] + temp[i];
	}
	lz_sync();
}

// ---------------------------------------------------------------------------------------------------
// match finder
LZ_INL const uint8_t *mf_cur(const Enc *e) { return e->src + (e->pos - 1); }
LZ_INL uint32_t mf_avail(const Enc *e) { return e->n - (e->pos - 1); }

// One step of the binary tree for the position e->pos (GetMatchesSpec1 with maxLen = 3): returns the
// number of uint32 written to d (pairs len, dist-1; len >= 4, strictly increasing).
LZ_FN inline uint32_t mf_tree_step(Enc *e, uint32_t *d)
{
	const uint32_t avail = mf_avail(e);
	if (avail < 4)
		return 0; // the last three positions are not searched and not inserted (LzFindMt.c:627-642)
	const uint8_t *cur = mf_cur(e);
	const uint32_t pos = e->pos;
	uint32_t hv;
	if (e->bigHash) // GetHeads4b (LzFindMt.c:386-394)
		hv = (e->crc[cur[0]] & e->hashMask) ^ ((uint32_t)cur[1] | ((uint32_t)cur[2] << 8) | ((uint32_t)cur[3] << 16));
	else            // GetHeads4 (LzFindMt.c:368-384) == HASH4_CALC (LzFind.c:49)
		hv = (e->crc[cur[0]] ^ cur[1] ^ ((uint32_t)cur[2] << 8) ^ (e->crc[cur[3]] << 5)) & e->hashMask;
	uint32_t curMatch = e->hash4[hv];
	e->hash4[hv] = pos;
	const uint32_t lenLimit = avail < e->fb ? avail : e->fb;
	const uint32_t cyc = e->cycPos, cbs = e->cyclicSize;
	uint32_t *son = e->son;
	uint32_t *ptr0 = son + ((size_t)cyc << 1) + 1, *ptr1 = son + ((size_t)cyc << 1);
	uint32_t len0 = 0, len1 = 0, maxLen = 3, cut = e->mc, nd = 0;
	const uint32_t cmCheck = pos <= cbs ? 0 : pos - cbs;
	if (cmCheck < curMatch) {
		do {
			const uint32_t delta = pos - curMatch;
			uint32_t *pair = son + ((size_t)(cyc - delta + (delta > cyc ? cbs : 0)) << 1);
			const uint8_t *pb = cur - delta;
			uint32_t len = len0 < len1 ? len0 : len1;
			const uint32_t pair0 = pair[0];
			if (pb[len] == cur[len]) {
				while (++len != lenLimit)
					if (pb[len] != cur[len])
						break;
				if (maxLen < len) {
					maxLen = len;
					d[nd++] = len;
					d[nd++] = delta - 1;
					if (len == lenLimit) {
						*ptr1 = pair0;
						*ptr0 = pair[1];
						return nd;
					}
				}
			}
			if (pb[len] < cur[len]) {
				*ptr1 = curMatch;
				curMatch = pair[1];
				ptr1 = pair + 1;
				len1 = len;
			} else {
				*ptr0 = curMatch;
				curMatch = pair0;
				ptr0 = pair;
				len0 = len;
			}
		} while (--cut && cmCheck < curMatch);
	}
	*ptr0 = *ptr1 = 0;
	return nd;
}

LZ_INL void mf_advance(Enc *e)
{
	e->pos++;
	if (++e->cycPos == e->cyclicSize)
		e->cycPos = 0;
}

LZ_INL void mf_hash23(const Enc *e, const uint8_t *cur, uint32_t &h2, uint32_t &h3)
{
	const uint32_t t = e->crc[cur[0]] ^ cur[1];
	h2 = t & (kHash2Size - 1);
	h3 = (t ^ ((uint32_t)cur[2] << 8)) & (kHash3Size - 1);
}

// MatchFinderMt_GetMatches + MixMatches3 (LzFindMt.c:1274-1317, 1093-1131)
LZ_FN inline uint32_t mf_get_matches(Enc *e, uint32_t *d)
{
	if (e->preRec) { // the data-parallel pre-pass already produced this position's list
		const uint32_t i0 = e->pos - 1;
#if defined(__CUDA_ARCH__)
		if (e->lkOn) { // staged in shared memory by the look-ahead warp?  (the same word for every lane)
			const uint32_t slot = e->pos & (kLkSlots - 1);
			if (*(const volatile uint32_t *)&e->lkHead[slot][0] == e->pos) {
				__threadfence_block();
				const uint32_t *b = e->lkRing + slot * kLkWords;
				const uint32_t nd = *(const volatile uint32_t *)&e->lkHead[slot][1] & 0xFFFFu; // the longest length sits above
				lz_sync();
				LZ_PFOR(i, nd)
					d[i] = b[1 + i];
				lz_sync();
				e->lkSlot = (int)slot;
				e->pos++;
				return nd;
			}
			e->lkSlot = -1;
		}
#endif
		uint64_t rec;
		bool live = false; // the walk is (or may be) still writing: no cached loads
#if defined(__CUDA_ARCH__)
		if (e->preWait) {
			live = true;
			rec = 0;
			if (i0 < e->preWait)
				while (!((rec = *(const volatile uint64_t *)(e->preRec + i0)) >> 63))
					__nanosleep(200);
		} else
#endif
			rec = e->preRec[i0];
		rec &= ~(1ull << 63);
		uint32_t nd = (uint32_t)rec & 1023u;
		if (live && (rec >> 10) + nd > e->prePoolCap)
			nd = 0; // pool overflow: the block is abandoned at the next symbol
		const uint32_t *s = e->prePool + (rec >> 10);
		lz_sync();
		LZ_PFOR(i, nd) {
#if defined(__CUDA_ARCH__)
			d[i] = live ? __ldcg(s + i) : s[i];
#else
			d[i] = s[i];
#endif
		}
		lz_sync();
		e->pos++;
		return nd;
	}
	uint32_t *bt = e->btTmp;
	const uint32_t nbt = mf_tree_step(e, bt);
	const uint32_t availAfter = mf_avail(e) - 1;
	const uint8_t *cur = mf_cur(e);
	const uint32_t m = e->pos;
	uint32_t nd = 0;
	bool mix = false;
	uint32_t minPos = 0;
	if (nbt == 0) {
		if (availAfter >= 3) {
			mix = true;
			minPos = m > e->historySize ? m - e->historySize : 1;
		}
	} else {
		mix = true;
		minPos = m - bt[1];
	}
	if (mix) {
		uint32_t h2, h3;
		mf_hash23(e, cur, h2, h3);
		const uint32_t c2 = e->hash2[h2], c3 = e->hash3[h3];
		e->hash2[h2] = m;
		e->hash3[h3] = m;
		bool done = false;
		if (c2 >= minPos && cur[(ptrdiff_t)c2 - (ptrdiff_t)m] == cur[0]) {
			d[nd + 1] = m - c2 - 1;
			if (cur[(ptrdiff_t)c2 - (ptrdiff_t)m + 2] == cur[2]) {
				d[nd] = 3;
				nd += 2;
				done = true;
			} else {
				d[nd] = 2;
				nd += 2;
			}
		}
		if (!done && c3 >= minPos && cur[(ptrdiff_t)c3 - (ptrdiff_t)m] == cur[0]) {
			d[nd++] = 3;
			d[nd++] = m - c3 - 1;
		}
	}
	for (uint32_t i = 0; i < nbt; i++)
		d[nd++] = bt[i];
	mf_advance(e);
	return nd;
}

// MatchFinderMt3_Skip (LzFindMt.c:1340-1350): the tree still sees every position
LZ_FN inline void mf_skip(Enc *e, uint32_t num)
{
	if (e->preRec) {
		e->pos += num;
		return;
	}
	while (num--) {
		mf_tree_step(e, e->btTmp);
		if (mf_avail(e) >= 3) {
			uint32_t h2, h3;
			mf_hash23(e, mf_cur(e), h2, h3);
			e->hash2[h2] = e->pos;
			e->hash3[h3] = e->pos;
		}
		mf_advance(e);
	}
}

// ReadMatchDistances (LzmaEnc.c:1083-1126)
LZ_FN inline uint32_t read_matches(Enc *e, uint32_t *numPairsRes)
{
	e->additionalOffset++;
	e->numAvail = mf_avail(e);
	const uint32_t numPairs = mf_get_matches(e, e->matches);
	*numPairsRes = numPairs;
	if (numPairs == 0)
		return 0;
	const uint32_t len = e->matches[numPairs - 2];
	if (len != e->fb)
		return len;
	uint32_t numAvail = e->numAvail;
	if (numAvail > kMatchMax)
		numAvail = kMatchMax;
	const uint8_t *p1 = mf_cur(e) - 1;
	const ptrdiff_t dif = (ptrdiff_t)-1 - (ptrdiff_t)e->matches[numPairs - 1];
	return lz_extend(p1 + dif, p1, len, numAvail);
}

LZ_INL void move_pos(Enc *e, uint32_t num)
{
	e->additionalOffset += num;
	mf_skip(e, num);
}

// ---------------------------------------------------------------------------------------------------
// optimal parser
LZ_INL uint32_t price_short_rep(const Enc *e, uint32_t state, uint32_t posState)
{
	return price0(e, e->isRepG0[state]) + price0(e, e->isRep0Long[state][posState]);
}

LZ_INL uint32_t price_rep0(const Enc *e, uint32_t state, uint32_t posState) // GetPrice_Rep_0
{
	return price1(e, e->isMatch[state][posState]) + price1(e, e->isRep0Long[state][posState]) + price1(e, e->isRep[state]) +
	       price0(e, e->isRepG0[state]);
}

LZ_FN inline uint32_t price_pure_rep(const Enc *e, uint32_t repIndex, uint32_t state, uint32_t posState)
{
	uint32_t price, prob = e->isRepG0[state];
	if (repIndex == 0) {
		price = price0(e, prob);
		price += price1(e, e->isRep0Long[state][posState]);
	} else {
		price = price1(e, prob);
		prob = e->isRepG1[state];
		if (repIndex == 1)
			price += price0(e, prob);
		else {
			price += price1(e, prob);
			price += price_bit(e, e->isRepG2[state], repIndex - 2);
		}
	}
	return price;
}

LZ_INL uint32_t len_price(const LenPrices *lp, uint32_t posState, uint32_t len) { return lp->prices[posState][len - kMatchMin]; }

// Backward (LzmaEnc.c:1167-1211)
LZ_FN inline uint32_t backward(Enc *e, uint32_t cur)
{
	uint32_t wr = cur + 1;
	e->optEnd = wr;
	for (;;) {
		uint32_t dist = e->opt[cur].dist;
		uint32_t len = e->opt[cur].len;
		const uint32_t extra = e->opt[cur].extra;
		cur -= len;
		if (extra) {
			wr--;
			e->opt[wr].len = len;
			cur -= extra;
			len = extra;
			if (extra == 1) {
				e->opt[wr].dist = dist;
				dist = kMarkLit;
			} else {
				e->opt[wr].dist = 0;
				len--;
				wr--;
				e->opt[wr].dist = kMarkLit;
				e->opt[wr].len = 1;
			}
		}
		if (cur == 0) {
			e->backRes = dist;
			e->optCur = wr;
			return len;
		}
		wr--;
		e->opt[wr].dist = dist;
		e->opt[wr].len = len;
	}
}

LZ_INL void opt_set(Opt *o, uint32_t price, uint32_t len, uint32_t dist, uint32_t extra)
{
	o->price = price;
	o->len = len;
	o->dist = dist;
	o->extra = (uint16_t)extra;
}

// price of a normal match (len, dist) given the price of reaching it (LzmaEnc.c:1402-1416, 1856-1872)
LZ_INL uint32_t match_price(const Enc *e, uint32_t normalMatchPrice, uint32_t posState, uint32_t len, uint32_t dist)
{
	uint32_t price = normalMatchPrice + len_price(&e->lenPrices, posState, len);
	uint32_t lenNorm = len - 2;
	lenNorm = lenNorm < kNumLenToPos - 1 ? lenNorm : kNumLenToPos - 1;
	if (dist < kNumFullDist)
		price += e->distPrices[lenNorm][dist & (kNumFullDist - 1)];
	else
		price += e->posSlotPrices[lenNorm][pos_slot(dist)] + e->alignPrices[dist & kAlignMask];
	return price;
}

// The cells [base + lo, base + hi] of the parse table get the rep-match price for their length.
LZ_FN inline void opt_rep_cells(Enc *e, uint32_t base, uint32_t lo, uint32_t hi, uint32_t price, uint32_t posState,
				uint32_t repIndex)
{
	lz_sync();
	LZ_PFOR(q, hi + 1 - lo) {
		const uint32_t len2 = lo + q;
		const uint32_t price2 = price + len_price(&e->repLenPrices, posState, len2);
		Opt *o = &e->opt[base + len2];
		if (price2 < o->price)
			opt_set(o, price2, len2, repIndex, 0);
	}
	lz_sync();
}

// The cells [base + startLen, base + newLen] get the price of the shortest-distance match that reaches them.
LZ_FN inline void opt_match_cells(Enc *e, uint32_t base, uint32_t startLen, uint32_t newLen, uint32_t normalMatchPrice,
				  uint32_t posState)
{
	const uint32_t *matches = e->matches;
	lz_sync();
	uint32_t offs = 0;
	LZ_PFOR(q, newLen + 1 - startLen) {
		const uint32_t len = startLen + q;
		while (len > matches[offs])
			offs += 2;
		const uint32_t dist = matches[offs + 1];
		const uint32_t price = match_price(e, normalMatchPrice, posState, len, dist);
		Opt *o = &e->opt[base + len];
		if (price < o->price)
			opt_set(o, price, len, dist + kNumReps, 0);
	}
	lz_sync();
}


#if defined(__CUDA_ARCH__)
LZ_INL void opt_set4(Opt *o, uint32_t price, uint32_t len, uint32_t dist, uint32_t extra)
{
	// one 16-byte store; the cell's state is rewritten when the parse reaches the cell, before anything reads it
	*reinterpret_cast<uint4 *>(o) = make_uint4(price, extra << 16, len, dist);
}

// One position of GetOptimum's main loop (LzmaEnc.c:1545-1949) for the common case, written for the warp instead of
// replicated: the position's match list, its distance-by-length table and the MATCH : LIT : REP_0 reaches are staged
// in shared memory by the look-ahead warp (b), nothing is clipped (numAvailFull >= fb, longest match < fb) and there
// are at most 32 pairs.  Lane k owns pair k; four groups of eight lanes compare the first eight bytes of the four
// reps; lanes 0-7 price the literal's eight decisions; the cells a match reaches are spread over the lanes.  Same
// decisions, same order of updates per cell as the loop below (which remains the general path and the host build).
__device__ __forceinline__ void opt_step_staged(Enc *e, const uint32_t *b, uint32_t hdr, uint32_t pos, uint32_t numAvailFull,
						uint32_t cur, uint32_t &last, uint32_t &position, uint32_t *reps,
						const uint8_t *srcAll, uint32_t pbMask, uint32_t fb, uint64_t &headNext)
{
	const uint32_t lane = lz_lane(), grp = lane >> 3, sub = lane & 7u;
	const uint32_t nd = hdr & 0xFFFFu, newLen = hdr >> 16, np = nd >> 1;
	const uint8_t *data = srcAll + (pos - 1);
	// ---- everything that does not depend on the parse table is requested first
	uint32_t pLen = 0, pDist = 0, pW = 0;
	if (lane < np) {
		pLen = b[1 + 2 * lane];
		pDist = b[2 + 2 * lane];
		pW = b[1 + kLkMaxList + lane];
	}
	const uint32_t dSub = data[sub], prevByte = *(data - 1), curByte = data[0];
	Opt *curOpt = &e->opt[cur], *nextOpt = curOpt + 1;
	const uint4 c0 = *reinterpret_cast<const uint4 *>(curOpt); // price, state | extra << 16, len, dist
	const uint4 n0 = *reinterpret_cast<const uint4 *>(nextOpt);
	const uint32_t statePrev1 = (curOpt - 1)->state; // the common predecessor (a literal or short rep): asked for before len is known
	LZ_T(15);
	// what ReadMatchDistances leaves behind
	e->additionalOffset++;
	e->numAvail = e->n - (pos - 1);
	e->pos = pos + 1;
	LZ_T(2);
	position++;
	const uint32_t curPrice = c0.x, curLen = c0.z, curDist = c0.w, curExtra = c0.y >> 16;
	uint32_t prev = cur - curLen, state;
	if (curLen == 1) {
		state = statePrev1;
		state = curDist == 0 ? st_shortrep(state) : st_lit(state);
	} else {
		if (curExtra) {
			prev -= curExtra;
			state = 8;
			if (curExtra == 1)
				state = curDist < kNumReps ? 8 : 7;
		} else {
			state = e->opt[prev].state;
			state = curDist < kNumReps ? st_rep(state) : st_match(state);
		}
		const uint4 pr = *reinterpret_cast<const uint4 *>(e->opt[prev].reps);
		if (curDist < kNumReps) {
			if (curDist == 0) {
				reps[0] = pr.x;
				reps[1] = pr.y;
				reps[2] = pr.z;
				reps[3] = pr.w;
			} else if (curDist == 1) {
				reps[0] = pr.y;
				reps[1] = pr.x;
				reps[2] = pr.z;
				reps[3] = pr.w;
			} else if (curDist == 2) {
				reps[0] = pr.z;
				reps[1] = pr.x;
				reps[2] = pr.y;
				reps[3] = pr.w;
			} else {
				reps[0] = pr.w;
				reps[1] = pr.x;
				reps[2] = pr.y;
				reps[3] = pr.z;
			}
		} else {
			reps[0] = curDist - kNumReps + 1;
			reps[1] = pr.x;
			reps[2] = pr.y;
			reps[3] = pr.z;
		}
	}
	curOpt->state = (uint16_t)state;
	*reinterpret_cast<uint4 *>(curOpt->reps) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
	*reinterpret_cast<uint4 *>(e->pubReps) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
	LZ_T(3);

	// ---- bytes from far back, all in flight at once: group g compares the first eight bytes of rep g
	uint32_t myRep = reps[0];
	myRep = grp == 1 ? reps[1] : myRep;
	myRep = grp == 2 ? reps[2] : myRep;
	myRep = grp == 3 ? reps[3] : myRep;
	const uint32_t rSub = (data - myRep)[sub];
	const uint32_t eq = lz_ballot(rSub == dSub); // bit 8 g + i: byte i of rep g equals byte i here
	const uint32_t matchByte = lz_shfl(rSub, 0);
	uint32_t repMask = 0;
	for (uint32_t q = 0; q < kNumReps; q++)
		repMask |= ((eq >> (8 * q)) & 3u) == 3u ? 1u << q : 0u;
	// MATCH : LIT : REP_0 reach of my pair (LzmaEnc.c:1876-1893), from the helper's byte comparison
	uint32_t pL2 = 0;
	if (lane < np) {
		uint32_t limit = pLen + 1 + fb;
		if (limit > numAvailFull)
			limit = numAvailFull;
		if ((pW >> 31) && pLen + 3 <= limit) {
			const uint32_t end = (pW & 0x7FFFFFFFu) < limit ? (pW & 0x7FFFFFFFu) : limit;
			pL2 = end - pLen;
		}
	}
	LZ_T(4);

	const uint32_t posState = position & pbMask;
	uint32_t matchPrice, litPrice, repMatchPrice;
	{
		const uint32_t prob = e->isMatch[state][posState];
		matchPrice = curPrice + price1(e, prob);
		litPrice = curPrice + price0(e, prob);
	}
	const uint32_t probRep = e->isRep[state];
	repMatchPrice = matchPrice + price1(e, probRep);
	const uint32_t normalMatchPrice = matchPrice + price0(e, probRep);
	uint32_t nPrice = n0.x, nLen = n0.z, nDist = n0.w;
	bool nextIsLit = false;
	if ((nPrice < kInfinity && matchByte == curByte) || litPrice > nPrice)
		litPrice = 0;
	else {
		const Prob *probs = lit_probs(e, position, prevByte);
		uint32_t v = 0;
		if (lane < 8) {
			const uint32_t node = (0x100u | curByte) >> (8 - lane), bit = (curByte >> (7 - lane)) & 1;
			if (is_lit_state(state))
				v = price_bit(e, probs[node], bit);
			else { // after a match: the match byte's bits pick the half while the bits above agree
				const uint32_t offs = (((matchByte ^ curByte) >> (8 - lane)) == 0) ? 0x100u : 0u;
				const uint32_t mb = ((matchByte >> (7 - lane)) & 1) << 8;
				v = price_bit(e, probs[offs + (mb & offs) + node], bit);
			}
		}
		litPrice += lz_sum(v);
		if (litPrice < nPrice) {
			opt_set4(nextOpt, litPrice, 1, kMarkLit, 0);
			nPrice = litPrice;
			nLen = 1;
			nDist = kMarkLit;
			nextIsLit = true;
		}
	}
	// SHORT_REP
	if (is_lit_state(state) && matchByte == curByte && repMatchPrice < nPrice && (nLen < 2 || nDist != 0)) {
		const uint32_t shortRepPrice = repMatchPrice + price_short_rep(e, state, posState);
		if (shortRepPrice < nPrice) {
			opt_set4(nextOpt, shortRepPrice, 1, 0, 0);
			nextIsLit = false;
		}
	}
	LZ_T(5);

	// LIT : REP_0 (bytes 1 and 2 behind rep0 equal; numAvailFull >= fb > 2)
	if (!nextIsLit && litPrice != 0 && matchByte != curByte && (eq & 6u) == 6u) {
		const uint8_t *data2 = data - reps[0];
		uint32_t len, limit = fb + 1;
		if (limit > numAvailFull)
			limit = numAvailFull;
		const uint32_t ne = ~eq & 0xF8u; // first of the bytes 3..7 that differs
		if (ne)
			len = lz_ffs(ne) - 1;
		else
			len = 8;
		if (len > limit)
			len = limit;
		else if (!ne && limit > 8)
			len = lz_extend(data2, data, 8, limit);
		const uint32_t state2 = st_lit(state), posState2 = (position + 1) & pbMask;
		const uint32_t price = litPrice + price_rep0(e, state2, posState2);
		const uint32_t offset = cur + len;
		if (last < offset)
			last = offset;
		len--;
		const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len);
		Opt *o = &e->opt[offset];
		if (price2 < o->price)
			opt_set4(o, price2, len, 0, 1);
	}
	uint32_t startLen = 2;
	LZ_T(6);
	// REP
	for (uint32_t rm = repMask; rm; rm &= rm - 1) {
		const uint32_t repIndex = lz_ffs(rm) - 1;
		const uint8_t *data2 = data - pick4(reps, repIndex);
		const uint32_t f = (eq >> (8 * repIndex)) & 0xFFu, nf = ~f & 0xFFu;
		uint32_t len = nf ? lz_ffs(nf) - 1 : 8; // equal leading bytes among the first eight; numAvail = fb
		if (len > fb)
			len = fb;
		else if (!nf && fb > 8)
			len = lz_extend(data2, data, 8, fb);
		if (last < cur + len)
			last = cur + len;
		uint32_t price = repMatchPrice + price_pure_rep(e, repIndex, state, posState);
		opt_rep_cells(e, cur, 2, len, price, posState, repIndex);
		if (repIndex == 0)
			startLen = len + 1;
		// REP : LIT : REP_0
		uint32_t len2 = len + 1, limit = len2 + fb;
		if (limit > numAvailFull)
			limit = numAvailFull;
		len2 += 2;
		bool two;
		if (len + 2 < 8)
			two = ((f >> (len + 1)) & 3u) == 3u;
		else
			two = len2 <= limit && data[len2 - 2] == data2[len2 - 2] && data[len2 - 1] == data2[len2 - 1];
		if (len2 <= limit && two) {
			uint32_t state2 = st_rep(state), posState2 = (position + len) & pbMask;
			price += len_price(&e->repLenPrices, posState, len) + price0(e, e->isMatch[state2][posState2]) +
				 lit_price_matched(e, lit_probs(e, position + len, data[len - 1]), data[len], data2[len]);
			state2 = 5; // kState_LitAfterRep
			posState2 = (posState2 + 1) & pbMask;
			price += price_rep0(e, state2, posState2);
			len2 = lz_extend(data2, data, len2, limit);
			len2 -= len;
			const uint32_t offset = cur + len + len2;
			if (last < offset)
				last = offset;
			len2--;
			const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len2);
			Opt *o = &e->opt[offset];
			if (price2 < o->price)
				opt_set4(o, price2, len2, repIndex, len + 1);
		}
	}
	LZ_T(7);
	// MATCH
	if (newLen >= startLen) {
		if (last < cur + newLen)
			last = cur + newLen;
		// MATCH : LIT : REP_0 of every pair that is long enough: priced by the pair's lane, applied in pair order
		const bool tr = lane < np && pLen >= startLen && pL2 != 0;
		const uint32_t trMask = lz_ballot(tr);
		if (trMask) {
			uint32_t tPrice = 0;
			if (tr) {
				const uint8_t *data2 = data - pDist - 1;
				uint32_t price = match_price(e, normalMatchPrice, posState, pLen, pDist);
				uint32_t state2 = st_match(state), posState2 = (position + pLen) & pbMask;
				price += price0(e, e->isMatch[state2][posState2]);
				price += lit_price_matched_1(e, lit_probs(e, position + pLen, data[pLen - 1]), data[pLen], data2[pLen]);
				state2 = 4; // kState_LitAfterMatch
				posState2 = (posState2 + 1) & pbMask;
				price += price_rep0(e, state2, posState2);
				tPrice = price + len_price(&e->repLenPrices, posState2, pL2 - 1);
			}
			for (uint32_t m = trMask; m; m &= m - 1) {
				const uint32_t k = lz_ffs(m) - 1;
				const uint32_t len = lz_shfl(pLen, k), l2 = lz_shfl(pL2, k), dist = lz_shfl(pDist, k), price2 = lz_shfl(tPrice, k);
				const uint32_t offset = cur + len + l2;
				if (last < offset)
					last = offset;
				Opt *o = &e->opt[offset];
				if (price2 < o->price)
					opt_set4(o, price2, l2 - 1, dist + kNumReps, len + 1);
			}
		}
		LZ_T(8);
		// the cells [cur + startLen, cur + newLen]: the shortest-distance pair that reaches each length (staged table)
		for (uint32_t len = startLen + lane; len <= newLen; len += LZ_W) {
			const uint32_t dist = b[kLkByLen + len];
			const uint32_t price = match_price(e, normalMatchPrice, posState, len, dist);
			Opt *o = &e->opt[cur + len];
			if (price < o->price)
				opt_set4(o, price, len, dist + kNumReps, 0);
		}
		LZ_T(9);
	}
	// the next position's ring entry, asked for now: the loop top finds it in a register (and asks again if the
	// look-ahead warp had not got there yet)
	headNext = *reinterpret_cast<const volatile uint64_t *>(&e->lkHead[(pos + 1) & (kLkSlots - 1)][0]);
	lz_sync();
	LZ_T(16);
}
// ---- the staged step on two warps ---------------------------------------------------------------------------------
// The literal / short-rep half of position cur + 1 (cell's state and reps, prices, the literal's eight decisions) does
// not depend on what the rep / match half of position cur writes -- matches reach cells cur + 2 and beyond -- except
// for the final comparison with cell cur + 2.  So warp A runs the first half of every position and warp B the second,
// one position behind: A works out position cur + 1 while B prices the reps and matches of position cur, waits for B,
// does its two compare-and-store updates of cell cur + 2 and hands position cur + 1 to B.  Per cell the order of updates
// is the reference's: everything of position cur (A's, then B's), then position cur + 1.  Named barriers 1 (a position
// for B) and 2 (B is done) pair one arriving warp with one waiting warp (PTX producer / consumer form).
LZ_INL void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
LZ_INL void bar_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
constexpr int kBarGo = 1, kBarDone = 2;

// A waits for B to finish the position it was given and takes over `last`.
LZ_INL void split_drain(Enc *e, bool &pending, uint32_t &last)
{
	if (!pending)
		return;
#if defined(LZ_PROF)
	const long long w0_ = clock64();
#endif
	bar_wait(kBarDone);
#if defined(LZ_PROF)
	e->prof[21] += (uint64_t)(clock64() - w0_);
	e->profN[21]++;
#endif
	pending = false;
	last = *(volatile uint32_t *)&e->lastB;
}

__device__ __forceinline__ void opt_step_a(Enc *e, uint32_t hdr, uint32_t pos, uint32_t numAvailFull, uint32_t cur,
					   uint32_t &last, uint32_t &position, uint32_t *reps, const uint8_t *srcAll,
					   uint32_t pbMask, uint64_t &headNext, bool &pending)
{
	const uint32_t lane = lz_lane(), grp = lane >> 3, sub = lane & 7u;
	const uint8_t *data = srcAll + (pos - 1);
	const uint32_t prevByte = *(data - 1), curByte = data[0], dSub = data[sub];
	Opt *curOpt = &e->opt[cur], *nextOpt = curOpt + 1;
	const uint4 c0 = *reinterpret_cast<const uint4 *>(curOpt); // price, state | extra << 16, len, dist
	const uint32_t statePrev1 = (curOpt - 1)->state;
	// what ReadMatchDistances leaves behind
	e->additionalOffset++;
	e->numAvail = e->n - (pos - 1);
	e->pos = pos + 1;
	LZ_T(2);
	position++;
	const uint32_t curPrice = c0.x, curLen = c0.z, curDist = c0.w, curExtra = c0.y >> 16;
	uint32_t prev = cur - curLen, state;
	if (curLen == 1) {
		state = statePrev1;
		state = curDist == 0 ? st_shortrep(state) : st_lit(state);
	} else {
		if (curExtra) {
			prev -= curExtra;
			state = 8;
			if (curExtra == 1)
				state = curDist < kNumReps ? 8 : 7;
		} else {
			state = e->opt[prev].state;
			state = curDist < kNumReps ? st_rep(state) : st_match(state);
		}
		const uint4 pr = *reinterpret_cast<const uint4 *>(e->opt[prev].reps);
		if (curDist < kNumReps) {
			if (curDist == 0) {
				reps[0] = pr.x;
				reps[1] = pr.y;
				reps[2] = pr.z;
				reps[3] = pr.w;
			} else if (curDist == 1) {
				reps[0] = pr.y;
				reps[1] = pr.x;
				reps[2] = pr.z;
				reps[3] = pr.w;
			} else if (curDist == 2) {
				reps[0] = pr.z;
				reps[1] = pr.x;
				reps[2] = pr.y;
				reps[3] = pr.w;
			} else {
				reps[0] = pr.w;
				reps[1] = pr.x;
				reps[2] = pr.y;
				reps[3] = pr.z;
			}
		} else {
			reps[0] = curDist - kNumReps + 1;
			reps[1] = pr.x;
			reps[2] = pr.y;
			reps[3] = pr.z;
		}
	}
	curOpt->state = (uint16_t)state;
	*reinterpret_cast<uint4 *>(curOpt->reps) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
	*reinterpret_cast<uint4 *>(e->pubReps) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
	// bytes from far back, all in flight at once: group g compares the first eight bytes of rep g (warp B works from
	// the resulting bits; the byte behind rep0 is the literal's match byte)
	uint32_t myRep = reps[0];
	myRep = grp == 1 ? reps[1] : myRep;
	myRep = grp == 2 ? reps[2] : myRep;
	myRep = grp == 3 ? reps[3] : myRep;
	const uint32_t rSub = (data - myRep)[sub];
	LZ_T(3);
	const uint32_t posState = position & pbMask;
	uint32_t matchPrice, litPrice, repMatchPrice;
	{
		const uint32_t prob = e->isMatch[state][posState];
		matchPrice = curPrice + price1(e, prob);
		litPrice = curPrice + price0(e, prob);
	}
	const uint32_t probRep = e->isRep[state];
	repMatchPrice = matchPrice + price1(e, probRep);
	const uint32_t normalMatchPrice = matchPrice + price0(e, probRep);
	const uint32_t eq = lz_ballot(rSub == dSub); // bit 8 g + i: byte i of rep g equals byte i here
	const uint32_t matchByte = lz_shfl(rSub, 0);
	LZ_T(4);
	// the literal's eight decisions, priced whether or not the comparison below will want them: B is still busy
	uint32_t litFull;
	{
		const Prob *probs = lit_probs(e, position, prevByte);
		uint32_t v = 0;
		if (lane < 8) {
			const uint32_t node = (0x100u | curByte) >> (8 - lane), bit = (curByte >> (7 - lane)) & 1;
			if (is_lit_state(state))
				v = price_bit(e, probs[node], bit);
			else {
				const uint32_t offs = (((matchByte ^ curByte) >> (8 - lane)) == 0) ? 0x100u : 0u;
				const uint32_t mb = ((matchByte >> (7 - lane)) & 1) << 8;
				v = price_bit(e, probs[offs + (mb & offs) + node], bit);
			}
		}
		litFull = litPrice + lz_sum(v);
	}
	const bool srCand = is_lit_state(state) && matchByte == curByte;
	const uint32_t shortRepPrice = srCand ? repMatchPrice + price_short_rep(e, state, posState) : 0;
	headNext = *reinterpret_cast<const volatile uint64_t *>(&e->lkHead[(pos + 1) & (kLkSlots - 1)][0]);

	LZ_T(5);
	// ---- cell cur + 1 must now hold everything position cur - 1 wrote
	split_drain(e, pending, last);
	LZ_T(6);
	const uint4 n0 = *reinterpret_cast<const uint4 *>(nextOpt);
	uint32_t nPrice = n0.x, nLen = n0.z, nDist = n0.w;
	bool nextIsLit = false;
	if ((nPrice < kInfinity && matchByte == curByte) || litPrice > nPrice)
		litPrice = 0;
	else {
		litPrice = litFull;
		if (litPrice < nPrice) {
			opt_set4(nextOpt, litPrice, 1, kMarkLit, 0);
			nPrice = litPrice;
			nLen = 1;
			nDist = kMarkLit;
			nextIsLit = true;
		}
	}
	// SHORT_REP
	if (srCand && repMatchPrice < nPrice && (nLen < 2 || nDist != 0)) {
		if (shortRepPrice < nPrice) {
			opt_set4(nextOpt, shortRepPrice, 1, 0, 0);
			nextIsLit = false;
		}
	}
	LZ_T(7);
	// the rep / match half goes to warp B
	StepPkt *k = &e->pkt;
	*reinterpret_cast<uint4 *>(&k->cmd) = make_uint4(1u, cur, pos, position);
	*reinterpret_cast<uint4 *>(&k->naf) = make_uint4(numAvailFull, hdr, state, posState);
	*reinterpret_cast<uint4 *>(k->reps) = make_uint4(reps[0], reps[1], reps[2], reps[3]);
	*reinterpret_cast<uint4 *>(&k->matchPrice) = make_uint4(matchPrice, repMatchPrice, normalMatchPrice, litPrice);
	const bool litRep0 = !nextIsLit && litPrice != 0 && matchByte != curByte && (eq & 6u) == 6u; // LIT : REP_0 is worth a look
	*reinterpret_cast<uint4 *>(&k->nextIsLit) = make_uint4(litRep0 ? 1u : 0u, eq, 0u, last);
	bar_arrive(kBarGo);
	pending = true;
	LZ_T(8);
}

// Warp B: the rep / match half of the position in e->pkt (LzmaEnc.c:1700-1949).
__device__ __forceinline__ void opt_step_b(Enc *e)
{
	const uint32_t lane = lz_lane();
	const StepPkt *k = &e->pkt;
	const uint4 k0 = *reinterpret_cast<const uint4 *>(&k->cmd), k1 = *reinterpret_cast<const uint4 *>(&k->naf);
	const uint4 k2 = *reinterpret_cast<const uint4 *>(k->reps), k3 = *reinterpret_cast<const uint4 *>(&k->matchPrice);
	const uint4 k4 = *reinterpret_cast<const uint4 *>(&k->nextIsLit);
	const uint32_t cur = k0.y, pos = k0.z, position = k0.w, numAvailFull = k1.x, hdr = k1.y, state = k1.z, posState = k1.w;
	uint32_t reps[kNumReps] = { k2.x, k2.y, k2.z, k2.w };
	const uint32_t repMatchPrice = k3.y, normalMatchPrice = k3.z, litPrice = k3.w;
	const bool litRep0 = k4.x != 0;
	const uint32_t eq = k4.y;
	uint32_t last = k4.w;
	const uint32_t fb = e->fb, pbMask = e->pbMask;
	const uint32_t *b = e->lkRing + (pos & (kLkSlots - 1)) * kLkWords;
	const uint32_t nd = hdr & 0xFFFFu, newLen = hdr >> 16, np = nd >> 1;
	const uint8_t *data = e->src + (pos - 1);
	uint32_t pLen = 0, pDist = 0, pW = 0;
	if (lane < np) {
		pLen = b[1 + 2 * lane];
		pDist = b[2 + 2 * lane];
		pW = b[1 + kLkMaxList + lane];
	}
	uint32_t repMask = 0;
	for (uint32_t q = 0; q < kNumReps; q++)
		repMask |= ((eq >> (8 * q)) & 3u) == 3u ? 1u << q : 0u;
	uint32_t pL2 = 0;
	if (lane < np) {
		uint32_t limit = pLen + 1 + fb;
		if (limit > numAvailFull)
			limit = numAvailFull;
		if ((pW >> 31) && pLen + 3 <= limit) {
			const uint32_t end = (pW & 0x7FFFFFFFu) < limit ? (pW & 0x7FFFFFFFu) : limit;
			pL2 = end - pLen;
		}
	}
	// LIT : REP_0 (bytes 1 and 2 behind rep0 equal; numAvailFull >= fb > 2)
	if (litRep0) {
		const uint8_t *data2 = data - reps[0];
		uint32_t len, limit = fb + 1;
		if (limit > numAvailFull)
			limit = numAvailFull;
		const uint32_t ne = ~eq & 0xF8u; // first of the bytes 3..7 that differs
		if (ne)
			len = lz_ffs(ne) - 1;
		else
			len = 8;
		if (len > limit)
			len = limit;
		else if (!ne && limit > 8)
			len = lz_extend(data2, data, 8, limit);
		const uint32_t state2 = st_lit(state), posState2 = (position + 1) & pbMask;
		const uint32_t price = litPrice + price_rep0(e, state2, posState2);
		const uint32_t offset = cur + len;
		if (last < offset)
			last = offset;
		len--;
		const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len);
		Opt *o = &e->opt[offset];
		if (price2 < o->price)
			opt_set4(o, price2, len, 0, 1);
	}
	uint32_t startLen = 2;
	// REP
	for (uint32_t rm = repMask; rm; rm &= rm - 1) {
		const uint32_t repIndex = lz_ffs(rm) - 1;
		const uint8_t *data2 = data - pick4(reps, repIndex);
		const uint32_t f = (eq >> (8 * repIndex)) & 0xFFu, nf = ~f & 0xFFu;
		uint32_t len = nf ? lz_ffs(nf) - 1 : 8; // equal leading bytes among the first eight; numAvail = fb
		if (len > fb)
			len = fb;
		else if (!nf && fb > 8)
			len = lz_extend(data2, data, 8, fb);
		if (last < cur + len)
			last = cur + len;
		uint32_t price = repMatchPrice + price_pure_rep(e, repIndex, state, posState);
		opt_rep_cells(e, cur, 2, len, price, posState, repIndex);
		if (repIndex == 0)
			startLen = len + 1;
		// REP : LIT : REP_0
		uint32_t len2 = len + 1, limit = len2 + fb;
		if (limit > numAvailFull)
			limit = numAvailFull;
		len2 += 2;
		bool two;
		if (len + 2 < 8)
			two = ((f >> (len + 1)) & 3u) == 3u;
		else
			two = len2 <= limit && data[len2 - 2] == data2[len2 - 2] && data[len2 - 1] == data2[len2 - 1];
		if (len2 <= limit && two) {
			uint32_t state2 = st_rep(state), posState2 = (position + len) & pbMask;
			price += len_price(&e->repLenPrices, posState, len) + price0(e, e->isMatch[state2][posState2]) +
				 lit_price_matched(e, lit_probs(e, position + len, data[len - 1]), data[len], data2[len]);
			state2 = 5; // kState_LitAfterRep
			posState2 = (posState2 + 1) & pbMask;
			price += price_rep0(e, state2, posState2);
			len2 = lz_extend(data2, data, len2, limit);
			len2 -= len;
			const uint32_t offset = cur + len + len2;
			if (last < offset)
				last = offset;
			len2--;
			const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len2);
			Opt *o = &e->opt[offset];
			if (price2 < o->price)
				opt_set4(o, price2, len2, repIndex, len + 1);
		}
	}
	// MATCH
	if (newLen >= startLen) {
		if (last < cur + newLen)
			last = cur + newLen;
		// MATCH : LIT : REP_0 of every pair that is long enough: priced by the pair's lane, applied in pair order
		const bool tr = lane < np && pLen >= startLen && pL2 != 0;
		const uint32_t trMask = lz_ballot(tr);
		if (trMask) {
			uint32_t tPrice = 0;
			if (tr) {
				const uint8_t *data2 = data - pDist - 1;
				uint32_t price = match_price(e, normalMatchPrice, posState, pLen, pDist);
				uint32_t state2 = st_match(state), posState2 = (position + pLen) & pbMask;
				price += price0(e, e->isMatch[state2][posState2]);
				price += lit_price_matched_1(e, lit_probs(e, position + pLen, data[pLen - 1]), data[pLen], data2[pLen]);
				state2 = 4; // kState_LitAfterMatch
				posState2 = (posState2 + 1) & pbMask;
				price += price_rep0(e, state2, posState2);
				tPrice = price + len_price(&e->repLenPrices, posState2, pL2 - 1);
			}
			for (uint32_t m = trMask; m; m &= m - 1) {
				const uint32_t k = lz_ffs(m) - 1;
				const uint32_t len = lz_shfl(pLen, k), l2 = lz_shfl(pL2, k), dist = lz_shfl(pDist, k), price2 = lz_shfl(tPrice, k);
				const uint32_t offset = cur + len + l2;
				if (last < offset)
					last = offset;
				Opt *o = &e->opt[offset];
				if (price2 < o->price)
					opt_set4(o, price2, l2 - 1, dist + kNumReps, len + 1);
			}
		}
		// the cells [cur + startLen, cur + newLen]: the shortest-distance pair that reaches each length (staged table)
		for (uint32_t len = startLen + lane; len <= newLen; len += LZ_W) {
			const uint32_t dist = b[kLkByLen + len];
			const uint32_t price = match_price(e, normalMatchPrice, posState, len, dist);
			Opt *o = &e->opt[cur + len];
			if (price < o->price)
				opt_set4(o, price, len, dist + kNumReps, 0);
		}
	}
	*(volatile uint32_t *)&e->lastB = last;
}
#endif

// GetOptimum (LzmaEnc.c:1219-1968).
//
// Order of updates: the reference interleaves, per position, "cell" updates (a rep or match of every
// length) with "x : LIT : REP_0" trials that write one cell further on.  Updates only replace a cell when
// strictly cheaper, so what must be kept is, per cell, the order of the updates that reach it.  A trial of
// rep i / pair k lands beyond that rep's / pair's own length, hence for any one cell the trials that reach
// it come before the normal match update of that cell and in pair order; reps are kept strictly in index
// order (cells, then trial).  Everything else (the cells of one loop are distinct) runs across lanes.
LZ_FN inline uint32_t get_optimum(Enc *e, uint32_t position)
{
	uint32_t last, cur;
	uint32_t reps[kNumReps], repLens[kNumReps];
	uint32_t *matches = e->matches;
	const uint32_t fb = e->fb;
	{
		uint32_t numAvail, numPairs, mainLen, repMaxIndex, i, posState, matchPrice, repMatchPrice;
		LZ_T(0);
		e->optCur = e->optEnd = 0;
		if (e->additionalOffset == 0)
			mainLen = read_matches(e, &numPairs);
		else {
			mainLen = e->longestMatchLen;
			numPairs = e->numPairs;
		}
		numAvail = e->numAvail;
		if (numAvail < 2) {
			e->backRes = kMarkLit;
			return 1;
		}
		if (numAvail > kMatchMax)
			numAvail = kMatchMax;
		const uint8_t *data = mf_cur(e) - 1;
		for (i = 0; i < kNumReps; i++)
			reps[i] = e->reps[i];
		lz_sync();
		LZ_PFOR(q, kNumReps) { // the four first-two-bytes checks at once (four cache misses overlap)
			const uint8_t *data2 = data - e->reps[q];
			e->xRepLen[q] = (data[0] == data2[0] && data[1] == data2[1]) ? 2 : 0;
		}
		lz_sync();
		repMaxIndex = 0;
		uint32_t repMaxLen = 0; // repLens[repMaxIndex]
		for (i = 0; i < kNumReps; i++)
			repLens[i] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (i = 0; i < kNumReps; i++) {
			if (e->xRepLen[i] == 0 || repMaxLen == kMatchMax) // a rep of the maximum length is returned right away
				continue;
			const uint32_t len = lz_extend(data - reps[i], data, 2, numAvail);
			repLens[i] = len;
			if (len > repMaxLen) {
				repMaxIndex = i;
				repMaxLen = len;
			}
		}
		if (repMaxLen >= fb) {
			e->backRes = repMaxIndex;
			const uint32_t len = repMaxLen;
			move_pos(e, len - 1);
			LZ_T(11);
			return len;
		}
		if (mainLen >= fb) {
			e->backRes = matches[numPairs - 1] + kNumReps;
			move_pos(e, mainLen - 1);
			LZ_T(11);
			return mainLen;
		}
		const uint8_t curByte = *data, matchByte = *(data - reps[0]);
		last = repMaxLen;
		if (last <= mainLen)
			last = mainLen;
		if (last < 2 && curByte != matchByte) {
			e->backRes = kMarkLit;
			LZ_T(11);
			return 1;
		}
		e->opt[0].state = (uint16_t)e->state;
		posState = position & e->pbMask;
		{
			const Prob *probs = lit_probs(e, position, *(data - 1));
			const uint32_t lp = !is_lit_state(e->state) ? lit_price_matched(e, probs, curByte, matchByte) : lit_price(e, probs, curByte);
			e->opt[1].price = price0(e, e->isMatch[e->state][posState]) + lp;
		}
		e->opt[1].dist = kMarkLit;
		e->opt[1].extra = 0;
		matchPrice = price1(e, e->isMatch[e->state][posState]);
		repMatchPrice = matchPrice + price1(e, e->isRep[e->state]);
		if (matchByte == curByte && repLens[0] == 0) {
			const uint32_t shortRepPrice = repMatchPrice + price_short_rep(e, e->state, posState);
			if (shortRepPrice < e->opt[1].price) {
				e->opt[1].price = shortRepPrice;
				e->opt[1].dist = 0;
				e->opt[1].extra = 0;
			}
			if (last < 2) {
				e->backRes = e->opt[1].dist;
				LZ_T(11);
				return 1;
			}
		}
		e->opt[1].len = 1;
		for (i = 0; i < kNumReps; i++)
			e->opt[0].reps[i] = reps[i];

#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (i = 0; i < kNumReps; i++) { // REP
			const uint32_t repLen = repLens[i];
			if (repLen < 2)
				continue;
			const uint32_t price = repMatchPrice + price_pure_rep(e, i, e->state, posState);
			opt_rep_cells(e, 0, 2, repLen, price, posState, i);
		}
		{ // MATCH
			uint32_t len = repLens[0] + 1;
			if (len <= mainLen) {
				const uint32_t normalMatchPrice = matchPrice + price0(e, e->isRep[e->state]);
				if (len < 2)
					len = 2;
				opt_match_cells(e, 0, len, mainLen, normalMatchPrice, posState);
			}
		}
		cur = 0;
		LZ_T(1);
	}
#if defined(__CUDA_ARCH__)
	// loop invariants the compiler cannot keep in registers by itself (every store into *e may alias them)
	const bool lkOn = e->lkOn != 0;
	uint64_t headNext = 0; // ring head word read ahead by the previous staged step (position in the low half)
	const bool splitOn = e->splitOn != 0;
	bool pending = false;  // warp B is working on the previous position
	const uint32_t nAll = e->n, pbMask = e->pbMask;
	const uint8_t *const srcAll = e->src;
#endif

	for (;;) {
		uint32_t numAvail, numAvailFull, newLen, numPairs, prev, state, posState, startLen;
		uint32_t litPrice, matchPrice, repMatchPrice;
		bool nextIsLit;
#if defined(__CUDA_ARCH__)
		if (++cur == last) {
			split_drain(e, pending, last); // B may still be extending `last`
			if (cur == last)
				break;
		}
		if (cur >= kNumOpts - 64)
			split_drain(e, pending, last);
#else
		if (++cur == last)
			break;
#endif
		if (cur >= kNumOpts - 64) {
			uint32_t best = cur, price = e->opt[cur].price;
			for (uint32_t j = cur + 1; j <= last; j++) {
				const uint32_t price2 = e->opt[j].price;
				if (price >= price2) {
					price = price2;
					best = j;
				}
			}
			const uint32_t delta = best - cur;
			if (delta != 0)
				move_pos(e, delta);
			cur = best;
			break;
		}
#if defined(__CUDA_ARCH__)
		LZ_T(19);
		if (lkOn) {
			// every iteration reads exactly one position: the finder stands two past `position` (1-based)
			const uint32_t pos = position + 2, slot = pos & (kLkSlots - 1);
			const uint32_t *b = e->lkRing + slot * kLkWords;
			const uint32_t availReal = nAll - (pos - 1);
			uint32_t naf = kNumOpts - 1 - cur; // numAvailFull
			if (naf > availReal)
				naf = availReal;
#if defined(LZ_PROF)
			if (e->pos != pos)
				e->profN[20]++;
#endif
			// one 8-byte load: the tag and the header the helper stored together after its fence; the entry's words
			// are read after it (shared-memory loads of one warp are served in order)
			uint64_t head = headNext;
			if ((uint32_t)head != pos)
				head = *reinterpret_cast<const volatile uint64_t *>(&e->lkHead[slot][0]);
			const uint32_t hdr = (uint32_t)(head >> 32);
			if (naf >= fb && naf >= 8 && (uint32_t)head == pos) {
				if ((hdr & 0xFFFFu) <= 64 && (hdr >> 16) < fb) {
					LZ_T(14);
					if (splitOn)
						opt_step_a(e, hdr, pos, naf, cur, last, position, reps, srcAll, pbMask, headNext, pending);
					else
						opt_step_staged(e, b, hdr, pos, naf, cur, last, position, reps, srcAll, pbMask, fb, headNext);
					continue;
				}
			}
		}
		split_drain(e, pending, last); // the general step below reads and writes the whole table
#endif
		newLen = read_matches(e, &numPairs);
		LZ_T(2);
		if (newLen >= fb) {
			e->numPairs = numPairs;
			e->longestMatchLen = newLen;
			break;
		}
		Opt *curOpt = &e->opt[cur];
		position++;
		prev = cur - curOpt->len;
		if (curOpt->len == 1) {
			state = e->opt[prev].state;
			state = curOpt->dist == 0 ? st_shortrep(state) : st_lit(state);
		} else {
			const uint32_t dist = curOpt->dist;
			if (curOpt->extra) {
				prev -= curOpt->extra;
				state = 8; // kState_RepAfterLit
				if (curOpt->extra == 1)
					state = dist < kNumReps ? 8 : 7; // RepAfterLit : MatchAfterLit
			} else {
				state = e->opt[prev].state;
				state = dist < kNumReps ? st_rep(state) : st_match(state);
			}
			const Opt *prevOpt = &e->opt[prev];
			uint32_t b0 = prevOpt->reps[0];
			if (dist < kNumReps) {
				if (dist == 0) {
					reps[0] = b0;
					reps[1] = prevOpt->reps[1];
					reps[2] = prevOpt->reps[2];
					reps[3] = prevOpt->reps[3];
				} else {
					reps[1] = b0;
					b0 = prevOpt->reps[1];
					if (dist == 1) {
						reps[0] = b0;
						reps[2] = prevOpt->reps[2];
						reps[3] = prevOpt->reps[3];
					} else {
						reps[2] = b0;
						reps[0] = prevOpt->reps[dist];
						reps[3] = prevOpt->reps[dist ^ 1];
					}
				}
			} else {
				reps[0] = dist - kNumReps + 1;
				reps[1] = b0;
				reps[2] = prevOpt->reps[1];
				reps[3] = prevOpt->reps[2];
			}
		}
		curOpt->state = (uint16_t)state;
		for (uint32_t i = 0; i < kNumReps; i++)
			curOpt->reps[i] = reps[i];

		LZ_T(3);
		const uint8_t *data = mf_cur(e) - 1;
		numAvailFull = e->numAvail;
		{
			const uint32_t temp = kNumOpts - 1 - cur;
			if (numAvailFull > temp)
				numAvailFull = temp;
		}
		numAvail = numAvailFull <= fb ? numAvailFull : fb;
		// MATCH list clipped to what is available (LzmaEnc.c:1830-1838; independent of the REP section)
		bool clipped = false;
		if (numAvailFull >= 2 && newLen > numAvail) {
			clipped = true;
			newLen = numAvail;
			for (numPairs = 0; newLen > matches[numPairs]; numPairs += 2) {
			}
			matches[numPairs] = newLen;
			numPairs += 2;
		}
		// ---- one pass over everything that needs bytes from far back in the block, so that the cache
		// misses overlap: first two bytes of each rep, and for each pair the bytes after a one-literal gap
		lz_sync();
		LZ_PFOR(q, kNumReps + (numAvailFull >= 2 ? (numPairs >> 1) : 0)) {
			if (q < kNumReps) {
				const uint8_t *data2 = data - curOpt->reps[q];
				e->xRepLen[q] = (numAvailFull >= 2 && data[0] == data2[0] && data[1] == data2[1]) ? 2 : 0;
			} else {
				// MATCH : LIT : REP_0 reach of pair k (LzmaEnc.c:1876-1893)
				const uint32_t k = q - kNumReps, len = matches[2 * k];
				uint32_t len2 = len + 1, limit = len2 + fb, res = 0;
				if (limit > numAvailFull)
					limit = numAvailFull;
				len2 += 2;
#if defined(__CUDA_ARCH__)
				if (e->lkSlot >= 0 && !clipped) {
					// the look-ahead warp compared the bytes already, up to min(len + 1 + fb, numAvail): its end,
					// cut at this position's own limit, is where the loop below would stop
					const uint32_t w = e->lkRing[(uint32_t)e->lkSlot * kLkWords + 1 + kLkMaxList + k];
					if ((w >> 31) && len2 <= limit) {
						const uint32_t end = (w & 0x7FFFFFFFu) < limit ? (w & 0x7FFFFFFFu) : limit;
						res = end - len;
					}
				} else
#endif
				{
					const uint8_t *data2 = data - matches[2 * k + 1] - 1;
					if (len2 <= limit && data[len2 - 2] == data2[len2 - 2] && data[len2 - 1] == data2[len2 - 1]) {
						while (len2 < limit && data[len2] == data2[len2])
							len2++;
						res = len2 - len;
					}
				}
				e->xPairLen2[k] = res;
			}
		}
		lz_sync();
		LZ_T(4);

		const uint8_t curByte = *data, matchByte = *(data - reps[0]);
		posState = position & e->pbMask;
		{
			const uint32_t curPrice = curOpt->price;
			const uint32_t prob = e->isMatch[state][posState];
			matchPrice = curPrice + price1(e, prob);
			litPrice = curPrice + price0(e, prob);
		}
		Opt *nextOpt = &e->opt[cur + 1];
		nextIsLit = false;
		if ((nextOpt->price < kInfinity && matchByte == curByte) || litPrice > nextOpt->price)
			litPrice = 0;
		else {
			const Prob *probs = lit_probs(e, position, *(data - 1));
			litPrice += !is_lit_state(state) ? lit_price_matched(e, probs, curByte, matchByte) : lit_price(e, probs, curByte);
			if (litPrice < nextOpt->price) {
				opt_set(nextOpt, litPrice, 1, kMarkLit, 0);
				nextIsLit = true;
			}
		}
		repMatchPrice = matchPrice + price1(e, e->isRep[state]);
		// SHORT_REP
		if (is_lit_state(state) && matchByte == curByte && repMatchPrice < nextOpt->price &&
		    (nextOpt->len < 2 || nextOpt->dist != 0)) {
			const uint32_t shortRepPrice = repMatchPrice + price_short_rep(e, state, posState);
			if (shortRepPrice < nextOpt->price) {
				opt_set(nextOpt, shortRepPrice, 1, 0, 0);
				nextIsLit = false;
			}
		}
		LZ_T(5);
		if (numAvailFull < 2)
			continue;

		// LIT : REP_0
		if (!nextIsLit && litPrice != 0 && matchByte != curByte && numAvailFull > 2) {
			const uint8_t *data2 = data - reps[0];
			if (data[1] == data2[1] && data[2] == data2[2]) {
				uint32_t len, limit = fb + 1;
				if (limit > numAvailFull)
					limit = numAvailFull;
				len = lz_extend(data2, data, 3, limit);
				const uint32_t state2 = st_lit(state), posState2 = (position + 1) & e->pbMask;
				const uint32_t price = litPrice + price_rep0(e, state2, posState2);
				const uint32_t offset = cur + len;
				if (last < offset)
					last = offset;
				len--;
				const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len);
				Opt *o = &e->opt[offset];
				if (price2 < o->price)
					opt_set(o, price2, len, 0, 1);
			}
		}
		startLen = 2;
		LZ_T(6);
		// REP
		for (uint32_t repIndex = 0; repIndex < kNumReps; repIndex++) {
			if (e->xRepLen[repIndex] == 0)
				continue;
			const uint8_t *data2 = data - pick4(reps, repIndex);
			const uint32_t len = lz_extend(data2, data, 2, numAvail);
			if (last < cur + len)
				last = cur + len;
			uint32_t price = repMatchPrice + price_pure_rep(e, repIndex, state, posState);
			opt_rep_cells(e, cur, 2, len, price, posState, repIndex);
			if (repIndex == 0)
				startLen = len + 1;
			// REP : LIT : REP_0
			uint32_t len2 = len + 1, limit = len2 + fb;
			if (limit > numAvailFull)
				limit = numAvailFull;
			len2 += 2;
			if (len2 <= limit && data[len2 - 2] == data2[len2 - 2] && data[len2 - 1] == data2[len2 - 1]) {
				uint32_t state2 = st_rep(state), posState2 = (position + len) & e->pbMask;
				price += len_price(&e->repLenPrices, posState, len) + price0(e, e->isMatch[state2][posState2]) +
					 lit_price_matched(e, lit_probs(e, position + len, data[len - 1]), data[len], data2[len]);
				state2 = 5; // kState_LitAfterRep
				posState2 = (posState2 + 1) & e->pbMask;
				price += price_rep0(e, state2, posState2);
				len2 = lz_extend(data2, data, len2, limit);
				len2 -= len;
				const uint32_t offset = cur + len + len2;
				if (last < offset)
					last = offset;
				len2--;
				const uint32_t price2 = price + len_price(&e->repLenPrices, posState2, len2);
				Opt *o = &e->opt[offset];
				if (price2 < o->price)
					opt_set(o, price2, len2, repIndex, len + 1);
			}
		}
		LZ_T(7);
		// MATCH
		if (newLen >= startLen) {
			const uint32_t normalMatchPrice = matchPrice + price0(e, e->isRep[state]);
			if (last < cur + newLen)
				last = cur + newLen;
			// MATCH : LIT : REP_0 of every pair that is long enough: prices across lanes, then applied in pair order
			const uint32_t np = numPairs >> 1;
			lz_sync();
			LZ_PFOR(k, np) {
				const uint32_t len = matches[2 * k], l2 = e->xPairLen2[k];
				if (len < startLen || l2 == 0)
					continue;
				const uint32_t dist = matches[2 * k + 1];
				const uint8_t *data2 = data - dist - 1;
				uint32_t price = match_price(e, normalMatchPrice, posState, len, dist);
				uint32_t state2 = st_match(state), posState2 = (position + len) & e->pbMask;
				price += price0(e, e->isMatch[state2][posState2]);
				price += lit_price_matched_1(e, lit_probs(e, position + len, data[len - 1]), data[len], data2[len]);
				state2 = 4; // kState_LitAfterMatch
				posState2 = (posState2 + 1) & e->pbMask;
				price += price_rep0(e, state2, posState2);
				e->xPairPrice[k] = price + len_price(&e->repLenPrices, posState2, l2 - 1);
			}
			lz_sync();
			for (uint32_t k = 0; k < np; k++) {
				const uint32_t len = matches[2 * k], l2 = e->xPairLen2[k];
				if (len < startLen || l2 == 0)
					continue;
				const uint32_t offset = cur + len + l2;
				if (last < offset)
					last = offset;
				Opt *o = &e->opt[offset];
				const uint32_t price2 = e->xPairPrice[k];
				if (price2 < o->price)
					opt_set(o, price2, l2 - 1, matches[2 * k + 1] + kNumReps, len + 1);
			}
			LZ_T(8);
			opt_match_cells(e, cur, startLen, newLen, normalMatchPrice, posState);
			LZ_T(9);
		}
	}
#if defined(__CUDA_ARCH__)
	split_drain(e, pending, last);
#endif
	lz_sync();
	LZ_PFOR(q, last)
		e->opt[1 + q].price = kInfinity;
	lz_sync();
	const uint32_t res_ = backward(e, cur);
	LZ_T(10);
	return res_;
}

// GetOptimumFast (LzmaEnc.c:1970-2098): levels 1-4.  No price tables; greedy with one position of look-ahead.
LZ_INL bool change_pair(uint32_t smallDist, uint32_t bigDist) { return (bigDist >> 7) > smallDist; }

LZ_FN inline uint32_t get_optimum_fast(Enc *e)
{
	uint32_t numAvail, mainDist, mainLen, numPairs, repIndex, repLen;
	uint32_t *matches = e->matches;
	if (e->additionalOffset == 0)
		mainLen = read_matches(e, &numPairs);
	else {
		mainLen = e->longestMatchLen;
		numPairs = e->numPairs;
	}
	numAvail = e->numAvail;
	e->backRes = kMarkLit;
	if (numAvail < 2)
		return 1;
	if (numAvail > kMatchMax)
		numAvail = kMatchMax;
	const uint8_t *data = mf_cur(e) - 1;
	repLen = repIndex = 0;
	lz_sync();
	LZ_PFOR(q, kNumReps) { // the four first-two-bytes checks at once
		const uint8_t *data2 = data - e->reps[q];
		e->xRepLen[q] = (data[0] == data2[0] && data[1] == data2[1]) ? 2 : 0;
	}
	lz_sync();
	for (uint32_t i = 0; i < kNumReps; i++) {
		if (e->xRepLen[i] == 0)
			continue;
		const uint32_t len = lz_extend(data - e->reps[i], data, 2, numAvail);
		if (len >= e->fb) {
			e->backRes = i;
			move_pos(e, len - 1);
			return len;
		}
		if (len > repLen) {
			repIndex = i;
			repLen = len;
		}
	}
	if (mainLen >= e->fb) {
		e->backRes = matches[numPairs - 1] + kNumReps;
		move_pos(e, mainLen - 1);
		return mainLen;
	}
	mainDist = 0;
	if (mainLen >= 2) {
		mainDist = matches[numPairs - 1];
		while (numPairs > 2) {
			if (mainLen != matches[numPairs - 4] + 1)
				break;
			const uint32_t dist2 = matches[numPairs - 3];
			if (!change_pair(dist2, mainDist))
				break;
			numPairs -= 2;
			mainLen--;
			mainDist = dist2;
		}
		if (mainLen == 2 && mainDist >= 0x80)
			mainLen = 1;
	}
	if (repLen >= 2)
		if (repLen + 1 >= mainLen || (repLen + 2 >= mainLen && mainDist >= (1u << 9)) ||
		    (repLen + 3 >= mainLen && mainDist >= (1u << 15))) {
			e->backRes = repIndex;
			move_pos(e, repLen - 1);
			return repLen;
		}
	if (mainLen < 2 || numAvail <= 2)
		return 1;
	{
		const uint32_t len1 = read_matches(e, &e->numPairs);
		e->longestMatchLen = len1;
		if (len1 >= 2) {
			const uint32_t newDist = matches[e->numPairs - 1];
			if ((len1 >= mainLen && newDist < mainDist) || (len1 == mainLen + 1 && !change_pair(mainDist, newDist)) ||
			    (len1 > mainLen + 1) || (len1 + 1 >= mainLen && mainLen >= 3 && change_pair(newDist, mainDist)))
				return 1;
		}
	}
	data = mf_cur(e) - 1;
	for (uint32_t i = 0; i < kNumReps; i++) {
		const uint8_t *data2 = data - e->reps[i];
		if (data[0] != data2[0] || data[1] != data2[1])
			continue;
		const uint32_t limit = mainLen - 1;
		for (uint32_t len = 2;; len++) {
			if (len >= limit)
				return 1;
			if (data[len] != data2[len])
				break;
		}
	}
	e->backRes = mainDist + kNumReps;
	if (mainLen != 2)
		move_pos(e, mainLen - 2);
	return mainLen;
}

// ---------------------------------------------------------------------------------------------------
// block driver
struct Config {
	uint32_t dictSize, fb, mc, lc, lp, pb;
	uint32_t historySize, cyclicSize, hashMask, bigHash, distTableSize;
	uint32_t fastMode; // levels 1-4: GetOptimumFast over the hc5 finder (LzmaEncProps_Normalize, LzmaEnc.c:68-108)
	uint64_t sonEntries, hash4Entries; // allocation sizes (uint32 counts)
};

// MatchFinder_GetHashMask (LzFind.c:347-373) for numHashBytes == 4
inline uint32_t hash_mask_for(uint32_t hs)
{
	if (hs != 0)
		hs--;
	hs |= hs >> 1;
	hs |= hs >> 2;
	hs |= hs >> 4;
	hs |= hs >> 8;
	hs >>= 1;
	if (hs >= (1u << 24))
		hs >>= 1;
	hs |= (1u << 16) - 1;
	return hs;
}

// LzmaEncProps_Normalize / LzmaEnc_SetProps / LzmaEnc_Alloc / MatchFinder_Create for the parameters
// lzma_compress_buf passes (src/stream.c:450-456): level, dictSize, lc3 lp0 pb2, fb, numThreads 2.
inline bool make_config(int level, uint32_t dictSize, uint32_t fb, uint64_t srcLen, Config &c)
{
	if (level < 1 || level > 9 || srcLen == 0 || srcLen >= 0xFFFF0000ull)
		return false;
	if (fb < 5)
		fb = 5;
	if (fb > kMatchMax)
		fb = kMatchMax;
	c.fastMode = level < 5 ? 1 : 0; // algo 0: hc5, single-threaded finder, mc halved
	c.dictSize = dictSize;
	c.fb = fb;
	c.mc = (16 + (fb >> 1)) >> (c.fastMode ? 1 : 0);
	c.lc = 3;
	c.lp = 0;
	c.pb = 2;
	uint32_t hist = dictSize;
	if (hist == (2u << 30) || hist == (3u << 30))
		hist -= 1;
	c.historySize = hist;
	c.cyclicSize = hist + 1;
	// numHashBytes 5 (hc5): the mask also covers the 18 bits the fifth byte's CRC is shifted into (LzFind.c:340-343)
	const uint32_t extra = c.fastMode ? ((256u << 10) - 1) : 0;
	const uint32_t hs = hash_mask_for(hist) | extra;
	uint32_t cur = hs;
	if (srcLen < hist) {
		cur = hash_mask_for((uint32_t)srcLen) | extra;
		if (cur > hs)
			cur = hs;
	}
	c.hashMask = cur;
	c.bigHash = (!c.fastMode && cur >= 0xFFFFFFu) ? 1 : 0; // LzmaEnc.c:2752 (two-thread match finder)
	uint32_t i;
	for (i = kEndPosModel / 2; i < 32; i++)
		if (dictSize <= (1u << i))
			break;
	c.distTableSize = i * 2;
	c.hash4Entries = (uint64_t)cur + 1;
	// a block shorter than the dictionary never wraps the cyclic buffer: only the first entries are touched
	const uint64_t need = srcLen + 2;
	c.sonEntries = 2 * (need < c.cyclicSize ? need : (uint64_t)c.cyclicSize);
	return true;
}

// LzmaEnc_Init + LzmaEnc_InitPrices (LzmaEnc.c:2769-2852); hash arrays must be zeroed by the caller.
LZ_FN inline void enc_init(Enc *e, const Config &c, const uint8_t *src, uint32_t n, uint8_t *out, uint64_t outCap,
			   uint32_t *hash2, uint32_t *hash3, uint32_t *hash4, uint32_t *son)
{
	e->src = src;
	e->n = n;
	e->fb = c.fb;
	e->mc = c.mc;
	e->historySize = c.historySize;
	e->cyclicSize = c.cyclicSize;
	e->hashMask = c.hashMask;
	e->bigHash = c.bigHash;
	e->distTableSize = c.distTableSize;
	e->fastMode = c.fastMode;
	e->lc = c.lc;
	e->pbMask = (1u << c.pb) - 1;
	e->lpMask = (0x100u << c.lp) - (0x100u >> c.lc);
	e->preRec = nullptr;
	e->prePool = nullptr;
	e->preWait = 0;
	e->prePoolCap = 0;
	e->mfOverflow = nullptr;
	e->gateState = nullptr;
	e->aborted = 0;
	e->lkOn = e->rcOn = 0;
	e->splitOn = 0;
	e->lastB = 0;
	e->pkt.cmd = 0;
	e->lkSlot = -1;
	for (uint32_t i = 0; i < kNumReps; i++)
		e->pubReps[i] = 1;
	e->rcTail = 0;
	e->rcTailPub = e->rcHeadPub = 0;
	e->rcDone = 0;
	e->hash2 = hash2;
	e->hash3 = hash3;
	e->hash4 = hash4;
	e->son = son;
	e->pos = 1;
	e->cycPos = 1;
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t r = i;
		for (int j = 0; j < 8; j++)
			r = (r >> 1) ^ (0xEDB88320u & (0u - (r & 1)));
		e->crc[i] = r;
	}
	e->low = 0;
	e->cacheSize = 0;
	e->range = 0xFFFFFFFFu;
	e->cache = 0;
	e->out = out;
	e->outPos = 0;
	e->outCap = outCap;
	e->overflow = 0;
	e->state = 0;
	for (uint32_t i = 0; i < kNumReps; i++)
		e->reps[i] = 1;
	for (uint32_t i = 0; i < kAlignSize; i++)
		e->posAlign[i] = kProbInit;
	for (uint32_t i = 0; i < kNumStates; i++) {
		for (uint32_t j = 0; j < kNumPosStatesMax; j++) {
			e->isMatch[i][j] = kProbInit;
			e->isRep0Long[i][j] = kProbInit;
		}
		e->isRep[i] = e->isRepG0[i] = e->isRepG1[i] = e->isRepG2[i] = kProbInit;
	}
	for (uint32_t i = 0; i < kNumLenToPos; i++)
		for (uint32_t j = 0; j < 64; j++)
			e->posSlot[i][j] = kProbInit;
	for (uint32_t i = 0; i < kNumFullDist; i++)
		e->posEnc[i] = kProbInit;
	LZ_PFOR(i, 0x300u << (c.lc + c.lp))
		e->lit[i] = kProbInit;
	for (uint32_t i = 0; i < (kNumPosStatesMax << 4); i++)
		e->lenProbs.low[i] = e->repLenProbs.low[i] = kProbInit;
	for (uint32_t i = 0; i < kLenHigh; i++)
		e->lenProbs.high[i] = e->repLenProbs.high[i] = kProbInit;
	e->optEnd = e->optCur = 0;
	// The parse table starts zeroed, like the fresh pages the reference's encoder object is allocated from
	// (defensive: shared memory starts with whatever the previous kernel left there).
	LZ_PFOR(i, kNumOpts) {
		Opt *o = &e->opt[i];
		o->price = kInfinity;
		o->state = 0;
		o->extra = 0;
		o->len = 0;
		o->dist = 0;
		for (uint32_t k = 0; k < kNumReps; k++)
			o->reps[k] = 0;
	}
	lz_sync();
	e->additionalOffset = 0;
	e->longestMatchLen = e->numPairs = e->numAvail = e->backRes = 0;
	init_prob_prices(e->probPrices);
	e->matchPriceCount = 0;
	if (!e->fastMode) { // LzmaEnc_InitPrices (LzmaEnc.c:2833-2850)
		fill_distance_prices(e);
		fill_align_prices(e);
	}
	e->lenPrices.tableSize = e->repLenPrices.tableSize = c.fb + 1 - kMatchMin;
	e->repLenCounter = kRepLenCount;
	len_update_prices(e, &e->lenPrices, 1u << c.pb, &e->lenProbs);
	len_update_prices(e, &e->repLenPrices, 1u << c.pb, &e->repLenProbs);
}

#if defined(__CUDA_ARCH__)
// ---- one symbol, its binary decisions in closed form, one per lane (device, queue mode) ----------------------
// Every decision of a symbol (which probability, which bit) follows from the symbol alone: decision i of a bit
// tree uses the node reached by the i bits above it.  Lane j therefore works out decision j by itself instead of all
// lanes walking the coder's loops; rcw_commit then applies the probability updates and queues the entries.
LZ_INL uint32_t enc_off(const Enc *e, const void *p) { return (uint32_t)(reinterpret_cast<const char *>(p) - reinterpret_cast<const char *>(e)); }

// decision j (0-based) of the 8-bit tree `probs` for symbol s (LitEnc_Encode, LzmaEnc.c:780-798)
LZ_INL void sym_tree8(const Enc *e, const Prob *probs, uint32_t s, uint32_t j, uint32_t &off, uint32_t &v)
{
	off = enc_off(e, probs + ((0x100u | s) >> (8 - j)));
	v = (s >> (7 - j)) & 1u;
}

// number of decisions of LenEnc_Encode for sym = len - 2 (LzmaEnc.c:928-960)
LZ_INL uint32_t sym_len_count(uint32_t sym) { return sym < kLenLow ? 4u : (sym < 2 * kLenLow ? 5u : 10u); }

LZ_INL void sym_len(const Enc *e, const LenProbs *lp, uint32_t sym, uint32_t posState, uint32_t j, uint32_t &off, uint32_t &v)
{
	const Prob *low = lp->low;
	if (sym >= 2 * kLenLow) { // choice 1, choice2 1, eight bits of the high tree
		if (j < 2) {
			off = enc_off(e, low + (j ? kLenLow : 0));
			v = 1;
		} else
			sym_tree8(e, lp->high, sym - 2 * kLenLow, j - 2, off, v);
		return;
	}
	const uint32_t mid = sym >= kLenLow ? 1u : 0u; // choice (1), choice2 0 in front of a 3-bit tree / choice 0
	if (j <= mid) {
		off = enc_off(e, low + (j ? kLenLow : 0));
		v = j < mid ? 1u : 0u;
		return;
	}
	const uint32_t s3 = sym - (mid ? kLenLow : 0), d = j - mid - 1; // depth 0..2
	const Prob *probs = low + (mid ? kLenLow : 0) + (posState << 4);
	off = enc_off(e, probs + ((8u | s3) >> (3 - d)));
	v = (s3 >> (2 - d)) & 1u;
}

// The whole symbol (len, dist as GetOptimum returns them: dist = kMarkLit literal, < kNumReps rep, else distance + 4).
// Leaves lane j's decision in w and the decision count in w.cnt; updates state, reps and the counters exactly as the
// serial coder below does.
LZ_INL void enc_symbol_warp(Enc *e, RcW &w, uint32_t nowPos, uint32_t len, uint32_t dist)
{
	const uint32_t j = lz_lane(), posState = nowPos & e->pbMask, state = e->state;
	uint32_t off = 0, v = 0, cnt;
	if (dist == kMarkLit) {
		const uint8_t *data = mf_cur(e) - e->additionalOffset;
		const uint32_t cur = *data, prevByte = *(data - 1);
		const Prob *probs = lit_probs(e, nowPos, prevByte);
		cnt = 9;
		if (j == 0) {
			off = enc_off(e, &e->isMatch[state][posState]);
			v = 0;
		} else if (j < 9) {
			const uint32_t i = j - 1;
			if (is_lit_state(state))
				sym_tree8(e, probs, cur, i, off, v);
			else { // LitEnc_MatchedEncode (LzmaEnc.c:800-826): the match byte's bit picks the half while the bits above agree
				const uint32_t mb = *(data - e->reps[0]);
				const uint32_t offs = (((mb ^ cur) >> (8 - i)) == 0) ? 0x100u : 0u;
				const uint32_t mbit = ((mb >> (7 - i)) & 1u) << 8;
				off = enc_off(e, probs + (offs + (mbit & offs) + ((0x100u | cur) >> (8 - i))));
				v = (cur >> (7 - i)) & 1u;
			}
		}
		e->state = st_lit(state);
	} else if (dist < kNumReps) {
		// header: isMatch 1, isRep 1, isRepG0 ..., (isRep0Long | isRepG1, isRepG2)
		const uint32_t nh = dist == 0 ? 4u : (dist == 1 ? 4u : 5u);
		const uint32_t nl = len != 1 ? sym_len_count(len - kMatchMin) : 0u;
		cnt = nh + nl;
		if (j == 0) {
			off = enc_off(e, &e->isMatch[state][posState]);
			v = 1;
		} else if (j == 1) {
			off = enc_off(e, &e->isRep[state]);
			v = 1;
		} else if (j == 2) {
			off = enc_off(e, &e->isRepG0[state]);
			v = dist == 0 ? 0u : 1u;
		} else if (j == 3) {
			if (dist == 0) {
				off = enc_off(e, &e->isRep0Long[state][posState]);
				v = len != 1 ? 1u : 0u;
			} else {
				off = enc_off(e, &e->isRepG1[state]);
				v = dist == 1 ? 0u : 1u;
			}
		} else if (j == 4 && dist >= 2) {
			off = enc_off(e, &e->isRepG2[state]);
			v = dist - 2;
		} else if (j < cnt)
			sym_len(e, &e->repLenProbs, len - kMatchMin, posState, j - nh, off, v);
		if (dist != 0) {
			const uint32_t r0 = e->reps[0], r1 = e->reps[1], r2 = e->reps[2], r3 = e->reps[3];
			const uint32_t d = dist == 1 ? r1 : (dist == 2 ? r2 : r3);
			if (dist == 3)
				e->reps[3] = r2;
			if (dist >= 2)
				e->reps[2] = r1;
			e->reps[1] = r0;
			e->reps[0] = d;
		}
		if (len == 1)
			e->state = st_shortrep(state);
		else {
			--e->repLenCounter;
			e->state = st_rep(state);
		}
	} else {
		const uint32_t d = dist - kNumReps, slot = pos_slot(d), nl = sym_len_count(len - kMatchMin);
		const uint32_t footer = (slot >> 1) - 1; // meaningful for d >= kStartPosModel
		const uint32_t nf = d < kStartPosModel ? 0u : (d < kNumFullDist ? footer : 1u + kNumAlignBits);
		const uint32_t slotBase = 2 + nl, footBase = slotBase + 6;
		cnt = footBase + nf;
		if (j == 0) {
			off = enc_off(e, &e->isMatch[state][posState]);
			v = 1;
		} else if (j == 1) {
			off = enc_off(e, &e->isRep[state]);
			v = 0;
		} else if (j < slotBase)
			sym_len(e, &e->lenProbs, len - kMatchMin, posState, j - 2, off, v);
		else if (j < footBase) { // the slot's six bits, most significant first
			const uint32_t i = j - slotBase;
			const Prob *probs = e->posSlot[len < kNumLenToPos + 1 ? len - 2 : kNumLenToPos - 1];
			off = enc_off(e, probs + ((64u | slot) >> (6 - i)));
			v = (slot >> (5 - i)) & 1u;
		} else if (j < cnt) {
			uint32_t i = j - footBase;
			const Prob *probs;
			uint32_t sym = d;
			bool direct = false;
			if (d < kNumFullDist) // reverse bit tree over the footer bits
				probs = e->posEnc + ((2u | (slot & 1u)) << footer);
			else if (i == 0) {
				direct = true;
				probs = nullptr;
			} else { // the four align bits, reverse tree
				i--;
				probs = e->posAlign;
				sym = d & kAlignMask;
			}
			if (direct) {
				off = 0;
				v = kRcDirect | ((footer - kNumAlignBits) << 26) | ((d & ((1u << footer) - 1)) >> kNumAlignBits);
			} else {
				const uint32_t m = (1u << i) | (i ? (__brev(sym) >> (32 - i)) : 0u); // 1, then the i lower bits, lowest first
				off = enc_off(e, probs + m);
				v = (sym >> i) & 1u;
			}
		}
		e->state = st_match(state);
		e->reps[3] = e->reps[2];
		e->reps[2] = e->reps[1];
		e->reps[1] = e->reps[0];
		e->reps[0] = d + 1;
		e->matchPriceCount++;
	}
	w.off = off;
	w.v = v;
	w.cnt = cnt;
}
#endif

// LzmaEnc_CodeOneBlock (LzmaEnc.c:2383-2680) run to the end of the block, then Flush (:2191-2200).
// Returns the number of output bytes (valid when !e->overflow).
LZ_FN inline uint64_t enc_run(Enc *e)
{
	uint32_t nowPos = 0;
	RcW w;
	w.off = w.v = w.cnt = 0;
	w.on = false;
#if defined(__CUDA_ARCH__)
	w.on = e->rcOn != 0;
#endif
	if (e->n == 0) {
		for (int i = 0; i < 5; i++)
			rc_shift_low(e);
		return e->outPos;
	}
	{
		uint32_t numPairs;
		read_matches(e, &numPairs);
		rc_bit(e, w, &e->isMatch[0][0], 0);
		const uint8_t curByte = *(mf_cur(e) - e->additionalOffset);
		lit_encode(e, w, e->lit, curByte);
		rcw_commit(e, w);
		e->additionalOffset--;
		nowPos++;
	}
	if (mf_avail(e) != 0) {
		for (;;) {
			uint32_t len;
			{
				const uint32_t oci = e->optCur;
				if (e->fastMode)
					len = get_optimum_fast(e);
				else if (e->optEnd == oci)
					len = get_optimum(e, nowPos);
				else {
					const Opt *o = &e->opt[oci];
					len = o->len;
					e->backRes = o->dist;
					e->optCur = oci + 1;
				}
			}
			const uint32_t posState = nowPos & e->pbMask;
			uint32_t dist = e->backRes;
			Prob *pm = &e->isMatch[e->state][posState];
#if defined(__CUDA_ARCH__)
			if (w.on)
				rcq_reserve(e);
#endif
			LZ_T(17);
#if defined(__CUDA_ARCH__)
			if (w.on)
				enc_symbol_warp(e, w, nowPos, len, dist);
			else
#endif
			if (dist == kMarkLit) {
				rc_bit(e, w, pm, 0);
				const uint8_t *data = mf_cur(e) - e->additionalOffset;
				Prob *probs = (Prob *)lit_probs(e, nowPos, *(data - 1));
				const uint32_t state = e->state;
				e->state = st_lit(state);
				if (is_lit_state(state))
					lit_encode(e, w, probs, *data);
				else
					lit_encode_matched(e, w, probs, *data, *(data - e->reps[0]));
			} else {
				rc_bit(e, w, pm, 1);
				if (dist < kNumReps) {
					rc_bit(e, w, &e->isRep[e->state], 1);
					if (dist == 0) {
						rc_bit(e, w, &e->isRepG0[e->state], 0);
						rc_bit(e, w, &e->isRep0Long[e->state][posState], len != 1 ? 1 : 0);
						if (len == 1)
							e->state = st_shortrep(e->state);
					} else {
						rc_bit(e, w, &e->isRepG0[e->state], 1);
						if (dist == 1) {
							rc_bit(e, w, &e->isRepG1[e->state], 0);
							dist = e->reps[1];
						} else {
							rc_bit(e, w, &e->isRepG1[e->state], 1);
							rc_bit(e, w, &e->isRepG2[e->state], dist - 2);
							if (dist == 2)
								dist = e->reps[2];
							else {
								dist = e->reps[3];
								e->reps[3] = e->reps[2];
							}
							e->reps[2] = e->reps[1];
						}
						e->reps[1] = e->reps[0];
						e->reps[0] = dist;
					}
					if (len != 1) {
						len_encode(e, w, &e->repLenProbs, len - kMatchMin, posState);
						--e->repLenCounter;
						e->state = st_rep(e->state);
					}
				} else {
					rc_bit(e, w, &e->isRep[e->state], 0);
					e->state = st_match(e->state);
					len_encode(e, w, &e->lenProbs, len - kMatchMin, posState);
					dist -= kNumReps;
					e->reps[3] = e->reps[2];
					e->reps[2] = e->reps[1];
					e->reps[1] = e->reps[0];
					e->reps[0] = dist + 1;
					e->matchPriceCount++;
					const uint32_t slot = pos_slot(dist);
					{
						uint32_t sym = slot + 64;
						Prob *probs = e->posSlot[len < kNumLenToPos + 1 ? len - 2 : kNumLenToPos - 1];
						do {
							Prob *prob = probs + (sym >> 6);
							const uint32_t bit = (sym >> 5) & 1;
							sym <<= 1;
							rc_bit(e, w, prob, bit);
						} while (sym < (1u << 12));
					}
					if (dist >= kStartPosModel) {
						const uint32_t footer = (slot >> 1) - 1;
						if (dist < kNumFullDist) {
							const uint32_t base = (2 | (slot & 1)) << footer;
							rc_reverse(e, w, e->posEnc + base, footer, dist);
						} else {
							rc_direct(e, w, (dist & ((1u << footer) - 1)) >> kNumAlignBits, footer - kNumAlignBits);
							rc_reverse(e, w, e->posAlign, kNumAlignBits, dist & kAlignMask);
						}
					}
				}
			}
			LZ_T(18);
			rcw_commit(e, w);
			LZ_T(12);
			nowPos += len;
			e->additionalOffset -= len;
			if (e->additionalOffset == 0) {
				if (!e->fastMode && e->matchPriceCount >= 64) {
					fill_align_prices(e);
					fill_distance_prices(e);
					len_update_prices(e, &e->lenPrices, e->pbMask + 1, &e->lenProbs);
				}
				if (!e->fastMode && e->repLenCounter <= 0) {
					e->repLenCounter = kRepLenCount;
					len_update_prices(e, &e->repLenPrices, e->pbMask + 1, &e->repLenProbs);
				}
				LZ_T(13);
				if (mf_avail(e) == 0)
					break;
				if (e->gateState && *(const volatile int *)e->gateState == 2) { // the same word for every lane
					e->aborted = 1;
					break;
				}
				if (e->mfOverflow && *(const volatile int *)e->mfOverflow) { // the match finder ran out of pool under us
					e->aborted = 2;
					break;
				}
			}
		}
	}
#if defined(__CUDA_ARCH__)
	if (w.on) { // the coder thread flushes (RangeEnc_FlushData) and reports the length
		rcq_reserve(e);
		rcq_push(e, kRcFlush);
		rcq_publish(e);
		while (*(volatile int *)&e->rcDone == 0)
			__nanosleep(200);
		__threadfence_block();
		return *(volatile uint64_t *)&e->outPos;
	}
#endif
	for (int i = 0; i < 5; i++) // RangeEnc_FlushData; no end marker (writeEndMark = 0)
		rc_shift_low(e);
	return e->outPos;
}

} // namespace lzma
} // namespace lrz
