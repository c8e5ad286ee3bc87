// lzma_mf.cuh -- the LZMA match finder as a data-parallel pre-pass (host/device shared pieces).
//
// Observation the design rests on: everything the 7-Zip bt4 match finder hands to the encoder is a
// pure function of the block's bytes, not of the encoder's decisions -- GetMatches and Skip update
// the hash heads and the binary trees identically (LzFind.c:1219-1285, 1554-1570; in the two-thread
// finder the BT thread visits every position regardless, LzFindMt.c:571-729).  And the structure is
// separable: there is one binary tree per 4-byte-hash bucket, a position is a node of exactly one
// tree, and a walk for position p only touches nodes of p's own tree.  So
//
//   * the positions of one bucket, taken in increasing order, can be inserted by ONE thread with no
//     knowledge of the other buckets -- all buckets of all blocks run concurrently;
//   * hash4[hv] at the time position p is inserted is simply the previous position of p's bucket;
//   * hash2 / hash3 heads (MixMatches3, LzFindMt.c:1093-1131) at position p are the previous
//     positions with the same 2- / 3-byte hash.
//
// "Previous position with the same hash" for all positions at once is a stable sort by hash value
// (lzma_mf.cu: LSD radix sort, 8-bit digits); neighbours in the sorted order are predecessor and
// successor inside the bucket.
//
// The tree array is indexed by position instead of cyclically: the reference overwrites slot
// p % cyclicBufferSize at time p + cyclicBufferSize, but no walk at an earlier time may be disturbed
// by that (buckets are not processed in global time order here), and no walk at a later time reads a
// node that old (cmCheck, LzFind.c:975-985) -- so a flat array of 2*(n+2) entries is equivalent.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MF_FN __host__ __device__
#define MF_INL __host__ __device__ __forceinline__
#else
#define MF_FN
#define MF_INL inline
#endif

namespace lrz {
namespace lzma {

constexpr uint32_t kMfCountBits = 10; // per-position record = (pool offset << 10) | number of uint32 (<= 2*273+2)
constexpr uint32_t kMfCountMask = (1u << kMfCountBits) - 1;
constexpr uint64_t kMfReady = 1ull << 63; // set by the tree walk when a record is final (readers mask it off)
// Before the walk reaches a position its record holds what a data-parallel pre-pass knows about it: the length of the
// common prefix with the previous position of its bucket (the first candidate the tree insertion compares with), so
// that a run of identical bytes -- one bucket, every insertion an immediate full-length hit -- costs the bucket's
// thread no byte comparison at all.
constexpr uint32_t kMfFirstValid = 1u << 16, kMfFirstNone = 0xFFFFFFFFu;

struct MfParams {
	uint32_t n;           // block length
	uint32_t fb, mc;      // numFastBytes (lenLimit), cutValue
	uint32_t hashMask, bigHash;
	uint32_t historySize, cyclicSize;
	uint32_t hc5;         // levels 1-4: hash-chain finder over 5-byte hashes (single-threaded in the reference)
};

MF_INL uint32_t mf_crc_entry(uint32_t i) // g_CrcTable[i] (7zCrc.c), reflected 0xEDB88320
{
	uint32_t r = i;
	for (int j = 0; j < 8; j++)
		r = (r >> 1) ^ (0xEDB88320u & (0u - (r & 1)));
	return r;
}

// HASH4_CALC (LzFind.c:49) / GetHeads4b (LzFindMt.c:386-394)
MF_INL uint32_t mf_hash4(const uint32_t *crc, const uint8_t *cur, uint32_t hashMask, uint32_t bigHash)
{
	if (bigHash)
		return (crc[cur[0]] & hashMask) ^ ((uint32_t)cur[1] | ((uint32_t)cur[2] << 8) | ((uint32_t)cur[3] << 16));
	return (crc[cur[0]] ^ cur[1] ^ ((uint32_t)cur[2] << 8) ^ (crc[cur[3]] << 5)) & hashMask;
}
MF_INL uint32_t mf_hash2(const uint32_t *crc, const uint8_t *cur) { return (crc[cur[0]] ^ cur[1]) & 0x3FFu; }
MF_INL uint32_t mf_hash3(const uint32_t *crc, const uint8_t *cur)
{
	return (crc[cur[0]] ^ cur[1] ^ ((uint32_t)cur[2] << 8)) & 0xFFFFu;
}

// HASH5_CALC (LzFind.c:56-63): kLzHash_CrcShift_1 = 5, kLzHash_CrcShift_2 = 10
MF_INL uint32_t mf_hash5(const uint32_t *crc, const uint8_t *cur, uint32_t hashMask)
{
	return (crc[cur[0]] ^ cur[1] ^ ((uint32_t)cur[2] << 8) ^ (crc[cur[3]] << 5) ^ (crc[cur[4]] << 10)) & hashMask;
}

// Length of the common prefix of a[len..) and b[len..), capped at limit (a = earlier copy, b = cur).
MF_INL uint32_t mf_extend(const uint8_t *a, const uint8_t *b, uint32_t len, uint32_t limit)
{
	while (len != limit && a[len] == b[len])
		len++;
	return len;
}

// One insertion of position `pos` (1-based: byte index pos-1) into its bucket's tree, with curMatch =
// the previous position of the bucket (0 = none): GetMatchesSpec1 (LzFind.c:962-1029) with maxLen = 3
// as the BT thread calls it (LzFindMt.c:627-700).  Writes (len, dist-1) pairs with strictly increasing
// len >= 4 to d and returns the number of uint32 written.  son is flat: node p at son[2p], son[2p+1].
MF_FN inline uint32_t mf_bt_insert(const uint8_t *src, const MfParams &P, uint32_t *son, uint32_t pos, uint32_t curMatch,
				   uint32_t *d, uint32_t firstLen = kMfFirstNone)
{
	const uint8_t *cur = src + (pos - 1);
	const uint32_t avail = P.n - (pos - 1);
	const uint32_t lenLimit = avail < P.fb ? avail : P.fb;
	uint32_t *ptr0 = son + ((size_t)pos << 1) + 1, *ptr1 = son + ((size_t)pos << 1);
	uint32_t len0 = 0, len1 = 0, maxLen = 3, cut = P.mc, nd = 0;
	const uint32_t cmCheck = pos <= P.cyclicSize ? 0 : pos - P.cyclicSize;
	if (cmCheck < curMatch) {
		do {
			const uint32_t delta = pos - curMatch;
			uint32_t *pair = son + ((size_t)curMatch << 1);
			const uint8_t *pb = cur - delta;
			uint32_t len = len0 < len1 ? len0 : len1;
			const uint32_t pair0 = pair[0], pair1 = pair[1];
			bool hit;
			if (firstLen != kMfFirstNone) { // first candidate: its common prefix (from 0, capped at lenLimit) is known
				hit = firstLen != 0;
				if (hit)
					len = firstLen;
				firstLen = kMfFirstNone;
			} else {
				hit = pb[len] == cur[len];
				if (hit)
					len = mf_extend(pb, cur, len + 1, lenLimit);
			}
			if (hit) {
				if (maxLen < len) {
					maxLen = len;
					d[nd++] = len;
					d[nd++] = delta - 1;
					if (len == lenLimit) {
						*ptr1 = pair0;
						*ptr0 = pair1;
						return nd;
					}
				}
			}
			if (pb[len] < cur[len]) {
				*ptr1 = curMatch;
				curMatch = pair1;
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

// MatchFinderMt_GetMatches + MixMatches3 (LzFindMt.c:1274-1317, 1093-1131) for position pos (1-based,
// with at least 4 bytes available): the 2- / 3-byte candidates c2 / c3 (previous positions with the same
// hash, 1-based, 0 = none) go in front of the tree's pairs when nearer than the first tree match.
// Tree pairs are expected at d + 4 (nbt uint32); the result is compacted to the front of d.
MF_FN inline uint32_t mf_mix(const uint8_t *src, const MfParams &P, uint32_t pos, uint32_t c2, uint32_t c3, uint32_t *d,
			     uint32_t nbt)
{
	const uint8_t *cur = src + (pos - 1);
	const uint32_t *bt = d + 4;
	const uint32_t minPos = nbt ? pos - bt[1] : (pos > P.historySize ? pos - P.historySize : 1);
	uint32_t nd = 0;
	bool done = false;
	if (c2 >= minPos && cur[(ptrdiff_t)c2 - (ptrdiff_t)pos] == cur[0]) {
		d[1] = pos - c2 - 1;
		if (cur[(ptrdiff_t)c2 - (ptrdiff_t)pos + 2] == cur[2]) {
			d[0] = 3;
			done = true;
		} else
			d[0] = 2;
		nd = 2;
	}
	if (!done && c3 >= minPos && cur[(ptrdiff_t)c3 - (ptrdiff_t)pos] == cur[0]) {
		d[nd++] = 3;
		d[nd++] = pos - c3 - 1;
	}
	if (nd != 4)
		for (uint32_t i = 0; i < nbt; i++)
			d[nd + i] = bt[i];
	return nd + nbt;
}

// ---- levels 1-4: Hc5_MatchFinder_GetMatches + Hc_GetMatchesSpec (LzFind.c:1431-1500, 880-960) -------------
// The hash-chain finder is data-parallel outright: son[p] is just the previous position with the same 5-byte
// hash, the 2- / 3-byte heads are the previous positions with the same 2- / 3-byte hash, GetMatches and Skip
// update all of them identically, and a position's list only READS the chain.  link[q] = previous position
// (1-based, 0 = none) with the same 5-byte hash as position q, for positions with at least 5 bytes available.
// The reference's cyclic son[] forgets links older than cyclicBufferSize; the delta test below does the same.
MF_FN inline uint32_t mf_hc5_matches(const uint8_t *src, const MfParams &P, const uint32_t *link, uint32_t pos, uint32_t c2,
				     uint32_t c3, uint32_t *d)
{
	const uint8_t *cur = src + (pos - 1);
	const uint32_t avail = P.n - (pos - 1);
	const uint32_t lenLimit = avail < P.fb ? avail : P.fb;
	const uint32_t mmm = pos < P.cyclicSize ? pos : P.cyclicSize; // SET_mmm
	uint32_t d2 = pos - c2, d3 = pos - c3; // an empty head (0) gives d == pos, which fails d < mmm
	uint32_t curMatch = link[pos];
	uint32_t nd = 0, maxLen = 4;
	for (;;) {
		if (d2 < mmm && *(cur - d2) == *cur) {
			d[nd] = 2;
			d[nd + 1] = d2 - 1;
			nd += 2;
			if (*(cur - d2 + 2) == cur[2]) {
			} else if (d3 < mmm && *(cur - d3) == *cur) {
				d[nd + 1] = d3 - 1;
				nd += 2;
				d2 = d3;
			} else
				break;
		} else if (d3 < mmm && *(cur - d3) == *cur) {
			d[nd + 1] = d3 - 1;
			nd += 2;
			d2 = d3;
		} else
			break;
		d[nd - 2] = 3;
		if (*(cur - d2 + 3) != cur[3])
			break;
		maxLen = mf_extend(cur - d2, cur, maxLen, lenLimit); // UPDATE_maxLen
		d[nd - 2] = maxLen;
		if (maxLen == lenLimit)
			return nd; // son[] gets its link all the same; no chain walk
		break;
	}
	// Hc_GetMatchesSpec
	uint32_t cut = P.mc;
	do {
		if (curMatch == 0)
			break;
		const uint32_t delta = pos - curMatch;
		if (delta >= P.cyclicSize)
			break;
		const uint8_t *pb = cur - delta;
		curMatch = link[curMatch];
		if (cur[maxLen] == pb[maxLen]) {
			uint32_t len = 0;
			while (cur[len] == pb[len]) {
				if (++len == lenLimit) {
					d[nd] = lenLimit;
					d[nd + 1] = delta - 1;
					return nd + 2;
				}
			}
			if (maxLen < len) {
				maxLen = len;
				d[nd] = len;
				d[nd + 1] = delta - 1;
				nd += 2;
			}
		}
	} while (--cut);
	return nd;
}

} // namespace lzma
} // namespace lrz
