// lz4_size.cuh -- size-only restatement of LZ4_compress_default() (liblz4 1.9.4, fast mode,
// acceleration 1) for lrzip-next's compressibility gate lz4_compresses() (src/stream.c:2325-2380).
//
// liblz4 is not vendored by the reference (system library; the oracle box has 1.9.4).  The gate only
// looks at the RETURN VALUE of LZ4_compress_default(src, dst, in_len, in_len + 1): the compressed
// size, or 0 when the output would not fit.  So nothing is written here; the sequence of matches the
// published algorithm finds is replayed and only the output cursor is advanced:
//
//   * one 4096-entry (byU32, inputs >= 64 KiB + 11) or 8192-entry (byU16) position table, zeroed,
//     multiplicative hash of 5 (resp. 4) bytes; a probe is a candidate when it is within 65535 bytes
//     and its first 4 bytes are equal;
//   * skip acceleration: step = (searchMatchNb++ >> 6) starting at 1 << 6;
//   * backward "catch up", literal-run and match-length token arithmetic, the immediate re-test at
//     the end of a match, the 12-byte match-finding limit and 5 last literals;
//   * the limitedOutput checks, which can reject a block a few bytes before it is really full.
//
// Compiled for the device (the gate kernel, one block per thread group) and for the host
// (tests/hostsim checks it against liblz4.so.1 on the oracle box).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LZ4S_FN __host__ __device__
#else
#define LZ4S_FN
#endif

namespace lrz {
namespace lz4s {

constexpr int kMinMatch = 4, kMfLimit = 12, kLastLiterals = 5, kMinLength = kMfLimit + 1;
constexpr uint32_t kDistMax = 65535;
constexpr int kHashLog = 12;
constexpr int k64KLimit = 65536 + (kMfLimit - 1);
constexpr uint32_t kRunMask = 15, kMlMask = 15;

LZ4S_FN inline uint32_t rd32(const uint8_t *p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

LZ4S_FN inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

LZ4S_FN inline uint32_t hash_pos(const uint8_t *p, bool u16)
{
	if (u16)
		return (rd32(p) * 2654435761u) >> (32 - (kHashLog + 1));
	return (uint32_t)(((rd64(p) << 24) * 889523592379ull) >> (64 - kHashLog));
}

// `table` must hold 4096 uint32 (it is used as 8192 uint16 for small inputs) and be zeroed.
// Returns what LZ4_compress_default(src, dst, n, cap) returns.
LZ4S_FN inline int compress_size(const uint8_t *src, int n, int cap, uint32_t *table)
{
	if (n < 0 || (uint32_t)n > 0x7E000000u)
		return 0;
	const bool u16 = n < k64KLimit;
	uint16_t *t16 = reinterpret_cast<uint16_t *>(table);
	const bool limited = cap < n + n / 255 + 16; // LZ4_compressBound
	int64_t op = 0;
	const int64_t olimit = cap;
	int ip = 0, anchor = 0;
	const int iend = n, mflimitPlusOne = iend - kMfLimit + 1, matchlimit = iend - kLastLiterals;
#define LZ4S_GET(h) (u16 ? (uint32_t)t16[h] : table[h])
#define LZ4S_PUT(h, v)                       \
	do {                                 \
		if (u16)                     \
			t16[h] = (uint16_t)(v); \
		else                         \
			table[h] = (uint32_t)(v); \
	} while (0)
	if (n >= kMinLength) {
		LZ4S_PUT(hash_pos(src, u16), 0);
		ip++;
		uint32_t forwardH = hash_pos(src + ip, u16);
		for (;;) {
			int match;
			int64_t token_lit; // literal length of the token being built (kept for symmetry with the format)
			{
				int forwardIp = ip, step = 1, searchMatchNb = 1 << 6;
				for (;;) {
					const uint32_t h = forwardH;
					const uint32_t current = (uint32_t)forwardIp;
					const uint32_t matchIndex = LZ4S_GET(h);
					ip = forwardIp;
					forwardIp += step;
					step = searchMatchNb++ >> 6;
					if (forwardIp > mflimitPlusOne)
						goto last_literals;
					match = (int)matchIndex;
					forwardH = hash_pos(src + forwardIp, u16);
					LZ4S_PUT(h, current);
					if (!u16 && matchIndex + kDistMax < current)
						continue; // too far
					if (rd32(src + match) == rd32(src + ip))
						break;
				}
			}
			while (ip > anchor && match > 0 && src[ip - 1] == src[match - 1]) { // catch up
				ip--;
				match--;
			}
			{
				const uint32_t litLength = (uint32_t)(ip - anchor);
				op++; // token
				if (limited && op + litLength + (2 + 1 + kLastLiterals) + (litLength / 255) > olimit)
					return 0;
				if (litLength >= kRunMask)
					op += (litLength - kRunMask) / 255 + 1;
				op += litLength;
				token_lit = litLength;
				(void)token_lit;
			}
		next_match:
			op += 2; // offset
			{
				// LZ4_count(ip + MINMATCH, match + MINMATCH, matchlimit)
				int a = ip + kMinMatch, b = match + kMinMatch;
				while (a < matchlimit && src[a] == src[b]) {
					a++;
					b++;
				}
				uint32_t matchCode = (uint32_t)(a - (ip + kMinMatch));
				ip += (int)matchCode + kMinMatch;
				if (limited && op + (1 + kLastLiterals) + (matchCode + 240) / 255 > olimit)
					return 0;
				if (matchCode >= kMlMask) {
					matchCode -= kMlMask;
					op += matchCode / 255 + 1;
				}
			}
			anchor = ip;
			if (ip >= mflimitPlusOne)
				break;
			LZ4S_PUT(hash_pos(src + ip - 2, u16), ip - 2);
			{
				const uint32_t h = hash_pos(src + ip, u16);
				const uint32_t current = (uint32_t)ip;
				const uint32_t matchIndex = LZ4S_GET(h);
				match = (int)matchIndex;
				LZ4S_PUT(h, current);
				if ((u16 || matchIndex + kDistMax >= current) && rd32(src + match) == rd32(src + ip)) {
					op++; // token with zero literals
					goto next_match;
				}
			}
			forwardH = hash_pos(src + ++ip, u16);
		}
	}
last_literals: {
	const int64_t lastRun = iend - anchor;
	if (limited && op + lastRun + 1 + ((lastRun + 255 - kRunMask) / 255) > olimit)
		return 0;
	if (lastRun >= kRunMask)
		op += 1 + (lastRun - kRunMask) / 255 + 1;
	else
		op += 1;
	op += lastRun;
}
#undef LZ4S_GET
#undef LZ4S_PUT
	return (int)op;
}

// lz4_compresses() (src/stream.c:2325-2380): progressively larger prefixes of the block are tested
// until one is compressible enough.  `scratch` = 4096 uint32.  Returns the function's return value
// (0 = incompressible => the block is stored).
LZ4S_FN inline int gate(const uint8_t *buf, int64_t s_len, int threshold, uint32_t *scratch)
{
	int64_t test_len = s_len;
	int in_len = (int)(test_len < 100ll * 1048576 ? test_len : 100ll * 1048576);
	int buftest_size = in_len;
	double pct = 101;
	while (test_len > 0) {
		for (int i = 0; i < 4096; i++)
			scratch[i] = 0;
		const int ret = compress_size(buf, in_len, in_len + 1, scratch);
		if (ret > 0) {
			pct = 100 * ((double)ret / (double)in_len);
			if (ret < in_len * ((double)threshold / 100))
				break;
		}
		test_len -= in_len;
		if (test_len > 0) {
			buftest_size += in_len;
			if (buftest_size < 10 * 1048576)
				buftest_size <<= 1;
			in_len = (int)(test_len < buftest_size ? test_len : buftest_size);
		}
	}
	return (int)(pct > threshold ? 0 : (pct < 1 ? pct + 1 : pct));
}

} // namespace lz4s
} // namespace lrz
