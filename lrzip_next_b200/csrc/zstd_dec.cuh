// zstd_dec.cuh -- Zstandard (RFC 8878) frame decoder for the decode path: the payload of a
// CTYPE_ZSTD stream block is one frame (src/stream.c:167-229 writes it with ZSTD_compress, :1989-2010 reads it back with
// ZSTD_decompress).  Written from the format specification: frame header, Raw / RLE / Compressed blocks, literals
// section (Raw, RLE, Huffman with one or four streams, tree description direct or FSE-compressed, Treeless), sequences
// section (Predefined / RLE / FSE-compressed / Repeat tables, backward bit stream, repeat offsets), sequence execution.
// The content checksum (XXH64) is skipped, not verified: the container has its own CRC-32 per chunk and MD5 per file.
// Dictionaries are not supported (the reference never uses one).
//
// A frame is decoded by ONE thread (blocks of a frame share history, repeat offsets and tables); the frames of a
// chunk -- one per stream block -- are decoded side by side.  The same source is compiled for the device (product) and
// for the host (tests/hostsim checks it against frames made by the system's libzstd at several levels).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ZD_FN __host__ __device__
#else
#define ZD_FN
#endif

namespace lrz {
namespace zd {

constexpr int kBlockMax = 128 * 1024;
constexpr int kMaxLLLog = 9, kMaxOFLog = 8, kMaxMLLog = 9, kMaxHufBits = 11, kMaxWeightLog = 6;

enum Err {
	kErrTruncated = -1, kErrMagic = -2, kErrHeader = -3, kErrBlock = -4, kErrLiterals = -5, kErrHuffman = -6,
	kErrFse = -7, kErrSequences = -8, kErrOutput = -9, kErrDict = -10,
};

struct FseEntry {
	uint8_t sym, nbits;
	uint16_t base;
};

struct FseTable {
	FseEntry e[512];
	int log;   // accuracy log; -1 = not set (a Repeat mode has nothing to repeat)
};

// Per-frame scratch (global memory on the device: one per decode job).
struct Work {
	FseTable ll, of, ml;
	uint16_t huf[1 << kMaxHufBits]; // sym | nbits << 8, indexed by the next huf_bits bits of the stream
	int huf_bits;                   // 0 = no Huffman table yet (Treeless needs one)
	uint8_t weights[260];
	int16_t freq[256];
	uint16_t sdesc[256];
	FseEntry wtab[1 << kMaxWeightLog];
	uint8_t lit[kBlockMax + 64];
};

ZD_FN inline int highbit(uint32_t v) // position of the highest set bit, v > 0
{
	int r = 0;
	while (v >>= 1)
		r++;
	return r;
}

// ---- forward bit reader (FSE table descriptions): LSB first
struct FwdBits {
	const uint8_t *p;
	int64_t len, bit; // bit = next bit to read
};
ZD_FN inline uint32_t fwd_read(FwdBits &b, int n)
{
	uint32_t v = 0;
	for (int i = 0; i < n; i++) {
		const int64_t at = b.bit + i;
		if ((at >> 3) < b.len)
			v |= (uint32_t)((b.p[at >> 3] >> (at & 7)) & 1) << i;
	}
	b.bit += n;
	return v;
}

// ---- backward bit reader (Huffman and FSE streams): the stream ends with a 1 bit followed by zero padding and is read
// from there towards its first byte; reading past the beginning yields zeros (`bit` goes negative)
struct BackBits {
	const uint8_t *p;
	int64_t bit; // number of unread bits
};
ZD_FN inline bool back_init(BackBits &b, const uint8_t *p, int64_t len)
{
	if (len < 1 || p[len - 1] == 0)
		return false;
	b.p = p;
	b.bit = len * 8 - (8 - highbit(p[len - 1]));
	return true;
}
ZD_FN inline uint32_t back_read(BackBits &b, int n) // n <= 32; the bits come out as a number, first-read bit highest
{
	if (n == 0)
		return 0;
	b.bit -= n;
	int64_t lo = b.bit; // lowest bit index of the field
	int take = n, shift = 0;
	if (lo < 0) {       // the part below the stream's first bit reads as zero
		shift = (int)(-lo > n ? n : -lo);
		take = n - shift;
		lo = 0;
	}
	if (take <= 0)
		return 0;
	uint64_t w = 0;
	const int64_t byte = lo >> 3;
	const int nb = (int)(((lo & 7) + take + 7) >> 3);
	for (int i = 0; i < nb; i++)
		w |= (uint64_t)b.p[byte + i] << (8 * i);
	const uint32_t v = (uint32_t)((w >> (lo & 7)) & ((1ull << take) - 1));
	return v << shift;
}

// ---- FSE decoding table from normalised frequencies (RFC 8878 4.1.1); -1 = "less than one"
ZD_FN inline bool fse_build(const int16_t *freq, int nsym, int log, FseEntry *tab, uint16_t *sdesc)
{
	const int size = 1 << log;
	int high = size;
	for (int s = 0; s < nsym; s++)
		if (freq[s] == -1) {
			tab[--high].sym = (uint8_t)s;
			sdesc[s] = 1;
		}
	const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
	int pos = 0;
	for (int s = 0; s < nsym; s++) {
		if (freq[s] <= 0)
			continue;
		sdesc[s] = (uint16_t)freq[s];
		for (int i = 0; i < freq[s]; i++) {
			tab[pos].sym = (uint8_t)s;
			do
				pos = (pos + step) & mask;
			while (pos >= high);
		}
	}
	if (pos != 0)
		return false;
	for (int i = 0; i < size; i++) {
		const int s = tab[i].sym;
		const uint32_t next = sdesc[s]++;
		const int nb = log - highbit(next);
		tab[i].nbits = (uint8_t)nb;
		tab[i].base = (uint16_t)((next << nb) - size);
	}
	return true;
}

// FSE table description (RFC 8878 4.1.1): returns the bytes consumed, < 0 on error
ZD_FN inline int64_t fse_read_description(const uint8_t *src, int64_t len, int max_log, int max_sym, int16_t *freq,
					  int &nsym, int &log)
{
	if (len < 1)
		return kErrTruncated;
	FwdBits b{ src, len, 0 };
	log = 5 + (int)fwd_read(b, 4);
	if (log > max_log)
		return kErrFse;
	int remaining = 1 << log, s = 0;
	while (remaining > 0 && s <= max_sym) {
		const int bits = highbit((uint32_t)remaining + 1) + 1;
		uint32_t val = fwd_read(b, bits);
		const uint32_t lower = (1u << (bits - 1)) - 1;
		const uint32_t threshold = (1u << bits) - 1 - ((uint32_t)remaining + 1);
		if ((val & lower) < threshold) {
			b.bit--;
			val &= lower;
		} else if (val > lower)
			val -= threshold;
		const int proba = (int)val - 1;
		remaining -= proba < 0 ? -proba : proba;
		freq[s++] = (int16_t)proba;
		if (proba == 0) {
			uint32_t rep = fwd_read(b, 2);
			for (;;) {
				for (uint32_t i = 0; i < rep && s <= max_sym; i++)
					freq[s++] = 0;
				if (rep == 3)
					rep = fwd_read(b, 2);
				else
					break;
			}
		}
		if ((b.bit + 7) / 8 > len)
			return kErrTruncated;
	}
	if (remaining != 0 || s > max_sym + 1)
		return kErrFse;
	nsym = s;
	return (b.bit + 7) / 8;
}

ZD_FN inline void fse_set_rle(FseTable &t, uint8_t sym)
{
	t.e[0].sym = sym;
	t.e[0].nbits = 0;
	t.e[0].base = 0;
	t.log = 0;
}

// ---- Huffman
// weights[0..n) known; the last symbol's weight completes the sum to a power of two (RFC 8878 4.2.1)
ZD_FN inline bool huf_build(Work *w, int n)
{
	uint32_t sum = 0;
	for (int i = 0; i < n; i++) {
		if (w->weights[i] > kMaxHufBits)
			return false;
		if (w->weights[i])
			sum += 1u << (w->weights[i] - 1);
	}
	if (sum == 0 || n >= 256)
		return false;
	const int max_bits = highbit(sum) + 1;
	if (max_bits > kMaxHufBits)
		return false;
	const uint32_t left = (1u << max_bits) - sum;
	if (left & (left - 1))
		return false;
	w->weights[n] = (uint8_t)(highbit(left) + 1);
	n++;
	// canonical order: longest codes (weight 1) first, symbols of one weight in ascending order
	uint32_t rank_cnt[kMaxHufBits + 2] = { 0 }, rank_idx[kMaxHufBits + 2] = { 0 };
	for (int i = 0; i < n; i++)
		if (w->weights[i])
			rank_cnt[max_bits + 1 - w->weights[i]]++; // indexed by code length
	rank_idx[max_bits] = 0;
	for (int b = max_bits; b >= 1; b--)
		rank_idx[b - 1] = rank_idx[b] + rank_cnt[b] * (1u << (max_bits - b));
	if (rank_idx[0] != (1u << max_bits))
		return false;
	for (int i = 0; i < n; i++) {
		if (!w->weights[i])
			continue;
		const int bits = max_bits + 1 - w->weights[i];
		const uint32_t len = 1u << (max_bits - bits), code = rank_idx[bits];
		for (uint32_t k = 0; k < len; k++)
			w->huf[code + k] = (uint16_t)(i | (bits << 8));
		rank_idx[bits] += len;
	}
	w->huf_bits = max_bits;
	return true;
}

// Huffman tree description: returns the bytes consumed
ZD_FN inline int64_t huf_read_tree(Work *w, const uint8_t *src, int64_t len)
{
	if (len < 1)
		return kErrTruncated;
	const int hb = src[0];
	int n = 0;
	int64_t used;
	if (hb >= 128) { // direct: 4 bits per weight
		n = hb - 127;
		used = 1 + (n + 1) / 2;
		if (used > len)
			return kErrTruncated;
		for (int i = 0; i < n; i++)
			w->weights[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
	} else { // FSE-compressed weights, two interleaved states
		used = 1 + hb;
		if (hb < 1 || used > len)
			return kErrTruncated;
		int nsym = 0, log = 0;
		const int64_t d = fse_read_description(src + 1, hb, kMaxWeightLog, 12, w->freq, nsym, log);
		if (d < 0)
			return d;
		if (!fse_build(w->freq, nsym, log, w->wtab, w->sdesc))
			return kErrFse;
		BackBits b;
		if (!back_init(b, src + 1 + d, hb - d))
			return kErrHuffman;
		uint32_t s1 = back_read(b, log), s2 = back_read(b, log);
		for (;;) {
			if (n > 253)
				return kErrHuffman;
			w->weights[n++] = w->wtab[s1].sym;
			s1 = w->wtab[s1].base + back_read(b, w->wtab[s1].nbits);
			if (b.bit < 0) {
				w->weights[n++] = w->wtab[s2].sym;
				break;
			}
			w->weights[n++] = w->wtab[s2].sym;
			s2 = w->wtab[s2].base + back_read(b, w->wtab[s2].nbits);
			if (b.bit < 0) {
				w->weights[n++] = w->wtab[s1].sym;
				break;
			}
		}
	}
	if (!huf_build(w, n))
		return kErrHuffman;
	return used;
}

ZD_FN inline int huf_decode_stream(const Work *w, const uint8_t *src, int64_t len, uint8_t *out, int64_t n)
{
	BackBits b;
	if (!back_init(b, src, len))
		return kErrHuffman;
	const int hb = w->huf_bits;
	const uint32_t mask = (1u << hb) - 1;
	uint32_t state = back_read(b, hb);
	int64_t i = 0;
	while (b.bit > -hb) {
		if (i >= n)
			return kErrHuffman;
		const uint16_t e = w->huf[state];
		out[i++] = (uint8_t)e;
		const int nb = e >> 8;
		state = ((state << nb) & mask) | back_read(b, nb);
	}
	return (i == n && b.bit == -hb) ? 0 : kErrHuffman;
}

// ---- literals section: fills w->lit[0..*nlit), returns the bytes consumed
ZD_FN inline int64_t read_literals(Work *w, const uint8_t *src, int64_t len, int64_t *nlit)
{
	if (len < 1)
		return kErrTruncated;
	const int type = src[0] & 3, sf = (src[0] >> 2) & 3;
	if (type < 2) { // Raw / RLE
		int64_t hdr, regen;
		if ((sf & 1) == 0) {
			hdr = 1;
			regen = src[0] >> 3;
		} else if (sf == 1) {
			hdr = 2;
			if (len < 2)
				return kErrTruncated;
			regen = (src[0] >> 4) | ((int64_t)src[1] << 4);
		} else {
			hdr = 3;
			if (len < 3)
				return kErrTruncated;
			regen = (src[0] >> 4) | ((int64_t)src[1] << 4) | ((int64_t)src[2] << 12);
		}
		if (regen > kBlockMax)
			return kErrLiterals;
		if (type == 0) {
			if (hdr + regen > len)
				return kErrTruncated;
			memcpy(w->lit, src + hdr, (size_t)regen);
			*nlit = regen;
			return hdr + regen;
		}
		if (hdr + 1 > len)
			return kErrTruncated;
		memset(w->lit, src[hdr], (size_t)regen);
		*nlit = regen;
		return hdr + 1;
	}
	// Compressed / Treeless
	int64_t hdr, regen, comp;
	int streams = 4;
	if (sf <= 1) {
		hdr = 3;
		if (len < 3)
			return kErrTruncated;
		const uint32_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16);
		regen = (v >> 4) & 0x3FF;
		comp = (v >> 14) & 0x3FF;
		if (sf == 0)
			streams = 1;
	} else if (sf == 2) {
		hdr = 4;
		if (len < 4)
			return kErrTruncated;
		const uint32_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
		regen = (v >> 4) & 0x3FFF;
		comp = (v >> 18) & 0x3FFF;
	} else {
		hdr = 5;
		if (len < 5)
			return kErrTruncated;
		const uint64_t v = src[0] | (src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
		regen = (int64_t)((v >> 4) & 0x3FFFF);
		comp = (int64_t)((v >> 22) & 0x3FFFF);
	}
	if (regen > kBlockMax || hdr + comp > len)
		return kErrLiterals;
	const uint8_t *p = src + hdr;
	int64_t left = comp;
	if (type == 2) {
		const int64_t t = huf_read_tree(w, p, left);
		if (t < 0)
			return t;
		p += t;
		left -= t;
	} else if (!w->huf_bits)
		return kErrHuffman; // Treeless without a previous table
	if (streams == 1) {
		const int rc = huf_decode_stream(w, p, left, w->lit, regen);
		if (rc)
			return rc;
	} else {
		if (left < 6)
			return kErrTruncated;
		const int64_t s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
		const int64_t s4 = left - 6 - s1 - s2 - s3;
		if (s4 < 1)
			return kErrLiterals;
		const int64_t q = (regen + 3) / 4;
		if (3 * q > regen)
			return kErrLiterals;
		const uint8_t *sp = p + 6;
		int rc = huf_decode_stream(w, sp, s1, w->lit, q);
		rc = rc ? rc : huf_decode_stream(w, sp + s1, s2, w->lit + q, q);
		rc = rc ? rc : huf_decode_stream(w, sp + s1 + s2, s3, w->lit + 2 * q, q);
		rc = rc ? rc : huf_decode_stream(w, sp + s1 + s2 + s3, s4, w->lit + 3 * q, regen - 3 * q);
		if (rc)
			return rc;
	}
	*nlit = regen;
	return hdr + comp;
}

// ---- sequences
// predefined distributions (RFC 8878 3.1.1.3.2.2.1-3)
ZD_FN inline void predefined(int which, int16_t *f, int &nsym, int &log)
{
	const int8_t ll[36] = { 4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1 };
	const int8_t ml[53] = { 1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
				1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1 };
	const int8_t of[29] = { 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };
	const int8_t *src = which == 0 ? ll : (which == 1 ? of : ml);
	nsym = which == 0 ? 36 : (which == 1 ? 29 : 53);
	log = which == 1 ? 5 : 6;
	for (int i = 0; i < nsym; i++)
		f[i] = src[i];
}

// one of the three tables of a block; which: 0 literal lengths, 1 offsets, 2 match lengths.  Returns bytes consumed.
ZD_FN inline int64_t read_seq_table(Work *w, FseTable &t, int which, int mode, const uint8_t *src, int64_t len)
{
	const int max_log = which == 0 ? kMaxLLLog : (which == 1 ? kMaxOFLog : kMaxMLLog);
	const int max_sym = which == 0 ? 35 : (which == 1 ? 31 : 52);
	int nsym = 0, log = 0;
	if (mode == 0) {
		predefined(which, w->freq, nsym, log);
		if (!fse_build(w->freq, nsym, log, t.e, w->sdesc))
			return kErrFse;
		t.log = log;
		return 0;
	}
	if (mode == 1) {
		if (len < 1)
			return kErrTruncated;
		if (src[0] > max_sym)
			return kErrFse;
		fse_set_rle(t, src[0]);
		return 1;
	}
	if (mode == 2) {
		const int64_t d = fse_read_description(src, len, max_log, max_sym, w->freq, nsym, log);
		if (d < 0)
			return d;
		if (!fse_build(w->freq, nsym, log, t.e, w->sdesc))
			return kErrFse;
		t.log = log;
		return d;
	}
	return t.log < 0 ? (int64_t)kErrFse : 0; // Repeat
}

struct FrameState {
	uint32_t rep[3];
};

// Decodes and executes the sequences of one block; out[0..*pos) is the frame's history.
ZD_FN inline int run_sequences(Work *w, const uint8_t *src, int64_t len, int64_t nlit, uint8_t *out, int64_t cap, int64_t *pos,
			       FrameState &fs)
{
	const uint32_t ll_base[36] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
				       48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536 };
	const uint8_t ll_bits[36] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
	const uint32_t ml_base[53] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30,
				       31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
				       16387, 32771, 65539 };
	const uint8_t ml_bits[53] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
				      1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
	int64_t lp = 0, op = *pos;
	if (len < 1)
		return kErrTruncated;
	int64_t nseq = src[0], at = 1;
	if (nseq >= 128) {
		if (nseq == 255) {
			if (len < 3)
				return kErrTruncated;
			nseq = src[1] + ((int64_t)src[2] << 8) + 0x7F00;
			at = 3;
		} else {
			if (len < 2)
				return kErrTruncated;
			nseq = ((nseq - 128) << 8) + src[1];
			at = 2;
		}
	}
	if (nseq > 0) {
		if (at >= len)
			return kErrTruncated;
		const int modes = src[at++];
		if (modes & 3)
			return kErrSequences;
		int64_t d = read_seq_table(w, w->ll, 0, (modes >> 6) & 3, src + at, len - at);
		if (d < 0)
			return (int)d;
		at += d;
		d = read_seq_table(w, w->of, 1, (modes >> 4) & 3, src + at, len - at);
		if (d < 0)
			return (int)d;
		at += d;
		d = read_seq_table(w, w->ml, 2, (modes >> 2) & 3, src + at, len - at);
		if (d < 0)
			return (int)d;
		at += d;
		BackBits b;
		if (!back_init(b, src + at, len - at))
			return kErrSequences;
		uint32_t sl = back_read(b, w->ll.log), so = back_read(b, w->of.log), sm = back_read(b, w->ml.log);
		for (int64_t i = 0; i < nseq; i++) {
			const int oc = w->of.e[so].sym, mc = w->ml.e[sm].sym, lc = w->ll.e[sl].sym;
			if (oc > 31 || mc > 52 || lc > 35)
				return kErrSequences;
			// offset, match length, literal length extra bits, in this order
			const uint32_t ov = (1u << oc) + back_read(b, oc);
			const uint32_t ml = ml_base[mc] + back_read(b, ml_bits[mc]);
			const uint32_t ll = ll_base[lc] + back_read(b, ll_bits[lc]);
			if (i + 1 < nseq) { // state updates: literal length, match length, offset
				sl = w->ll.e[sl].base + back_read(b, w->ll.e[sl].nbits);
				sm = w->ml.e[sm].base + back_read(b, w->ml.e[sm].nbits);
				so = w->of.e[so].base + back_read(b, w->of.e[so].nbits);
			}
			if (b.bit < 0)
				return kErrSequences;
			uint32_t off;
			if (ov > 3) {
				off = ov - 3;
				fs.rep[2] = fs.rep[1];
				fs.rep[1] = fs.rep[0];
				fs.rep[0] = off;
			} else {
				const uint32_t idx = ov - 1 + (ll == 0 ? 1 : 0); // 0 .. 3
				if (idx == 0)
					off = fs.rep[0];
				else {
					off = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
					if (idx > 1)
						fs.rep[2] = fs.rep[1];
					fs.rep[1] = fs.rep[0];
					fs.rep[0] = off;
				}
			}
			if (lp + ll > nlit || op + ll + ml > cap || off == 0 || off > op + ll)
				return kErrOutput;
			memcpy(out + op, w->lit + lp, ll);
			lp += ll;
			op += ll;
			if (off >= ml)
				memcpy(out + op, out + op - off, ml);
			else
				for (uint32_t k = 0; k < ml; k++)
					out[op + k] = out[op + k - off];
			op += ml;
		}
		if (b.bit != 0)
			return kErrSequences;
	}
	if (op + (nlit - lp) > cap)
		return kErrOutput;
	memcpy(out + op, w->lit + lp, (size_t)(nlit - lp));
	op += nlit - lp;
	*pos = op;
	return 0;
}

// One frame -> out[0..cap).  Returns the bytes produced (the caller checks them against the block's u_len), < 0 = Err.
ZD_FN inline int64_t decode_frame(const uint8_t *src, int64_t len, uint8_t *out, int64_t cap, Work *w)
{
	if (len < 6)
		return kErrTruncated;
	if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD))
		return kErrMagic;
	const int fhd = src[4];
	const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
	if (fhd & 8)
		return kErrHeader;
	int64_t at = 5;
	if (!single)
		at += 1; // window descriptor: the whole block is in memory anyway
	if (did)
		return kErrDict;
	const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
	if (at + fcs_bytes > len)
		return kErrTruncated;
	int64_t fcs = -1;
	if (fcs_bytes) {
		uint64_t v = 0;
		for (int i = 0; i < fcs_bytes; i++)
			v |= (uint64_t)src[at + i] << (8 * i);
		fcs = (int64_t)v + (fcs_bytes == 2 ? 256 : 0);
		at += fcs_bytes;
	}
	w->ll.log = w->of.log = w->ml.log = -1;
	w->huf_bits = 0;
	FrameState fs;
	fs.rep[0] = 1;
	fs.rep[1] = 4;
	fs.rep[2] = 8;
	int64_t pos = 0;
	for (;;) {
		if (at + 3 > len)
			return kErrTruncated;
		const uint32_t bh = src[at] | (src[at + 1] << 8) | ((uint32_t)src[at + 2] << 16);
		at += 3;
		const int last = bh & 1, type = (bh >> 1) & 3;
		const int64_t bsize = bh >> 3;
		if (type == 0) { // Raw
			if (at + bsize > len)
				return kErrTruncated;
			if (pos + bsize > cap)
				return kErrOutput;
			memcpy(out + pos, src + at, (size_t)bsize);
			pos += bsize;
			at += bsize;
		} else if (type == 1) { // RLE
			if (at + 1 > len)
				return kErrTruncated;
			if (pos + bsize > cap)
				return kErrOutput;
			memset(out + pos, src[at], (size_t)bsize);
			pos += bsize;
			at += 1;
		} else if (type == 2) {
			if (bsize > kBlockMax || at + bsize > len)
				return kErrBlock;
			int64_t nlit = 0;
			const int64_t l = read_literals(w, src + at, bsize, &nlit);
			if (l < 0)
				return l;
			const int rc = run_sequences(w, src + at + l, bsize - l, nlit, out, cap, &pos, fs);
			if (rc)
				return rc;
			at += bsize;
		} else
			return kErrBlock;
		if (last)
			break;
	}
	if (checksum)
		at += 4;
	if (at > len)
		return kErrTruncated;
	if (fcs >= 0 && fcs != pos)
		return kErrOutput;
	return pos;
}

} // namespace zd
} // namespace lrz
