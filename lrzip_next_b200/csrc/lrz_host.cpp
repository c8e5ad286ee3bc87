// lrz_host.cpp -- see lrz_host.h.  Host control plane: sizing rules, block flush plan, magic, MD5.
#include "lrz_host.h"

#include <string.h>

namespace lrz {

static int64_t round_up_page(int64_t len, int page) // src/util.c:197-204
{
	const int64_t rem = len % page;
	return rem ? len + page - rem : len;
}

static uint32_t lzma2_dic_from_prop(unsigned p) // LZMA2_DIC_SIZE_FROM_PROP, lrzip_private.h:236
{
	return p == 40 ? 0xFFFFFFFFu : (((uint32_t)2 | (p & 1)) << (p / 2 + 11));
}

unsigned lzma2_prop_from_dic(uint32_t dict) // lrzip_private.h:238-245
{
	unsigned i = 0;
	while (i <= 40 && dict > lzma2_dic_from_prop(i))
		i++;
	return i;
}

int chunk_bytes_for(int64_t chunk_size)
{
	int bits = 8;
	while (chunk_size >> bits > 0)
		bits++;
	return bits / 8 + (bits % 8 ? 1 : 0);
}

int zstd_level_for(int level)
{
	static const int map[10] = { -1, 2, 4, 5, 7, 12, 15, 17, 18, 22 };
	return map[level < 1 ? 1 : (level > 9 ? 9 : level)];
}

static uint32_t lzma_dict_for_level(int level) // src/util.c:108-126
{
	if (level >= 1 && level <= 3)
		return 1u << (level * 2 + 16);
	if (level >= 4 && level <= 6)
		return 1u << (level + 19);
	if (level == 7)
		return 1u << 25;
	if (level == 8)
		return 1u << 26;
	if (level == 9)
		return 1u << 27;
	return 1u << 24;
}

static int64_t lzma_overhead_for(uint32_t dict) // src/util.c:130
{
	return (int64_t)dict * 23 / 2 + 6 * kOneMB + 16384;
}

// Window policy src/rzip.c:995-1013 + 1046-1049, thread count src/stream.c:1090-1102, block size
// src/stream.c:1169-1323 (evaluated once per archive with the first chunk's size as chunk_limit; the
// test malloc of :1291-1306 is taken to succeed).
int compute_sizing(const lrzgpu_params &p, int64_t st_size_in, lrzgpu_sizing_t &o)
{
	int64_t st_size = st_size_in;
	// STDIN (src/rzip.c:969-972, 1001-1013, mmap_stdin :800-836): the size is not known in advance.  Every chunk is one
	// anonymous mmap of max_mmap = min(maxram, max_chunk) bytes filled from the pipe (the last one shrunk to what was
	// read), and when the streams are opened control->st_size holds the bytes read so far = the first chunk.
	int64_t stdin_chunk = 0;
	if (p.stdin_mode) {
		if (p.unlimited || p.ramsize <= 0 || p.page_size <= 0)
			return LRZGPU_EUNSUPPORTED;
		int64_t maxram = p.ramsize / 3; // setup_ram, src/util.c:179-188 (not STDOUT)
		maxram -= maxram % p.page_size;
		if (!maxram)
			maxram = p.page_size;
		const int64_t mc = p.window ? (int64_t)p.window * kChunkMultiple : p.ramsize / 3 * 2;
		stdin_chunk = maxram < mc ? maxram : mc;
		if (st_size >= stdin_chunk && st_size % stdin_chunk == 0)
			return LRZGPU_EUNSUPPORTED; // the reference would append an empty chunk after the last full one
		if (st_size > stdin_chunk)
			st_size = stdin_chunk;
	}
	if (p.level < 1 || p.level > 9 || p.rzip_level < 0 || p.rzip_level > 9 || p.page_size <= 0 || p.threads < 1 ||
	    p.ramsize <= 0)
		return LRZGPU_EINVAL;
	// windows after the first start at multiples of max_chunk, which is only rounded to page_size: the tag scan's
	// TMA loads (cp.async.bulk) need 16-byte aligned sources, so the page must keep that alignment
	if (p.page_size % 16)
		return LRZGPU_EINVAL;
	// --nobemt with more than one thread selects the single-threaded bt4 finder (numThreads = 1, src/stream.c:456),
	// whose hash2/hash3 handling differs from the two-thread finder this library reproduces
	if (p.nobemt && p.backend == LRZGPU_BACKEND_LZMA && p.level >= 5 && p.threads > 1)
		return LRZGPU_EUNSUPPORTED;
	if (p.backend != LRZGPU_BACKEND_NONE && p.backend != LRZGPU_BACKEND_LZMA && p.backend != LRZGPU_BACKEND_ZSTD)
		return LRZGPU_EUNSUPPORTED;
	// filters (src/main.c:700-755): the RISC-V converter is not built; --delta takes 1..16 or a multiple of 16 up to 256
	switch (p.filter) {
	case LRZGPU_FILTER_NONE:
	case LRZGPU_FILTER_X86:
	case LRZGPU_FILTER_ARM:
	case LRZGPU_FILTER_ARMT:
	case LRZGPU_FILTER_IA64:
	case LRZGPU_FILTER_PPC:
	case LRZGPU_FILTER_SPARC:
	case LRZGPU_FILTER_ARM64:
	case LRZGPU_FILTER_RISCV:
		break;
	case LRZGPU_FILTER_DELTA:
		if (p.delta < 1 || p.delta > 256 || (p.delta > 16 && p.delta % 16))
			return LRZGPU_EINVAL;
		break;
	default:
		return LRZGPU_EINVAL;
	}
	const bool stored = p.backend == LRZGPU_BACKEND_NONE, lzma = p.backend == LRZGPU_BACKEND_LZMA;
	const int testbufs = stored ? 1 : 2;
	const int64_t usable_ram = p.ramsize / 3; // setup_ram, src/util.c:179-188

	int64_t max_chunk = p.unlimited ? st_size : (p.window ? (int64_t)p.window * kChunkMultiple : p.ramsize / 3 * 2);
	if (max_chunk < st_size) { // round_to_page, src/util.c:190-195
		max_chunk -= max_chunk % p.page_size;
		if (!max_chunk)
			max_chunk = p.page_size;
	}
	if (stdin_chunk)
		max_chunk = stdin_chunk;
	int64_t chunk_limit = max_chunk < st_size ? max_chunk : st_size;
	if (chunk_limit < p.page_size)
		chunk_limit = p.page_size;

	int threads = p.threads > 1 ? p.threads + 1 : p.threads;
	if (stored)
		threads = 1;

	uint32_t dict = lzma ? lzma_dict_for_level(p.level) : 0;
	int64_t overhead = lzma ? lzma_overhead_for(dict) : 0;
	int64_t limit = usable_ram / testbufs;
	if (lzma) {
		const int save_threads = threads;
		int thread_limit = threads >= p.processors / 2 ? threads / 2 : threads;
		const uint32_t save_dict = dict;
		unsigned exponent = lzma2_prop_from_dic(dict);
		const unsigned save_exponent = exponent;
		bool set = false;
		for (;;) {
			do {
				for (threads = save_threads; threads >= thread_limit; threads--)
					if (limit >= overhead * threads / testbufs) {
						set = true;
						break;
					}
				if (set)
					break;
				exponent -= 1;
				dict = lzma2_dic_from_prop(exponent);
				overhead = lzma_overhead_for(dict);
			} while (dict > (1u << 24));
			if (set || thread_limit <= 1)
				break;
			thread_limit--;
			dict = save_dict; // NB: like the reference (src/stream.c:1213-1217) the overhead of the
			exponent = save_exponent; // last tried dictionary stays in place until the next reduction
		}
	}
	if (threads < 1)
		return LRZGPU_EINVAL; // not even one encoder fits the RAM budget: the reference divides by zero here
	if (st_size > 0 && st_size < limit)
		limit = st_size > kStreamBufsize ? st_size : kStreamBufsize;
	else if (limit > chunk_limit)
		limit = chunk_limit;

	if (lzma && limit / threads > kStreamBufsize) {
		const int64_t a = overhead - (int64_t)dict;
		o.bufsize = round_up_page((limit > a ? limit : a) / threads, p.page_size);
	} else {
		const int64_t a = limit / threads > kStreamBufsize ? limit / threads : kStreamBufsize;
		o.bufsize = round_up_page(limit < a ? limit : a, p.page_size);
	}
	o.threads = threads;
	o.dict_size = dict;
	o.overhead = overhead;
	o.max_chunk = max_chunk;
	return LRZGPU_OK;
}

void plan_blocks(int64_t s0_len, int64_t s1_len, int64_t bufsize, const int64_t *w1, std::vector<BlockPlan> &out)
{
	const int64_t nb0 = s0_len / bufsize, nb1 = s1_len / bufsize;
	int64_t j = 0, k = 1;
	out.clear();
	while (j < nb0 || k <= nb1) {
		// stream-1 block k filled before stream-0 block j iff k*bufsize bytes of stream 1 were out by then
		const bool take1 = (j >= nb0) || (k <= nb1 && k * bufsize <= w1[j]);
		if (take1) {
			out.push_back({ 1, (k - 1) * bufsize, bufsize });
			k++;
		} else {
			out.push_back({ 0, j * bufsize, bufsize });
			j++;
		}
	}
	out.push_back({ 0, nb0 * bufsize, s0_len - nb0 * bufsize }); // close_stream_out: stream 0 tail, then stream 1 tail
	out.push_back({ 1, nb1 * bufsize, s1_len - nb1 * bufsize });
}

void put_le(uint8_t *at, int64_t v, int width)
{
	for (int i = 0; i < width; i++)
		at[i] = (uint8_t)((uint64_t)v >> (8 * i));
}

// src/lrzip.c:131-208 write_magic (file -> file, MD5, no encryption, no filter, no comment)
void make_magic(uint8_t magic[21], const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t st_size)
{
	const int rzl = p.rzip_level ? p.rzip_level : p.level;
	memset(magic, 0, 21);
	memcpy(magic, "LRZI", 4);
	magic[4] = 0;
	magic[5] = 14;
	put_le(magic + 6, st_size, 8);
	magic[14] = 1; // MD5
	// filter byte (src/lrzip.c:150-155): the flag, or 128 + the coded delta distance
	if (p.filter == LRZGPU_FILTER_DELTA)
		magic[16] = (uint8_t)(128 + (p.delta <= 16 ? p.delta : (p.delta >> 4) + 15));
	else
		magic[16] = (uint8_t)p.filter;
	if (p.backend == LRZGPU_BACKEND_LZMA) {
		magic[17] = 1;
		magic[18] = (uint8_t)lzma2_prop_from_dic(sz.dict_size);
	} else if (p.backend == LRZGPU_BACKEND_ZSTD) {
		magic[17] = (uint8_t)((p.level << 4) | 4);
		magic[18] = (uint8_t)zstd_level_for(p.level);
	}
	magic[19] = (uint8_t)((rzl << 4) + p.level);
}

// ---- archive walker (get_fileinfo, src/lrzip.c:1069-1459) -----------------------------------------------------------
static int64_t get_le_w(const uint8_t *p, int width)
{
	int64_t v = 0;
	for (int i = 0; i < width; i++)
		v |= (int64_t)p[i] << (8 * i);
	return v;
}

extern "C" int lrzgpu_info(const uint8_t *arc, int64_t len, lrzgpu_archive_info *info, lrzgpu_block_info *blocks, int64_t cap,
			   int64_t *nblocks)
{
	if (!arc || !info || len < 21 + 16 || memcmp(arc, "LRZI", 4))
		return LRZGPU_EINVAL;
	memset(info, 0, sizeof(*info));
	info->major = arc[4];
	info->minor = arc[5];
	info->expected_size = get_le_w(arc + 6, 8);
	info->hash_type = arc[14];
	info->encrypted = arc[15];
	// filter byte (src/lrzip.c:323-340 for 0.13+ archives): 128 + coded distance = Delta, else the flag
	if (arc[16] & 128) {
		const int i = arc[16] & 127;
		info->filter = LRZGPU_FILTER_DELTA;
		info->delta = i <= 16 ? i : (i - 15) * 16;
	} else
		info->filter = arc[16];
	info->backend_code = arc[17] & 15;
	info->backend_prop = arc[18];
	if (info->backend_code == 1) { // LZMA2-style dictionary code (src/lrzip.c:245-250): 2^(n/2 + 12), odd: x1.5
		info->lzma_dict_size = lzma2_dic_from_prop(arc[18]);
	}
	info->rzip_level = arc[19] >> 4;
	info->level = arc[19] & 15;
	info->archive_bytes = len;
	if (info->encrypted)
		return LRZGPU_EUNSUPPORTED;
	const int64_t tail = info->hash_type ? 16 : 0; // MD5; other digests of the reference are longer and not walked
	if (info->hash_type > 1)
		return LRZGPU_EUNSUPPORTED;
	const int64_t end = len - tail;
	int64_t pos = 21 + arc[20], nb = 0;
	for (bool last = false; !last;) {
		if (pos + 2 > end)
			return LRZGPU_EINVAL;
		const int cb = arc[pos], eof = arc[pos + 1];
		const int64_t hdr = 1 + 3 * (int64_t)cb;
		if (cb < 1 || cb > 8 || pos + 2 + cb + 2 * hdr > end)
			return LRZGPU_EINVAL;
		pos += 2 + cb;
		const int64_t initial = pos;
		int64_t chunk_end = initial + 2 * hdr;
		for (int s = 0; s < 2; s++) {
			int64_t head = get_le_w(arc + initial + s * hdr + 1 + 2 * cb, cb);
			while (head) {
				const int64_t at = initial + head;
				if (at < initial || at + hdr > end)
					return LRZGPU_EINVAL;
				lrzgpu_block_info b;
				b.chunk = info->chunks;
				b.stream = s;
				b.ctype = arc[at];
				b.c_len = get_le_w(arc + at + 1, cb);
				b.u_len = get_le_w(arc + at + 1 + cb, cb);
				b.offset = at;
				b.next_head = get_le_w(arc + at + 1 + 2 * cb, cb);
				if (b.c_len < 0 || b.u_len < 0 || at + hdr + b.c_len > end)
					return LRZGPU_EINVAL;
				if (blocks && nb < cap)
					blocks[nb] = b;
				nb++;
				info->stream_c_bytes[s] += b.c_len;
				info->stream_u_bytes[s] += b.u_len;
				info->blocks_by_ctype[b.ctype & 15]++;
				if (at + hdr + b.c_len > chunk_end)
					chunk_end = at + hdr + b.c_len;
				head = b.next_head;
			}
		}
		info->chunks++;
		pos = chunk_end;
		last = eof != 0;
	}
	if (pos != end)
		return LRZGPU_EINVAL;
	info->blocks = nb;
	if (tail)
		memcpy(info->md5, arc + end, 16);
	if (nblocks)
		*nblocks = nb;
	return LRZGPU_OK;
}

// ---- MD5 ---------------------------------------------------------------------------------------
Md5::Md5() : a(0x67452301u), b(0xefcdab89u), c(0x98badcfeu), d(0x10325476u), len(0), fill(0) {}

static inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

#define MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define MD5_STEP(f, w, x, y, z, m, k, s) w = x + rol(w + f(x, y, z) + (m) + (k), s)

static void md5_blocks(uint32_t st[4], const uint8_t *p, size_t nblocks)
{
	uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
	while (nblocks--) {
		uint32_t m[16];
		memcpy(m, p, 64); // little-endian host
		const uint32_t sa = a, sb = b, sc = c, sd = d;
		MD5_STEP(MD5_F, a, b, c, d, m[0], 0xd76aa478, 7);
		MD5_STEP(MD5_F, d, a, b, c, m[1], 0xe8c7b756, 12);
		MD5_STEP(MD5_F, c, d, a, b, m[2], 0x242070db, 17);
		MD5_STEP(MD5_F, b, c, d, a, m[3], 0xc1bdceee, 22);
		MD5_STEP(MD5_F, a, b, c, d, m[4], 0xf57c0faf, 7);
		MD5_STEP(MD5_F, d, a, b, c, m[5], 0x4787c62a, 12);
		MD5_STEP(MD5_F, c, d, a, b, m[6], 0xa8304613, 17);
		MD5_STEP(MD5_F, b, c, d, a, m[7], 0xfd469501, 22);
		MD5_STEP(MD5_F, a, b, c, d, m[8], 0x698098d8, 7);
		MD5_STEP(MD5_F, d, a, b, c, m[9], 0x8b44f7af, 12);
		MD5_STEP(MD5_F, c, d, a, b, m[10], 0xffff5bb1, 17);
		MD5_STEP(MD5_F, b, c, d, a, m[11], 0x895cd7be, 22);
		MD5_STEP(MD5_F, a, b, c, d, m[12], 0x6b901122, 7);
		MD5_STEP(MD5_F, d, a, b, c, m[13], 0xfd987193, 12);
		MD5_STEP(MD5_F, c, d, a, b, m[14], 0xa679438e, 17);
		MD5_STEP(MD5_F, b, c, d, a, m[15], 0x49b40821, 22);
		MD5_STEP(MD5_G, a, b, c, d, m[1], 0xf61e2562, 5);
		MD5_STEP(MD5_G, d, a, b, c, m[6], 0xc040b340, 9);
		MD5_STEP(MD5_G, c, d, a, b, m[11], 0x265e5a51, 14);
		MD5_STEP(MD5_G, b, c, d, a, m[0], 0xe9b6c7aa, 20);
		MD5_STEP(MD5_G, a, b, c, d, m[5], 0xd62f105d, 5);
		MD5_STEP(MD5_G, d, a, b, c, m[10], 0x02441453, 9);
		MD5_STEP(MD5_G, c, d, a, b, m[15], 0xd8a1e681, 14);
		MD5_STEP(MD5_G, b, c, d, a, m[4], 0xe7d3fbc8, 20);
		MD5_STEP(MD5_G, a, b, c, d, m[9], 0x21e1cde6, 5);
		MD5_STEP(MD5_G, d, a, b, c, m[14], 0xc33707d6, 9);
		MD5_STEP(MD5_G, c, d, a, b, m[3], 0xf4d50d87, 14);
		MD5_STEP(MD5_G, b, c, d, a, m[8], 0x455a14ed, 20);
		MD5_STEP(MD5_G, a, b, c, d, m[13], 0xa9e3e905, 5);
		MD5_STEP(MD5_G, d, a, b, c, m[2], 0xfcefa3f8, 9);
		MD5_STEP(MD5_G, c, d, a, b, m[7], 0x676f02d9, 14);
		MD5_STEP(MD5_G, b, c, d, a, m[12], 0x8d2a4c8a, 20);
		MD5_STEP(MD5_H, a, b, c, d, m[5], 0xfffa3942, 4);
		MD5_STEP(MD5_H, d, a, b, c, m[8], 0x8771f681, 11);
		MD5_STEP(MD5_H, c, d, a, b, m[11], 0x6d9d6122, 16);
		MD5_STEP(MD5_H, b, c, d, a, m[14], 0xfde5380c, 23);
		MD5_STEP(MD5_H, a, b, c, d, m[1], 0xa4beea44, 4);
		MD5_STEP(MD5_H, d, a, b, c, m[4], 0x4bdecfa9, 11);
		MD5_STEP(MD5_H, c, d, a, b, m[7], 0xf6bb4b60, 16);
		MD5_STEP(MD5_H, b, c, d, a, m[10], 0xbebfbc70, 23);
		MD5_STEP(MD5_H, a, b, c, d, m[13], 0x289b7ec6, 4);
		MD5_STEP(MD5_H, d, a, b, c, m[0], 0xeaa127fa, 11);
		MD5_STEP(MD5_H, c, d, a, b, m[3], 0xd4ef3085, 16);
		MD5_STEP(MD5_H, b, c, d, a, m[6], 0x04881d05, 23);
		MD5_STEP(MD5_H, a, b, c, d, m[9], 0xd9d4d039, 4);
		MD5_STEP(MD5_H, d, a, b, c, m[12], 0xe6db99e5, 11);
		MD5_STEP(MD5_H, c, d, a, b, m[15], 0x1fa27cf8, 16);
		MD5_STEP(MD5_H, b, c, d, a, m[2], 0xc4ac5665, 23);
		MD5_STEP(MD5_I, a, b, c, d, m[0], 0xf4292244, 6);
		MD5_STEP(MD5_I, d, a, b, c, m[7], 0x432aff97, 10);
		MD5_STEP(MD5_I, c, d, a, b, m[14], 0xab9423a7, 15);
		MD5_STEP(MD5_I, b, c, d, a, m[5], 0xfc93a039, 21);
		MD5_STEP(MD5_I, a, b, c, d, m[12], 0x655b59c3, 6);
		MD5_STEP(MD5_I, d, a, b, c, m[3], 0x8f0ccc92, 10);
		MD5_STEP(MD5_I, c, d, a, b, m[10], 0xffeff47d, 15);
		MD5_STEP(MD5_I, b, c, d, a, m[1], 0x85845dd1, 21);
		MD5_STEP(MD5_I, a, b, c, d, m[8], 0x6fa87e4f, 6);
		MD5_STEP(MD5_I, d, a, b, c, m[15], 0xfe2ce6e0, 10);
		MD5_STEP(MD5_I, c, d, a, b, m[6], 0xa3014314, 15);
		MD5_STEP(MD5_I, b, c, d, a, m[13], 0x4e0811a1, 21);
		MD5_STEP(MD5_I, a, b, c, d, m[4], 0xf7537e82, 6);
		MD5_STEP(MD5_I, d, a, b, c, m[11], 0xbd3af235, 10);
		MD5_STEP(MD5_I, c, d, a, b, m[2], 0x2ad7d2bb, 15);
		MD5_STEP(MD5_I, b, c, d, a, m[9], 0xeb86d391, 21);
		a += sa;
		b += sb;
		c += sc;
		d += sd;
		p += 64;
	}
	st[0] = a;
	st[1] = b;
	st[2] = c;
	st[3] = d;
}

void Md5::update(const uint8_t *p, size_t n)
{
	uint32_t st[4] = { a, b, c, d };
	len += n;
	if (fill) {
		const size_t take = (size_t)(64 - fill) < n ? (size_t)(64 - fill) : n;
		memcpy(buf + fill, p, take);
		fill += (int)take;
		p += take;
		n -= take;
		if (fill == 64) {
			md5_blocks(st, buf, 1);
			fill = 0;
		}
	}
	if (n >= 64) {
		md5_blocks(st, p, n / 64);
		p += n & ~(size_t)63;
		n &= 63;
	}
	if (n) {
		memcpy(buf, p, n);
		fill = (int)n;
	}
	a = st[0];
	b = st[1];
	c = st[2];
	d = st[3];
}

void Md5::final(uint8_t digest[16])
{
	uint8_t tail[128];
	const uint64_t bits = len * 8;
	const int tl = fill < 56 ? 64 : 128;
	uint32_t st[4] = { a, b, c, d };
	memset(tail, 0, sizeof(tail));
	memcpy(tail, buf, (size_t)fill);
	tail[fill] = 0x80;
	for (int i = 0; i < 8; i++)
		tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
	md5_blocks(st, tail, (size_t)tl / 64);
	for (int i = 0; i < 16; i++)
		digest[i] = (uint8_t)(st[i / 4] >> (8 * (i % 4)));
}

} // namespace lrz
