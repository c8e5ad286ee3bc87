/* oracle/lrz_format.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the parts of lrzip-next that turn rzip's two byte streams into a .lrz
 * archive: window (chunk) policy, stream block sizing, block flush order, chunk / stream / block
 * headers with next_head patching, trailing MD5 and the 21-byte magic header.
 * Pinned against oracle/_ref archives by tests/test_oracle.py and tests/golden/.
 * References are to /root/reference.
 */
#include "rzip_oracle.h"

#include <stdlib.h>
#include <string.h>

#define ONE_MB 1048576LL
#define STREAM_BUFSIZE (10 * ONE_MB)     /* src/include/lrzip_private.h:16 */
#define CHUNK_MULTIPLE (100 * ONE_MB)    /* src/rzip.c:48 */
#define CTYPE_NONE 3                     /* src/include/lrzip_private.h:287 */

/* ---- MD5 (RFC 1321), the reference's default whole-file hash (src/main.c:789) ------------ */
typedef struct { uint32_t a, b, c, d; uint64_t len; uint8_t buf[64]; int fill; } md5_ctx;

static const uint32_t md5_k[64] = {
	0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
	0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
	0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
	0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
	0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
	0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
	0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
	0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
static const uint8_t md5_s[64] = {
	7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
	4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21 };

static void md5_block(md5_ctx *c, const uint8_t *p)
{
	uint32_t m[16], a = c->a, b = c->b, cc = c->c, d = c->d;
	int i;
	for (i = 0; i < 16; i++)
		m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
	for (i = 0; i < 64; i++) {
		uint32_t f, g, t;
		if (i < 16) { f = (b & cc) | (~b & d); g = i; }
		else if (i < 32) { f = (d & b) | (~d & cc); g = (5 * i + 1) & 15; }
		else if (i < 48) { f = b ^ cc ^ d; g = (3 * i + 5) & 15; }
		else { f = cc ^ (b | ~d); g = (7 * i) & 15; }
		t = a + f + md5_k[i] + m[g];
		a = d; d = cc; cc = b;
		b = b + ((t << md5_s[i]) | (t >> (32 - md5_s[i])));
	}
	c->a += a; c->b += b; c->c += cc; c->d += d;
}

void rzo_md5(const uint8_t *in, int64_t n, uint8_t digest[16])
{
	md5_ctx c = { 0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476, 0, {0}, 0 };
	uint8_t tail[128];
	int64_t i, full = n / 64;
	int rem = (int)(n % 64), tl;
	uint64_t bits = (uint64_t)n * 8;
	uint32_t v[4];
	for (i = 0; i < full; i++)
		md5_block(&c, in + 64 * i);
	memset(tail, 0, sizeof(tail));
	memcpy(tail, in + 64 * full, (size_t)rem);
	tail[rem] = 0x80;
	tl = rem < 56 ? 64 : 128;
	for (i = 0; i < 8; i++)
		tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
	md5_block(&c, tail);
	if (tl == 128)
		md5_block(&c, tail + 64);
	v[0] = c.a; v[1] = c.b; v[2] = c.c; v[3] = c.d;
	for (i = 0; i < 16; i++)
		digest[i] = (uint8_t)(v[i / 4] >> (8 * (i % 4)));
}

/* ---- sizing ------------------------------------------------------------------------------ */
static int64_t round_up_page(int64_t len, int page) /* src/util.c:197-204 */
{
	int64_t rem = len % page;
	return rem ? len + page - rem : len;
}

static void round_to_page(int64_t *size, int page) /* src/util.c:190-195 */
{
	*size -= *size % page;
	if (!*size)
		*size = page;
}

/* src/include/lrzip_private.h:236-245 */
static uint32_t lzma2_dic_from_prop(unsigned p)
{
	return p == 40 ? 0xFFFFFFFFu : (((uint32_t)2 | (p & 1)) << (p / 2 + 11));
}

static unsigned lzma2_prop_from_dic(uint32_t d)
{
	unsigned i;
	for (i = 0; i <= 40; i++)
		if (d <= lzma2_dic_from_prop(i))
			break;
	return i;
}

/* src/util.c:103-131 setup_overhead (LZMA branch) */
static uint32_t lzma_default_dict(int level)
{
	switch (level) {
	case 1: case 2: case 3: return 1u << (level * 2 + 16);
	case 4: case 5: case 6: return 1u << (level + 19);
	case 7: return 1u << 25;
	case 8: return 1u << 26;
	case 9: return 1u << 27;
	default: return 1u << 24;
	}
}

static int64_t lzma_overhead(uint32_t dict) { return (int64_t)dict * 23 / 2 + 6 * ONE_MB + 16384; }

/* src/main.c:779-780, src/util.c:179-188 setup_ram, src/rzip.c:995-1013 window policy,
 * src/stream.c:1090-1118 prepare_streamout_threads, :1169-1331 open_stream_out sizing
 * (evaluated once, with chunk_limit = the first chunk's size). */
int rzo_sizing_compute(const rzo_params *p, int64_t st_size, rzo_sizing *o)
{
	const int nocomp = p->backend == RZO_BACKEND_NONE;
	const int page = p->page_size;
	int64_t usable_ram = p->ramsize / 3, limit, chunk_limit, max_chunk, overhead = 0;
	int threads = p->threads, testbufs = nocomp ? 1 : 2;
	uint32_t dict = 0;

	if (p->unlimited)
		max_chunk = st_size;
	else if (p->window)
		max_chunk = (int64_t)p->window * CHUNK_MULTIPLE;
	else
		max_chunk = p->ramsize / 3 * 2;
	if (max_chunk < st_size)
		round_to_page(&max_chunk, page);
	chunk_limit = max_chunk < st_size ? max_chunk : st_size; /* src/rzip.c:1046-1049 */
	if (chunk_limit < page)                                   /* src/stream.c:1150-1151 */
		chunk_limit = page;

	if (threads > 1)
		++threads;
	if (nocomp)
		threads = 1;

	if (p->backend == RZO_BACKEND_LZMA) {
		dict = lzma_default_dict(p->level);
		overhead = lzma_overhead(dict);
	}
	limit = usable_ram / testbufs;
	{
		int save_threads = threads;
		int thread_limit = threads >= p->processors / 2 ? threads / 2 : threads;
		if (p->backend == RZO_BACKEND_LZMA) { /* src/stream.c:1187-1216 */
			unsigned exponent = lzma2_prop_from_dic(dict), save_exponent = exponent;
			uint32_t save_dict = dict;
			int set = 0;
retry_lzma:
			do {
				for (threads = save_threads; threads >= thread_limit; threads--)
					if (limit >= overhead * threads / testbufs) {
						set = 1;
						break;
					}
				if (set)
					break;
				exponent -= 1;
				dict = lzma2_dic_from_prop(exponent);
				overhead = lzma_overhead(dict);
			} while (dict > (1u << 24));
			if (!set && thread_limit > 1) {
				thread_limit--;
				dict = save_dict;
				exponent = save_exponent;
				goto retry_lzma;
			}
		}
	}
	if (threads < 1)
		return -1; /* the reference would divide by zero at src/stream.c:1316 */
	if (st_size > 0 && st_size < limit) /* src/stream.c:1286-1290 */
		limit = st_size > STREAM_BUFSIZE ? st_size : STREAM_BUFSIZE;
	else if (limit > chunk_limit)
		limit = chunk_limit;
	/* :1291-1306 the test malloc is assumed to succeed */
	if (p->backend == RZO_BACKEND_LZMA && limit / threads > STREAM_BUFSIZE) { /* :1316-1321 */
		int64_t a = overhead - (int64_t)dict;
		o->bufsize = round_up_page((limit > a ? limit : a) / threads, page);
	} else { /* :1322-1323 */
		int64_t a = limit / threads > STREAM_BUFSIZE ? limit / threads : STREAM_BUFSIZE;
		o->bufsize = round_up_page(limit < a ? limit : a, page);
	}
	o->threads = threads;
	o->dict_size = dict;
	o->overhead = overhead;
	o->max_chunk = max_chunk;
	return 0;
}

/* ---- archive assembly -------------------------------------------------------------------- */
typedef struct { uint8_t *p; int64_t len, cap; } obuf;

static int ob_reserve(obuf *o, int64_t extra)
{
	if (o->len + extra > o->cap) {
		int64_t nc = o->cap ? o->cap * 2 : 65536;
		uint8_t *np;
		while (nc < o->len + extra)
			nc *= 2;
		np = realloc(o->p, (size_t)nc);
		if (!np)
			return -1;
		o->p = np;
		o->cap = nc;
	}
	return 0;
}

static int ob_put(obuf *o, const void *src, int64_t n)
{
	if (ob_reserve(o, n))
		return -1;
	memcpy(o->p + o->len, src, (size_t)n);
	o->len += n;
	return 0;
}

static int ob_val(obuf *o, int64_t v, int width) /* src/stream.c write_val: LE, `width` bytes */
{
	uint8_t b[8];
	int i;
	for (i = 0; i < 8; i++)
		b[i] = (uint8_t)((uint64_t)v >> (8 * i));
	return ob_put(o, b, width);
}

static void patch_val(uint8_t *at, int64_t v, int width)
{
	int i;
	for (i = 0; i < width; i++)
		at[i] = (uint8_t)((uint64_t)v >> (8 * i));
}

typedef struct { int stream; int64_t off, len; } blk;

/* Re-derive the global flush order of stream blocks (src/stream.c:2198-2216 write_stream,
 * src/rzip.c:229-246 write_sbstream, :1878 flush_buffer, :2253-2259 close_stream_out) by
 * replaying the self-describing stream-0 records: a buffer is flushed the moment it holds
 * exactly `bufsize` bytes; a literal record writes its 3 header bytes to stream 0 before its
 * payload goes to stream 1; at close stream 0's tail is flushed before stream 1's. */
static int plan_blocks(const uint8_t *s0, int64_t s0_len, int64_t s1_len, int cb, int64_t bufsize,
		       blk **out, int *nout)
{
	int cap = 16, n = 0;
	blk *b = malloc(sizeof(blk) * cap);
	int64_t o0 = 0, fl0 = 0, o1 = 0, fl1 = 0; /* written / flushed offsets per stream */
	int64_t pos = 0;
	int done_records = 0;
	if (!b)
		return -1;
#define PUSH(S, OFF, LEN) do { if (n == cap) { cap *= 2; b = realloc(b, sizeof(blk) * cap); if (!b) return -1; } \
		b[n].stream = (S); b[n].off = (OFF); b[n].len = (LEN); n++; } while (0)
#define ADV0(K) do { int64_t k_ = (K); while (k_) { int64_t r_ = bufsize - (o0 - fl0); int64_t t_ = k_ < r_ ? k_ : r_; \
		o0 += t_; k_ -= t_; if (o0 - fl0 == bufsize) { PUSH(0, fl0, bufsize); fl0 = o0; } } } while (0)
#define ADV1(K) do { int64_t k_ = (K); while (k_) { int64_t r_ = bufsize - (o1 - fl1); int64_t t_ = k_ < r_ ? k_ : r_; \
		o1 += t_; k_ -= t_; if (o1 - fl1 == bufsize) { PUSH(1, fl1, bufsize); fl1 = o1; } } } while (0)
	while (pos < s0_len) {
		if (!done_records) {
			int head, len;
			if (pos + 3 > s0_len) { free(b); return -2; }
			head = s0[pos];
			len = s0[pos + 1] | (s0[pos + 2] << 8);
			ADV0(3);
			pos += 3;
			if (head == 1) {
				ADV0(cb);
				pos += cb;
			} else if (len) {
				ADV1(len);
			} else {
				done_records = 1; /* terminator (0,0); the 4 CRC bytes follow */
			}
		} else {
			ADV0(s0_len - pos);
			pos = s0_len;
		}
	}
	if (o1 != s1_len) { free(b); return -3; }
	PUSH(0, fl0, o0 - fl0);
	PUSH(1, fl1, o1 - fl1);
#undef PUSH
#undef ADV0
#undef ADV1
	*out = b;
	*nout = n;
	return 0;
}

/* src/lrzip.c:1464-1591 compress_file -> src/rzip.c:922 rzip_fd -> src/stream.c:1550 compthread,
 * src/lrzip.c:131-208 write_magic; file -> file, no encryption, no comment, MD5 hash. */
int rzo_compress(const rzo_params *p, const uint8_t *in, int64_t n, rzo_block_fn fn, void *user,
		 uint8_t **out, int64_t *out_len, rzo_stats *sum)
{
	rzo_sizing sz;
	obuf o = { 0, 0, 0 };
	int rzl = p->rzip_level ? p->rzip_level : p->level;
	int64_t left = n, victim_round = 0;
	uint8_t magic[21];
	uint8_t md5[16];
	int pass = 0;
	uint8_t *scratch = NULL;
	int64_t scratch_cap = 0;

	if (sum)
		memset(sum, 0, sizeof(*sum));
	rzo_sizing_compute(p, n, &sz);
	memset(magic, 0, sizeof(magic));
	if (ob_put(&o, magic, 21))
		return -2;

	while (!pass || left > 0) { /* src/rzip.c:1041 */
		int64_t offset = n - left;
		int64_t chunk = sz.max_chunk < left ? sz.max_chunk : left;
		int64_t size_field = chunk < p->page_size ? p->page_size : chunk; /* src/stream.c:1150-1152 */
		int bits = 8, cb, eof, nb = 0, i;
		uint8_t *s0 = NULL, *s1 = NULL;
		int64_t s0_len = 0, s1_len = 0, initial_pos, cur_pos = 0, last_head[2];
		blk *blocks = NULL;
		rzo_stats st;

		while (chunk >> bits > 0) /* src/rzip.c:1129-1133 */
			bits++;
		cb = bits / 8 + (bits % 8 ? 1 : 0);
		eof = chunk == left;
		pass++;
		if (rzo_rzip_chunk(in + offset, chunk, rzl, cb, &victim_round, &s0, &s0_len, &s1, &s1_len, &st))
			return -2;
		if (sum) {
			sum->matches += st.matches; sum->match_bytes += st.match_bytes;
			sum->literals += st.literals; sum->literal_bytes += st.literal_bytes;
			sum->tag_hits += st.tag_hits; sum->tag_misses += st.tag_misses;
			sum->inserts += st.inserts; sum->lookups += st.lookups;
			sum->chain_evictions += st.chain_evictions; sum->sweeps += st.sweeps;
			sum->displacements += st.displacements; sum->insert_probes += st.insert_probes;
			sum->lookup_probes += st.lookup_probes; sum->probes_ge32 += st.probes_ge32;
			if (st.max_depth > sum->max_depth) sum->max_depth = st.max_depth;
			if (st.max_probe > sum->max_probe) sum->max_probe = st.max_probe;
		}
		if (plan_blocks(s0, s0_len, s1_len, cb, sz.bufsize, &blocks, &nb))
			return -3;

		/* src/stream.c:1737-1769 chunk preamble + two empty stream headers */
		ob_val(&o, cb, 1);
		ob_val(&o, eof, 1);
		ob_val(&o, size_field, cb);
		initial_pos = o.len;
		for (i = 0; i < 2; i++) {
			last_head[i] = cur_pos + 1 + 2 * cb;
			ob_val(&o, CTYPE_NONE, 1);
			ob_val(&o, 0, cb);
			ob_val(&o, 0, cb);
			ob_val(&o, 0, cb);
			cur_pos += 1 + 3 * cb;
		}
		for (i = 0; i < nb; i++) { /* src/stream.c:1571-1572, 1633-1650, 1774-1821 */
			const uint8_t *src = (blocks[i].stream ? s1 : s0) + blocks[i].off;
			const uint8_t *payload = src;
			int64_t u_len = blocks[i].len, c_len = u_len;
			int c_type = CTYPE_NONE;
			if (fn && p->backend != RZO_BACKEND_NONE && u_len >= 64) {
				if (scratch_cap < u_len + 65536) {
					free(scratch);
					scratch_cap = u_len + 65536;
					scratch = malloc((size_t)scratch_cap);
					if (!scratch)
						return -2;
				}
				if (fn(user, src, u_len, blocks[i].stream, scratch, scratch_cap, &c_len, &c_type))
					return -4;
				if (c_type != CTYPE_NONE)
					payload = scratch;
			}
			patch_val(o.p + initial_pos + last_head[blocks[i].stream], cur_pos, cb);
			last_head[blocks[i].stream] = cur_pos + 1 + 2 * cb;
			ob_val(&o, c_type, 1);
			ob_val(&o, c_len, cb);
			ob_val(&o, u_len, cb);
			ob_val(&o, 0, cb);
			if (ob_put(&o, payload, c_len))
				return -2;
			cur_pos += 1 + 3 * cb + c_len;
		}
		free(blocks);
		rzo_free(s0);
		rzo_free(s1);
		left -= chunk;
	}
	free(scratch);

	rzo_md5(in, n, md5); /* src/rzip.c:1195-1218 */
	if (ob_put(&o, md5, 16))
		return -2;

	/* src/lrzip.c:131-208 write_magic */
	memcpy(magic, "LRZI", 4);
	magic[4] = 0;
	magic[5] = 14;
	patch_val(magic + 6, n, 8);
	magic[14] = 1; /* MD5 */
	if (p->backend == RZO_BACKEND_LZMA) {
		magic[17] = 1;
		magic[18] = (uint8_t)lzma2_prop_from_dic(sz.dict_size);
	} else if (p->backend == RZO_BACKEND_ZSTD) {
		static const int zl[10] = { -1, 2, 4, 5, 7, 12, 15, 17, 18, 22 }; /* src/main.c:87 */
		magic[17] = (uint8_t)((p->level << 4) + 4);
		magic[18] = (uint8_t)zl[p->level];
	}
	magic[19] = (uint8_t)((rzl << 4) + p->level);
	memcpy(o.p, magic, 21);
	*out = o.p;
	*out_len = o.len;
	return 0;
}
