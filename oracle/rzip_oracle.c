/* oracle/rzip_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain-C, single-threaded CPU restatement of lrzip-next's rzip pre-processor for ONE
 * chunk (window): rolling 31-byte XOR tag, open-addressed hash table with bitness-ordered
 * displacement / sweep cleaning / round-robin chain victims, greedy match selection and the
 * stream-0 / stream-1 record encoding.  It exists so that tests/ and __graft_entry__.smoke()
 * can check the CUDA path bit-for-bit, and so bench.py has a "port" CPU baseline when the
 * compiled reference (oracle/_ref) is unavailable.
 *
 * PARITY PINNING: this restatement is itself checked against the unmodified reference built
 * into oracle/_ref (tests/test_oracle.py) and against the committed golden vectors
 * under tests/golden/ that were produced by that reference (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "rzip_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define MINIMUM_MATCH 31   /* src/include/lrzip_private.h (MINIMUM_MATCH) */
#define GREAT_MATCH 1024   /* src/include/lrzip_private.h (GREAT_MATCH)   */

/* src/rzip.c:67-82  levels[]: {hash table MiB, initial_freq, max_chain_len} */
static const struct { unsigned mb; unsigned initial_freq; unsigned max_chain; } k_levels[10] = {
	{1, 4, 1}, {2, 4, 2}, {4, 4, 2}, {8, 4, 2}, {16, 4, 3},
	{32, 4, 4}, {32, 2, 6}, {64, 1, 16}, {64, 1, 32}, {64, 1, 128},
};

/* ---- hash_index[] ------------------------------------------------------------------------
 * src/rzip.c:765-771 fills hash_index[i] = (random() << 16) ^ random() with the never-seeded
 * glibc random(), i.e. the TYPE_3 additive-feedback generator with seed 1.  Restated here
 * from the published algorithm so the table does not depend on process-global libc state. */
static void glibc_random_seed1(uint32_t *out, int count)
{
	int32_t r[34 + 310 + 2 * 256 + 8];
	int i, n = 0;
	r[0] = 1;
	for (i = 1; i < 31; i++) {
		int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
		if (v < 0)
			v += 2147483647LL;
		r[i] = (int32_t)v;
	}
	for (i = 31; i < 34; i++)
		r[i] = r[i - 31];
	for (i = 34; i < 344; i++)
		r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
	for (i = 344; n < count; i++, n++) {
		r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
		out[n] = ((uint32_t)r[i]) >> 1;
	}
}

void rzo_hash_index(int64_t hi[256])
{
	uint32_t v[512];
	int i;
	glibc_random_seed1(v, 512);
	for (i = 0; i < 256; i++)
		hi[i] = ((int64_t)v[2 * i] << 16) ^ (int64_t)v[2 * i + 1];
}

/* ---- growable byte sink --------------------------------------------------------------- */
typedef struct { uint8_t *p; int64_t len, cap; } sink;

static int sink_put(sink *s, const void *src, int64_t n)
{
	if (s->len + n > s->cap) {
		int64_t ncap = s->cap ? s->cap * 2 : 4096;
		uint8_t *np;
		while (ncap < s->len + n)
			ncap *= 2;
		np = realloc(s->p, (size_t)ncap);
		if (!np)
			return -1;
		s->p = np;
		s->cap = ncap;
	}
	memcpy(s->p + s->len, src, (size_t)n);
	s->len += n;
	return 0;
}

/* ---- per-chunk state (src/include/lrzip_private.h:441-469 struct rzip_state) ------------ */
typedef struct { int64_t offset; int64_t t; } hentry; /* src/rzip.c:61-64 */

typedef struct {
	const uint8_t *buf;
	int64_t n;
	int64_t hi[256];
	hentry *tab;
	int bits;
	int64_t hash_count, hash_limit;
	int64_t min_mask;
	int64_t clean_ptr;
	int64_t last_match;
	unsigned max_chain;
	int64_t *victim_round; /* the reference's function-static counter, src/rzip.c:308 */
	int chunk_bytes;
	sink s0, s1;
	rzo_stats st;
	int oom;
} rz;

static inline int empty(const hentry *he) { return !(he->offset | he->t); } /* src/rzip.c:268-271 */

/* src/rzip.c:295-300 lesser_bitness: ffsll(~a) < ffsll(~b), i.e. a has fewer trailing 1 bits */
static inline int lesser_bitness(int64_t a, int64_t b)
{
	return __builtin_ffsll(~a) < __builtin_ffsll(~b);
}

/* src/rzip.c:304-353 insert_hash */
static void insert_hash_d(rz *z, int64_t t, int64_t offset, int depth);
static void insert_hash(rz *z, int64_t t, int64_t offset) { insert_hash_d(z, t, offset, 0); }
static void insert_hash_d(rz *z, int64_t t, int64_t offset, int depth)
{
	const int64_t mask = ((int64_t)1 << z->bits) - 1;
	const int64_t better = (z->min_mask << 1) | 1; /* :284-291 minimum_bitness */
	int64_t h = t & mask, victim_h = 0, round = 0;
	hentry *he = &z->tab[h];

	while (!empty(he)) {
		if ((he->t & better) != better) { /* due for cleaning: just replace (:316-319) */
			z->hash_count--;
			break;
		}
		if (lesser_bitness(he->t, t)) { /* re-home the weaker occupant, take its slot (:324-328) */
			z->st.displacements++;
			if (depth + 1 > z->st.max_depth)
				z->st.max_depth = depth + 1;
			insert_hash_d(z, he->t, he->offset, depth + 1);
			break;
		}
		if (he->t == t) { /* equal-tag chain cap with round-robin victim (:332-343) */
			if (round == *z->victim_round)
				victim_h = h;
			if (++round == (int64_t)z->max_chain) {
				h = victim_h;
				he = &z->tab[h];
				z->hash_count--;
				if (++*z->victim_round == (int64_t)z->max_chain)
					*z->victim_round = 0;
				z->st.chain_evictions++;
				break;
			}
		}
		h = (h + 1) & mask;
		he = &z->tab[h];
		z->st.insert_probes++;
	}
	he->t = t;
	he->offset = offset;
}

/* src/rzip.c:357-383 clean_one_from_hash: delete the next entry (sweep order) whose tag lacks
 * the next-stricter mask; on wrap-around promote the mask.  Returns the new insert mask. */
static int64_t clean_one(rz *z)
{
	const int64_t size = (int64_t)1 << z->bits;
	for (;;) {
		const int64_t better = (z->min_mask << 1) | 1;
		for (; z->clean_ptr < size; z->clean_ptr++) {
			hentry *he = &z->tab[z->clean_ptr];
			if (empty(he))
				continue;
			if ((he->t & better) != better) {
				he->offset = 0;
				he->t = 0;
				z->hash_count--;
				return better;
			}
		}
		z->min_mask = better;
		z->clean_ptr = 0;
		z->st.sweeps++;
	}
}

/* src/rzip.c:431-461 single_match_len */
static int64_t match_len(const rz *z, int64_t p0, int64_t op, int64_t end, int64_t *rev)
{
	const uint8_t *b = z->buf;
	int64_t p = p0, len, lo;

	if (op >= p0)
		return 0;
	while (p < end && b[p] == b[op]) {
		p++;
		op++;
	}
	len = p - p0;
	p = p0;
	op -= len;
	lo = z->last_match > 0 ? z->last_match : 0;
	while (p > lo && op > 0 && b[op - 1] == b[p - 1]) {
		op--;
		p--;
	}
	*rev = p0 - p;
	len += *rev;
	return len < MINIMUM_MATCH ? 0 : len;
}

/* src/rzip.c:495-534 find_best_match */
static int64_t find_best_match(rz *z, int64_t t, int64_t p, int64_t end, int64_t *offset, int64_t *reverse)
{
	const int64_t mask = ((int64_t)1 << z->bits) - 1;
	int64_t h = t & mask, length = 0, rev = 0, nprobe = 0;
	const hentry *he = &z->tab[h];

	*reverse = 0;
	while (!empty(he)) {
		if (he->t == t) {
			int64_t mlen = match_len(z, p, he->offset, end, &rev);
			if (mlen) {
				if (mlen > length) {
					length = mlen;
					*offset = he->offset - rev;
					*reverse = rev;
				}
				z->st.tag_hits++;
			} else
				z->st.tag_misses++;
		}
		h = (h + 1) & mask;
		he = &z->tab[h];
		nprobe++;
	}
	z->st.lookup_probes += nprobe;
	if (nprobe > z->st.max_probe)
		z->st.max_probe = nprobe;
	if (nprobe >= 32)
		z->st.probes_ge32++;
	return length;
}

/* src/rzip.c:202-206 put_header: u8 head, u16 length (LE) on stream 0 */
static void put_header(rz *z, uint8_t head, int64_t len)
{
	uint8_t b[3] = { head, (uint8_t)(len & 0xff), (uint8_t)((len >> 8) & 0xff) };
	if (sink_put(&z->s0, b, 3))
		z->oom = 1;
}

/* src/rzip.c:208-226 put_match */
static void put_match(rz *z, int64_t p, int64_t offset, int64_t len)
{
	do {
		int64_t n = len > 0xFFFF ? 0xFFFF : len;
		uint64_t ofs = (uint64_t)(p - offset);
		uint8_t b[8];
		int i;
		put_header(z, 1, n);
		for (i = 0; i < 8; i++)
			b[i] = (uint8_t)(ofs >> (8 * i));
		if (sink_put(&z->s0, b, z->chunk_bytes)) /* :196-200 put_vchars */
			z->oom = 1;
		z->st.matches++;
		z->st.match_bytes += n;
		len -= n;
		p += n;
		offset += n;
	} while (len);
}

/* src/rzip.c:248-265 put_literal (+ :229-246 write_sbstream: literal bytes go to stream 1) */
static void put_literal(rz *z, int64_t last, int64_t p)
{
	do {
		int64_t len = p - last;
		if (len > 0xFFFF)
			len = 0xFFFF;
		z->st.literals++;
		z->st.literal_bytes += len;
		put_header(z, 0, len);
		if (len && sink_put(&z->s1, z->buf + last, len))
			z->oom = 1;
		last += len;
	} while (p > last);
}

static int64_t full_tag(const rz *z, int64_t p) /* src/rzip.c:405-416 */
{
	int64_t t = 0;
	int i;
	for (i = 0; i < MINIMUM_MATCH; i++)
		t ^= z->hi[z->buf[p + i]];
	return t;
}

/* src/rzip.c:586-762 hash_search, for one chunk held fully in memory */
int rzo_rzip_chunk(const uint8_t *buf, int64_t n, int rzip_level, int chunk_bytes,
		   int64_t *victim_round, uint8_t **s0, int64_t *s0_len,
		   uint8_t **s1, int64_t *s1_len, rzo_stats *stats)
{
	rz z;
	int64_t hashsize, p, end, t = 0, tag_mask;
	struct { int64_t p, ofs, len; } cur;
	uint32_t crc;
	uint8_t crcb[4];
	int64_t vr_local = 0;

	if (rzip_level < 0 || rzip_level > 9 || chunk_bytes < 1 || chunk_bytes > 8)
		return -1;
	memset(&z, 0, sizeof(z));
	z.buf = buf;
	z.n = n;
	z.chunk_bytes = chunk_bytes;
	z.max_chain = k_levels[rzip_level].max_chain;
	z.victim_round = victim_round ? victim_round : &vr_local;
	rzo_hash_index(z.hi);

	/* :602-611 table size and 66 % limit */
	hashsize = (int64_t)k_levels[rzip_level].mb * (1048576 / (int64_t)sizeof(hentry));
	for (z.bits = 0; ((int64_t)1 << z.bits) < hashsize; z.bits++)
		;
	z.hash_limit = ((int64_t)1 << z.bits) / 3 * 2;
	z.tab = calloc((size_t)1 << z.bits, sizeof(hentry));
	if (!z.tab)
		return -2;

	tag_mask = ((int64_t)1 << k_levels[rzip_level].initial_freq) - 1; /* :590 */
	z.min_mask = tag_mask;                                            /* :616-619 */
	p = 0;
	end = n - MINIMUM_MATCH;
	cur.p = 0;
	cur.ofs = 0;
	cur.len = 0;
	if (end > 0)
		t = full_tag(&z, 0);

	while (p < end) { /* :631-705 */
		int64_t reverse, mlen, offset = 0;

		++p;
		t ^= z.hi[buf[p - 1]] ^ z.hi[buf[p + MINIMUM_MATCH - 1]]; /* :385-393 next_tag */
		if ((t & z.min_mask) != z.min_mask)
			continue;
		z.st.lookups++;
		mlen = find_best_match(&z, t, p, end, &offset, &reverse);
		if ((t & tag_mask) == tag_mask) {
			z.st.inserts++;
			z.hash_count++;
			insert_hash(&z, t, p);
			if (z.hash_count > z.hash_limit)
				tag_mask = clean_one(&z);
		}
		if (mlen > cur.len) {
			cur.p = p - reverse;
			cur.len = mlen;
			cur.ofs = offset;
		}
		if ((cur.len >= GREAT_MATCH || p >= cur.p + MINIMUM_MATCH) && cur.len >= MINIMUM_MATCH) {
			if (z.last_match < cur.p)
				put_literal(&z, z.last_match, cur.p);
			put_match(&z, cur.p, cur.ofs, cur.len);
			z.last_match = cur.p + cur.len;
			cur.p = p = z.last_match;
			cur.len = 0;
			t = full_tag(&z, p);
		}
	}
	if (z.last_match < n) /* :710-711 */
		put_literal(&z, z.last_match, n);

	/* :713-760 chunk CRC32 (gcrypt digest byte order = big endian), terminator record */
	crc = (uint32_t)crc32(0L, Z_NULL, 0);
	{
		int64_t done = 0;
		while (done < n) {
			int64_t m = n - done > (1 << 30) ? (1 << 30) : n - done;
			crc = (uint32_t)crc32(crc, buf + done, (uInt)m);
			done += m;
		}
	}
	put_literal(&z, 0, 0);
	crcb[0] = (uint8_t)(crc >> 24);
	crcb[1] = (uint8_t)(crc >> 16);
	crcb[2] = (uint8_t)(crc >> 8);
	crcb[3] = (uint8_t)crc;
	if (sink_put(&z.s0, crcb, 4))
		z.oom = 1;

	z.st.hash_count = z.hash_count;
	z.st.final_min_mask = z.min_mask;
	z.st.final_tag_mask = tag_mask;
	z.st.crc32 = crc;
	free(z.tab);
	if (z.oom) {
		free(z.s0.p);
		free(z.s1.p);
		return -2;
	}
	*s0 = z.s0.p;
	*s0_len = z.s0.len;
	*s1 = z.s1.p;
	*s1_len = z.s1.len;
	if (stats)
		*stats = z.st;
	return 0;
}

void rzo_free(void *p) { free(p); }

/* Tag of the 31-byte window at p (src/rzip.c:405-416), exposed for kernel tag-scan tests. */
int64_t rzo_full_tag(const uint8_t *buf, int64_t p)
{
	static int64_t hi[256];
	static int init;
	int64_t t = 0;
	int i;
	if (!init) {
		rzo_hash_index(hi);
		init = 1;
	}
	for (i = 0; i < MINIMUM_MATCH; i++)
		t ^= hi[buf[p + i]];
	return t;
}
