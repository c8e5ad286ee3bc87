/* oracle/rzip_oracle.h -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the reference's rzip + .lrz framing; see rzip_oracle.c / lrz_format.c.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may
 * load liboracle.so; the product (liblrzgpu.so) never does. */
#ifndef RZIP_ORACLE_H
#define RZIP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rzo_stats {
	int64_t matches, match_bytes, literals, literal_bytes; /* src/rzip.c:1238-1241 */
	int64_t tag_hits, tag_misses, inserts;                 /* src/rzip.c:1242-1244 */
	int64_t lookups, chain_evictions, sweeps;
	int64_t displacements, max_depth, insert_probes, lookup_probes, max_probe, probes_ge32;
	int64_t hash_count, final_min_mask, final_tag_mask;
	uint32_t crc32;
} rzo_stats;

void rzo_hash_index(int64_t hi[256]);
int64_t rzo_full_tag(const uint8_t *buf, int64_t p);
int rzo_rzip_chunk(const uint8_t *buf, int64_t n, int rzip_level, int chunk_bytes,
		   int64_t *victim_round, uint8_t **s0, int64_t *s0_len,
		   uint8_t **s1, int64_t *s1_len, rzo_stats *stats);
void rzo_free(void *p);

/* ---- .lrz framing (lrz_format.c) ---- */
enum { RZO_BACKEND_NONE = 0, RZO_BACKEND_LZMA = 1, RZO_BACKEND_ZSTD = 4 };

typedef struct rzo_params {
	int level;        /* -L, 1..9 (default 7) */
	int rzip_level;   /* -R, 0 = same as level */
	int backend;      /* RZO_BACKEND_* */
	int threads;      /* -p */
	int window;       /* -w (x 100 MiB), 0 = unset */
	int unlimited;    /* -U */
	int64_t ramsize;  /* -m N  => N * 100 MiB */
	int page_size;    /* 4096 */
	int processors;   /* sysconf(_SC_NPROCESSORS_ONLN) of the machine the reference ran on */
	int threshold;    /* lz4 gate: 0 = off (-T), else percent (default 100) */
	int nobemt;
} rzo_params;

typedef struct rzo_sizing {
	int threads;        /* after prepare_streamout_threads / open_stream_out */
	uint32_t dict_size; /* lzma dictionary after possible reduction */
	int64_t overhead;
	int64_t bufsize;    /* stream block size */
	int64_t max_chunk;  /* rzip window */
} rzo_sizing;

/* block compressor callback: return 0 and leave *c_len = u_len for "stored" */
typedef int (*rzo_block_fn)(void *user, const uint8_t *in, int64_t u_len, int stream,
			    uint8_t *out, int64_t out_cap, int64_t *c_len, int *c_type);

int rzo_sizing_compute(const rzo_params *p, int64_t st_size, rzo_sizing *out);
int rzo_compress(const rzo_params *p, const uint8_t *in, int64_t n, rzo_block_fn fn, void *user,
		 uint8_t **out, int64_t *out_len, rzo_stats *stats_sum);
void rzo_md5(const uint8_t *in, int64_t n, uint8_t digest[16]);

#ifdef __cplusplus
}
#endif
#endif
