/* lrzgpu.h -- C ABI of liblrzgpu.so: a B200 (sm_100a) implementation of lrzip-next's compress hot
 * path (rzip long-range pre-processor -> stream/block framing -> per-block backend), writing
 * lrzip-next v0.14 archives.
 *
 * lrzip-next has no library ABI of its own (its liblrzip was abandoned, src/libdemo/README); the
 * seams this header replaces are internal C functions.  Each entry point names the reference
 * interface it stands in for (paths relative to the lrzip-next tree):
 *
 *   lrzgpu_compress / _file / _device   rzip_fd()            src/include/rzip.h:12, src/rzip.c:922
 *                                       + write_magic()      src/lrzip.c:131 (called from
 *                                                            compress_file(), src/lrzip.c:1549-1553)
 *   lrzgpu_compress_multi               the same with the windows of the file dealt to several GPUs
 *   lrzgpu_compress_chunk               one pass of the chunk loop src/rzip.c:1041-1186
 *   lrzgpu_chunk_begin / _finish        (rzip_chunk :873 + close_stream_out, src/stream.c:2253); the two-call form
 *                                       splits it at the point where rzip_chunk returns
 *   lrzgpu_chunk_begin_all / _select    the same window for every value of insert_hash()'s static victim_round
 *                                       (src/rzip.c:308), so that windows need not wait for their predecessor
 *   lrzgpu_decompress                   runzip_fd()          src/runzip.c:372 (the inverse path, 8(f1))
 *   lrzgpu_info                         get_fileinfo()       src/lrzip.c:1069 (`-i`, 8(f4))
 *   lrzgpu_rzip_chunk                   hash_search()        src/rzip.c:586 with the scan primitives
 *                                       full_tag/next_tag/match_len, lrzip_private.h:573-576
 *   lrzgpu_tag_scan                     single_full_tag()/single_next_tag()  src/rzip.c:385-416
 *   lrzgpu_crc32                        cksum_update()/cksumthread()  src/rzip.c:564-584, 713-757
 *   lrzgpu_sizing                       setup_overhead/setup_ram src/util.c:103-188,
 *                                       open_stream_out sizing  src/stream.c:1169-1331
 *   lrzgpu_block_compress               lzma_compress_buf()/zstd_compress_buf()  src/stream.c:429/167
 *                                       (struct compress_thread {s_buf,c_type,s_len,c_len}, :67-76)
 *   lrzgpu_lz4_gate                     lz4_compresses()     src/stream.c:2325-2380
 *
 * Conventions: plain pointers and sizes, caller-owned inputs, outputs allocated by the library and
 * released with lrzgpu_free(); every function returns 0 on success or a negative LRZGPU_E* code and
 * never calls exit() (the reference calls fatal(), src/include/util.h:34-46).  A context is bound to
 * one CUDA device and is not thread-safe; use one context per thread / per GPU.
 * There is no CPU fallback: without a CUDA device lrzgpu_create() fails with LRZGPU_ENODEV.
 */
#ifndef LRZGPU_H
#define LRZGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
	LRZGPU_OK = 0,
	LRZGPU_EINVAL = -1,
	LRZGPU_ENOMEM = -2,
	LRZGPU_ECUDA = -3,
	LRZGPU_ENODEV = -4,
	LRZGPU_EIO = -5,
	LRZGPU_EINTERNAL = -6,
	LRZGPU_EUNSUPPORTED = -7,
};

/* c_type byte of a block header, src/include/lrzip_private.h:287-295 */
enum { LRZGPU_CTYPE_NONE = 3, LRZGPU_CTYPE_LZMA = 6, LRZGPU_CTYPE_ZSTD = 10 };

/* pre-compression filter of the stream-1 blocks (control->filter_flag, src/include/lrzip_private.h:389-397;
 * applied in compthread, src/stream.c:1587-1628).  All of them are built (RISC-V: z7_BranchConv_RISCV_Enc,
 * src/lzma/C/Bra.c:423-720). */
enum {
	LRZGPU_FILTER_NONE = 0, LRZGPU_FILTER_X86 = 1, LRZGPU_FILTER_ARM = 2, LRZGPU_FILTER_ARMT = 3, LRZGPU_FILTER_PPC = 4,
	LRZGPU_FILTER_SPARC = 5, LRZGPU_FILTER_IA64 = 6, LRZGPU_FILTER_ARM64 = 7, LRZGPU_FILTER_RISCV = 8, LRZGPU_FILTER_DELTA = 128
};

/* backend selector (the reference's FLAG_NO_COMPRESS / default lzma / FLAG_ZSTD_COMPRESS) */
enum { LRZGPU_BACKEND_NONE = 0, LRZGPU_BACKEND_LZMA = 1, LRZGPU_BACKEND_ZSTD = 4 };

/* Every option that changes archive bytes (rzip_control fields, src/include/lrzip_private.h:472-581). */
typedef struct lrzgpu_params {
	int level;        /* -L 1..9 (compression_level), default 7 */
	int rzip_level;   /* -R, 0 = same as level (rzip_compression_level) */
	int backend;      /* LRZGPU_BACKEND_* */
	int threads;      /* -p (control->threads before prepare_streamout_threads) */
	int window;       /* -w, units of 100 MiB, 0 = unset */
	int unlimited;    /* -U */
	int64_t ramsize;  /* control->ramsize in bytes (-m N  => N * 100 MiB) */
	int page_size;    /* control->page_size, 4096 (must be a multiple of 16) */
	int processors;   /* sysconf(_SC_NPROCESSORS_ONLN) of the machine being reproduced */
	int threshold;    /* lz4 gate: 0 = off (-T), else percent (default 100) */
	int nobemt;       /* --nobemt: with LZMA level >= 5 and threads > 1 the reference switches to the single-threaded
			     bt4 finder (src/stream.c:456); not reproduced => LRZGPU_EUNSUPPORTED */
	int filter;       /* LRZGPU_FILTER_* (--x86 --arm --armt --arm64 --ppc --sparc --ia64 --riscv --delta), 0 = none */
	int delta;        /* --delta: the distance in bytes, 1..16 or a multiple of 16 up to 256 (control->delta) */
	int stdin_mode;   /* reproduce `... | lrzip-next -o out`: the input's size is unknown while it is read, so every
			     chunk is one mmap buffer of min(ramsize / 3, max_chunk) bytes (src/rzip.c:995-1013, 800-836) and the
			     block size is derived from the first chunk alone; not with -U */
} lrzgpu_params;

typedef struct lrzgpu_sizing_t {
	int threads;        /* after prepare_streamout_threads / open_stream_out */
	uint32_t dict_size; /* lzma dictionary after possible reduction (magic byte 18) */
	int64_t overhead;
	int64_t bufsize;    /* stream block size (stream_bufsize) */
	int64_t max_chunk;  /* rzip window */
} lrzgpu_sizing_t;

/* rzip statistics, src/rzip.c:1238-1246, summed over chunks, plus stage timings */
typedef struct lrzgpu_stats {
	int64_t matches, match_bytes, literals, literal_bytes;
	int64_t tag_hits, tag_misses, inserts, lookups;
	int64_t chain_evictions, sweeps, displacements;
	int64_t hash_count, final_min_mask, final_tag_mask; /* of the last chunk */
	int64_t chunks, blocks, blocks_stored;
	int64_t stream0_bytes, stream1_bytes;
	uint32_t crc32;      /* of the last chunk */
	uint32_t pad;
	double ms_h2d, ms_rzip, ms_emit, ms_backend, ms_d2h, ms_md5, ms_total;
	int64_t kernel_launches;
} lrzgpu_stats;

typedef struct lrzgpu_ctx lrzgpu_ctx;

int lrzgpu_create(int device, lrzgpu_ctx **ctx);
void lrzgpu_destroy(lrzgpu_ctx *ctx);
const char *lrzgpu_last_error(const lrzgpu_ctx *ctx);
void lrzgpu_free(void *p);
const char *lrzgpu_version(void);

/* Host-side sizing only (no device work): block size, threads, dictionary, window. */
int lrzgpu_sizing(const lrzgpu_params *p, int64_t st_size, lrzgpu_sizing_t *out);

/* Whole file, host buffers: in[0..n) -> malloc'ed .lrz archive. */
int lrzgpu_compress(lrzgpu_ctx *ctx, const lrzgpu_params *p, const uint8_t *in, int64_t n,
		    uint8_t **out, int64_t *out_len, lrzgpu_stats *stats);
/* Whole file, file -> file. */
int lrzgpu_compress_file(lrzgpu_ctx *ctx, const lrzgpu_params *p, const char *in_path, const char *out_path,
			 lrzgpu_stats *stats);
/* Whole file, input already resident in device memory: d_in must be 16-byte aligned, preceded by
 * at least 32 readable bytes and followed by at least 8192 readable bytes.  The trailing MD5 is
 * computed by a host thread from a device->host stream of the input unless md5 is given. */
int lrzgpu_compress_device(lrzgpu_ctx *ctx, const lrzgpu_params *p, const void *d_in, int64_t n,
			   const uint8_t *md5_or_null, uint8_t **out, int64_t *out_len, lrzgpu_stats *stats);

/* Whole file on several GPUs of one box from ONE process (what a patched lrzip-next would call; one-process-per-GPU
 * callers use the chunk entry points below and gather the blobs themselves, lrzip_next_b200/multigpu.py).  The file's
 * rzip windows (src/rzip.c:1041-1186; more than one window needs -w or a file larger than 2/3 of the RAM) are dealt to
 * the contexts round-robin, ctxs[i mod nctx]; every window after the first is scanned for all values of the
 * cross-window counter at once (lrzgpu_chunk_begin_all), so the windows run concurrently; blobs are appended in
 * file order between magic and MD5.  Contexts should sit on distinct devices (the same device works, for tests).
 * The archive is byte-identical to lrzgpu_compress() with the same parameters. */
int lrzgpu_compress_multi(lrzgpu_ctx **ctxs, int nctx, const lrzgpu_params *p, const uint8_t *in, int64_t n,
			  uint8_t **out, int64_t *out_len, lrzgpu_stats *stats);

/* One chunk (window) -> its position-independent blob (chunk preamble, stream headers, blocks), the
 * unit that is sharded across GPUs.  chunk_bytes/eof as in src/rzip.c:1129-1143; victim_round carries
 * the reference's static counter (src/rzip.c:308) in and out; *chain_evictions reports whether it was
 * consulted. */
int lrzgpu_compress_chunk(lrzgpu_ctx *ctx, const lrzgpu_params *p, const lrzgpu_sizing_t *sz,
			  const uint8_t *in, int64_t n, int eof, int64_t *victim_round,
			  uint8_t **blob, int64_t *blob_len, lrzgpu_stats *stats);

/* The same window in two calls, for chained windows (multi-GPU): _begin runs the rzip stage and returns the
 * outgoing victim_round as soon as the commit is done -- the only thing the NEXT window needs, so its rank can
 * start while this one is still in the backend; _finish runs blocks -> backend -> framing of the pending
 * window and returns the same blob lrzgpu_compress_chunk would.  One pending window per context. */
int lrzgpu_chunk_begin(lrzgpu_ctx *ctx, const lrzgpu_params *p, const lrzgpu_sizing_t *sz, const uint8_t *in, int64_t n,
		       int eof, int64_t *victim_round, lrzgpu_stats *stats);
int lrzgpu_chunk_finish(lrzgpu_ctx *ctx, uint8_t **blob, int64_t *blob_len, lrzgpu_stats *stats);

/* All-values speculation of the cross-window counter, so that the windows of one file are independent and can
 * run on different GPUs at the same time.  insert_hash()'s function-static victim_round (src/rzip.c:308, 332-343)
 * is the only state that leaks from one window into the next, and it takes max_chain_len values
 * (lrzgpu_victim_values(): 16 at rzip level 7).  The commit stage of a window occupies one SM of 148, so
 * _begin_all runs the window once per possible incoming value, concurrently, and reports victim_out[v] = the
 * counter after the window when it started at v.  Once the true incoming value is known (from the predecessor's
 * table: v_i = victim_out_{i-1}[v_{i-1}], v_0 = 0), _select emits the streams of that variant and
 * lrzgpu_chunk_finish() produces the blob -- byte-identical to lrzgpu_compress_chunk(victim_round = v). */
int lrzgpu_victim_values(const lrzgpu_params *p);
int lrzgpu_chunk_begin_all(lrzgpu_ctx *ctx, const lrzgpu_params *p, const lrzgpu_sizing_t *sz, const uint8_t *in, int64_t n,
			   int eof, int64_t *victim_out, int nvalues, lrzgpu_stats *stats);
int lrzgpu_chunk_select(lrzgpu_ctx *ctx, int64_t victim_in, lrzgpu_stats *stats);

/* Decode (runzip_fd / runzip_chunk, src/runzip.c:261-470; block chains src/stream.c:1883-2016): a whole archive in
 * host memory -> the original bytes (malloc'ed).  Container walk on the host; LZMA blocks (lzma_decompress_buf,
 * src/stream.c:556-616) decoded on the device, one thread per block; stream 0 parsed into records and replayed on
 * the device (literals scattered by all SMs, matches in order); chunk CRC-32 and the trailing MD5 are verified.
 * zstd blocks (zstd_decompress_buf, src/stream.c:1989-2010) are decoded on the device too, one thread per frame.
 * Stored, LZMA and zstd blocks (the other back ends: LRZGPU_EUNSUPPORTED); filtered archives are unfiltered per
 * stream-1 block (every filter of the reference); no encryption. */
int lrzgpu_decompress(lrzgpu_ctx *ctx, const uint8_t *archive, int64_t archive_len, uint8_t **out, int64_t *out_len);

/* Archive walker (get_fileinfo, src/lrzip.c:1069-1459, what `lrzip-next -i [-vv]` prints): host only, no device.
 * Walks magic, chunks, the two streams' block chains and the trailing MD5 of an archive in memory, fills `info`
 * and, when `blocks` is given, up to `cap` block records in the order the reference lists them (per chunk: stream 0's
 * chain, then stream 1's).  *nblocks receives the number of blocks in the archive.  LRZGPU_EINVAL on a malformed
 * container, LRZGPU_EUNSUPPORTED for encrypted archives (their block headers are encrypted). */
typedef struct lrzgpu_archive_info {
	int major, minor;          /* magic bytes 4-5 */
	int64_t expected_size;     /* magic bytes 6-13 (st_size) */
	int hash_type;             /* magic byte 14: 1 = MD5 ... */
	int encrypted;             /* magic byte 15 */
	int filter, delta;         /* magic byte 16 decoded: LRZGPU_FILTER_*, delta distance */
	int backend_code;          /* magic byte 17 & 15: 0 none, 1 lzma, 4 zstd, ... (src/lrzip.c:160-200) */
	int backend_prop;          /* magic byte 18: lzma dictionary code / zstd level */
	uint32_t lzma_dict_size;   /* decoded from backend_prop when backend_code == 1 */
	int rzip_level, level;     /* magic byte 19 */
	int64_t chunks, blocks;
	int64_t stream_c_bytes[2]; /* compressed payload bytes per stream (sum of c_len) */
	int64_t stream_u_bytes[2]; /* uncompressed bytes per stream (sum of u_len) */
	int64_t blocks_by_ctype[16];
	int64_t archive_bytes;
	uint8_t md5[16];           /* trailing digest (hash_type == 1) */
} lrzgpu_archive_info;

typedef struct lrzgpu_block_info {
	int64_t chunk;     /* 0-based */
	int stream;        /* 0 or 1 */
	int ctype;         /* LRZGPU_CTYPE_* and the reference's other codes */
	int64_t c_len, u_len;
	int64_t offset;    /* of the block header in the archive */
	int64_t next_head; /* as stored (relative to the chunk's first stream header), 0 = last of its stream */
} lrzgpu_block_info;

int lrzgpu_info(const uint8_t *archive, int64_t archive_len, lrzgpu_archive_info *info, lrzgpu_block_info *blocks,
		int64_t cap, int64_t *nblocks);

/* rzip of one chunk -> stream 0 / stream 1 bytes (malloc'ed). */
int lrzgpu_rzip_chunk(lrzgpu_ctx *ctx, const uint8_t *in, int64_t n, int rzip_level, int chunk_bytes,
		      int64_t *victim_round, uint8_t **s0, int64_t *s0_len, uint8_t **s1, int64_t *s1_len,
		      lrzgpu_stats *stats);

/* Tag scan of positions [pos_lo, pos_hi) of in[0..n): candidates with (tag & mask) == mask, in
 * position order, into caller arrays of capacity cap; *count receives the total found. */
int lrzgpu_tag_scan(lrzgpu_ctx *ctx, const uint8_t *in, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask,
		    int64_t *out_pos, int64_t *out_tag, int64_t cap, int64_t *count);

int lrzgpu_crc32(lrzgpu_ctx *ctx, const uint8_t *in, int64_t n, uint32_t *crc);

/* One stream block through a backend: out (malloc'ed) holds the payload as it would be written
 * after the block header; *c_type is LRZGPU_CTYPE_NONE when the block is left stored. */
int lrzgpu_block_compress(lrzgpu_ctx *ctx, const lrzgpu_params *p, uint32_t dict_size, const uint8_t *in,
			  int64_t u_len, uint8_t **out, int64_t *c_len, int *c_type);
/* lz4 compressibility gate: *compressible = lz4_compresses() result. */
int lrzgpu_lz4_gate(lrzgpu_ctx *ctx, const uint8_t *in, int64_t len, int threshold, int *compressible);

/* ---- measurement hooks (bench.py): kernels on caller-provided device buffers and stream ------- */
/* K1 over a device-resident chunk: d_cand needs 16 * round_up(n, 512) bytes, d_tile_count
 * 4 * ceil(n / 512) bytes (candidate tiles are 512 positions).  `stream` is a cudaStream_t (0 = default stream). */
int lrzgpu_k1_launch(lrzgpu_ctx *ctx, const void *d_buf, int64_t n, int64_t mask, void *d_cand,
		     void *d_tile_count, void *stream);
int lrzgpu_crc32_launch(lrzgpu_ctx *ctx, const void *d_buf, int64_t n, void *d_crc, void *stream);
int lrzgpu_sm_count(const lrzgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
