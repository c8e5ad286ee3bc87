// kernels.h -- host launchers of the sm_100a kernels (one .cu file per kernel family).
#pragma once
#if defined(LRZ_SIMT_HOST) // tests/hostsim: kernels compiled for the CPU under the SIMT emulator (simt.h)
#include "simt.h"
#else
#include <cuda_runtime.h>
#endif
// A kernel launch that the emulator can take over (host orchestration that is worth running in the CPU tests as it is):
// on the device exactly `kernel<<<grid, block, smem, stream>>>(args...)`.
#if !defined(LRZ_LAUNCH)
#if defined(LRZ_SIMT_HOST)
#define LRZ_LAUNCH(grid, block, smem, stream, kernel, ...) \
	((void)(stream), (void)simt::run_grid(dim3(grid), (int)(block), [&]() { kernel(__VA_ARGS__); }, 64 << 10, (size_t)(smem)))
#else
#define LRZ_LAUNCH(grid, block, smem, stream, kernel, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#endif
#include "lrz_common.h"

namespace lrz {

// ---- K1 tag scan (k1_tagscan.cu) ------------------------------------------------------------
int k1_init_tables();
int k1_preload(); // load the kernels' code now (see k1_tagscan.cu)
int k2_preload();
int k4_preload();
// Scan positions [pos_lo, pos_hi) (pos_lo tile-aligned, kTile = 512) of the n-byte chunk at d_buf (followed by at
// least kInputPad readable bytes) and write the candidates with (tag & mask) == mask, tile-strided:
// candidates of tile T (positions [T*512, (T+1)*512)) start at d_cand[(T - pos_lo/512) * 512].
// When d_state is non-null the mask is read on the device from d_state[0..nstates)->min_mask (the loosest of
// them; so the launch needs no host round trip) and the launch is skipped if every scan already moved past pos_hi.
int k1_launch(const uint8_t *d_buf, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask,
	      const ScanState *d_state, int nstates, Cand *d_cand, uint32_t *d_tile_count, int num_sms, cudaStream_t stream);

// ---- K2 commit (k2_commit.cu) ---------------------------------------------------------------
// nvar > 1: all-values speculation of the reference's cross-window counter (victim_round, src/rzip.c:308).  CTA v
// commits the same candidates into its own table / state / record array (d_state[v], d_tab + v * tab_stride
// entries, d_recs + v * rec_stride records); the variants differ only in the counter value they started from.
int k2_launch(const uint8_t *d_buf, ScanState *d_state, HEntry *d_tab, const Cand *d_cand,
	      const uint32_t *d_tile_count, int64_t pos_lo, int64_t pos_hi, MatchRec *d_recs, bool last_segment,
	      int nvar, int64_t tab_stride, int64_t rec_stride, cudaStream_t stream);

// ---- K4 emit + CRC (k4_emit.cu) -------------------------------------------------------------
int k4_init_tables();
// CRC-32 (IEEE, as gcrypt GCRY_MD_CRC32 / zlib) of d_buf[0..n): *d_crc must be zeroed by the caller
// (crc32_launch does it on `stream`); the finished value is *d_crc ^ 0xFFFFFFFF, applied by k4.
int crc32_launch(const uint8_t *d_buf, int64_t n, uint32_t *d_crc, int num_sms, cudaStream_t stream);
// Stream 0 (record headers, terminator, CRC) from the match records.
int k4_headers_launch(const MatchRec *d_recs, int64_t n_rec, int chunk_bytes, const uint32_t *d_crc,
		      uint8_t *d_s0, cudaStream_t stream);
// Stream 1 (literal bytes) gathered from the chunk: bytes [s1_from, s1_len) of the stream (whole tiles: bytes
// before s1_from in its tile are written again with the same values), from the records [0, n_rec) known so far.
int k4_literals_launch(const uint8_t *d_buf, const MatchRec *d_recs, int64_t n_rec, int64_t s1_from, int64_t s1_len,
		       uint8_t *d_s1, int num_sms, cudaStream_t stream);
// For each stream-0 block boundary j (byte offset (j+1)*bufsize), the number of stream-1 bytes that
// had been written when that byte was written: decides the global flush order of blocks.
int k4_flush_order_launch(const MatchRec *d_recs, int64_t n_rec, int chunk_bytes, int64_t bufsize,
			  int64_t n_bounds, int64_t *d_w1, cudaStream_t stream);

// ---- decode side (unrzip.cu) -------------------------------------------------------
struct DecLit {   // one literal record of stream 0: `len` bytes of stream 1 from lit_off go to out_off
	int64_t out_off, lit_off, len;
};
struct DecMatch { // one match record: `len` bytes copied from `dist` back
	int64_t out_off, len, dist;
};
struct DecSummary {
	int64_t n_lit, n_match, out_len, lit_len;
	uint32_t crc; // the chunk CRC stored behind the terminator
	int32_t status; // 0 ok, < 0 malformed stream
};
struct LzmaDecJob {
	const uint8_t *src;
	int64_t c_len;
	uint8_t *out;
	int64_t u_len;
	int64_t produced;
	int32_t status;
	int32_t pad;
};
// stream 0 -> literal / match records with their output offsets (one thread: a running sum over 3..3+cb byte records)
int unrzip_parse_launch(const uint8_t *d_s0, int64_t s0_len, int cb, int64_t chunk_size, DecLit *d_lits, DecMatch *d_matches,
			int64_t cap, DecSummary *d_sum, cudaStream_t stream);
// literals scattered by all SMs, then the matches in record order by one CTA
int unrzip_replay_launch(const uint8_t *d_s1, int64_t s1_len, const DecLit *d_lits, int64_t n_lit, const DecMatch *d_matches,
			 int64_t n_match, uint8_t *d_out, int num_sms, cudaStream_t stream);
// ---- pre-compression filters of the stream-1 blocks (filters.cu) -----------------------------------
// Converts in place the stream blocks that make up s[from, to) (from is a multiple of the block size bs); `side` is
// scratch of filter_side_bytes() bytes (Delta only).
size_t filter_side_bytes(int filter, int64_t span, int64_t bs);
int filter_blocks_launch(int filter, int delta, uint8_t *s, int64_t from, int64_t to, int64_t bs, uint8_t *side,
			 cudaStream_t stream, int64_t *launches, bool enc = true); // enc false: the decode side's inverse
int filter_preload();
size_t lzma_dec_prob_bytes(int njobs);
int lzma_dec_launch(LzmaDecJob *d_jobs, int njobs, void *d_probs, cudaStream_t stream);
// zstd frames (CTYPE_ZSTD blocks): same job record, `d_work` = zstd_dec_work_bytes(njobs) bytes of scratch
size_t zstd_dec_work_bytes(int njobs);
int zstd_dec_launch(LzmaDecJob *d_jobs, int njobs, void *d_work, cudaStream_t stream);
int unrzip_preload();

} // namespace lrz
