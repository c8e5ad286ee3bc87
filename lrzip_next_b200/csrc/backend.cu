// backend.cu -- per-block backends on the device (src/stream.c:1633-1650 compthread -> *_compress_buf).
//
//   lz4 gate   lz4_compresses()   src/stream.c:2325-2380   -> lz4_gate_kernel   (lz4_size.cuh)
//   LZMA       lzma_compress_buf  src/stream.c:429-494     -> lzma_block_kernel (lzma_enc.cuh)
//   zstd       zstd_compress_buf  src/stream.c:167-229     -> zstd_* kernels below
//
// Every stream block of a chunk is an independent job ("each CUDA block owns one rzip output block"):
// the kernels are launched once over all jobs of the chunk.  The LZMA encoder is an adaptive, strictly
// sequential coder per block, so its parallelism is the number of blocks in flight.
//
// zstd: libzstd is not vendored by the reference and its source is not available here, so frames
// byte-identical to ZSTD_compress(level 17) are out of reach ("parity unpinned").  What
// is produced is a valid Zstandard frame (RFC 8878) of genuinely compressed blocks (zstd_enc.cuh: LZ
// sequences from the same data-parallel match finder the LZMA backend uses, FSE-coded with the format's
// predefined tables, Huffman literals), which the reference's ZSTD_decompress() accepts; a frame that is
// not smaller than the block leaves the block stored, like the reference does.
#if !defined(LRZ_SIMT_HOST) // (tests/hostsim builds the kernels of this file for the SIMT emulator, not the pipeline)
#include "backend.h"
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <vector>

#include "lz4_size.cuh"
#include "lzma_enc.cuh"
#include "lzma_mf.h"
#include "zstd_enc.cuh"

namespace lrz {

namespace {

#if !defined(LRZ_SIMT_HOST)
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t n)
	{
		if (n <= cap)
			return cudaSuccess;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc(&p, n + (n >> 5) + 4096);
		if (e == cudaSuccess)
			cap = n + (n >> 5) + 4096;
		return e;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};
#endif

// ---- lz4 gate -------------------------------------------------------------------------------------
struct GateJob {
	const uint8_t *src;
	int64_t len;
	int result;
};

__global__ void __launch_bounds__(32) lz4_gate_kernel(GateJob *jobs, int threshold)
{
	__shared__ uint32_t table[4096];
	if (threadIdx.x == 0) {
		GateJob &j = jobs[blockIdx.x];
		j.result = lz4s::gate(j.src, j.len, threshold, table);
	}
}

// ---- LZMA -----------------------------------------------------------------------------------------
struct LzmaJob {
	const uint8_t *src;
	uint32_t n;
	uint8_t *out;
	uint64_t outCap;
	const uint64_t *rec;  // match lists of the data-parallel finder (lzma_mf.cu)
	const uint32_t *pool;
	lzma::Config cfg;
	int threshold;          // lz4 gate (lz4_compresses, src/stream.c:2325-2380): 0 = off
	const int *mf_overflow; // the match finder ran out of pool for this block: encode it again with a larger one
	uint32_t wait_count;    // > 0: the tree walk runs beside this kernel; records [0, wait_count) carry a "final" bit
	uint64_t pool_cap;      // uint32 in the pool (a record of an overflowed walk may point past it)
	uint64_t outLen;
	int overflow;
	int skipped;            // 1: gate said incompressible, 2: match-list pool overflow, 3: zstd frame not smaller
	// zstd: per 128 KiB zstd block scratch carved from the match finder's arrays the walk no longer needs
	zs::Seq *zseq;          // nzb x (kBlockMax / 3 + 2) sequences           (over `son`)
	uint8_t *zlit;          // nzb x kBlockMax literal bytes                  (over `c2`)
	uint8_t *zstage;        // nzb x (kBlockMax + 64) encoded block contents  (over `c3`)
	uint32_t *zsize;        // [2 * nzb] (type << 28 | payload size), then offsets in the frame  (over `sorted`)
	uint32_t nzb;
	const int *gate_result; // zstd: verdict of the lz4 gate kernel that ran before (device), or null
};

constexpr uint32_t kZsSeqPerBlock = zs::kBlockMax / 3 + 2;
constexpr uint32_t kZsStagePerBlock = zs::kBlockMax + 64;

// zstd, step 1: every 128 KiB zstd block of every stream block, one warp each (lane 0 parses and codes: both are
// serial bit-stream work; the blocks are the parallelism -- 80 per 10 MiB stream block).
__global__ void __launch_bounds__(32) zstd_encode_kernel(LzmaJob *jobs, const zs::Tables *T, uint32_t fb)
{
	LzmaJob &j = jobs[blockIdx.y];
	const uint32_t x = blockIdx.x;
	if (x >= j.nzb)
		return;
	if ((j.gate_result && *j.gate_result == 0) || *j.mf_overflow)
		return; // decided in the assemble kernel
	const uint32_t lo = x * zs::kBlockMax, hi = lo + zs::kBlockMax < j.n ? lo + zs::kBlockMax : j.n, size = hi - lo;
	int same = 1;
	for (uint32_t i = lo + threadIdx.x; i < hi; i += 32)
		if (j.src[i] != j.src[lo])
			same = 0;
	same = __all_sync(0xffffffffu, same);
	if (threadIdx.x != 0)
		return;
	uint32_t v;
	if (same && size > 1)
		v = (1u << 28) | 1u; // RLE block
	else {
		const uint32_t cs = zs::encode_block(*T, j.src, j.n, lo, hi, j.rec, j.pool, fb, j.zseq + (size_t)x * kZsSeqPerBlock,
						     j.zlit + (size_t)x * zs::kBlockMax, j.zstage + (size_t)x * kZsStagePerBlock,
						     size + 64); // the last block's share of the scratch array is only that long
		v = cs ? ((2u << 28) | cs) : size; // compressed, else raw
	}
	j.zsize[x] = v;
}

// zstd, step 2: frame header, block offsets, block headers and payloads of one stream block's frame.
__global__ void __launch_bounds__(256) zstd_assemble_kernel(LzmaJob *jobs)
{
	LzmaJob &j = jobs[blockIdx.x];
	__shared__ uint64_t total_s;
	__shared__ int why_s;
	uint32_t *off = j.zsize + j.nzb;
	if (threadIdx.x == 0) {
		int why = (j.gate_result && *j.gate_result == 0) ? 1 : (*j.mf_overflow ? 2 : 0);
		uint64_t total = 0;
		if (!why) {
			uint8_t hdr[16];
			const uint32_t hl = zs::frame_header(j.n, hdr);
			total = hl;
			for (uint32_t x = 0; x < j.nzb; x++) {
				off[x] = (uint32_t)total;
				total += 3 + (j.zsize[x] & 0x0FFFFFFFu);
			}
			if (total >= j.n || total > j.outCap) // src/stream.c:205-221: not smaller (or no room) => stays stored
				why = 3;
			else
				for (uint32_t i = 0; i < hl; i++)
					j.out[i] = hdr[i];
		}
		total_s = total;
		why_s = why;
		j.outLen = total;
		j.skipped = why;
		j.overflow = 0;
	}
	__syncthreads();
	if (why_s)
		return;
	for (uint32_t x = 0; x < j.nzb; x++) {
		const uint32_t v = j.zsize[x], type = v >> 28, payload = v & 0x0FFFFFFFu;
		const uint32_t lo = x * zs::kBlockMax, size = (lo + zs::kBlockMax < j.n ? lo + zs::kBlockMax : j.n) - lo;
		uint8_t *w = j.out + off[x];
		if (threadIdx.x == 0) {
			const uint32_t h = (uint32_t)(x == j.nzb - 1) | (type << 1) | ((type == 1 ? size : payload) << 3);
			w[0] = (uint8_t)h;
			w[1] = (uint8_t)(h >> 8);
			w[2] = (uint8_t)(h >> 16);
		}
		const uint8_t *from = type == 2 ? j.zstage + (size_t)x * kZsStagePerBlock : j.src + lo;
		for (uint32_t i = threadIdx.x; i < payload; i += 256)
			w[3 + i] = from[i];
	}
}

// K7b: optimal parser + range coder, one block per CTA, over the precomputed match lists.
// The whole encoder state (probabilities, price tables, the 2048-cell parse table: 135 KB) lives in the
// SM's shared memory; warp 0 runs the encoder cooperatively (lzma_enc.cuh: replicated scalar code,
// lane-split loops).  Warp 1 is a pure look-ahead: it follows the encoder's position (e->pos in shared
// memory) and, for the next few positions, pulls into L1 what the parser will read from far away in HBM --
// the position's record, its match list in the pool, and for every (len, dist) pair the bytes at the end of
// the match that the MATCH : LIT : REP0 trial compares (LzmaEnc.c:1876-1893).  It writes nothing, so it
// cannot change the output; it only turns the parser's dependent HBM/L2 round trips into L1 hits.
constexpr uint32_t kLzmaAhead = 6; // positions

// Warp 1.  Besides pulling the far bytes towards L1 it STAGES the lists of the coming positions in a ring in
// shared memory (lzma_enc.cuh: kLkSlots entries), together with the byte comparison of every pair's
// MATCH : LIT : REP0 trial (LzmaEnc.c:1876-1893), so that the encoder warp finds both without touching HBM.
// A staged entry is a cache: the encoder checks its tag and falls back to HBM, so nothing here can change the output.
__device__ void lzma_lookahead_warp(lzma::Enc *e, const LzmaJob &j, const volatile int *done)
{
	uint32_t *const lkData = e->lkRing;
	const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
	const volatile uint32_t *ppos = &e->pos;
	const uint8_t *src = j.src;
	const uint32_t n = j.n, fb = j.cfg.fb;
	const bool live = j.wait_count != 0;
	uint32_t upto = 0; // positions <= upto (1-based, like e->pos) are already staged
	while (!*done) {
		const uint32_t pos = *ppos;
		uint32_t lo = pos + 1 > upto + 1 ? pos + 1 : upto + 1;
		uint32_t hi = pos + kLzmaAhead < n ? pos + kLzmaAhead : n;
		if (lo > hi) {
			__nanosleep(100);
			continue;
		}
		if (hi > lo + 3)
			hi = lo + 3;
		// four positions in flight, eight lanes each: the dependent far loads (record -> list -> bytes) of the four overlap
		const uint32_t q = lo + grp;
		uint64_t rec = 0;
		bool ready = true;
		if (q <= hi) {
			if (live) { // the tree walk may not have reached this position yet
				if (q - 1 < j.wait_count) {
					rec = *(const volatile uint64_t *)(j.rec + (q - 1));
					ready = (rec >> 63) != 0;
				}
			} else
				rec = j.rec[q - 1];
			rec &= ~lzma::kMfReady;
		}
		if (!__all_sync(0xffffffffu, ready)) {
			__nanosleep(200);
			continue;
		}
		const bool usable = q <= hi && (!live || (rec >> 10) + ((uint32_t)rec & 1023u) <= j.pool_cap);
		if (usable) {
			const uint32_t i0 = q - 1;
			const uint32_t nd = (uint32_t)rec & 1023u;
			const uint32_t *lst = j.pool + (rec >> 10);
			const bool stage = nd <= lzma::kLkMaxList;
			uint32_t *b = lkData + (q & (lzma::kLkSlots - 1)) * lzma::kLkWords;
			const uint32_t numAvail = n - i0; // what ReadMatchDistances will see at this position
			const uint8_t *data = src + i0;
			if (sub < lzma::kNumReps) { // the bytes behind the reps, as they stand now (they seldom change from one cell to the next)
				const uint32_t r = *(const volatile uint32_t *)&e->pubReps[sub];
				if (r <= i0)
					lzma::lz_prefetch(data - r);
			}
			for (uint32_t k = sub; 2 * k < nd; k += 8) {
				// (beside a running walk the pool is read past L1: a cached line may predate a neighbouring list)
				const uint32_t len = live ? __ldcg(lst + 2 * k) : lst[2 * k], dist = live ? __ldcg(lst + 2 * k + 1) : lst[2 * k + 1];
				uint32_t w = 0;
				if (dist < i0) {
					const uint8_t *data2 = data - dist - 1;
					lzma::lz_prefetch(data2);
					uint32_t len2 = len + 3, limit = len + 1 + fb;
					if (limit > numAvail)
						limit = numAvail;
					if (len2 <= limit && data[len2 - 2] == data2[len2 - 2] && data[len2 - 1] == data2[len2 - 1]) {
						while (len2 < limit && data[len2] == data2[len2])
							len2++;
						w = 0x80000000u | len2;
					}
				}
				if (stage) {
					b[1 + 2 * k] = len;
					b[2 + 2 * k] = dist;
					b[1 + lzma::kLkMaxList + k] = w;
					// distance by length: this pair serves the lengths above the previous pair's up to its own
					uint32_t from = k ? (live ? __ldcg(lst + 2 * k - 2) : lst[2 * k - 2]) + 1 : 2;
					for (; from <= len && from <= lzma::kMatchMax; from++)
						b[lzma::kLkByLen + from] = dist;
				}
			}
		}
		__syncwarp();
		__threadfence_block();
		__syncwarp();
		if (usable && sub == 0) {
			const uint32_t nd = (uint32_t)rec & 1023u;
			if (nd <= lzma::kLkMaxList) { // position, count and the longest length (ReadMatchDistances' result) in one store
				const uint32_t *lp = j.pool + (rec >> 10) + nd - 2;
				const uint32_t hdr = nd | ((nd ? (live ? __ldcg(lp) : *lp) : 0u) << 16);
				*reinterpret_cast<volatile uint64_t *>(&e->lkHead[q & (lzma::kLkSlots - 1)][0]) = (uint64_t)q | ((uint64_t)hdr << 32);
			}
		}
		upto = hi;
		__syncwarp();
	}
}

// Warp 3, one thread: the range arithmetic, carries and byte output for the (probability, bit) pairs the encoder
// warp queues (lzma_enc.cuh: rc_bit with a queue).  Same bytes as the in-line coder: the queue preserves the order
// of the binary decisions and carries the probability each one was coded with.
__device__ void lzma_coder_thread(lzma::Enc *e)
{
	const volatile uint32_t *q = e->rcQueue;
	volatile uint32_t *tailPub = &e->rcTailPub, *headPub = &e->rcHeadPub;
	volatile int *rcDone = &e->rcDone;
	uint32_t head = 0;
	for (;;) {
		const uint32_t tail = *tailPub;
		if (head == tail) {
			__nanosleep(50);
			continue;
		}
		__threadfence_block();
		bool flush = false;
		while (head != tail) {
			const uint32_t op = q[head & (lzma::kRcQ - 1)];
			head++;
			if (op == lzma::kRcFlush) {
				flush = true;
				break;
			}
			if (op & lzma::kRcDirect) {
				uint32_t nbits = (op >> 26) & 31u;
				const uint32_t value = op & 0x3FFFFFFu;
				while (nbits--) {
					e->range >>= 1;
					if ((value >> nbits) & 1)
						e->low += e->range;
					lzma::rc_norm(e);
				}
			} else
				lzma::rc_bit_value(e, op >> 1, op & 1u);
		}
		*headPub = head;
		if (flush) {
			for (int i = 0; i < 5; i++) // RangeEnc_FlushData; no end marker (writeEndMark = 0)
				lzma::rc_shift_low(e);
			__threadfence_block();
			*rcDone = 1;
			return;
		}
	}
}

// K7b kernel: one block per CTA, four warps.  Warp 0 is the encoder (optimal parser, probability model, symbol
// decisions: lzma_enc.cuh, replicated scalar code with lane-split loops); warp 1 looks ahead and stages match lists;
// warp 2 runs the lz4 compressibility gate beside it; one thread of warp 3 is the range coder.
__global__ void __launch_bounds__(160, 1) lzma_block_kernel(LzmaJob *jobs)
{
#if defined(LRZ_SIMT_HOST)
	uint8_t *lzma_smem = simt::dyn_smem();
#else
	extern __shared__ __align__(16) uint8_t lzma_smem[];
#endif
	__shared__ uint32_t gate_table[4096];
	__shared__ int done, gate_state;
	lzma::Enc *e = reinterpret_cast<lzma::Enc *>(lzma_smem);
	LzmaJob &j = jobs[blockIdx.x];
	if (*j.mf_overflow) { // uniform over the CTA: the host encodes this block again with a larger pool
		if (threadIdx.x == 0)
			j.skipped = 2;
		return;
	}
	if (threadIdx.x < lzma::kLkSlots)
		e->lkHead[threadIdx.x][0] = e->lkHead[threadIdx.x][1] = 0; // no position 0: positions count from 1
	if (threadIdx.x == 0) {
		done = 0;
		gate_state = 0;
	}
	__syncthreads();
	if (threadIdx.x < 32) {
		lzma::enc_init(e, j.cfg, j.src, j.n, j.out, j.outCap, nullptr, nullptr, nullptr, nullptr);
		e->preRec = j.rec;
		e->prePool = j.pool;
		e->preWait = j.wait_count;
		e->prePoolCap = j.pool_cap;
		e->mfOverflow = j.wait_count ? j.mf_overflow : nullptr;
		e->gateState = j.threshold ? &gate_state : nullptr;
		e->lkOn = e->rcOn = 1;
		e->splitOn = 1;
	}
	__syncthreads(); // the encoder state is initialised: the helpers may read it
#if defined(__CUDA_ARCH__) // (the helpers exist in the device pass only)
	if (threadIdx.x >= 128) {
		// warp 4: the rep / match half of every staged position, one position behind warp 0 (lzma_enc.cuh: opt_step_b)
		for (;;) {
#if defined(LZ_PROF)
			const long long w0 = clock64();
#endif
			lzma::bar_wait(lzma::kBarGo);
			if (*(volatile uint32_t *)&e->pkt.cmd == 0)
				return;
#if defined(LZ_PROF)
			const long long w1 = clock64();
#endif
			lzma::opt_step_b(e);
#if defined(LZ_PROF)
			if (threadIdx.x == 128) {
				e->prof[22] += (uint64_t)(w1 - w0);
				e->prof[23] += (uint64_t)(clock64() - w1);
				e->profN[22]++;
			}
#endif
			lzma::bar_arrive(lzma::kBarDone);
		}
	}
#endif
	if (threadIdx.x >= 96) {
		if (threadIdx.x == 96)
			lzma_coder_thread(e);
		return;
	}
	if (threadIdx.x >= 64) {
		// warp 2: the lz4 compressibility gate (LZ4_TEST) of this block, beside the encoder instead of in front of
		// it.  lz4_compresses() is a serial LZ4 emulation: one lane.  A block it rejects costs the encoder the
		// gate's run time (it gives up when it reads the verdict); every other block starts encoding at once.
		if (threadIdx.x == 64) {
			int v = 1;
			if (j.threshold)
				v = lz4s::gate(j.src, (int64_t)j.n, j.threshold, gate_table) ? 1 : 2;
			__threadfence_block();
			*(volatile int *)&gate_state = v;
		}
		return;
	}
	if (threadIdx.x >= 32) {
		lzma_lookahead_warp(e, j, &done);
		return;
	}
#if defined(LZ_PROF)
	for (int i = 0; i < 24; i++)
		e->prof[i] = e->profN[i] = 0;
	e->profT = clock64();
	const long long prof_t0 = e->profT;
#endif
	const uint64_t len = lzma::enc_run(e);
	__syncwarp();
#if defined(__CUDA_ARCH__)
	*(volatile uint32_t *)&e->pkt.cmd = 0; // warp 4 leaves
	lzma::bar_arrive(lzma::kBarGo);
#endif
#if defined(LZ_PROF)
	if (threadIdx.x == 0) {
		const long long tot = clock64() - prof_t0;
		printf("[lzprof] n=%u total=%lld cyc (%.0f / byte)\n", j.n, tot, (double)tot / j.n);
		for (int i = 0; i < 20; i++)
			printf("[lzprof] s%-2d %6.2f%%  %12llu cyc  %10llu calls  %8.0f cyc/call\n", i, 100.0 * e->prof[i] / tot,
			       (unsigned long long)e->prof[i], (unsigned long long)e->profN[i],
			       e->profN[i] ? (double)e->prof[i] / e->profN[i] : 0.0);
		printf("[lzprof] pos mismatch %llu\n", (unsigned long long)e->profN[20]);
		printf("[lzprof] warp A waits for B: %.0f cyc / position (%llu waits); warp B waits for A: %.0f, works: %.0f cyc / position\n",
		       e->profN[21] ? (double)e->prof[21] / e->profN[21] : 0.0, (unsigned long long)e->profN[21],
		       e->profN[22] ? (double)e->prof[22] / e->profN[22] : 0.0, e->profN[22] ? (double)e->prof[23] / e->profN[22] : 0.0);
	}
#endif
	int verdict = 1;
	if (j.threshold) // a block shorter than the gate's run time: wait for the verdict
		while ((verdict = *(volatile int *)&gate_state) == 0)
			__nanosleep(500);
	if (threadIdx.x == 0) {
		done = 1;
		j.outLen = len;
		j.overflow = *(volatile int *)&e->overflow;
		j.skipped = verdict == 2 ? 1 : (*(volatile int *)&e->aborted == 2 ? 2 : 0);
	}
}

int64_t round_up_page(int64_t v, int page) { return v % page ? v + page - v % page : v; }

} // namespace

#if !defined(LRZ_SIMT_HOST)
// ---- LZMA pipeline ------------------------------------------------------------------------------------
// Blocks are SUBMITTED (lz4 gate, match finder, parser enqueued on the backend's own streams, no host wait)
// and later DRAINED (wait, read verdicts and lengths).  The rzip stage submits stream-1 blocks while it is
// still scanning (src/stream.c:1836-1875: the reference hands a block to a compthread the moment it fills);
// the synchronous entry point is submit-everything-then-drain.
struct SlotLay {
	size_t son, c2, c3, sorted, ctl, rec, pool, end;
	uint64_t poolCap;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static SlotLay lay_for(size_t n, bool hc5, unsigned poolMul)
{
	const size_t minAvail = hc5 ? 5 : 4, count = n >= minAvail ? n - (minAvail - 1) : 0;
	SlotLay L;
	size_t o = 0;
	L.son = o;
	o += hc5 ? 256 : align_up(8 * (n + 2), 256); // the hash-chain finder keeps its links in `sorted`
	L.c2 = o;
	o += align_up(4 * count + 4, 256);
	L.c3 = o;
	o += align_up(4 * count + 4, 256);
	L.sorted = o;
	o += align_up(4 * count + 4, 256);
	L.ctl = o; // cursor, overflow flag; zeroed together with rec
	o += 256;
	L.rec = o;
	o += align_up(8 * n, 256);
	L.pool = o;
	L.poolCap = (uint64_t)poolMul * n + 65536;
	o += align_up(4 * L.poolCap, 256);
	L.end = o;
	return L;
}

constexpr unsigned kPoolMul = 12; // uint32 of match-list pool per input byte; text needs ~6-9

struct AsyncSub {
	BlockJob job;
	lzma::Config cfg;
	size_t oofs = 0;     // payload offset in BackendCtx::out
	int group = -1;
	int index_in_group = 0;
	bool done = false;
};

struct AsyncGroup {
	size_t first = 0, count = 0; // subs[first .. first + count)
	size_t meta_off = 0;         // [LzmaJob x m][MfBlock x m][segBase x (m + 1)][GateJob x m] in BackendCtx::meta
	cudaStream_t ps = nullptr;   // parser stream
	cudaStream_t ws = nullptr;   // tree-walk stream (LZMA levels 5-9: the walk runs beside the parser)
	cudaEvent_t evWalk = nullptr;
	cudaEvent_t evMF = nullptr, evDone = nullptr; // sorts done / parser done
	uint64_t walk_total = 0;                       // positions of the group (grid of the tree walk)
	uint32_t max_nzb = 0;                          // zstd: most 128 KiB blocks in one stream block of the group
	bool hc5 = false, gated = false, launched = false, finished = false;
};

struct BackendCtx {
	DevBuf jobs, work, out, flags, offs, scratch, meta, big, zs_tables;
	bool zstd = false; // the chunk's backend (else LZMA)
	cudaStream_t sMF = nullptr, sGate = nullptr;
	std::vector<cudaStream_t> pstreams, wstreams;
	std::vector<cudaEvent_t> events;
	size_t ev_next = 0;
	// state of the chunk being encoded
	bool active = false;
	lrzgpu_params p;
	lrzgpu_sizing_t sz;
	uint32_t fb = 0;
	size_t slot_bytes = 0, max_block = 0;
	int nslots = 0, slots_used = 0;
	size_t out_used = 0, meta_used = 0;
	std::vector<AsyncSub> subs;
	std::vector<AsyncGroup> groups; // enqueued since the last drain
	size_t next_launch = 0;         // groups[next_launch ..) still wait for their parser kernel
	int inflight = 0, inflight_max = 100; // blocks whose parser kernel is running
};

BackendCtx *backend_create() { return new BackendCtx(); }

void backend_destroy(BackendCtx *b)
{
	if (!b)
		return;
	DevBuf *bufs[] = { &b->jobs, &b->work, &b->out, &b->flags, &b->offs, &b->scratch, &b->meta, &b->big, &b->zs_tables };
	for (DevBuf *d : bufs)
		d->release();
	for (cudaStream_t s : b->pstreams)
		cudaStreamDestroy(s);
	for (cudaStream_t s : b->wstreams)
		cudaStreamDestroy(s);
	for (cudaEvent_t e : b->events)
		cudaEventDestroy(e);
	if (b->sMF)
		cudaStreamDestroy(b->sMF);
	if (b->sGate)
		cudaStreamDestroy(b->sGate);
	delete b;
}

static cudaEvent_t next_event(BackendCtx *b)
{
	if (b->ev_next == b->events.size()) {
		cudaEvent_t e = nullptr;
		if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
			return nullptr;
		b->events.push_back(e);
	}
	return b->events[b->ev_next++];
}

int backend_lz4_gate(BackendCtx *b, const uint8_t *d_src, int64_t len, int threshold, int *result, cudaStream_t stream,
		     int64_t *launches)
{
	GateJob g = { d_src, len, 0 };
	if (b->jobs.ensure(sizeof(GateJob)) != cudaSuccess)
		return LRZGPU_ENOMEM;
	if (cudaMemcpyAsync(b->jobs.p, &g, sizeof(g), cudaMemcpyHostToDevice, stream) != cudaSuccess)
		return LRZGPU_ECUDA;
	lz4_gate_kernel<<<1, 32, 0, stream>>>((GateJob *)b->jobs.p, threshold);
	if (launches)
		(*launches)++;
	if (cudaMemcpyAsync(&g, b->jobs.p, sizeof(g), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
	    cudaStreamSynchronize(stream) != cudaSuccess)
		return LRZGPU_ECUDA;
	*result = g.result;
	return LRZGPU_OK;
}

int backend_async_begin(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t max_block_len,
			int64_t max_blocks, int64_t payload_bytes_upper, char *err, size_t errlen)
{
	if (b->active && !b->groups.empty()) // an abandoned chunk: its kernels still use the work space
		cudaDeviceSynchronize();
	b->groups.clear();
	b->active = false;
	if (p.backend != LRZGPU_BACKEND_LZMA && p.backend != LRZGPU_BACKEND_ZSTD) {
		snprintf(err, errlen, "the block pipeline serves the LZMA and zstd backends");
		return LRZGPU_EUNSUPPORTED;
	}
	b->zstd = p.backend == LRZGPU_BACKEND_ZSTD;
	if (b->zstd && !b->zs_tables.p) { // the format's predefined FSE tables, built once on the host
		zs::Tables T;
		zs::build_tables(T);
		if (b->zs_tables.ensure(sizeof(T)) != cudaSuccess || cudaMemcpy(b->zs_tables.p, &T, sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
			snprintf(err, errlen, "zstd tables: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
	}
	if (lzma::mf_init_tables()) {
		snprintf(err, errlen, "LZMA match finder tables: %s", cudaGetErrorString(cudaGetLastError()));
		return LRZGPU_ECUDA;
	}
	if (!b->sMF && (cudaStreamCreateWithFlags(&b->sMF, cudaStreamNonBlocking) != cudaSuccess ||
			cudaStreamCreateWithFlags(&b->sGate, cudaStreamNonBlocking) != cudaSuccess)) {
		snprintf(err, errlen, "backend streams: %s", cudaGetErrorString(cudaGetLastError()));
		return LRZGPU_ECUDA;
	}
	b->p = p;
	b->sz = sz;
	b->fb = (p.level < 7 && !b->zstd) ? 32 : 64; // src/stream.c:455
	if (b->zstd) { // the match finder in its binary-tree form with a 32 MiB window, whatever the level
		b->sz.dict_size = 1u << 25;
		b->p.level = p.level < 5 ? 5 : p.level;
	}
	lzma::Config cfg;
	if (max_block_len < 1)
		max_block_len = 1;
	if (!lzma::make_config(b->p.level, b->sz.dict_size, b->fb, (uint64_t)max_block_len, cfg) || b->fb > lzma::kMfMaxFb) {
		snprintf(err, errlen, "LZMA level %d / block of %lld bytes is not supported by the device encoder", p.level,
			 (long long)max_block_len);
		return LRZGPU_EUNSUPPORTED;
	}
	const bool hc5 = cfg.fastMode != 0;
	const size_t minAvail = hc5 ? 5 : 4;
	b->max_block = (size_t)max_block_len;
	b->slot_bytes = align_up(lay_for(b->max_block, hc5, kPoolMul).end, 4096);
	const uint32_t maxCount = b->max_block >= minAvail ? (uint32_t)(b->max_block - (minAvail - 1)) : 0;
	const size_t scratch_bytes = align_up(lzma::mf_sort_scratch_bytes(maxCount), 256);
	const size_t out_bytes = (size_t)payload_bytes_upper + (size_t)payload_bytes_upper / 50 + (size_t)(max_blocks + 1) * 8192;
	const size_t meta_bytes = (size_t)(max_blocks + 1) * (sizeof(LzmaJob) + sizeof(lzma::MfBlock) + sizeof(GateJob) + 16 + 1024);
	if (b->out.ensure(out_bytes) != cudaSuccess || b->scratch.ensure(scratch_bytes) != cudaSuccess ||
	    b->meta.ensure(meta_bytes) != cudaSuccess) {
		snprintf(err, errlen, "out of device memory for LZMA payloads / sort scratch (%zu MiB)", (out_bytes + scratch_bytes) >> 20);
		return LRZGPU_ENOMEM;
	}
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	const size_t have = free_b + b->work.cap;
	const size_t budget = have > (3ull << 30) ? have - (2ull << 30) : have / 2;
	int64_t slots = (int64_t)(budget / b->slot_bytes);
	if (slots > max_blocks)
		slots = max_blocks;
	if (slots < 1)
		slots = 1;
	if (b->work.ensure((size_t)slots * b->slot_bytes) != cudaSuccess) {
		snprintf(err, errlen, "out of device memory for %lld LZMA block encoders (%zu MiB)", (long long)slots,
			 ((size_t)slots * b->slot_bytes) >> 20);
		return LRZGPU_ENOMEM;
	}
	// streams and events of the groups are made now: nothing is created or freed while block encoders run
	const size_t want_streams = (size_t)(max_blocks < 96 ? max_blocks : 96);
	while (b->pstreams.size() < want_streams) {
		cudaStream_t st = nullptr;
		if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess)
			return LRZGPU_ECUDA;
		b->pstreams.push_back(st);
	}
	while (b->wstreams.size() < want_streams) {
		cudaStream_t st = nullptr;
		if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess)
			return LRZGPU_ECUDA;
		b->wstreams.push_back(st);
	}
	while (b->events.size() < 3 * want_streams) {
		cudaEvent_t ev = nullptr;
		if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
			return LRZGPU_ECUDA;
		b->events.push_back(ev);
	}
	b->nslots = (int)slots;
	b->slots_used = 0;
	b->out_used = b->meta_used = 0;
	b->subs.clear();
	b->groups.clear();
	b->ev_next = 0;
	b->next_launch = 0;
	b->inflight = 0;
	{
		int dev = 0, sms = 148;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		b->inflight_max = sms > 48 ? sms - 24 : sms / 2; // one parser CTA per SM; the rest stays free for the scan
	}
	if (const char *im = getenv("LRZGPU_INFLIGHT")) // development switch: cap on block encoders running at once
		b->inflight_max = atoi(im) > 0 ? atoi(im) : b->inflight_max;
	if (cudaFuncSetAttribute(lzma_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lzma::Enc)) != cudaSuccess) {
		snprintf(err, errlen, "LZMA encoder state (%zu bytes) does not fit in shared memory", sizeof(lzma::Enc));
		return LRZGPU_ECUDA;
	}
	b->active = true;
	return LRZGPU_OK;
}

// Enqueue gate + match finder + parser for subs[first .. first + m) whose work arrays start at `bases[i]`.
static int enqueue_group(BackendCtx *b, size_t first, size_t m, const std::vector<uint8_t *> &bases, unsigned poolMul,
			 bool gate, cudaEvent_t ready, int64_t *launches, char *err, size_t errlen)
{
	AsyncGroup G;
	G.first = first;
	G.count = m;
	const size_t o_mb = m * sizeof(LzmaJob), o_seg = align_up(o_mb + m * sizeof(lzma::MfBlock), 8);
	const size_t o_gate = align_up(o_seg + (m + 1) * 8, 8);
	const size_t total = align_up(o_gate + m * sizeof(GateJob), 256);
	if (b->meta_used + total > b->meta.cap)
		return 1; // drain first
	G.meta_off = b->meta_used;
	b->meta_used += total;
	uint8_t *J = (uint8_t *)b->meta.p + G.meta_off;
	// parser stream of this group: its kernel runs for the whole encode of the blocks, so every group that is
	// in flight at the same time needs its own
	const size_t gi = b->groups.size();
	while (b->pstreams.size() <= gi) {
		cudaStream_t s = nullptr;
		if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
			snprintf(err, errlen, "parser stream: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
		b->pstreams.push_back(s);
	}
	G.ps = b->pstreams[gi];
	while (b->wstreams.size() <= gi) {
		cudaStream_t s = nullptr;
		if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
			snprintf(err, errlen, "walk stream: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
		b->wstreams.push_back(s);
	}
	G.ws = b->wstreams[gi];
	G.evWalk = next_event(b);
	if (!G.evWalk)
		return LRZGPU_ECUDA;
	cudaEvent_t evMF = next_event(b);
	if (!evMF)
		return LRZGPU_ECUDA;
	if (ready && cudaEventSynchronize(ready) != cudaSuccess) // see pump_groups: no device-side waits on these streams
		return LRZGPU_ECUDA;
	std::vector<LzmaJob> lj(m);
	std::vector<lzma::MfBlock> mb(m);
	std::vector<uint64_t> seg(m + 1, 0);
	std::vector<GateJob> gj(m);
	bool hc5 = false;
	for (size_t i = 0; i < m; i++) {
		AsyncSub &S = b->subs[first + i];
		const BlockJob &bj = S.job;
		const lzma::Config &c = S.cfg;
		hc5 = c.fastMode != 0;
		const size_t minAvail = hc5 ? 5 : 4;
		const SlotLay L = lay_for((size_t)bj.u_len, hc5, poolMul);
		uint8_t *W = bases[i];
		lzma::MfBlock &B = mb[i];
		B.src = bj.d_src;
		B.P.n = (uint32_t)bj.u_len;
		B.P.fb = c.fb;
		B.P.mc = c.mc;
		B.P.hashMask = c.hashMask;
		B.P.bigHash = c.bigHash;
		B.P.historySize = c.historySize;
		B.P.cyclicSize = c.cyclicSize;
		B.P.hc5 = c.fastMode;
		B.count = (size_t)bj.u_len >= minAvail ? (uint32_t)bj.u_len - (uint32_t)(minAvail - 1) : 0;
		B.son = (uint32_t *)(W + L.son);
		B.c2 = (uint32_t *)(W + L.c2);
		B.c3 = (uint32_t *)(W + L.c3);
		B.sorted = (uint32_t *)(W + L.sorted);
		B.cursor = (unsigned long long *)(W + L.ctl);
		B.overflow = (int *)(W + L.ctl + 8);
		B.rec = (uint64_t *)(W + L.rec);
		B.pool = (uint32_t *)(W + L.pool);
		B.poolCap = L.poolCap;
		seg[i + 1] = seg[i] + B.count;
		LzmaJob &j = lj[i];
		memset(&j, 0, sizeof(j));
		j.src = bj.d_src;
		j.n = (uint32_t)bj.u_len;
		j.out = (uint8_t *)b->out.p + S.oofs;
		j.outCap = (uint64_t)round_up_page((int64_t)((double)bj.u_len * 1.02), b->p.page_size);
		j.rec = B.rec;
		j.pool = B.pool;
		j.cfg = c;
		j.threshold = gate ? b->p.threshold : 0;
		j.mf_overflow = B.overflow;
		j.pool_cap = B.poolCap;
		j.wait_count = (!b->zstd && !hc5) ? B.count : 0; // the parser starts beside the tree walk (pump_groups)
		if (b->zstd) {
			j.outCap = (uint64_t)round_up_page(bj.u_len, b->p.page_size); // src/stream.c:169
			j.nzb = (uint32_t)(((uint64_t)bj.u_len + zs::kBlockMax - 1) / zs::kBlockMax);
			j.zseq = (zs::Seq *)(W + L.son);
			j.zlit = W + L.c2;
			j.zstage = W + L.c3;
			j.zsize = (uint32_t *)(W + L.sorted);
			j.gate_result = gate ? &((GateJob *)(J + o_gate))[i].result : nullptr;
			gj[i].src = bj.d_src;
			gj[i].len = bj.u_len;
			gj[i].result = 0;
		}
		S.group = (int)gi;
		S.index_in_group = (int)i;
		if (cudaMemsetAsync(W + L.ctl, 0, 256 + 8 * (size_t)bj.u_len, b->sMF) != cudaSuccess)
			return LRZGPU_ECUDA;
		if (lzma::mf_prepare_block(B, b->scratch.p, b->sMF, launches)) {
			snprintf(err, errlen, "LZMA match finder sort: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
	}
	if (cudaMemcpyAsync(J, lj.data(), o_mb, cudaMemcpyHostToDevice, b->sMF) != cudaSuccess ||
	    cudaMemcpyAsync(J + o_mb, mb.data(), m * sizeof(lzma::MfBlock), cudaMemcpyHostToDevice, b->sMF) != cudaSuccess ||
	    cudaMemcpyAsync(J + o_seg, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, b->sMF) != cudaSuccess ||
	    (b->zstd && cudaMemcpyAsync(J + o_gate, gj.data(), m * sizeof(GateJob), cudaMemcpyHostToDevice, b->sMF) != cudaSuccess))
		return LRZGPU_ECUDA;
	if (cudaEventRecord(evMF, b->sMF) != cudaSuccess)
		return LRZGPU_ECUDA;
	G.evMF = evMF;
	G.walk_total = seg.back();
	G.hc5 = hc5;
	G.gated = gate;
	for (size_t i = 0; i < m; i++)
		if (lj[i].nzb > G.max_nzb)
			G.max_nzb = lj[i].nzb;
	G.evDone = next_event(b);
	if (!G.evDone)
		return LRZGPU_ECUDA;
	b->groups.push_back(G);
	return LRZGPU_OK;
}

// Launch the tree walk and the parser kernel of the groups whose sorts have finished.  The host does the
// waiting: a parser stream that waited on the device would sit at the head of a hardware work queue it shares
// with other streams (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) and hold up their launches -- the scan's
// among them -- and so would a parser kernel that finds no SM with 135 KB of shared memory free, hence the cap
// on blocks in flight.  wait = false: launch what is ready, return at once.
static int pump_groups(BackendCtx *b, bool wait, int64_t *launches, char *err, size_t errlen)
{
	for (AsyncGroup &G : b->groups) // retire finished kernels
		if (G.launched && !G.finished && cudaEventQuery(G.evDone) == cudaSuccess) {
			G.finished = true;
			b->inflight -= (int)G.count;
		}
	while (b->next_launch < b->groups.size()) {
		AsyncGroup &G = b->groups[b->next_launch];
		if (b->inflight > 0 && b->inflight + (int)G.count > b->inflight_max) {
			if (!wait)
				return LRZGPU_OK;
			for (AsyncGroup &O : b->groups) // the oldest running group
				if (O.launched && !O.finished) {
					if (cudaEventSynchronize(O.evDone) != cudaSuccess)
						return LRZGPU_ECUDA;
					O.finished = true;
					b->inflight -= (int)O.count;
					break;
				}
			continue;
		}
		if (wait) {
			if (cudaEventSynchronize(G.evMF) != cudaSuccess) {
				snprintf(err, errlen, "LZMA match finder sorts failed: %s", cudaGetErrorString(cudaGetLastError()));
				return LRZGPU_ECUDA;
			}
		} else if (cudaEventQuery(G.evMF) != cudaSuccess)
			return LRZGPU_OK;
		// the tree walk's run time is that of the block's longest hash bucket (one thread per bucket) whatever the
		// number of blocks in the launch: it goes on the group's own stream, so that the walks of successive
		// groups overlap, and the parser kernel follows it in stream order
		uint8_t *J = (uint8_t *)b->meta.p + G.meta_off;
		const size_t o_mb = G.count * sizeof(LzmaJob), o_seg = align_up(o_mb + G.count * sizeof(lzma::MfBlock), 8);
		// LZMA levels 5-9: the parser does not wait for the walk to end.  A record is final once its bucket's thread
		// has passed the position, and the slowest bucket needs ~1 s for a block the parser needs ~10 s for, so the
		// parser (look-ahead warp) only waits for the "final" bit of the records right in front of it.
		const bool beside = !b->zstd && !G.hc5;
		if (lzma::mf_walk_launch((const lzma::MfBlock *)(J + o_mb), (int)G.count, (const uint64_t *)(J + o_seg), G.walk_total,
					 G.hc5, beside ? G.ws : G.ps, launches)) {
			snprintf(err, errlen, "LZMA match finder walk: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
		if (b->zstd) {
			const size_t o_gate = align_up(o_seg + (G.count + 1) * 8, 8);
			if (G.gated) {
				lz4_gate_kernel<<<(unsigned)G.count, 32, 0, G.ps>>>((GateJob *)(J + o_gate), b->p.threshold);
				if (launches)
					(*launches)++;
			}
			zstd_encode_kernel<<<dim3(G.max_nzb, (unsigned)G.count), 32, 0, G.ps>>>((LzmaJob *)J, (const zs::Tables *)b->zs_tables.p, b->fb);
			zstd_assemble_kernel<<<(unsigned)G.count, 256, 0, G.ps>>>((LzmaJob *)J);
			if (launches)
				(*launches) += 2;
		} else {
			lzma_block_kernel<<<(unsigned)G.count, 160, sizeof(lzma::Enc), G.ps>>>((LzmaJob *)J);
			if (launches)
				(*launches)++;
			if (beside && (cudaEventRecord(G.evWalk, G.ws) != cudaSuccess || cudaStreamWaitEvent(G.ps, G.evWalk, 0) != cudaSuccess)) {
				snprintf(err, errlen, "walk event: %s", cudaGetErrorString(cudaGetLastError()));
				return LRZGPU_ECUDA;
			}
		}
		if (cudaGetLastError() != cudaSuccess || cudaEventRecord(G.evDone, G.ps) != cudaSuccess) {
			snprintf(err, errlen, "LZMA kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
		G.launched = true;
		b->inflight += (int)G.count;
		b->next_launch++;
	}
	return LRZGPU_OK;
}

int backend_async_poll(BackendCtx *b, int64_t *launches, char *err, size_t errlen)
{
	return b->active ? pump_groups(b, false, launches, err, errlen) : LRZGPU_OK;
}

// Submit `n` blocks (u_len >= 64) as one group.  `ready`: event after which their bytes are in place (or null).
// Returns 0, a negative error, or 1 when the work space is used up: drain, then submit again.
int backend_async_submit(BackendCtx *b, const BlockJob *jobs, int n, cudaEvent_t ready, int64_t *launches, char *err,
			 size_t errlen)
{
	if (!b->active || n <= 0)
		return b->active ? LRZGPU_OK : LRZGPU_EINVAL;
	if (b->slots_used + n > b->nslots)
		return 1;
	size_t osum = b->out_used;
	std::vector<AsyncSub> add((size_t)n);
	for (int i = 0; i < n; i++) {
		AsyncSub &S = add[(size_t)i];
		S.job = jobs[i];
		S.job.c_type = LRZGPU_CTYPE_NONE;
		S.job.c_len = jobs[i].u_len;
		S.job.d_payload = jobs[i].d_src;
		if ((size_t)jobs[i].u_len > b->max_block ||
		    !lzma::make_config(b->p.level, b->sz.dict_size, b->fb, (uint64_t)jobs[i].u_len, S.cfg)) {
			snprintf(err, errlen, "block of %lld bytes is not supported by the device encoder", (long long)jobs[i].u_len);
			return LRZGPU_EUNSUPPORTED;
		}
		S.oofs = osum;
		osum += align_up((size_t)round_up_page((int64_t)((double)jobs[i].u_len * 1.02), b->p.page_size), 256);
	}
	if (osum > b->out.cap) {
		snprintf(err, errlen, "LZMA payload area exhausted");
		return LRZGPU_EINTERNAL;
	}
	const size_t first = b->subs.size();
	std::vector<uint8_t *> bases((size_t)n);
	for (int i = 0; i < n; i++)
		bases[(size_t)i] = (uint8_t *)b->work.p + (size_t)(b->slots_used + i) * b->slot_bytes;
	b->subs.insert(b->subs.end(), add.begin(), add.end());
	const int rc = enqueue_group(b, first, (size_t)n, bases, kPoolMul, b->p.threshold != 0, ready, launches, err, errlen);
	if (rc) {
		b->subs.resize(first);
		return rc;
	}
	b->slots_used += n;
	b->out_used = osum;
	return pump_groups(b, false, launches, err, errlen);
}

static int collect_groups(BackendCtx *b, std::vector<size_t> &redo, int64_t *launches, char *err, size_t errlen)
{
	int prc = pump_groups(b, true, launches, err, errlen);
	if (prc)
		return prc;
	for (const AsyncGroup &G : b->groups) {
		cudaError_t ce = cudaStreamSynchronize(G.ps);
		if (ce != cudaSuccess) {
			snprintf(err, errlen, "LZMA kernel failed: %s", cudaGetErrorString(ce));
			return LRZGPU_ECUDA;
		}
		std::vector<LzmaJob> lj(G.count);
		if (cudaMemcpy(lj.data(), (uint8_t *)b->meta.p + G.meta_off, G.count * sizeof(LzmaJob), cudaMemcpyDeviceToHost) != cudaSuccess)
			return LRZGPU_ECUDA;
		for (size_t i = 0; i < G.count; i++) {
			AsyncSub &S = b->subs[G.first + i];
			S.done = true;
			if (lj[i].skipped == 1 || lj[i].skipped == 3)
				continue; // LZ4_TEST (FLAG_THRESHOLD): incompressible blocks stay stored; zstd frame not smaller
			if (lj[i].skipped == 2) {
				S.done = false;
				redo.push_back(G.first + i);
				continue;
			}
			// src/stream.c:482-487: kept only when smaller; SZ_ERROR_OUTPUT_EOF leaves the block stored
			if (!lj[i].overflow && (int64_t)lj[i].outLen < S.job.u_len) {
				S.job.c_type = b->zstd ? LRZGPU_CTYPE_ZSTD : LRZGPU_CTYPE_LZMA;
				S.job.c_len = (int64_t)lj[i].outLen;
				S.job.d_payload = lj[i].out;
			}
		}
	}
	b->groups.clear();
	b->next_launch = 0;
	b->inflight = 0;
	b->slots_used = 0;
	b->meta_used = 0;
	b->ev_next = 0;
	return LRZGPU_OK;
}

// Wait for everything submitted so far; afterwards the work space is free again and backend_async_result()
// answers for every submitted block.
int backend_async_drain(BackendCtx *b, int64_t *launches, char *err, size_t errlen)
{
	if (!b->active)
		return LRZGPU_EINVAL;
	std::vector<size_t> redo;
	int rc = collect_groups(b, redo, launches, err, errlen);
	if (rc)
		return rc;
	// blocks whose match lists did not fit the pool (worst case 2 * (fb - 3) + 4 words per position): again, one at
	// a time, with a pool that cannot overflow
	for (size_t k : redo) {
		AsyncSub &S = b->subs[k];
		const unsigned mul = 2 * b->fb + 8;
		const SlotLay L = lay_for((size_t)S.job.u_len, S.cfg.fastMode != 0, mul);
		if (b->big.ensure(L.end) != cudaSuccess) {
			snprintf(err, errlen, "out of device memory for a %zu MiB LZMA match pool", L.end >> 20);
			return LRZGPU_ENOMEM;
		}
		std::vector<uint8_t *> bases(1, (uint8_t *)b->big.p);
		rc = enqueue_group(b, k, 1, bases, mul, false, nullptr, launches, err, errlen);
		if (rc)
			return rc < 0 ? rc : LRZGPU_EINTERNAL;
		std::vector<size_t> again;
		rc = collect_groups(b, again, launches, err, errlen);
		if (rc)
			return rc;
		if (!again.empty()) {
			snprintf(err, errlen, "LZMA match pool overflow with the worst-case pool");
			return LRZGPU_EINTERNAL;
		}
	}
	return LRZGPU_OK;
}

int backend_async_count(const BackendCtx *b) { return (int)b->subs.size(); }

const BlockJob *backend_async_result(const BackendCtx *b, int i)
{
	if (i < 0 || (size_t)i >= b->subs.size() || !b->subs[(size_t)i].done)
		return nullptr;
	return &b->subs[(size_t)i].job;
}

// All LZMA blocks of a chunk, synchronously: submit in as large groups as the work space takes, drain, repeat.
static int run_lzma(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, std::vector<BlockJob> &jobs,
		    const std::vector<int> &idx, cudaStream_t stream, int64_t *launches, char *err, size_t errlen)
{
	int64_t max_len = 0, sum = 0;
	for (int i : idx) {
		max_len = jobs[(size_t)i].u_len > max_len ? jobs[(size_t)i].u_len : max_len;
		sum += jobs[(size_t)i].u_len;
	}
	if (cudaStreamSynchronize(stream) != cudaSuccess) // the blocks' bytes were produced on the caller's stream
		return LRZGPU_ECUDA;
	int rc = backend_async_begin(b, p, sz, max_len, (int64_t)idx.size(), sum, err, errlen);
	if (rc)
		return rc;
	size_t at = 0;
	while (at < idx.size()) {
		size_t take = idx.size() - at;
		if (take > (size_t)b->nslots)
			take = (size_t)b->nslots;
		std::vector<BlockJob> grp;
		for (size_t k = 0; k < take; k++)
			grp.push_back(jobs[(size_t)idx[at + k]]);
		const auto t0 = std::chrono::steady_clock::now();
		rc = backend_async_submit(b, grp.data(), (int)take, nullptr, launches, err, errlen);
		if (rc == 1) {
			snprintf(err, errlen, "LZMA work space exhausted on an empty pipeline");
			rc = LRZGPU_EINTERNAL;
		}
		if (!rc)
			rc = backend_async_drain(b, launches, err, errlen);
		if (rc)
			return rc;
		if (getenv("LRZGPU_DEBUG"))
			fprintf(stderr, "[lrzgpu] lzma wave: %zu blocks in %.1f ms\n", take,
				std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		at += take;
	}
	for (size_t k = 0; k < idx.size(); k++) {
		const BlockJob *r = backend_async_result(b, (int)k);
		if (!r)
			return LRZGPU_EINTERNAL;
		BlockJob &bj = jobs[(size_t)idx[k]];
		bj.c_type = r->c_type;
		bj.c_len = r->c_len;
		bj.d_payload = r->d_payload;
	}
	b->active = false;
	return LRZGPU_OK;
}

int backend_encode_blocks(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, std::vector<BlockJob> &jobs,
			  int num_sms, cudaStream_t stream, int64_t *launches, char *err, size_t errlen)
{
	(void)num_sms;
	std::vector<int> idx;
	for (size_t i = 0; i < jobs.size(); i++)
		if (jobs[i].u_len >= 64) // src/stream.c:1633
			idx.push_back((int)i);
	if (idx.empty())
		return LRZGPU_OK;
	if (p.backend == LRZGPU_BACKEND_LZMA || p.backend == LRZGPU_BACKEND_ZSTD) // gate, match finder, block coder: pipelined
		return run_lzma(b, p, sz, jobs, idx, stream, launches, err, errlen);
	snprintf(err, errlen, "backend %d is not supported", p.backend);
	return LRZGPU_EUNSUPPORTED;
}

// Load this file's kernels now (CUDA loads a kernel's code at its first launch, and that load waits for every kernel
// that is running -- block encoders run for tens of seconds).
int backend_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, lz4_gate_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, lzma_block_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, zstd_assemble_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, zstd_encode_kernel) == cudaSuccess;
	ok = ok && lzma::mf_preload() == 0;
	return ok ? 0 : -1;
}
#endif // !LRZ_SIMT_HOST

} // namespace lrz
