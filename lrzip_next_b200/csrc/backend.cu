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
// is produced is a valid Zstandard frame (RFC 8878) of Raw and RLE blocks, which the reference's
// ZSTD_decompress() accepts; blocks that do not shrink are stored, like the reference does.
#include "backend.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <vector>

#include "lz4_size.cuh"
#include "lzma_enc.cuh"
#include "lzma_mf.h"

namespace lrz {

namespace {

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
	uint64_t outLen;
	int overflow;
};

// K7b: optimal parser + range coder, one block per CTA, over the precomputed match lists.
// The whole encoder state (probabilities, price tables, the 2048-cell parse table: 135 KB) lives in the
// SM's shared memory; warp 0 runs the encoder cooperatively (lzma_enc.cuh: replicated scalar code,
// lane-split loops).  Warp 1 is a pure look-ahead: it follows the encoder's position (e->pos in shared
// memory) and, for the next few positions, pulls into L1 what the parser will read from far away in HBM --
// the position's record, its match list in the pool, and for every (len, dist) pair the bytes at the end of
// the match that the MATCH : LIT : REP0 trial compares (LzmaEnc.c:1876-1893).  It writes nothing, so it
// cannot change the output; it only turns the parser's dependent HBM/L2 round trips into L1 hits.
constexpr uint32_t kLzmaAhead = 6; // positions

__device__ void lzma_lookahead_warp(const lzma::Enc *e, const LzmaJob &j, const volatile int *done)
{
	const uint32_t lane = threadIdx.x & 31u;
	const volatile uint32_t *ppos = &e->pos;
	const uint8_t *src = j.src;
	const uint32_t n = j.n;
	uint32_t upto = 0; // positions <= upto (1-based, like e->pos) are already prefetched
	while (!*done) {
		const uint32_t pos = *ppos;
		uint32_t lo = pos + 1 > upto + 1 ? pos + 1 : upto + 1;
		uint32_t hi = pos + kLzmaAhead < n ? pos + kLzmaAhead : n;
		if (lo > hi) {
			__nanosleep(200);
			continue;
		}
		for (uint32_t q = lo; q <= hi; q++) {
			const uint32_t i0 = q - 1;
			const uint64_t rec = j.rec[i0];
			const uint32_t nd = (uint32_t)rec & 1023u;
			const uint32_t *lst = j.pool + (rec >> 10);
			for (uint32_t k = lane; 2 * k < nd; k += 32) {
				const uint32_t len = lst[2 * k], dist = lst[2 * k + 1];
				if (dist < i0) {
					const uint8_t *far = src + i0 - dist - 1;
					lzma::lz_prefetch(far);
					lzma::lz_prefetch(far + len + 2);
				}
			}
		}
		upto = hi;
		__syncwarp();
	}
}

__global__ void __launch_bounds__(64, 1) lzma_block_kernel(LzmaJob *jobs)
{
	extern __shared__ __align__(16) uint8_t lzma_smem[];
	__shared__ int done;
	lzma::Enc *e = reinterpret_cast<lzma::Enc *>(lzma_smem);
	LzmaJob &j = jobs[blockIdx.x];
	if (threadIdx.x >= 32) {
		if (threadIdx.x == 32)
			done = 0;
		__syncthreads(); // the encoder state (e->pos) is initialised
		lzma_lookahead_warp(e, j, &done);
		return;
	}
	lzma::enc_init(e, j.cfg, j.src, j.n, j.out, j.outCap, nullptr, nullptr, nullptr, nullptr);
	e->preRec = j.rec;
	e->prePool = j.pool;
	__syncthreads();
	const uint64_t len = lzma::enc_run(e);
	__syncwarp();
	if (threadIdx.x == 0) {
		done = 1;
		j.outLen = len;
		j.overflow = e->overflow;
	}
}

// ---- zstd (Raw / RLE blocks) ----------------------------------------------------------------------
constexpr int64_t kZstdBlock = 128 * 1024;

// flags[c] = 1 when chunk c of the job is one repeated byte
__global__ void __launch_bounds__(256) zstd_scan_kernel(const uint8_t *src, int64_t len, uint8_t *flags)
{
	const int64_t c = blockIdx.x;
	const int64_t lo = c * kZstdBlock, hi = (lo + kZstdBlock < len) ? lo + kZstdBlock : len;
	const uint8_t first = src[lo];
	int same = 1;
	for (int64_t i = lo + threadIdx.x; i < hi; i += 256)
		if (src[i] != first)
			same = 0;
	same = __syncthreads_and(same);
	if (threadIdx.x == 0)
		flags[c] = (uint8_t)same;
}

// writes block c (header + payload) at out + offs[c]
__global__ void __launch_bounds__(256) zstd_emit_kernel(const uint8_t *src, int64_t len, const uint8_t *flags,
							 const int64_t *offs, int64_t nchunks, uint8_t *out)
{
	const int64_t c = blockIdx.x;
	const int64_t lo = c * kZstdBlock, hi = (lo + kZstdBlock < len) ? lo + kZstdBlock : len, size = hi - lo;
	uint8_t *w = out + offs[c];
	const int rle = flags[c];
	if (threadIdx.x == 0) {
		const uint32_t hdr = (uint32_t)(c == nchunks - 1) | ((rle ? 1u : 0u) << 1) | ((uint32_t)size << 3);
		w[0] = (uint8_t)hdr;
		w[1] = (uint8_t)(hdr >> 8);
		w[2] = (uint8_t)(hdr >> 16);
		if (rle)
			w[3] = src[lo];
	}
	if (!rle)
		for (int64_t i = threadIdx.x; i < size; i += 256)
			w[3 + i] = src[lo + i];
}

int64_t round_up_page(int64_t v, int page) { return v % page ? v + page - v % page : v; }

} // namespace

struct BackendCtx {
	DevBuf jobs, work, out, flags, offs, scratch;
};

BackendCtx *backend_create() { return new BackendCtx(); }

void backend_destroy(BackendCtx *b)
{
	if (!b)
		return;
	b->jobs.release();
	b->work.release();
	b->out.release();
	b->flags.release();
	b->offs.release();
	b->scratch.release();
	delete b;
}

static int run_gate(BackendCtx *b, std::vector<BlockJob> &jobs, const std::vector<int> &idx, int threshold,
		    std::vector<int> &pass, cudaStream_t stream, int64_t *launches)
{
	std::vector<GateJob> g(idx.size());
	for (size_t i = 0; i < idx.size(); i++) {
		g[i].src = jobs[idx[i]].d_src;
		g[i].len = jobs[idx[i]].u_len;
		g[i].result = 0;
	}
	if (b->jobs.ensure(g.size() * sizeof(GateJob)) != cudaSuccess)
		return LRZGPU_ENOMEM;
	if (cudaMemcpyAsync(b->jobs.p, g.data(), g.size() * sizeof(GateJob), cudaMemcpyHostToDevice, stream) != cudaSuccess)
		return LRZGPU_ECUDA;
	lz4_gate_kernel<<<(unsigned)g.size(), 32, 0, stream>>>((GateJob *)b->jobs.p, threshold);
	if (launches)
		(*launches)++;
	if (cudaMemcpyAsync(g.data(), b->jobs.p, g.size() * sizeof(GateJob), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
	    cudaStreamSynchronize(stream) != cudaSuccess)
		return LRZGPU_ECUDA;
	pass.clear();
	for (size_t i = 0; i < idx.size(); i++)
		if (g[i].result)
			pass.push_back(idx[i]);
	return LRZGPU_OK;
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

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// All LZMA blocks of a chunk: K7a (match finder, data-parallel over positions and buckets) then K7b
// (parser + range coder, one CTA per block).  Work arrays are laid out per block in one arena; blocks
// are taken in waves when the arena does not fit in free HBM.  Payloads of all waves stay in b->out.
static int run_lzma(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, std::vector<BlockJob> &jobs,
		    const std::vector<int> &idx, cudaStream_t stream, int64_t *launches, char *err, size_t errlen)
{
	const uint32_t fb = p.level < 7 ? 32 : 64; // src/stream.c:455
	if (lzma::mf_init_tables()) {
		snprintf(err, errlen, "LZMA match finder tables: %s", cudaGetErrorString(cudaGetLastError()));
		return LRZGPU_ECUDA;
	}
	// payload area for every block of the chunk
	std::vector<size_t> oofs(idx.size());
	std::vector<lzma::Config> cfgs(idx.size());
	size_t osum = 0;
	uint32_t maxCount = 0;
	for (size_t k = 0; k < idx.size(); k++) {
		const BlockJob &bj = jobs[idx[k]];
		if (!lzma::make_config(p.level, sz.dict_size, fb, (uint64_t)bj.u_len, cfgs[k]) || fb > lzma::kMfMaxFb) {
			snprintf(err, errlen, "LZMA level %d / block of %lld bytes is not supported by the device encoder", p.level,
				 (long long)bj.u_len);
			return LRZGPU_EUNSUPPORTED;
		}
		oofs[k] = osum;
		osum += align_up((size_t)round_up_page((int64_t)((double)bj.u_len * 1.02), p.page_size), 256);
		const uint32_t minAvail = cfgs[k].fastMode ? 5 : 4;
		const uint32_t count = bj.u_len >= minAvail ? (uint32_t)bj.u_len - (minAvail - 1) : 0;
		if (count > maxCount)
			maxCount = count;
	}
	const bool hc5 = cfgs[0].fastMode != 0; // one level per call: all blocks use the same finder
	const size_t minAvail = hc5 ? 5 : 4;
	const size_t scratch_bytes = align_up(lzma::mf_sort_scratch_bytes(maxCount), 256);
	if (b->out.ensure(osum) != cudaSuccess || b->scratch.ensure(scratch_bytes) != cudaSuccess) {
		snprintf(err, errlen, "out of device memory for LZMA payloads / sort scratch (%zu MiB)", (osum + scratch_bytes) >> 20);
		return LRZGPU_ENOMEM;
	}
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	const size_t budget = free_b + b->work.cap > (2ull << 30) ? free_b + b->work.cap - (1ull << 30) : free_b + b->work.cap;

	struct Lay {
		size_t son, c2, c3, sorted, ctl, rec, pool, end;
		uint64_t poolCap;
	};
	size_t at = 0;
	unsigned poolMul = 12; // uint32 of match-list pool per input byte; text needs ~6-9, raised on overflow
	while (at < idx.size()) {
		std::vector<Lay> lay;
		size_t wsum = 0, first = at;
		const auto t_wave = std::chrono::steady_clock::now();
		for (; at < idx.size(); at++) {
			const size_t n = (size_t)jobs[idx[at]].u_len, count = n >= minAvail ? n - (minAvail - 1) : 0;
			Lay L;
			size_t o = wsum;
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
			if (!lay.empty() && o > budget)
				break;
			lay.push_back(L);
			wsum = o;
		}
		if (b->work.ensure(wsum) != cudaSuccess || b->jobs.ensure(lay.size() * (sizeof(LzmaJob) + sizeof(lzma::MfBlock) + 8) + 64) != cudaSuccess) {
			snprintf(err, errlen, "out of device memory for %zu LZMA block encoders (%zu MiB)", lay.size(), wsum >> 20);
			return LRZGPU_ENOMEM;
		}
		uint8_t *W = (uint8_t *)b->work.p;
		std::vector<LzmaJob> lj(lay.size());
		std::vector<lzma::MfBlock> mb(lay.size());
		std::vector<uint64_t> seg(lay.size() + 1, 0);
		for (size_t i = 0; i < lay.size(); i++) {
			const BlockJob &bj = jobs[idx[first + i]];
			const Lay &L = lay[i];
			const lzma::Config &c = cfgs[first + i];
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
			j.out = (uint8_t *)b->out.p + oofs[first + i];
			j.outCap = (uint64_t)round_up_page((int64_t)((double)bj.u_len * 1.02), p.page_size);
			j.rec = B.rec;
			j.pool = B.pool;
			j.cfg = c;
			if (cudaMemsetAsync(W + L.ctl, 0, 256 + 8 * (size_t)bj.u_len, stream) != cudaSuccess)
				return LRZGPU_ECUDA;
			if (lzma::mf_prepare_block(B, b->scratch.p, stream, launches)) {
				snprintf(err, errlen, "LZMA match finder sort: %s", cudaGetErrorString(cudaGetLastError()));
				return LRZGPU_ECUDA;
			}
		}
		// job table: [LzmaJob x m][MfBlock x m][segBase x (m + 1)]
		uint8_t *J = (uint8_t *)b->jobs.p;
		const size_t o_mb = lay.size() * sizeof(LzmaJob), o_seg = o_mb + lay.size() * sizeof(lzma::MfBlock);
		if (cudaMemcpyAsync(J, lj.data(), o_mb, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
		    cudaMemcpyAsync(J + o_mb, mb.data(), lay.size() * sizeof(lzma::MfBlock), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
		    cudaMemcpyAsync(J + o_seg, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		if (lzma::mf_walk_launch((const lzma::MfBlock *)(J + o_mb), (int)lay.size(), (const uint64_t *)(J + o_seg), seg.back(),
					 hc5, stream, launches)) {
			snprintf(err, errlen, "LZMA match finder walk: %s", cudaGetErrorString(cudaGetLastError()));
			return LRZGPU_ECUDA;
		}
		// pool overflow check before the parser consumes the lists
		std::vector<int> ovf(lay.size(), 0);
		for (size_t i = 0; i < lay.size(); i++)
			if (cudaMemcpyAsync(&ovf[i], mb[i].overflow, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
				return LRZGPU_ECUDA;
		cudaError_t ce = cudaStreamSynchronize(stream);
		if (ce != cudaSuccess) {
			snprintf(err, errlen, "LZMA match finder failed: %s", cudaGetErrorString(ce));
			return LRZGPU_ECUDA;
		}
		const auto t_mf = std::chrono::steady_clock::now();
		bool any_ovf = false;
		for (int v : ovf)
			any_ovf = any_ovf || v;
		if (any_ovf) { // redo this wave with a larger pool (worst case 2 * (fb - 3) + 4 words per position)
			if (poolMul >= 2 * fb)
				return LRZGPU_EINTERNAL;
			poolMul = poolMul < 40 ? 40 : 2 * fb;
			at = first;
			continue;
		}
		if (cudaFuncSetAttribute(lzma_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lzma::Enc)) != cudaSuccess) {
			snprintf(err, errlen, "LZMA encoder state (%zu bytes) does not fit in shared memory", sizeof(lzma::Enc));
			return LRZGPU_ECUDA;
		}
		lzma_block_kernel<<<(unsigned)lj.size(), 64, sizeof(lzma::Enc), stream>>>((LzmaJob *)J);
		if (launches)
			(*launches)++;
		if (cudaMemcpyAsync(lj.data(), J, o_mb, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		ce = cudaStreamSynchronize(stream);
		if (ce != cudaSuccess) {
			snprintf(err, errlen, "LZMA kernel failed: %s", cudaGetErrorString(ce));
			return LRZGPU_ECUDA;
		}
		if (getenv("LRZGPU_DEBUG")) {
			const auto t_end = std::chrono::steady_clock::now();
			fprintf(stderr, "[lrzgpu] lzma wave: %zu blocks, %llu positions, arena %zu MiB, match finder %.1f ms, parser %.1f ms\n",
				lj.size(), (unsigned long long)seg.back(), wsum >> 20,
				std::chrono::duration<double, std::milli>(t_mf - t_wave).count(),
				std::chrono::duration<double, std::milli>(t_end - t_mf).count());
		}
		for (size_t i = 0; i < lj.size(); i++) {
			BlockJob &bj = jobs[idx[first + i]];
			// src/stream.c:482-487: kept only when smaller; SZ_ERROR_OUTPUT_EOF leaves the block stored
			if (!lj[i].overflow && (int64_t)lj[i].outLen < bj.u_len) {
				bj.c_type = LRZGPU_CTYPE_LZMA;
				bj.c_len = (int64_t)lj[i].outLen;
				bj.d_payload = lj[i].out;
			}
		}
	}
	return LRZGPU_OK;
}

static int run_zstd(BackendCtx *b, std::vector<BlockJob> &jobs, const std::vector<int> &idx, cudaStream_t stream,
		    int64_t *launches)
{
	// layout of all frames in b->out
	size_t osum = 0;
	std::vector<size_t> oofs;
	for (int i : idx) {
		oofs.push_back(osum);
		osum += align_up((size_t)jobs[i].u_len + 64, 256);
	}
	if (b->out.ensure(osum) != cudaSuccess)
		return LRZGPU_ENOMEM;
	for (size_t k = 0; k < idx.size(); k++) {
		BlockJob &bj = jobs[idx[k]];
		const int64_t n = bj.u_len, nchunks = (n + kZstdBlock - 1) / kZstdBlock;
		if (b->flags.ensure((size_t)nchunks) != cudaSuccess || b->offs.ensure((size_t)nchunks * 8) != cudaSuccess)
			return LRZGPU_ENOMEM;
		zstd_scan_kernel<<<(unsigned)nchunks, 256, 0, stream>>>(bj.d_src, n, (uint8_t *)b->flags.p);
		std::vector<uint8_t> flags((size_t)nchunks);
		if (cudaMemcpyAsync(flags.data(), b->flags.p, (size_t)nchunks, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
		    cudaStreamSynchronize(stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		if (launches)
			(*launches)++;
		// frame header: magic, FHD (single segment, FCS size by value), FCS
		uint8_t hdr[16];
		int hl = 0;
		hdr[hl++] = 0x28;
		hdr[hl++] = 0xB5;
		hdr[hl++] = 0x2F;
		hdr[hl++] = 0xFD;
		if (n < 256) {
			hdr[hl++] = 0x20;
			hdr[hl++] = (uint8_t)n;
		} else if (n < 65536 + 256) {
			hdr[hl++] = 0x60;
			hdr[hl++] = (uint8_t)(n - 256);
			hdr[hl++] = (uint8_t)((n - 256) >> 8);
		} else if (n <= 0xFFFFFFFFll) {
			hdr[hl++] = 0xA0;
			for (int i = 0; i < 4; i++)
				hdr[hl++] = (uint8_t)(n >> (8 * i));
		} else {
			hdr[hl++] = 0xE0;
			for (int i = 0; i < 8; i++)
				hdr[hl++] = (uint8_t)(n >> (8 * i));
		}
		std::vector<int64_t> offs((size_t)nchunks);
		int64_t total = hl;
		for (int64_t c = 0; c < nchunks; c++) {
			const int64_t size = (c == nchunks - 1) ? n - c * kZstdBlock : kZstdBlock;
			offs[(size_t)c] = total;
			total += 3 + (flags[(size_t)c] ? 1 : size);
		}
		if (total >= n)
			continue; // src/stream.c:215-221: not smaller => stays CTYPE_NONE
		uint8_t *out = (uint8_t *)b->out.p + oofs[k];
		if (cudaMemcpyAsync(out, hdr, (size_t)hl, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
		    cudaMemcpyAsync(b->offs.p, offs.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		zstd_emit_kernel<<<(unsigned)nchunks, 256, 0, stream>>>(bj.d_src, n, (const uint8_t *)b->flags.p,
									 (const int64_t *)b->offs.p, nchunks, out);
		if (launches)
			(*launches)++;
		if (cudaStreamSynchronize(stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		bj.c_type = LRZGPU_CTYPE_ZSTD;
		bj.c_len = total;
		bj.d_payload = out;
	}
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
	if (p.threshold) { // LZ4_TEST (FLAG_THRESHOLD): incompressible blocks stay stored
		std::vector<int> pass;
		int rc = run_gate(b, jobs, idx, p.threshold, pass, stream, launches);
		if (rc) {
			snprintf(err, errlen, "lz4 gate kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
			return rc;
		}
		idx.swap(pass);
		if (idx.empty())
			return LRZGPU_OK;
	}
	if (p.backend == LRZGPU_BACKEND_LZMA)
		return run_lzma(b, p, sz, jobs, idx, stream, launches, err, errlen);
	if (p.backend == LRZGPU_BACKEND_ZSTD) {
		int rc = run_zstd(b, jobs, idx, stream, launches);
		if (rc)
			snprintf(err, errlen, "zstd backend failed: %s", cudaGetErrorString(cudaGetLastError()));
		return rc;
	}
	snprintf(err, errlen, "backend %d is not supported", p.backend);
	return LRZGPU_EUNSUPPORTED;
}

} // namespace lrz
