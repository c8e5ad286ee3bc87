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
#include <string.h>

#include <vector>

#include "lz4_size.cuh"
#include "lzma_enc.cuh"

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
	lzma::Enc *enc;
	uint32_t *h2, *h3, *h4, *son;
	lzma::Config cfg;
	uint64_t outLen;
	int overflow;
};

__global__ void __launch_bounds__(32) lzma_block_kernel(LzmaJob *jobs)
{
	if (threadIdx.x == 0) {
		LzmaJob &j = jobs[blockIdx.x];
		lzma::enc_init(j.enc, j.cfg, j.src, j.n, j.out, j.outCap, j.h2, j.h3, j.h4, j.son);
		j.outLen = lzma::enc_run(j.enc);
		j.overflow = j.enc->overflow;
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
	DevBuf jobs, work, out, flags, offs;
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

static int run_lzma(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, std::vector<BlockJob> &jobs,
		    const std::vector<int> &idx, cudaStream_t stream, int64_t *launches, char *err, size_t errlen)
{
	const uint32_t fb = p.level < 7 ? 32 : 64; // src/stream.c:455
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	const size_t budget = free_b + b->work.cap + b->out.cap - (1ull << 30);

	size_t at = 0;
	while (at < idx.size()) {
		// one wave = as many blocks as fit the memory budget
		std::vector<LzmaJob> lj;
		std::vector<size_t> wofs, oofs;
		size_t wsum = 0, osum = 0, first = at;
		for (; at < idx.size(); at++) {
			const BlockJob &bj = jobs[idx[at]];
			LzmaJob j;
			memset(&j, 0, sizeof(j));
			if (!lzma::make_config(p.level, sz.dict_size, fb, (uint64_t)bj.u_len, j.cfg)) {
				snprintf(err, errlen, "LZMA level %d / block of %lld bytes is not supported by the device encoder",
					 p.level, (long long)bj.u_len);
				return LRZGPU_EUNSUPPORTED;
			}
			const size_t wneed = align_up(sizeof(lzma::Enc), 256) + align_up((lzma::kHash2Size + lzma::kHash3Size) * 4, 256) +
					     align_up(j.cfg.hash4Entries * 4, 256) + align_up(j.cfg.sonEntries * 4, 256);
			const size_t oneed = align_up((size_t)round_up_page((int64_t)((double)bj.u_len * 1.02), p.page_size), 256);
			if (!lj.empty() && wsum + osum + wneed + oneed > budget)
				break;
			j.src = bj.d_src;
			j.n = (uint32_t)bj.u_len;
			j.outCap = (uint64_t)round_up_page((int64_t)((double)bj.u_len * 1.02), p.page_size);
			lj.push_back(j);
			wofs.push_back(wsum);
			oofs.push_back(osum);
			wsum += wneed;
			osum += oneed;
		}
		if (b->work.ensure(wsum) != cudaSuccess || b->out.ensure(osum) != cudaSuccess ||
		    b->jobs.ensure(lj.size() * sizeof(LzmaJob)) != cudaSuccess) {
			snprintf(err, errlen, "out of device memory for %zu LZMA block encoders (%zu MiB)", lj.size(),
				 (wsum + osum) >> 20);
			return LRZGPU_ENOMEM;
		}
		for (size_t i = 0; i < lj.size(); i++) {
			uint8_t *w = (uint8_t *)b->work.p + wofs[i];
			lj[i].enc = (lzma::Enc *)w;
			w += align_up(sizeof(lzma::Enc), 256);
			lj[i].h2 = (uint32_t *)w;
			lj[i].h3 = lj[i].h2 + lzma::kHash2Size;
			w += align_up((lzma::kHash2Size + lzma::kHash3Size) * 4, 256);
			lj[i].h4 = (uint32_t *)w;
			w += align_up(lj[i].cfg.hash4Entries * 4, 256);
			lj[i].son = (uint32_t *)w;
			lj[i].out = (uint8_t *)b->out.p + oofs[i];
			// MatchFinder_Init_HighHash / _LowHash: hash heads start empty (son entries are written before use)
			const size_t hz = align_up((lzma::kHash2Size + lzma::kHash3Size) * 4, 256) + lj[i].cfg.hash4Entries * 4;
			if (cudaMemsetAsync(lj[i].h2, 0, hz, stream) != cudaSuccess)
				return LRZGPU_ECUDA;
		}
		if (cudaMemcpyAsync(b->jobs.p, lj.data(), lj.size() * sizeof(LzmaJob), cudaMemcpyHostToDevice, stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		lzma_block_kernel<<<(unsigned)lj.size(), 32, 0, stream>>>((LzmaJob *)b->jobs.p);
		if (launches)
			(*launches)++;
		if (cudaMemcpyAsync(lj.data(), b->jobs.p, lj.size() * sizeof(LzmaJob), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
			return LRZGPU_ECUDA;
		cudaError_t ce = cudaStreamSynchronize(stream);
		if (ce != cudaSuccess) {
			snprintf(err, errlen, "LZMA kernel failed: %s", cudaGetErrorString(ce));
			return LRZGPU_ECUDA;
		}
		const bool more_waves = at < idx.size();
		for (size_t i = 0; i < lj.size(); i++) {
			BlockJob &bj = jobs[idx[first + i]];
			// src/stream.c:482-487: kept only when smaller; SZ_ERROR_OUTPUT_EOF leaves the block stored
			if (!lj[i].overflow && (int64_t)lj[i].outLen < bj.u_len) {
				bj.c_type = LRZGPU_CTYPE_LZMA;
				bj.c_len = (int64_t)lj[i].outLen;
				bj.d_payload = lj[i].out;
			}
		}
		if (more_waves) {
			snprintf(err, errlen, "LZMA blocks of this chunk do not fit in device memory at once (%zu of %zu)", lj.size(),
				 idx.size());
			return LRZGPU_ENOMEM; // payloads of a wave would be overwritten by the next one
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
