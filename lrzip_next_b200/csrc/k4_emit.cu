// k4_emit.cu -- K4: turn the committed match records into rzip's two output streams, plus the
// per-chunk CRC-32.
//
//   stream 0 (src/rzip.c:184-265): for every record, literal headers  [00 len16]*  then match
//            headers [01 len16 dist(cb)]*, runs longer than 0xFFFF split; the chunk closes with the
//            terminator 00 00 00 and the CRC-32 in gcrypt digest order (big endian) (:757-760)
//   stream 1: the literal bytes, i.e. the concatenation of the unmatched input ranges (:229-246)
//   CRC-32  : src/rzip.c:564-584, 690-757 feed the whole chunk through GCRY_MD_CRC32 (IEEE 802.3,
//            reflected, init/xorout 0xFFFFFFFF)
//
// The commit kernel already stored each record's offsets in both streams, so all three kernels here
// are embarrassingly parallel and HBM bound: the literal gather reads and writes |stream 1| bytes,
// the CRC reads the chunk once.
#include "kernels.h"

namespace lrz {

// ------------------------------------------------------------------------------------------------
// CRC-32: per-thread table-driven remainders of 128-byte pieces, combined with carry-less
// multiplications by x^(8*bytes_after) mod P (the algebra behind zlib's crc32_combine).
static constexpr uint32_t kCrcPoly = 0xedb88320u;
static constexpr int CRC_THREADS = 256;
static constexpr int CRC_PIECE = 128;                       // bytes per thread per tile
static constexpr int CRC_TILE = CRC_THREADS * CRC_PIECE;    // 32 KiB
static constexpr int CRC_SMEM_TABLE = 256 * 32 * 4;         // bank-replicated byte table
static constexpr int CRC_SMEM_DATA = CRC_THREADS * 33 * 4;  // pieces padded to 33 words
static constexpr int CRC_SMEM = CRC_SMEM_TABLE + CRC_SMEM_DATA + 64;

__constant__ uint32_t c_crc_table[256];
__constant__ uint32_t c_x2n[32];        // x^(2^k) mod P, reflected
__constant__ uint32_t c_piece_shift[CRC_THREADS]; // x^(8*128*(255-l)) mod P

static uint32_t h_multmodp(uint32_t a, uint32_t b)
{
	uint32_t m = 1u << 31, p = 0;
	for (;;) {
		if (a & m) {
			p ^= b;
			if ((a & (m - 1)) == 0)
				break;
		}
		m >>= 1;
		b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
	}
	return p;
}

__device__ __forceinline__ uint32_t d_multmodp(uint32_t a, uint32_t b)
{
	uint32_t p = 0;
#pragma unroll 4
	for (int i = 31; i >= 0; i--) {
		if ((a >> i) & 1)
			p ^= b;
		b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
	}
	return p;
}

// x^(8*nbytes) mod P
__device__ __forceinline__ uint32_t d_xpow_bytes(int64_t nbytes)
{
	uint32_t p = 1u << 31;
	unsigned k = 3;
	while (nbytes) {
		if (nbytes & 1)
			p = d_multmodp(c_x2n[k & 31], p);
		nbytes >>= 1;
		k++;
	}
	return p;
}

__global__ void __launch_bounds__(CRC_THREADS)
crc32_kernel(const uint8_t *__restrict__ buf, int64_t n, uint32_t *__restrict__ acc)
{
#if defined(LRZ_SIMT_HOST)
	uint8_t *smem = simt::dyn_smem();
#else
	extern __shared__ __align__(16) uint8_t smem[];
#endif
	uint32_t *tab = reinterpret_cast<uint32_t *>(smem);                   // tab[b * 32 + lane]
	uint32_t *data = reinterpret_cast<uint32_t *>(smem + CRC_SMEM_TABLE); // piece l at word l * 33
	uint32_t *red = reinterpret_cast<uint32_t *>(smem + CRC_SMEM_TABLE + CRC_SMEM_DATA);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	for (int i = tid; i < 256 * 32; i += CRC_THREADS)
		tab[i] = c_crc_table[i >> 5];
	__syncthreads();
	const uint32_t *my_tab = tab + lane;

	const int64_t num_tiles = (n + CRC_TILE - 1) / CRC_TILE;
	for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
		const int64_t t0 = t * CRC_TILE;
		const int64_t t1 = (t0 + CRC_TILE < n) ? t0 + CRC_TILE : n;
		// coalesced 16-byte loads (the chunk is followed by zero padding), scattered to padded rows
#pragma unroll
		for (int r = 0; r < CRC_TILE / (16 * CRC_THREADS); r++) {
			const int v = r * CRC_THREADS + tid; // 16-byte vector index inside the tile
			uint4 q = make_uint4(0, 0, 0, 0);
			if (t0 + (int64_t)v * 16 < n)
				q = __ldcs(reinterpret_cast<const uint4 *>(buf + t0) + v);
			const int piece = v >> 3, w = (v & 7) * 4;
			uint32_t *d = data + piece * 33 + w;
			d[0] = q.x;
			d[1] = q.y;
			d[2] = q.z;
			d[3] = q.w;
		}
		__syncthreads();
		const int64_t p0 = t0 + (int64_t)tid * CRC_PIECE;
		int64_t plen = t1 - p0;
		if (plen > CRC_PIECE)
			plen = CRC_PIECE;
		uint32_t contrib = 0;
		if (plen > 0) {
			uint32_t crc = (p0 == 0) ? 0xffffffffu : 0u;
			const uint32_t *d = data + tid * 33;
			if (plen == CRC_PIECE) {
#pragma unroll 8
				for (int w = 0; w < CRC_PIECE / 4; w++) {
					uint32_t x = d[w] ^ crc;
					crc = 0;
#pragma unroll
					for (int b = 0; b < 4; b++)
						x = my_tab[(x & 0xff) * 32] ^ (x >> 8);
					crc = x;
				}
			} else {
				for (int i = 0; i < (int)plen; i++) {
					const uint32_t byte = (d[i >> 2] >> ((i & 3) * 8)) & 0xff;
					crc = my_tab[((crc ^ byte) & 0xff) * 32] ^ (crc >> 8);
				}
			}
			// shift to the end of the tile
			const int64_t after = t1 - (p0 + plen);
			const uint32_t sh = (t1 - t0 == CRC_TILE) ? c_piece_shift[tid] : d_xpow_bytes(after);
			contrib = (after == 0) ? crc : d_multmodp(sh, crc);
		}
#pragma unroll
		for (int o = 16; o; o >>= 1)
			contrib ^= __shfl_xor_sync(0xffffffffu, contrib, o);
		if (lane == 0)
			red[warp] = contrib;
		__syncthreads();
		if (tid == 0) {
			uint32_t tile_crc = 0;
			for (int w = 0; w < CRC_THREADS / 32; w++)
				tile_crc ^= red[w];
			const int64_t after = n - t1;
			if (after)
				tile_crc = d_multmodp(d_xpow_bytes(after), tile_crc);
			atomicXor(acc, tile_crc);
		}
		__syncthreads();
	}
}

int k4_init_tables()
{
	uint32_t table[256], x2n[32], shift[CRC_THREADS];
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t c = i;
		for (int k = 0; k < 8; k++)
			c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
		table[i] = c;
	}
	x2n[0] = 1u << 30;
	for (int k = 1; k < 32; k++)
		x2n[k] = h_multmodp(x2n[k - 1], x2n[k - 1]);
	for (int l = 0; l < CRC_THREADS; l++) {
		int64_t nbytes = (int64_t)CRC_PIECE * (CRC_THREADS - 1 - l);
		uint32_t p = 1u << 31;
		unsigned k = 3;
		while (nbytes) {
			if (nbytes & 1)
				p = h_multmodp(x2n[k & 31], p);
			nbytes >>= 1;
			k++;
		}
		shift[l] = p;
	}
	if (cudaMemcpyToSymbol(c_crc_table, table, sizeof(table)) != cudaSuccess ||
	    cudaMemcpyToSymbol(c_x2n, x2n, sizeof(x2n)) != cudaSuccess ||
	    cudaMemcpyToSymbol(c_piece_shift, shift, sizeof(shift)) != cudaSuccess)
		return -1;
#if !defined(LRZ_SIMT_HOST)
	if (cudaFuncSetAttribute(crc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CRC_SMEM) != cudaSuccess)
		return -1;
#endif
	return 0;
}

#if !defined(LRZ_SIMT_HOST) // launchers: the emulator (tests/hostsim) calls the kernels directly
int crc32_launch(const uint8_t *d_buf, int64_t n, uint32_t *d_crc, int num_sms, cudaStream_t stream)
{
	if (cudaMemsetAsync(d_crc, 0, sizeof(uint32_t), stream) != cudaSuccess)
		return -1;
	if (n <= 0)
		return 0;
	int64_t tiles = (n + CRC_TILE - 1) / CRC_TILE;
	int64_t grid = (int64_t)num_sms * 3;
	if (grid > tiles)
		grid = tiles;
	crc32_kernel<<<(unsigned)grid, CRC_THREADS, CRC_SMEM, stream>>>(d_buf, n, d_crc);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#endif

// ------------------------------------------------------------------------------------------------
// Stream 0: one warp per record, lanes stride over the record's 0xFFFF-byte pieces.
__device__ __forceinline__ void put_hdr(uint8_t *d, uint8_t head, int64_t len)
{
	d[0] = head;
	d[1] = (uint8_t)(len & 0xff);
	d[2] = (uint8_t)((len >> 8) & 0xff);
}

__global__ void __launch_bounds__(256)
k4_headers_kernel(const MatchRec *__restrict__ recs, int64_t n_rec, int cb, const uint32_t *__restrict__ crc_acc,
		  uint8_t *__restrict__ s0)
{
	const int lane = threadIdx.x & 31;
	const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t i = warp0; i < n_rec; i += nwarps) {
		const MatchRec r = recs[i];
		const int64_t lp = pieces_of(r.lit_len), mp = pieces_of(r.len);
		uint8_t *base = s0 + r.s0_off;
		for (int64_t k = lane; k < lp; k += 32) {
			const int64_t len = (k < lp - 1) ? 0xFFFF : r.lit_len - 0xFFFF * (lp - 1);
			put_hdr(base + 3 * k, 0, len);
		}
		base += 3 * lp;
		const uint64_t dist = (uint64_t)(r.p - r.ofs);
		for (int64_t k = lane; k < mp; k += 32) {
			const int64_t len = (k < mp - 1) ? 0xFFFF : r.len - 0xFFFF * (mp - 1);
			uint8_t *d = base + (3 + cb) * k;
			put_hdr(d, 1, len);
			for (int b = 0; b < cb; b++)
				d[3 + b] = (uint8_t)(dist >> (8 * b));
		}
		if (i == n_rec - 1 && lane == 0) { // closing record: terminator + CRC
			uint8_t *d = base + (3 + cb) * mp;
			const uint32_t crc = *crc_acc ^ 0xffffffffu;
			d[0] = d[1] = d[2] = 0;
			d[3] = (uint8_t)(crc >> 24);
			d[4] = (uint8_t)(crc >> 16);
			d[5] = (uint8_t)(crc >> 8);
			d[6] = (uint8_t)crc;
		}
	}
}

#if !defined(LRZ_SIMT_HOST)
int k4_headers_launch(const MatchRec *d_recs, int64_t n_rec, int chunk_bytes, const uint32_t *d_crc,
		      uint8_t *d_s0, cudaStream_t stream)
{
	if (n_rec <= 0)
		return 0;
	int64_t blocks = (n_rec + 7) / 8;
	if (blocks > 148 * 8)
		blocks = 148 * 8;
	k4_headers_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_recs, n_rec, chunk_bytes, d_crc, d_s0);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#endif

// ------------------------------------------------------------------------------------------------
// Stream 1: output-driven gather.  Each thread owns 16 consecutive output bytes (aligned 16-byte
// store); the owning record is found by binary search on the records' s1_off inside the small range
// of records that overlap the CTA's tile.
static constexpr int LIT_THREADS = 256;
static constexpr int LIT_TILE = LIT_THREADS * 16 * 4; // 16 KiB per CTA iteration

// last record index in [lo, hi] with s1_off <= x
__device__ __forceinline__ int64_t rec_of(const MatchRec *recs, int64_t lo, int64_t hi, int64_t x)
{
	while (lo < hi) {
		const int64_t mid = (lo + hi + 1) >> 1;
		if (recs[mid].s1_off <= x)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ uint4 load16u_g(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
	const unsigned sh = (unsigned)(a & 3) * 8;
	const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
	if (!sh)
		return make_uint4(w0, w1, w2, w3);
	const uint32_t w4 = __ldg(q + 4);
	return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
			  __funnelshift_r(w3, w4, sh));
}

__global__ void __launch_bounds__(LIT_THREADS)
k4_literals_kernel(const uint8_t *__restrict__ buf, const MatchRec *__restrict__ recs, int64_t n_rec, int64_t s1_from,
		   int64_t s1_len, uint8_t *__restrict__ s1)
{
	__shared__ int64_t s_range[2];
	const int64_t first_tile = s1_from / LIT_TILE; // s1_from is a multiple of 16: tiles are re-done whole
	const int64_t num_tiles = (s1_len + LIT_TILE - 1) / LIT_TILE;
	for (int64_t t = first_tile + blockIdx.x; t < num_tiles; t += gridDim.x) {
		const int64_t t0 = t * LIT_TILE;
		const int64_t t1 = (t0 + LIT_TILE < s1_len) ? t0 + LIT_TILE : s1_len;
		if (threadIdx.x < 2)
			s_range[threadIdx.x] = rec_of(recs, 0, n_rec - 1, threadIdx.x ? t1 - 1 : t0);
		__syncthreads();
		const int64_t rlo = s_range[0], rhi = s_range[1];
		for (int64_t x = t0 + threadIdx.x * 16; x < t1; x += LIT_THREADS * 16) {
			int64_t i = rec_of(recs, rlo, rhi, x);
			int64_t so = recs[i].s1_off, ll = recs[i].lit_len, src = recs[i].p - ll;
			if (x + 16 <= so + ll && x + 16 <= s1_len) {
				const uint4 v = load16u_g(buf + src + (x - so));
				*reinterpret_cast<uint4 *>(s1 + x) = v;
			} else {
				const int64_t xe = (x + 16 < s1_len) ? x + 16 : s1_len;
				for (int64_t y = x; y < xe; y++) {
					while (y >= so + ll) {
						i++;
						so = recs[i].s1_off;
						ll = recs[i].lit_len;
						src = recs[i].p - ll;
					}
					s1[y] = buf[src + (y - so)];
				}
			}
		}
		__syncthreads();
	}
}

#if !defined(LRZ_SIMT_HOST)
int k4_literals_launch(const uint8_t *d_buf, const MatchRec *d_recs, int64_t n_rec, int64_t s1_from, int64_t s1_len,
		       uint8_t *d_s1, int num_sms, cudaStream_t stream)
{
	if (s1_len <= s1_from || n_rec <= 0)
		return 0;
	int64_t tiles = (s1_len + LIT_TILE - 1) / LIT_TILE - s1_from / LIT_TILE;
	int64_t grid = (int64_t)num_sms * 8;
	if (grid > tiles)
		grid = tiles;
	k4_literals_kernel<<<(unsigned)grid, LIT_THREADS, 0, stream>>>(d_buf, d_recs, n_rec, s1_from, s1_len, d_s1);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#endif

// ------------------------------------------------------------------------------------------------
// Flush order (src/stream.c:2198-2216 write_stream flushes a stream buffer the moment it is full;
// src/rzip.c:248-265 put_literal writes a piece's 3 header bytes to stream 0 before its payload goes
// to stream 1).  For the stream-0 byte at offset (j+1)*bufsize - 1 report how many stream-1 bytes
// were already written when it was written.
__global__ void k4_flush_order_kernel(const MatchRec *__restrict__ recs, int64_t n_rec, int cb, int64_t bufsize,
				      int64_t n_bounds, int64_t *__restrict__ w1)
{
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_bounds)
		return;
	const int64_t x = (j + 1) * bufsize - 1; // offset of the byte that fills block j
	int64_t lo = 0, hi = n_rec - 1;
	while (lo < hi) {
		const int64_t mid = (lo + hi + 1) >> 1;
		if (recs[mid].s0_off <= x)
			lo = mid;
		else
			hi = mid - 1;
	}
	const MatchRec r = recs[lo];
	const int64_t lp = pieces_of(r.lit_len);
	const int64_t rel = x - r.s0_off;
	int64_t written;
	if (rel < 3 * lp) {
		const int64_t piece = rel / 3; // header of literal piece `piece`: earlier pieces' payloads are out
		written = r.s1_off + piece * 0xFFFF;
	} else
		written = r.s1_off + r.lit_len;
	w1[j] = written;
}

#if !defined(LRZ_SIMT_HOST)
int k4_flush_order_launch(const MatchRec *d_recs, int64_t n_rec, int chunk_bytes, int64_t bufsize,
			  int64_t n_bounds, int64_t *d_w1, cudaStream_t stream)
{
	if (n_bounds <= 0)
		return 0;
	k4_flush_order_kernel<<<(unsigned)((n_bounds + 127) / 128), 128, 0, stream>>>(d_recs, n_rec, chunk_bytes, bufsize,
										       n_bounds, d_w1);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Load this file's kernels now (CUDA loads a kernel's code at its first launch, and that load waits for every kernel
// that is running -- block encoders run for tens of seconds).
int k4_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, crc32_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, k4_flush_order_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, k4_headers_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, k4_literals_kernel) == cudaSuccess;
	return ok ? 0 : -1;
}
#endif // !LRZ_SIMT_HOST

} // namespace lrz
