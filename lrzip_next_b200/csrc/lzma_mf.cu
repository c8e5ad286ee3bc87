// lzma_mf.cu -- K7a: the LZMA match finder of all stream blocks of a chunk as data-parallel kernels
// (see lzma_mf.cuh for why this is exact).  Per block:
//
//   1. three stable LSD radix sorts (8-bit digits) of the block's positions by their 2-, 3- and 4-byte
//      hash (LzFind.c:49, LzFindMt.c:368-394); keys of the first pass are computed from the bytes;
//   2. neighbours in sorted order give c2[p] / c3[p] (the hash2 / hash3 heads position p would see,
//      LzFindMt.c:1093-1131) and the 4-byte-hash buckets with their positions in increasing order;
// then for all blocks in ONE launch:
//   3. mf_walk_kernel: one thread per bucket inserts the bucket's positions into the bucket's binary
//      tree (GetMatchesSpec1, LzFind.c:962-1029), mixes in the 2-/3-byte candidates and appends the
//      position's match list to the block's pool; rec[p] = (pool offset << 10) | count.
//
// Roofline class: the sorts are HBM streaming passes (16 B per position and pass); the walk is a
// latency-bound pointer chase whose parallelism is the number of buckets (millions per block).
#include "lzma_mf.h"

#include <stdio.h>

namespace lrz {
namespace lzma {

namespace {

__constant__ uint32_t c_crc[256];

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS; // 4096 keys per CTA; warp w owns keys [512 w, 512 (w+1)) of the tile

enum { KEY_H4 = 0, KEY_H3 = 1, KEY_H2 = 2, KEY_H5 = 3 };

struct KeySrc { // first pass: keys are hashes of the block bytes, values are the positions themselves
	const uint8_t *src;
	uint32_t hashMask, bigHash;
	int kind;
};

__device__ __forceinline__ uint32_t key_of(const KeySrc &ks, uint32_t i)
{
	const uint8_t *cur = ks.src + i;
	if (ks.kind == KEY_H4)
		return mf_hash4(c_crc, cur, ks.hashMask, ks.bigHash);
	if (ks.kind == KEY_H5)
		return mf_hash5(c_crc, cur, ks.hashMask);
	if (ks.kind == KEY_H3)
		return mf_hash3(c_crc, cur);
	return mf_hash2(c_crc, cur);
}

// hist[digit * numTiles + tile] = number of keys of the tile with that digit
template <bool kFromSrc>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t *__restrict__ keys, KeySrc ks, uint32_t count,
							      int shift, uint32_t *__restrict__ hist, uint32_t numTiles)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RS_TILE + (threadIdx.x >> 5) * (32 * RS_ROUNDS) + (threadIdx.x & 31);
#pragma unroll 4
	for (int r = 0; r < RS_ROUNDS; r++) {
		const uint32_t i = base + r * 32;
		if (i < count) {
			const uint32_t k = kFromSrc ? key_of(ks, i) : keys[i];
			atomicAdd(&h[(k >> shift) & 255u], 1u);
		}
	}
	__syncthreads();
	hist[threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of a[0..m) in place, one CTA of 1024 threads
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *a, uint32_t m)
{
	__shared__ uint32_t wsum[32];
	__shared__ uint32_t carry_s;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0)
		carry_s = 0;
	__syncthreads();
	for (uint32_t base = 0; base < m; base += 4096) {
		const uint32_t i0 = base + tid * 4;
		uint32_t v[4];
#pragma unroll
		for (int j = 0; j < 4; j++)
			v[j] = (i0 + j < m) ? a[i0 + j] : 0;
		const uint32_t mine = v[0] + v[1] + v[2] + v[3];
		uint32_t inc = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d)
				inc += o;
		}
		if (lane == 31)
			wsum[warp] = inc;
		__syncthreads();
		uint32_t woff = 0;
		for (int w = 0; w < warp; w++)
			woff += wsum[w];
		uint32_t total = 0;
		if (tid == 1023)
			total = woff + inc;
		uint32_t run = carry_s + woff + inc - mine;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (i0 + j < m)
				a[i0 + j] = run;
			run += v[j];
		}
		__syncthreads();
		if (tid == 1023)
			carry_s += total;
		__syncthreads();
	}
}

// stable scatter of the tile's keys (and values) to their sorted places for this digit
template <bool kFromSrc>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t *__restrict__ keys_in,
								 const uint32_t *__restrict__ vals_in, KeySrc ks, uint32_t count,
								 int shift, const uint32_t *__restrict__ hist, uint32_t numTiles,
								 uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	__shared__ uint32_t wcnt[RS_WARPS][256];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS)
		(&wcnt[0][0])[i] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RS_TILE + warp * (32 * RS_ROUNDS) + lane;
	uint32_t k[RS_ROUNDS], v[RS_ROUNDS];
#pragma unroll
	for (int r = 0; r < RS_ROUNDS; r++) {
		const uint32_t i = base + r * 32;
		if (i < count) {
			k[r] = kFromSrc ? key_of(ks, i) : keys_in[i];
			v[r] = kFromSrc ? i : vals_in[i];
			atomicAdd(&wcnt[warp][(k[r] >> shift) & 255u], 1u);
		}
	}
	__syncthreads();
	{ // thread d: bases of digit d for the 8 warps, in warp order
		uint32_t run = hist[tid * numTiles + blockIdx.x];
#pragma unroll
		for (int w = 0; w < RS_WARPS; w++) {
			const uint32_t c = wcnt[w][tid];
			wcnt[w][tid] = run;
			run += c;
		}
	}
	__syncthreads();
	const uint32_t lt = (1u << lane) - 1;
#pragma unroll
	for (int r = 0; r < RS_ROUNDS; r++) {
		const uint32_t i = base + r * 32;
		const bool valid = i < count;
		const uint32_t dg = valid ? ((k[r] >> shift) & 255u) : (0x10000u | (uint32_t)lane);
		const uint32_t peers = __match_any_sync(0xffffffffu, dg);
		const uint32_t rank = __popc(peers & lt);
		uint32_t b = 0;
		if (valid)
			b = wcnt[warp][dg];
		__syncwarp();
		if (valid && rank == 0)
			wcnt[warp][dg] = b + __popc(peers);
		__syncwarp();
		if (valid) {
			keys_out[b + rank] = k[r];
			vals_out[b + rank] = v[r];
		}
	}
}

// c[V[s]] = previous position (1-based) with the same key, 0 if none
__global__ void __launch_bounds__(256) mf_prev_kernel(const uint32_t *__restrict__ K, const uint32_t *__restrict__ V,
						       uint32_t count, uint32_t *__restrict__ c)
{
	const uint32_t s = blockIdx.x * 256u + threadIdx.x;
	if (s >= count)
		return;
	c[V[s]] = (s > 0 && K[s - 1] == K[s]) ? V[s - 1] + 1 : 0;
}

// The sorted positions go to the block's `sorted` array, and every position that has a predecessor in its bucket gets
// the length of its common prefix with that predecessor -- the first comparison of its tree insertion -- into its
// (still unused) record: one thread per position instead of the bucket's thread (lzma_mf.cuh, kMfFirstValid).
__global__ void __launch_bounds__(256) mf_first_kernel(const uint32_t *__restrict__ K, const uint32_t *__restrict__ V, uint32_t count,
							const uint8_t *__restrict__ src, MfParams P, uint32_t *__restrict__ sorted,
							uint64_t *__restrict__ rec)
{
	const uint32_t s = blockIdx.x * 256u + threadIdx.x;
	if (s >= count)
		return;
	const uint32_t i = V[s];
	sorted[s] = i;
	if (s == 0)
		return;
	const uint32_t pi = V[s - 1], pos = i + 1, curMatch = pi + 1;
	// the same bucket test as the walk's (mf_walk_kernel)
	if (mf_hash4(c_crc, src + pi, P.hashMask, P.bigHash) != mf_hash4(c_crc, src + i, P.hashMask, P.bigHash))
		return;
	const uint32_t cmCheck = pos <= P.cyclicSize ? 0 : pos - P.cyclicSize;
	if (cmCheck >= curMatch)
		return; // outside the dictionary: the insertion does not look at it
	const uint32_t avail = P.n - i, lenLimit = avail < P.fb ? avail : P.fb;
	const uint8_t *a = src + pi, *b = src + i;
	uint32_t len = 0;
	while (len != lenLimit && a[len] == b[len])
		len++;
	rec[i] = kMfFirstValid | len;
}

__global__ void __launch_bounds__(256) mf_copy_kernel(const uint32_t *__restrict__ a, uint32_t count, uint32_t *__restrict__ b)
{
	const uint32_t s = blockIdx.x * 256u + threadIdx.x;
	if (s < count)
		b[s] = a[s];
}

// One thread per sorted element of the wave; the thread of a bucket's first element owns the bucket.
__global__ void __launch_bounds__(128) mf_walk_kernel(const MfBlock *__restrict__ blocks, int nblocks,
						       const uint64_t *__restrict__ segBase)
{
	const uint64_t g = (uint64_t)blockIdx.x * 128u + threadIdx.x;
	int lo = 0, hi = nblocks; // segBase[lo] <= g < segBase[lo + 1]
	if (g >= segBase[nblocks])
		return;
	while (hi - lo > 1) {
		const int mid = (lo + hi) >> 1;
		if (segBase[mid] <= g)
			lo = mid;
		else
			hi = mid;
	}
	const MfBlock &B = blocks[lo];
	uint32_t s = (uint32_t)(g - segBase[lo]);
	const uint32_t count = B.count;
	const uint32_t *__restrict__ V = B.sorted;
	const uint8_t *src = B.src;
	uint32_t i = V[s];
	const uint32_t hv = mf_hash4(c_crc, src + i, B.P.hashMask, B.P.bigHash);
	if (s > 0 && mf_hash4(c_crc, src + V[s - 1], B.P.hashMask, B.P.bigHash) == hv)
		return; // not the first of its bucket
	const MfParams P = B.P;
	uint32_t d[2 * kMfMaxFb + 6];
	uint32_t prev = 0;
	// A bucket of thousands of positions (a run of identical bytes is ONE bucket as long as the run) is a serial chain on
	// this thread, so what every link of the chain would wait for is taken out of it once the bucket proves long:
	//  * a full-length hit on the first candidate (known from the pre-pass) is GetMatchesSpec1's early exit -- record the
	//    pair, take over the candidate's children -- and the candidate is the node this thread wrote one step ago: its
	//    children are kept in registers instead of read back through L2;
	//  * pool space is taken 512 words at a time instead of one atomic round trip per position;
	//  * records are published 16 at a time behind one fence.
	constexpr uint32_t kLongBucket = 256, kChunk = 512, kBatch = 16;
	uint32_t steps = 0, cpos = 0, cp0 = 0, cp1 = 0;
	uint64_t chunkOff = 0;
	uint32_t chunkLeft = 0, nb = 0;
	uint32_t bi[kBatch];
	uint64_t bv[kBatch];
	for (;;) {
		const uint32_t pos = i + 1;
		const uint32_t r0 = (uint32_t)B.rec[i]; // the pre-pass's note about the first candidate (mf_first_kernel)
		const uint32_t avail = P.n - i, lenLimit = avail < P.fb ? avail : P.fb;
		uint32_t nbt;
		if ((r0 & kMfFirstValid) && (r0 & 0xFFFFu) == lenLimit && P.mc) {
			uint32_t p0, p1;
			if (prev == cpos) {
				p0 = cp0;
				p1 = cp1;
			} else {
				p0 = B.son[(size_t)prev << 1];
				p1 = B.son[((size_t)prev << 1) + 1];
			}
			B.son[(size_t)pos << 1] = p0;
			B.son[((size_t)pos << 1) + 1] = p1;
			d[4] = lenLimit;
			d[5] = pos - prev - 1;
			nbt = 2;
			cpos = pos;
			cp0 = p0;
			cp1 = p1;
		} else {
			nbt = mf_bt_insert(src, P, B.son, pos, prev, d + 4, (r0 & kMfFirstValid) ? (r0 & 0xFFFFu) : kMfFirstNone);
			cpos = 0;
		}
		const uint32_t nd = mf_mix(src, P, pos, B.c2[i], B.c3[i], d, nbt);
		uint64_t off = 0;
		if (nd) {
			if (steps < kLongBucket)
				off = atomicAdd(B.cursor, (unsigned long long)nd);
			else {
				if (chunkLeft < nd) {
					const uint32_t grab = nd > kChunk ? nd : kChunk;
					chunkOff = atomicAdd(B.cursor, (unsigned long long)grab);
					chunkLeft = grab;
				}
				off = chunkOff;
				chunkOff += nd;
				chunkLeft -= nd;
			}
			if (off + nd <= B.poolCap) {
				uint32_t *w = B.pool + off;
				for (uint32_t j = 0; j < nd; j++)
					w[j] = d[j];
			} else
				*B.overflow = 1;
		}
		// the block's parser may already be running and waiting for this very record (backend.cu): the list (or the
		// overflow flag) must be visible before the record that announces it
		const uint64_t rv = kMfReady | (off << kMfCountBits) | nd;
		bool last = false;
		prev = pos;
		steps++;
		if (++s == count)
			last = true;
		else {
			i = V[s];
			if (mf_hash4(c_crc, src + i, P.hashMask, P.bigHash) != hv)
				last = true;
		}
		if (steps <= kLongBucket) {
			__threadfence();
			*(volatile uint64_t *)&B.rec[pos - 1] = rv;
		} else {
			bi[nb] = pos - 1;
			bv[nb] = rv;
			nb++;
		}
		if (nb == kBatch || (last && nb)) {
			__threadfence();
			for (uint32_t j = 0; j < nb; j++)
				*(volatile uint64_t *)&B.rec[bi[j]] = bv[j];
			nb = 0;
		}
		if (last)
			break;
	}
}

// Levels 1-4 (hash chains): one thread per position of the wave; nothing is carried between positions.
// B.sorted holds c5[i] = previous position (1-based) with the same 5-byte hash, i.e. the reference's son[].
__global__ void __launch_bounds__(128) mf_hc_kernel(const MfBlock *__restrict__ blocks, int nblocks,
						     const uint64_t *__restrict__ segBase)
{
	const uint64_t g = (uint64_t)blockIdx.x * 128u + threadIdx.x;
	int lo = 0, hi = nblocks;
	if (g >= segBase[nblocks])
		return;
	while (hi - lo > 1) {
		const int mid = (lo + hi) >> 1;
		if (segBase[mid] <= g)
			lo = mid;
		else
			hi = mid;
	}
	const MfBlock &B = blocks[lo];
	const uint32_t i = (uint32_t)(g - segBase[lo]);
	const MfParams P = B.P;
	uint32_t d[2 * kMfMaxFb + 6];
	const uint32_t nd = mf_hc5_matches(B.src, P, B.sorted - 1, i + 1, B.c2[i], B.c3[i], d);
	uint64_t off = 0;
	if (nd) {
		off = atomicAdd(B.cursor, (unsigned long long)nd);
		if (off + nd <= B.poolCap) {
			uint32_t *w = B.pool + off;
			for (uint32_t j = 0; j < nd; j++)
				w[j] = d[j];
		} else
			*B.overflow = 1;
	}
	B.rec[i] = (off << kMfCountBits) | nd;
}

bool g_init = false;

} // namespace

int mf_init_tables()
{
	uint32_t crc[256];
	for (uint32_t i = 0; i < 256; i++)
		crc[i] = mf_crc_entry(i);
	g_init = cudaMemcpyToSymbol(c_crc, crc, sizeof(crc)) == cudaSuccess;
	return g_init ? 0 : -1;
}

size_t mf_sort_scratch_bytes(uint32_t maxCount)
{
	const size_t tiles = ((size_t)maxCount + RS_TILE - 1) / RS_TILE;
	return 4 * (size_t)maxCount * 4 + 256 * tiles * 4 + 1024;
}

// Sorts positions [0, count) of the block by the chosen hash; leaves keys/values in (*K, *V).
static int sort_by(const MfBlock &B, int kind, int bits, uint32_t *bufs[4], uint32_t *hist, uint32_t **K, uint32_t **V,
		   cudaStream_t st, int64_t *launches)
{
	const uint32_t count = B.count;
	const uint32_t tiles = (count + RS_TILE - 1) / RS_TILE;
	KeySrc ks = { B.src, B.P.hashMask, B.P.bigHash, kind };
	uint32_t *ki = nullptr, *vi = nullptr, *ko = bufs[0], *vo = bufs[1];
	for (int shift = 0, pass = 0; shift < bits; shift += 8, pass++) {
		if (pass == 0) {
			LRZ_LAUNCH(tiles, RS_THREADS, 0, st, rs_hist_kernel<true>, nullptr, ks, count, shift, hist, tiles);
			LRZ_LAUNCH(1, 1024, 0, st, rs_scan_kernel, hist, 256 * tiles);
			LRZ_LAUNCH(tiles, RS_THREADS, 0, st, rs_scatter_kernel<true>, nullptr, nullptr, ks, count, shift, hist, tiles, ko, vo);
		} else {
			LRZ_LAUNCH(tiles, RS_THREADS, 0, st, rs_hist_kernel<false>, ki, ks, count, shift, hist, tiles);
			LRZ_LAUNCH(1, 1024, 0, st, rs_scan_kernel, hist, 256 * tiles);
			LRZ_LAUNCH(tiles, RS_THREADS, 0, st, rs_scatter_kernel<false>, ki, vi, ks, count, shift, hist, tiles, ko, vo);
		}
		if (launches)
			*launches += 3;
		ki = ko;
		vi = vo;
		ko = (ki == bufs[0]) ? bufs[2] : bufs[0];
		vo = (vi == bufs[1]) ? bufs[3] : bufs[1];
	}
	*K = ki;
	*V = vi;
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mf_prepare_block(const MfBlock &B, void *scratch, cudaStream_t st, int64_t *launches)
{
	if (!g_init || B.count == 0)
		return B.count == 0 ? 0 : -1;
	const uint32_t count = B.count;
	uint32_t *bufs[4];
	for (int i = 0; i < 4; i++)
		bufs[i] = (uint32_t *)scratch + (size_t)i * count;
	uint32_t *hist = (uint32_t *)scratch + 4 * (size_t)count;
	const uint32_t grid = (count + 255) / 256;
	uint32_t *K, *V;
	if (sort_by(B, KEY_H2, 10, bufs, hist, &K, &V, st, launches))
		return -1;
	LRZ_LAUNCH(grid, 256, 0, st, mf_prev_kernel, K, V, count, B.c2);
	if (sort_by(B, KEY_H3, 16, bufs, hist, &K, &V, st, launches))
		return -1;
	LRZ_LAUNCH(grid, 256, 0, st, mf_prev_kernel, K, V, count, B.c3);
	int bits = 0;
	while (bits < 32 && (B.P.hashMask >> bits))
		bits++;
	if (B.P.hc5) { // hash chains: the predecessor in the 5-byte-hash order IS the chain link (son[])
		if (sort_by(B, KEY_H5, bits, bufs, hist, &K, &V, st, launches))
			return -1;
		LRZ_LAUNCH(grid, 256, 0, st, mf_prev_kernel, K, V, count, B.sorted);
		if (launches)
			*launches += 3;
		return cudaGetLastError() == cudaSuccess ? 0 : -1;
	}
	if (sort_by(B, KEY_H4, bits, bufs, hist, &K, &V, st, launches))
		return -1;
	LRZ_LAUNCH(grid, 256, 0, st, mf_first_kernel, K, V, count, B.src, B.P, B.sorted, B.rec);
	if (launches)
		*launches += 3;
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mf_walk_launch(const MfBlock *d_blocks, int nblocks, const uint64_t *d_segBase, uint64_t total, bool hc5,
		   cudaStream_t st, int64_t *launches)
{
	if (total == 0)
		return 0;
	const uint64_t grid = (total + 127) / 128;
	if (grid > 0x7fffffffull)
		return -1;
	if (hc5)
		LRZ_LAUNCH((unsigned)grid, 128, 0, st, mf_hc_kernel, d_blocks, nblocks, d_segBase);
	else
		LRZ_LAUNCH((unsigned)grid, 128, 0, st, mf_walk_kernel, d_blocks, nblocks, d_segBase);
	if (launches)
		*launches += 1;
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Load this file's kernels now (CUDA loads a kernel's code at its first launch, and that load waits for every kernel
// that is running -- block encoders run for tens of seconds).
#if !defined(LRZ_SIMT_HOST)
int mf_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, mf_copy_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, mf_first_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, mf_hc_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, mf_prev_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, mf_walk_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, rs_hist_kernel<true>) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, rs_hist_kernel<false>) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, rs_scan_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, rs_scatter_kernel<true>) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, rs_scatter_kernel<false>) == cudaSuccess;
	return ok ? 0 : -1;
}
#endif

} // namespace lzma
} // namespace lrz
