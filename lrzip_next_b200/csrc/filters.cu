// filters.cu -- device side of the pre-compression filters (filters.cuh; src/stream.c:1587-1628): every stream-1 block
// is converted in place, from position 0 of the block, before the lz4 gate and the backend see it.
#include "filters.cuh"
#include "kernels.h"

namespace lrz {

namespace {

constexpr int kDeltaTile = 32 * 1024; // bytes of a block one CTA converts
constexpr int kDeltaMax = 256;        // largest delta distance (DELTA_STATE_SIZE, src/lzma/include/Delta.h)

// block k = bytes [from + k * bs, min(from + (k + 1) * bs, to)) of the stream
__device__ __forceinline__ int64_t block_len(int64_t k, int64_t from, int64_t to, int64_t bs)
{
	const int64_t lo = from + k * bs, hi = lo + bs < to ? lo + bs : to;
	return hi - lo;
}

// One thread per aligned 32-bit word of a block (the tail of 1-3 bytes is left alone, like the reference's size &= ~3).
__global__ void __launch_bounds__(256) filter_words_kernel(uint8_t *s, int64_t from, int64_t to, int64_t bs, int filter, bool enc)
{
	const int64_t k = blockIdx.y;
	uint8_t *b = s + from + k * bs;
	const int64_t words = block_len(k, from, to, bs) >> 2;
	for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < words; w += (int64_t)gridDim.x * 256) {
		uint8_t *p = b + 4 * w; // a block starts at a multiple of the block size inside a 256-byte aligned buffer
		const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
		const uint32_t c = flt::conv_word(filter, v, (uint32_t)(4 * w), enc);
		if (c != v) {
			p[0] = (uint8_t)c;
			p[1] = (uint8_t)(c >> 8);
			p[2] = (uint8_t)(c >> 16);
			p[3] = (uint8_t)(c >> 24);
		}
	}
}

// x86, ARM Thumb and RISC-V: the scan's state runs through the whole block: one thread per block.
__global__ void filter_serial_kernel(uint8_t *s, int64_t from, int64_t to, int64_t bs, int filter, bool enc)
{
	if (threadIdx.x)
		return;
	const int64_t k = blockIdx.x;
	if (filter == flt::kX86)
		flt::x86_convert(s + from + k * bs, (size_t)block_len(k, from, to, bs), enc);
	else if (filter == flt::kRISCV)
		flt::riscv_convert(s + from + k * bs, (size_t)block_len(k, from, to, bs), enc);
	else
		flt::armt_convert(s + from + k * bs, (size_t)block_len(k, from, to, bs), enc);
}

// Delta, decode side (Delta_Decode): out[i] = in[i] + out[i - delta] is a running sum along each residue class
// modulo delta: one thread per class and block.
__global__ void delta_decode_kernel(uint8_t *s, int64_t from, int64_t to, int64_t bs, int delta)
{
	const int64_t k = blockIdx.x;
	uint8_t *b = s + from + k * bs;
	const int64_t n = block_len(k, from, to, bs);
	const int r = threadIdx.x;
	if (r >= delta)
		return;
	uint8_t acc = r < n ? b[r] : 0;
	for (int64_t i = (int64_t)r + delta; i < n; i += delta) {
		acc = (uint8_t)(acc + b[i]);
		b[i] = acc;
	}
}

// IA64: one thread per 16-byte bundle.
__global__ void __launch_bounds__(256) filter_ia64_kernel(uint8_t *s, int64_t from, int64_t to, int64_t bs, bool enc)
{
	const int64_t k = blockIdx.y;
	uint8_t *b = s + from + k * bs;
	const int64_t bundles = block_len(k, from, to, bs) >> 4;
	for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < bundles; w += (int64_t)gridDim.x * 256)
		flt::ia64_bundle(b + 16 * w, (uint32_t)(16 * w), enc);
}

// Delta in place, two passes so that no CTA reads bytes another CTA has already replaced: first every tile's `delta`
// predecessor bytes are saved, then every tile is converted from its own (still original) bytes and the saved ones.
__global__ void __launch_bounds__(256) delta_save_kernel(const uint8_t *s, int64_t from, int64_t to, int64_t bs, int delta,
							 int tiles_per_block, uint8_t *side)
{
	const int64_t k = blockIdx.y, t = blockIdx.x;
	const int64_t len = block_len(k, from, to, bs), t0 = t * kDeltaTile;
	if (t0 >= len || (int)threadIdx.x >= delta)
		return;
	const uint8_t *b = s + from + k * bs;
	const int64_t src = t0 - delta + threadIdx.x; // before the block: zero (Delta_Init)
	side[((size_t)k * tiles_per_block + t) * kDeltaMax + threadIdx.x] = src >= 0 ? b[src] : 0;
}

__global__ void __launch_bounds__(256) delta_apply_kernel(uint8_t *s, int64_t from, int64_t to, int64_t bs, int delta,
							  int tiles_per_block, const uint8_t *side)
{
#if defined(LRZ_SIMT_HOST)
	uint8_t *tile = simt::dyn_smem();
#else
	extern __shared__ uint8_t tile[]; // delta saved bytes, then the tile
#endif
	const int64_t k = blockIdx.y, t = blockIdx.x;
	const int64_t len = block_len(k, from, to, bs), t0 = t * kDeltaTile;
	if (t0 >= len)
		return;
	uint8_t *b = s + from + k * bs + t0;
	const int n = (int)(len - t0 < kDeltaTile ? len - t0 : kDeltaTile);
	const uint8_t *sv = side + ((size_t)k * tiles_per_block + t) * kDeltaMax;
	for (int i = threadIdx.x; i < delta; i += 256)
		tile[i] = sv[i];
	for (int i = threadIdx.x; i < n; i += 256)
		tile[delta + i] = b[i];
	__syncthreads();
	for (int i = threadIdx.x; i < n; i += 256)
		b[i] = (uint8_t)(tile[delta + i] - tile[i]);
}

} // namespace

size_t filter_side_bytes(int filter, int64_t span, int64_t bs)
{
	if (filter != flt::kDelta || span <= 0)
		return 0;
	const int64_t nblk = (span + bs - 1) / bs, tpb = (bs + kDeltaTile - 1) / kDeltaTile;
	return (size_t)nblk * (size_t)tpb * kDeltaMax;
}

// Convert the stream blocks that make up s[from, to) in place (from is a block boundary).  side: filter_side_bytes().
int filter_blocks_launch(int filter, int delta, uint8_t *s, int64_t from, int64_t to, int64_t bs, uint8_t *side,
			 cudaStream_t stream, int64_t *launches, bool enc)
{
	if (filter == flt::kNone || to <= from)
		return 0;
	if (!flt::supported(filter) || bs <= 0)
		return -1;
	const int64_t nblk = (to - from + bs - 1) / bs;
	if (nblk > 65535)
		return -1;
	if (flt::wordwise(filter)) {
		const int64_t words = (bs < to - from ? bs : to - from) >> 2;
		unsigned gx = (unsigned)((words + 255) / 256);
		if (gx > 1184)
			gx = 1184; // 8 CTAs on each of 148 SMs, grid-stride
		if (gx == 0)
			gx = 1;
		LRZ_LAUNCH(dim3(gx, (unsigned)nblk), 256, 0, stream, filter_words_kernel, s, from, to, bs, filter, enc);
		if (launches)
			*launches += 1;
	} else if (flt::serial(filter)) {
		LRZ_LAUNCH((unsigned)nblk, 32, 0, stream, filter_serial_kernel, s, from, to, bs, filter, enc);
		if (launches)
			*launches += 1;
	} else if (filter == flt::kIA64) {
		const int64_t bundles = (bs < to - from ? bs : to - from) >> 4;
		unsigned gx = (unsigned)((bundles + 255) / 256);
		gx = gx > 1184 ? 1184 : (gx ? gx : 1);
		LRZ_LAUNCH(dim3(gx, (unsigned)nblk), 256, 0, stream, filter_ia64_kernel, s, from, to, bs, enc);
		if (launches)
			*launches += 1;
	} else if (!enc) { // delta, decode side
		if (delta < 1 || delta > kDeltaMax)
			return -1;
		LRZ_LAUNCH((unsigned)nblk, 256, 0, stream, delta_decode_kernel, s, from, to, bs, delta);
		if (launches)
			*launches += 1;
	} else { // delta
		if (delta < 1 || delta > kDeltaMax || !side)
			return -1;
		const int tpb = (int)((bs + kDeltaTile - 1) / kDeltaTile);
		LRZ_LAUNCH(dim3((unsigned)tpb, (unsigned)nblk), 256, 0, stream, delta_save_kernel, s, from, to, bs, delta, tpb, side);
		LRZ_LAUNCH(dim3((unsigned)tpb, (unsigned)nblk), 256, kDeltaTile + kDeltaMax, stream, delta_apply_kernel, s, from, to, bs, delta, tpb, side);
		if (launches)
			*launches += 2;
	}
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#if !defined(LRZ_SIMT_HOST)
int filter_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, filter_words_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, filter_serial_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, filter_ia64_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, delta_save_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, delta_decode_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, delta_apply_kernel) == cudaSuccess;
	return ok ? 0 : -1;
}

#endif

} // namespace lrz
