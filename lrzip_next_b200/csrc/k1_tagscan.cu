// k1_tagscan.cu -- K1: rzip rolling-tag scan + candidate compaction (sm_100a).
//
// Replaces, for a whole window segment at once, the per-position work of the reference's hash_search
// loop head: next_tag / full_tag (src/rzip.c:385-416) and the tag-mask gate (src/rzip.c:658).
//
//   tag(p) = XOR_{i<31} hash_index[buf[p+i]]                      (src/rzip.c:405-416)
//
// Every position p in [1, n-31] whose tag satisfies (tag & mask) == mask becomes one 16-byte
// candidate record {pos, tag} (layout of struct hash_entry, src/rzip.c:61-64).  Records of tile
// T (positions [T*512, (T+1)*512)) are written, in position order, to the tile-strided region
// cand[T*512 ...] and their number to tile_count[T]; no cross-CTA (or cross-warp) dependency exists.
//
// HBM traffic (the algorithmic bytes of SURVEY.md 8(d)): 1 byte read per position + 16 bytes
// written per candidate = 1 + 16 * 2^-initial_freq bytes per input byte (9 B/B at rzip level 7).
//
// Structure (persistent CTAs, 256 threads = 8 warps, one 4096-byte step per iteration; a candidate
// tile is 512 positions = one warp's share of the step, so tiles never need a cross-warp prefix):
//   * the step's bytes (+32 B halo) are staged global -> shared by the TMA engine (cp.async.bulk with
//     an mbarrier transaction count), double buffered, so the load of step i+1 overlaps step i;
//   * phase A: thread t owns the 16 bytes of slot t, XOR-accumulates hash_index over them through a
//     16-way replicated 64-bit table in shared memory (lanes l and l+16 are in different 64-bit
//     phases, so lookups are bank-conflict free); a shuffle XOR-scan inside the warp turns the local
//     prefixes into the WARP-relative running XOR Z[i] = XOR_{512w <= b < i} hash_index[byte b],
//     stored with one pad word per 16 entries so that both phases are conflict free;
//   * phase B, fused with the output: lane l of warp w takes position 512w + 32j + l, so that the
//     survivors of one ballot are consecutive positions and their 16-byte records form one contiguous
//     run in HBM:   tag(q) = Z[q + 31] ^ Z[q]  (^ the warp's total when q + 31 is in the next warp's
//     share); ballot, 16 B store per surviving lane at the warp's running count.
#include "kernels.h"

#if !defined(LRZ_SIMT_HOST)
#include <cuda_runtime.h>
#endif

namespace lrz {

__constant__ int64_t c_hash_index[256];

static constexpr int K1_THREADS = 256;
static constexpr int K1_WARPS = K1_THREADS / 32;
static constexpr int K1_STEP = K1_WARPS * kTile;     // 4096 bytes per CTA iteration
static constexpr int K1_IN_BYTES = K1_STEP + 32;     // step + halo, multiple of 16
static constexpr int K1_Z = K1_STEP + 32 + 1;        // running XORs Z[0..4128]
static constexpr int K1_ZPAD = K1_Z + K1_Z / 16 + 2; // one pad word per 16 entries: conflict-free columns
static constexpr int K1_SMEM_TABLE = 256 * 16 * 8;
static constexpr int K1_SMEM_Z = ((K1_ZPAD * 8 + 127) / 128) * 128;
static constexpr int K1_SMEM_IN = 2 * K1_IN_BYTES;
static constexpr int K1_OFF_Z = K1_SMEM_TABLE;
static constexpr int K1_OFF_IN = K1_OFF_Z + K1_SMEM_Z;
static constexpr int K1_OFF_BAR = ((K1_OFF_IN + K1_SMEM_IN + 15) / 16) * 16;
static constexpr int K1_SMEM = K1_OFF_BAR + 32;
static_assert(kTile == 512, "one candidate tile per warp and step");

#if defined(LRZ_SIMT_HOST)
// tests/hostsim (simt.h): an mbarrier is a phase bit, a bulk copy completes the moment it is issued
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	while ((*(volatile uint64_t *)bar & 1) == parity)
		simt::yield();
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	memcpy(dst, src, bytes);
	*bar ^= 1;
}
__device__ __forceinline__ void st_cand(Cand *dst, uint32_t plo, uint32_t phi, uint32_t tlo, uint32_t thi)
{
	dst->pos = (int64_t)(((uint64_t)phi << 32) | plo);
	dst->tag = (int64_t)(((uint64_t)thi << 32) | tlo);
}
#define K1_FENCE_MBAR_INIT() ((void)0)
#else
#define K1_FENCE_MBAR_INIT() asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory")
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		     "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted on the mbarrier.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void st_cand(Cand *dst, uint32_t plo, uint32_t phi, uint32_t tlo, uint32_t thi)
{
	asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(plo), "r"(phi), "r"(tlo), "r"(thi) : "memory");
}
#endif

__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d)
{
	const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
	const uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
	return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ int zidx(int j) { return j + (j >> 4); }

// Phase B for one warp: tags of positions base + 32 j + lane (base = start of the warp's tile), survivors
// stored in position order at dst; returns their number.  z points at the padded Z entry of the tile start.
template <bool kInterior>
__device__ __forceinline__ uint32_t k1_phase_b(const uint64_t *__restrict__ z, int zq0, int lane, int64_t base, int64_t lo,
					       int64_t hi, uint32_t mlo, uint32_t mhi, Cand *__restrict__ dst)
{
	const uint64_t *za = z + zidx(zq0 + lane), *zb = z + zidx(zq0 + lane + 31);
	const uint64_t wtot = z[zidx(zq0 + kTile)]; // XOR over the warp's whole share
	const int l0 = (int)(lo - base), h0 = (int)(hi - base);
	const uint32_t lt = (1u << lane) - 1;
	const uint32_t pos_hi = (uint32_t)((uint64_t)base >> 32), pos_lo = (uint32_t)base + (uint32_t)lane;
	uint32_t cnt = 0;
#pragma unroll
	for (int j = 0; j < 16; j++) {
		// 32 positions further = 34 padded words further
		uint64_t a = za[34 * j];
		const uint64_t b = zb[34 * j];
		if (j == 0 && lane == 0)
			a = 0; // Z at the tile start holds the previous warp's total
		uint64_t tg = a ^ b;
		if (j == 15 && lane >= 2)
			tg ^= wtot; // q + 31 lies in the next warp's share, whose Z restarts at 0
		const uint32_t tl = (uint32_t)tg, th = (uint32_t)(tg >> 32);
		bool ok = (tl & mlo) == mlo && (th & mhi) == mhi;
		if (!kInterior) {
			const int q = 32 * j + lane;
			ok = ok && q >= l0 && q < h0;
		}
		const uint32_t bm = __ballot_sync(0xffffffffu, ok);
		if (ok)
			st_cand(dst + cnt + __popc(bm & lt), pos_lo + 32 * j, pos_hi, tl, th);
		cnt += __popc(bm);
	}
	return cnt;
}

__global__ void __launch_bounds__(K1_THREADS, 3)
k1_tagscan_kernel(const uint8_t *__restrict__ buf, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask_arg,
		  const ScanState *__restrict__ state, int nstates, Cand *__restrict__ cand, uint32_t *__restrict__ tile_count,
		  int64_t first_tile, int64_t num_tiles, int64_t first_step, int64_t num_steps)
{
#if defined(LRZ_SIMT_HOST)
	uint8_t *smem = simt::dyn_smem();
#else
	extern __shared__ __align__(128) uint8_t smem[];
#endif
	uint64_t *tab = reinterpret_cast<uint64_t *>(smem);                // tab[b * 16 + (lane & 15)]
	uint64_t *z = reinterpret_cast<uint64_t *>(smem + K1_OFF_Z);       // padded warp-relative running XORs
	uint8_t *in = smem + K1_OFF_IN;
	uint64_t *bar = reinterpret_cast<uint64_t *>(smem + K1_OFF_BAR);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	int64_t mask = mask_arg;
	if (state) {
		// several scan states = the variants of one window run for every incoming victim_round value
		// (all-values speculation): the loosest gate of them decides what is a candidate (gates are nested
		// 2^b - 1 masks, so AND gives the loosest), each commit filters by its own gate again
		mask = state[0].min_mask;
		int64_t least = state[0].scan_pos;
		for (int v = 1; v < nstates; v++) {
			mask &= state[v].min_mask;
			if (state[v].scan_pos < least)
				least = state[v].scan_pos;
		}
		if (least + 1 >= pos_hi) { // every scan already jumped past this segment
			for (int64_t t = blockIdx.x * (int64_t)K1_THREADS + tid; t < num_tiles; t += (int64_t)gridDim.x * K1_THREADS)
				tile_count[t] = 0;
			return;
		}
	}
	const uint32_t mlo = (uint32_t)mask, mhi = (uint32_t)((uint64_t)mask >> 32);

	for (int i = tid; i < 256 * 16; i += K1_THREADS)
		tab[i] = (uint64_t)c_hash_index[i >> 4];
	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		K1_FENCE_MBAR_INIT();
	}
	__syncthreads();

	// valid positions: [max(pos_lo, 1), min(pos_hi, n - 31 + 1))   (src/rzip.c:622, 631-634)
	const int64_t vlo = pos_lo > 1 ? pos_lo : 1;
	const int64_t vhi = (pos_hi < n - kMinMatch + 1) ? pos_hi : n - kMinMatch + 1;
	const uint64_t *my_tab = tab + (lane & 15);

	int64_t s = blockIdx.x;
	if (s < num_steps && tid == 0) {
		mbar_expect_tx(&bar[0], K1_IN_BYTES);
		tma_load_1d(in, buf + (first_step + s) * (int64_t)K1_STEP, K1_IN_BYTES, &bar[0]);
	}
	for (int it = 0; s < num_steps; s += gridDim.x, it++) {
		const int stage = it & 1;
		const unsigned parity = (it >> 1) & 1;
		const int64_t step = first_step + s;
		if (tid == 0 && s + gridDim.x < num_steps) { // prefetch the next step into the other stage
			mbar_expect_tx(&bar[stage ^ 1], K1_IN_BYTES);
			tma_load_1d(in + (stage ^ 1) * K1_IN_BYTES, buf + (step + gridDim.x) * (int64_t)K1_STEP, K1_IN_BYTES,
				    &bar[stage ^ 1]);
		}
		mbar_wait(&bar[stage], parity);

		// ---- phase A: warp-relative running XOR of hash_index over the step (+ halo)
		{
			uint64_t x[16];
			const uint4 v = *reinterpret_cast<const uint4 *>(in + stage * K1_IN_BYTES + tid * 16);
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
			uint64_t acc = 0;
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const uint32_t b = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
				acc ^= my_tab[b * 16];
				x[j] = acc;
			}
			uint64_t incl = acc; // inclusive XOR scan of the slot totals inside the warp
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint64_t o = shfl_up64(incl, d);
				if (lane >= d)
					incl ^= o;
			}
			const uint64_t off = incl ^ acc; // exclusive
			uint64_t *dst = z + 1 + tid * 16;
			const int pad = (tid * 16 + 1) >> 4; // = tid
#pragma unroll
			for (int j = 0; j < 16; j++)
				dst[pad + (j == 15 ? 1 : 0) + j] = off ^ x[j];
		}
		if (warp == K1_WARPS - 1) { // halo: one byte per lane, XOR scan, Z relative to the end of the step
			const uint32_t b = in[stage * K1_IN_BYTES + K1_STEP + lane];
			uint64_t incl = my_tab[b * 16];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint64_t o = shfl_up64(incl, d);
				if (lane >= d)
					incl ^= o;
			}
			z[zidx(K1_STEP + 1 + lane)] = incl;
		}
		__syncthreads();

		// ---- phase B: warp w owns candidate tile step * 8 + w
		const int64_t t = step * K1_WARPS + warp - first_tile;
		if (t >= 0 && t < num_tiles) {
			const int64_t base = (first_tile + t) * (int64_t)kTile;
			Cand *dst = cand + t * (int64_t)kTile;
			uint32_t cnt;
			if (base >= vlo && base + kTile <= vhi)
				cnt = k1_phase_b<true>(z, warp * kTile, lane, base, vlo, vhi, mlo, mhi, dst);
			else
				cnt = k1_phase_b<false>(z, warp * kTile, lane, base, vlo, vhi, mlo, mhi, dst);
			if (lane == 0)
				tile_count[t] = cnt;
		}
		__syncthreads(); // every read of z / in[stage] is done before they are overwritten
	}
}

int k1_init_tables()
{
	int64_t hi[256];
	make_hash_index(hi);
#if defined(LRZ_SIMT_HOST)
	memcpy(c_hash_index, hi, sizeof(hi));
	return 0;
#else
	cudaError_t e = cudaMemcpyToSymbol(c_hash_index, hi, sizeof(hi));
	if (e != cudaSuccess)
		return -1;
	e = cudaFuncSetAttribute(k1_tagscan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM);
	return e == cudaSuccess ? 0 : -1;
#endif
}

#if !defined(LRZ_SIMT_HOST) // launchers: the emulator (tests/hostsim) calls the kernel directly

int k1_launch(const uint8_t *d_buf, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask,
	      const ScanState *d_state, int nstates, Cand *d_cand, uint32_t *d_tile_count, int num_sms, cudaStream_t stream)
{
	if (pos_hi <= pos_lo)
		return 0;
	const int64_t first_tile = pos_lo / kTile;
	const int64_t last_tile = (pos_hi - 1) / kTile;
	const int64_t num_tiles = last_tile - first_tile + 1;
	const int64_t first_step = pos_lo / K1_STEP;
	const int64_t num_steps = (pos_hi - 1) / K1_STEP - first_step + 1;
	int64_t grid = (int64_t)num_sms * 3;
	if (grid > num_steps)
		grid = num_steps;
	k1_tagscan_kernel<<<(unsigned)grid, K1_THREADS, K1_SMEM, stream>>>(
		d_buf, n, pos_lo, pos_hi, mask, d_state, nstates, d_cand, d_tile_count, first_tile, num_tiles, first_step, num_steps);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Load this file's kernels now (CUDA loads a kernel's code at its first launch, and that load waits for every kernel
// that is running -- block encoders run for tens of seconds).
int k1_preload()
{
	cudaFuncAttributes a;
	bool ok = true;
	ok = ok && cudaFuncGetAttributes(&a, k1_tagscan_kernel) == cudaSuccess;
	return ok ? 0 : -1;
}
#endif // !LRZ_SIMT_HOST

} // namespace lrz
