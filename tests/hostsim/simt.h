// tests/hostsim/simt.h -- TEST INFRASTRUCTURE ONLY.
// A small SIMT emulator: one CUDA thread block run on the CPU, every CUDA thread a fiber (its own stack, switched
// by hand), so that a kernel written with warp collectives (__ballot_sync, __shfl_sync, __any_sync, __syncwarp)
// and block / named barriers compiles as plain C++ and executes with the same meaning.  A fiber runs until it
// reaches a collective, deposits its operand and yields; the last lane to arrive completes the collective and every
// lane picks its result up when it is scheduled again.  One OS thread, cooperative scheduling: memory is
// sequentially consistent, and a collective that not every lane of a warp reaches is reported as a deadlock.
//
// Restrictions (all met by the kernels built with it): full-warp masks only, block size a multiple of 32, all lanes
// of a warp leave the kernel together, polling loops must call __nanosleep / simt::yield.  Lanes run one after the other
// between collectives, so a kernel whose lanes execute REPLICATED scalar code on shared state in lockstep (every lane
// doing `state->x++` and relying on the warp doing it once: K7b, lzma_enc.cuh) cannot be emulated -- tried, and that is
// why the LZMA parser's device-only paths stay covered by the GPU tests alone.
// The product never includes this header: a kernel source opts in with `#if defined(LRZ_SIMT_HOST)`.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <vector>

#if !defined(__x86_64__)
#error "simt.h switches fibers with x86-64 assembly"
#endif

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))

#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct longlong2 {
	long long x, y;
};
struct uint4 {
	unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w)
{
	uint4 v;
	v.x = x;
	v.y = y;
	v.z = z;
	v.w = w;
	return v;
}
static inline longlong2 make_longlong2(long long x, long long y)
{
	longlong2 v;
	v.x = x;
	v.y = y;
	return v;
}

// void simt_switch(void **save_sp, void *load_sp): callee-saved registers on the old stack, then the new one's
extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "	pushq %rbp\n	pushq %rbx\n	pushq %r12\n	pushq %r13\n	pushq %r14\n	pushq %r15\n"
    "	movq %rsp, (%rdi)\n"
    "	movq %rsi, %rsp\n"
    "	popq %r15\n	popq %r14\n	popq %r13\n	popq %r12\n	popq %rbx\n	popq %rbp\n"
    "	ret\n"
    ".size simt_switch,.-simt_switch\n");

namespace simt {

struct Dim {
	unsigned x, y, z;
};

struct Fiber {
	void *sp = nullptr;
	char *stack = nullptr;
	bool done = false;
	const volatile unsigned *wait_on = nullptr; // blocked while *wait_on == wait_val
	unsigned wait_val = 0;
};

struct WarpSync {
	int arrived = 0;
	unsigned gen = 0;
	uint64_t in[32];
	uint64_t out[2][32];
};

struct BarSync {
	int arrived = 0;
	unsigned gen = 0;
};

struct Block {
	int nthreads = 0, cur = 0;
	Dim tidx{ 0, 0, 0 }, bidx{ 0, 0, 0 }, bdim{ 1, 1, 1 }, gdim{ 1, 1, 1 };
	void *sched_sp = nullptr;
	std::vector<Fiber> fibers;
	std::vector<WarpSync> warps;
	BarSync bars[16];
	std::function<void()> entry;
	unsigned long long collectives = 0;
	std::vector<uint8_t> dyn_smem; // `extern __shared__` of the kernel
};

static Block *B = nullptr;

static inline void block_on(const volatile unsigned *p, unsigned v)
{
	Fiber &f = B->fibers[(size_t)B->cur];
	f.wait_on = p;
	f.wait_val = v;
	simt_switch(&f.sp, B->sched_sp);
	f.wait_on = nullptr;
}

// gives the other threads a turn (a polling loop must call it: scheduling is cooperative)
static inline void yield()
{
	Fiber &f = B->fibers[(size_t)B->cur];
	f.wait_on = nullptr;
	simt_switch(&f.sp, B->sched_sp);
}

// every lane of the calling warp deposits v; returns the 32 deposited values
static inline const uint64_t *warp_exchange(uint64_t v)
{
	const int tid = B->cur, lane = tid & 31;
	WarpSync &w = B->warps[(size_t)(tid >> 5)];
	const unsigned g = w.gen;
	w.in[lane] = v;
	B->collectives++;
	if (++w.arrived == 32) {
		memcpy(w.out[g & 1], w.in, sizeof(w.in));
		w.arrived = 0;
		w.gen = g + 1;
	} else
		block_on(&w.gen, g);
	return w.out[g & 1];
}

static inline void bar_sync(int id, int count)
{
	BarSync &b = B->bars[id];
	const unsigned g = b.gen;
	if (++b.arrived == count) {
		b.arrived = 0;
		b.gen = g + 1;
	} else
		block_on(&b.gen, g);
}

// bar.arrive: counts towards the barrier without waiting for it (the producer half of a producer / consumer pair)
static inline void bar_arrive(int id, int count)
{
	BarSync &b = B->bars[id];
	if (++b.arrived == count) {
		b.arrived = 0;
		b.gen++;
	}
}

static void fiber_main()
{
	B->entry();
	Fiber &f = B->fibers[(size_t)B->cur];
	f.done = true;
	simt_switch(&f.sp, B->sched_sp);
	abort(); // a finished fiber is never resumed
}

// Runs `kernel` as one block of `nthreads` threads with blockIdx.x = block_x.  Returns false on a deadlock.
static inline bool run_block(int nthreads, unsigned block_x, std::function<void()> kernel, size_t stack_bytes = 256 << 10,
			     unsigned grid_x = 1, size_t dyn_smem_bytes = 0, unsigned block_y = 0, unsigned grid_y = 1)
{
	Block blk;
	blk.bidx.y = block_y;
	blk.gdim.y = grid_y;
	blk.dyn_smem.assign(dyn_smem_bytes + 256, 0xCD); // uninitialised on the device: a pattern, not zeros
	blk.nthreads = nthreads;
	blk.bidx.x = block_x;
	blk.bdim.x = (unsigned)nthreads;
	blk.gdim.x = grid_x;
	blk.fibers.resize((size_t)nthreads);
	blk.warps.resize((size_t)(nthreads + 31) / 32);
	blk.entry = std::move(kernel);
	char *arena = (char *)malloc(stack_bytes * (size_t)nthreads + 64);
	for (int t = 0; t < nthreads; t++) {
		Fiber &f = blk.fibers[(size_t)t];
		f.stack = arena + stack_bytes * (size_t)t;
		uintptr_t top = ((uintptr_t)f.stack + stack_bytes) & ~(uintptr_t)15;
		void **sp = (void **)top;
		*--sp = nullptr;              // where a caller's return address would be: keeps rsp = 8 (mod 16) at entry
		*--sp = (void *)&fiber_main;  // popped by simt_switch's ret
		for (int i = 0; i < 6; i++)
			*--sp = nullptr;          // rbp rbx r12 r13 r14 r15
		f.sp = (void *)sp;
	}
	Block *outer = B;
	B = &blk;
	bool ok = true;
	for (;;) {
		bool progress = false, alive = false;
		for (int t = 0; t < nthreads; t++) {
			Fiber &f = blk.fibers[(size_t)t];
			if (f.done)
				continue;
			alive = true;
			if (f.wait_on && *f.wait_on == f.wait_val)
				continue;
			blk.cur = t;
			blk.tidx.x = (unsigned)t;
			simt_switch(&blk.sched_sp, f.sp);
			progress = true;
		}
		if (!alive)
			break;
		if (!progress) {
			fprintf(stderr, "simt: deadlock (a collective or barrier that not every thread reaches)\n");
			ok = false;
			break;
		}
	}
	B = outer;
	free(arena);
	return ok;
}

// A one-dimensional grid, block after block (blocks of a grid do not synchronise with each other).
static inline bool run_grid(dim3 grid, int nthreads, std::function<void()> kernel, size_t stack_bytes = 64 << 10,
			    size_t dyn_smem_bytes = 0)
{
	for (unsigned by = 0; by < grid.y; by++)
		for (unsigned b = 0; b < grid.x; b++)
			if (!run_block(nthreads, b, kernel, stack_bytes, grid.x, dyn_smem_bytes, by, grid.y))
				return false;
	return true;
}
static inline uint8_t *dyn_smem() // 128-byte aligned
{
	uint8_t *p = B->dyn_smem.data();
	return p + ((128 - ((uintptr_t)p & 127)) & 127);
}

} // namespace simt

#define threadIdx (simt::B->tidx)
#define blockIdx (simt::B->bidx)
#define blockDim (simt::B->bdim)
#define gridDim (simt::B->gdim)

static inline unsigned __ballot_sync(unsigned, bool pred)
{
	const uint64_t *v = simt::warp_exchange(pred ? 1 : 0);
	unsigned m = 0;
	for (int i = 0; i < 32; i++)
		m |= (unsigned)(v[i] & 1) << i;
	return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
static inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_exchange(0); }
static inline void __syncthreads() { simt::bar_sync(0, simt::B->nthreads); }
template <class T>
static inline T __shfl_sync(unsigned, T val, int src)
{
	static_assert(sizeof(T) <= 8, "shuffle operand");
	uint64_t raw = 0;
	memcpy(&raw, &val, sizeof(T));
	const uint64_t *v = simt::warp_exchange(raw);
	T r;
	memcpy(&r, &v[src & 31], sizeof(T));
	return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T val, int lane_mask)
{
	uint64_t raw = 0;
	memcpy(&raw, &val, sizeof(T));
	const int lane = simt::B->cur & 31;
	const uint64_t *v = simt::warp_exchange(raw);
	T r;
	memcpy(&r, &v[(lane ^ lane_mask) & 31], sizeof(T));
	return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T val, unsigned delta)
{
	uint64_t raw = 0;
	memcpy(&raw, &val, sizeof(T));
	const int lane = simt::B->cur & 31;
	const uint64_t *v = simt::warp_exchange(raw);
	T r;
	memcpy(&r, &v[lane >= (int)delta ? lane - (int)delta : lane], sizeof(T));
	return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T val, unsigned delta)
{
	uint64_t raw = 0;
	memcpy(&raw, &val, sizeof(T));
	const int lane = simt::B->cur & 31;
	const uint64_t *v = simt::warp_exchange(raw);
	T r;
	memcpy(&r, &v[lane + (int)delta < 32 ? lane + (int)delta : lane], sizeof(T));
	return r;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) { simt::yield(); }
static inline void __threadfence_block() {}
template <class T>
static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
static inline unsigned __brev(unsigned v)
{
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		r |= ((v >> i) & 1u) << (31 - i);
	return r;
}
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __reduce_add_sync(unsigned, unsigned val)
{
	const uint64_t *v = simt::warp_exchange(val);
	unsigned s = 0;
	for (int i = 0; i < 32; i++)
		s += (unsigned)v[i];
	return s;
}
template <class T>
static inline T __ldcs(const T *p) { return *p; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) // low 32 bits of (hi:lo) >> (sh & 31)
{
	sh &= 31;
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline unsigned atomicXor(unsigned *p, unsigned v)
{
	const unsigned o = *p;
	*p = o ^ v;
	return o;
}
// the few runtime calls that kernel files make next to their kernels (table set-up)
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
static inline int cudaGetLastError() { return cudaSuccess; }
static inline void __threadfence() {}
static inline unsigned atomicAdd(unsigned *p, unsigned v)
{
	const unsigned o = *p;
	*p = o + v;
	return o;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
	const unsigned long long o = *p;
	*p = o + v;
	return o;
}
static inline unsigned __match_any_sync(unsigned, unsigned val) // lanes of the warp holding the same value
{
	const uint64_t *v = simt::warp_exchange(val);
	unsigned m = 0;
	for (int i = 0; i < 32; i++)
		m |= (unsigned)((unsigned)v[i] == val) << i;
	return m;
}
#define cudaMemcpyToSymbol(sym, src, bytes) (memcpy((void *)&(sym), (src), (bytes)), cudaSuccess)
static inline unsigned atomicOr(unsigned *p, unsigned v)
{
	const unsigned o = *p;
	*p = o | v;
	return o;
}
static inline int atomicAdd(int *p, int v)
{
	const int o = *p;
	*p = o + v;
	return o;
}
