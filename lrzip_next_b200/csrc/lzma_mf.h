// lzma_mf.h -- host interface of the data-parallel LZMA match finder (lzma_mf.cu).
#pragma once
#include "kernels.h" // cuda_runtime.h (or the emulator's stand-ins, tests/hostsim) and LRZ_LAUNCH
#include <stddef.h>
#include <stdint.h>

#include "lzma_mf.cuh"

namespace lrz {
namespace lzma {

constexpr uint32_t kMfMaxFb = 64; // lrzip-next passes fb = 32 or 64 (src/stream.c:455)

// One stream block's match-finder arrays in HBM.  Positions are 0-based byte indices i here; the
// reference's (and mf_bt_insert's) positions are i + 1.
struct MfBlock {
	const uint8_t *src;
	MfParams P;
	uint32_t count;       // positions that are searched and inserted: at least 4 (bt4) / 5 (hc5) bytes available
	uint32_t *son;        // 2 * (n + 2): node p at son[2p], son[2p + 1]
	uint32_t *c2, *c3;    // [count] previous position (1-based) with the same 2- / 3-byte hash
	uint32_t *sorted;     // [count] bt4: positions ordered by (hash4, position); hc5: previous position with the same hash5
	uint64_t *rec;        // [n] (pool offset << 10) | number of uint32 of the position's match list; zeroed by the caller
	uint32_t *pool;
	uint64_t poolCap;     // in uint32
	unsigned long long *cursor; // zeroed by the caller
	int *overflow;        // zeroed by the caller
};

int mf_init_tables();
int mf_preload();
// scratch needed by mf_prepare_block for blocks of up to maxCount positions
size_t mf_sort_scratch_bytes(uint32_t maxCount);
// sorts + c2/c3 + bucket order of one block (stream-ordered; scratch may be reused by the next block)
int mf_prepare_block(const MfBlock &B, void *scratch, cudaStream_t st, int64_t *launches);
// the tree walk of all prepared blocks in one launch; d_segBase[b] = sum of count of blocks < b (nblocks + 1 entries)
// hc5: the blocks use the hash-chain finder (levels 1-4): one thread per position instead of one per bucket
int mf_walk_launch(const MfBlock *d_blocks, int nblocks, const uint64_t *d_segBase, uint64_t total, bool hc5,
		   cudaStream_t st, int64_t *launches);

} // namespace lzma
} // namespace lrz
