// k1_tagscan.h -- host launchers for the K1 tag-scan kernel (see k1_tagscan.cu).
#pragma once
#include <cuda_runtime.h>
#include "lrz_common.h"

namespace lrz {
int k1_init_tables();
// Scan positions [pos_lo, pos_hi) (pos_lo tile-aligned) of the n-byte chunk at d_buf (padded by
// at least kTile + kInputPad zero bytes) and write candidates with (tag & mask) == mask.
int k1_launch(const uint8_t *d_buf, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask,
	      Cand *d_cand, uint32_t *d_tile_count, int num_sms, cudaStream_t stream);
} // namespace lrz
