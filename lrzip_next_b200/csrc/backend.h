// backend.h -- per-block backends (device side).  A backend takes the stream blocks of one chunk,
// already resident in HBM, and produces for each a c_type, a compressed length and a device payload.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/lrzgpu.h"

namespace lrz {

struct BlockJob {
	const uint8_t *d_src; // block bytes in HBM (inside stream 0 / stream 1)
	int64_t u_len;
	int stream;
	// results
	int c_type;           // LRZGPU_CTYPE_*
	int64_t c_len;
	const uint8_t *d_payload; // == d_src when stored
};

struct BackendCtx; // opaque per-context scratch, owned by lrzgpu_ctx

BackendCtx *backend_create();
int backend_preload(); // load every backend kernel's code now
void backend_destroy(BackendCtx *b);
// Runs the lz4 gate (when threshold != 0) and the backend on every job with u_len >= 64
// (src/stream.c:1633), synchronously on `stream`.  Jobs left stored keep c_type NONE / c_len u_len.
int backend_encode_blocks(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, std::vector<BlockJob> &jobs,
			  int num_sms, cudaStream_t stream, int64_t *launches, char *err, size_t errlen);
// ---- the LZMA backend as a pipeline (what backend_encode_blocks runs underneath): blocks are submitted while
// the rzip stage is still producing later ones (the reference hands a block to a compthread the moment it
// fills, src/stream.c:1836-1875) and drained at the end of the chunk.  Results come back in submission order.
int backend_async_begin(BackendCtx *b, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t max_block_len,
			int64_t max_blocks, int64_t payload_bytes_upper, char *err, size_t errlen);
// `ready`: an event after which the blocks' bytes are in place (or null).  Returns 0, a negative LRZGPU_E* code, or
// 1 when the work space is used up (drain, then submit again).  Blocks must have u_len >= 64.
int backend_async_submit(BackendCtx *b, const BlockJob *jobs, int n, cudaEvent_t ready, int64_t *launches, char *err,
			 size_t errlen);
// launch whatever has become ready since (never blocks); call it now and then while blocks are in flight
int backend_async_poll(BackendCtx *b, int64_t *launches, char *err, size_t errlen);
int backend_async_drain(BackendCtx *b, int64_t *launches, char *err, size_t errlen);
int backend_async_count(const BackendCtx *b);
const BlockJob *backend_async_result(const BackendCtx *b, int i); // null until drained

// lz4_compresses() of src/stream.c:2325-2380 on a device-resident buffer.
int backend_lz4_gate(BackendCtx *b, const uint8_t *d_src, int64_t len, int threshold, int *result, cudaStream_t stream,
		     int64_t *launches);

} // namespace lrz
