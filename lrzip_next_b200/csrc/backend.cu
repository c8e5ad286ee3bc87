// backend.cu -- per-block backend dispatch (src/stream.c:1633-1650 compthread -> *_compress_buf).
#include "backend.h"

#include <stdio.h>

namespace lrz {

struct BackendCtx {
	int dummy;
};

BackendCtx *backend_create() { return new BackendCtx(); }
void backend_destroy(BackendCtx *b) { delete b; }

int backend_encode_blocks(BackendCtx *, const lrzgpu_params &p, const lrzgpu_sizing_t &, std::vector<BlockJob> &, int,
			  cudaStream_t, int64_t *, char *err, size_t errlen)
{
	snprintf(err, errlen, "backend %d is not built yet", p.backend);
	return LRZGPU_EUNSUPPORTED;
}

int backend_lz4_gate(BackendCtx *, const uint8_t *, int64_t, int, int *, cudaStream_t, int64_t *)
{
	return LRZGPU_EUNSUPPORTED;
}

} // namespace lrz
