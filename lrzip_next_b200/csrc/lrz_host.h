// lrz_host.h -- host-side (CPU, control-plane only) pieces of the product: sizing rules, block
// flush plan, archive framing helpers and the whole-file MD5.  None of the data-path arithmetic
// (tags, matches, streams, CRC, block payloads) lives here; that is all in the kernels.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/lrzgpu.h"

namespace lrz {

constexpr int64_t kOneMB = 1048576;
constexpr int64_t kStreamBufsize = 10 * kOneMB;   // STREAM_BUFSIZE, src/include/lrzip_private.h:16
constexpr int64_t kChunkMultiple = 100 * kOneMB;  // CHUNK_MULTIPLE, src/rzip.c:48

int compute_sizing(const lrzgpu_params &p, int64_t st_size, lrzgpu_sizing_t &out);
unsigned lzma2_prop_from_dic(uint32_t dict);
int chunk_bytes_for(int64_t chunk_size); // src/rzip.c:1129-1133
int zstd_level_for(int level);           // src/main.c:87

// One stream block in flush order.
struct BlockPlan {
	int stream;
	int64_t off, u_len;
};
// Global flush order of a chunk's blocks (src/stream.c:2198-2216, 2253-2259): w1[j] = stream-1 bytes
// written when stream-0 block j filled (from k4_flush_order).
void plan_blocks(int64_t s0_len, int64_t s1_len, int64_t bufsize, const int64_t *w1, std::vector<BlockPlan> &out);

void put_le(uint8_t *at, int64_t v, int width);
void make_magic(uint8_t magic[21], const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t st_size);

// Streaming MD5 (RFC 1321): the reference's default whole-file hash (src/main.c:789), fed in file
// order by a host thread while the GPU works.
struct Md5 {
	uint32_t a, b, c, d;
	uint64_t len;
	uint8_t buf[64];
	int fill;
	Md5();
	void update(const uint8_t *p, size_t n);
	void final(uint8_t digest[16]);
};

} // namespace lrz
