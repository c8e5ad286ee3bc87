// unrzip.cu -- the decode side on the device: rzip stream replay and the LZMA block decoder.
//
//   runzip_chunk()      src/runzip.c:261-370   stream 0 = records, stream 1 = literal bytes
//     unzip_literal()   src/runzip.c:139-176   `len` bytes from stream 1 to the output
//     unzip_match()     src/runzip.c:178-241   `len` bytes copied from `dist` back in the output (copies may
//                                              overlap: the reference replicates min(len, dist) bytes at a time)
//   lzma_decompress_buf src/stream.c:556-616 -> LzmaUncompress -> LzmaDec (src/lzma/C/LzmaDec.c): raw LZMA
//                                              stream, lc3 lp0 pb2, known output size, no end marker
//   zstd_decompress_buf src/stream.c:1989-2010 -> ZSTD_decompress: one RFC 8878 frame per block (zstd_dec.cuh)
//
// The replay is a gather once the record offsets are known: one thread walks stream 0 and writes, per record,
// where its bytes go (a running sum -- the only serial part, 3 to 3+cb bytes per record); every literal byte is
// then independent (HBM-bound scatter over all SMs); matches read bytes that earlier records produced, so they are
// replayed in record order by one CTA, each copy spread over its threads (a copy that overlaps itself reads
// periodically from the dist bytes before it, so it is parallel as well).
#include "kernels.h"
#include "zstd_dec.cuh"

namespace lrz {

namespace {

// ---- stream 0 -> records -----------------------------------------------------------------------------------
__global__ void s0_parse_kernel(const uint8_t *__restrict__ s0, int64_t s0_len, int cb, int64_t chunk_size, DecLit *lits,
				DecMatch *matches, int64_t cap, DecSummary *sum)
{
	if (threadIdx.x || blockIdx.x)
		return;
	int64_t pos = 0, o = 0, l = 0, nl = 0, nm = 0;
	int status = 0;
	for (;;) {
		if (pos + 3 > s0_len) {
			status = -1; // ran off the end without a terminator
			break;
		}
		const int head = s0[pos];
		const int64_t len = (int64_t)s0[pos + 1] | ((int64_t)s0[pos + 2] << 8); // fixed width 2 (src/runzip.c:315)
		pos += 3;
		if (head == 0) {
			if (len == 0)
				break; // terminator (src/runzip.c:322)
			if (nl >= cap) {
				status = -2;
				break;
			}
			lits[nl].out_off = o;
			lits[nl].lit_off = l;
			lits[nl].len = len;
			nl++;
			o += len;
			l += len;
		} else {
			if (pos + cb > s0_len) {
				status = -1;
				break;
			}
			int64_t dist = 0;
			for (int i = 0; i < cb; i++)
				dist |= (int64_t)s0[pos + i] << (8 * i);
			pos += cb;
			if (nm >= cap || dist < 1 || dist > o) {
				status = dist < 1 || dist > o ? -3 : -2; // a match may not reach before the chunk
				break;
			}
			matches[nm].out_off = o;
			matches[nm].len = len;
			matches[nm].dist = dist;
			nm++;
			o += len;
		}
		if (o > chunk_size) {
			status = -4;
			break;
		}
	}
	uint32_t crc = 0;
	if (!status) {
		if (pos + 4 > s0_len)
			status = -1;
		else // gcrypt's digest order: most significant byte first
			crc = ((uint32_t)s0[pos] << 24) | ((uint32_t)s0[pos + 1] << 16) | ((uint32_t)s0[pos + 2] << 8) | s0[pos + 3];
	}
	sum->n_lit = nl;
	sum->n_match = nm;
	sum->out_len = o;
	sum->lit_len = l;
	sum->crc = crc;
	sum->status = status;
}

// ---- literals: stream 1 -> output, 16 bytes per thread -------------------------------------------------------
__device__ __forceinline__ int64_t lit_of(const DecLit *lits, int64_t n, int64_t x) // last record with lit_off <= x
{
	int64_t lo = 0, hi = n - 1;
	while (lo < hi) {
		const int64_t mid = (lo + hi + 1) >> 1;
		if (lits[mid].lit_off <= x)
			lo = mid;
		else
			hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(256) lit_scatter_kernel(const uint8_t *__restrict__ s1, int64_t s1_len, const DecLit *__restrict__ lits,
							   int64_t n_lit, uint8_t *__restrict__ out)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
	for (int64_t x = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; x < s1_len; x += stride) {
		int64_t r = lit_of(lits, n_lit, x);
		const int64_t xe = x + 16 < s1_len ? x + 16 : s1_len;
		for (int64_t y = x; y < xe; y++) {
			while (y >= lits[r].lit_off + lits[r].len)
				r++;
			out[lits[r].out_off + (y - lits[r].lit_off)] = s1[y];
		}
	}
}

// ---- matches, in record order, one CTA -----------------------------------------------------------------------
__global__ void __launch_bounds__(1024) match_replay_kernel(const DecMatch *__restrict__ matches, int64_t n_match, uint8_t *out)
{
	for (int64_t k = 0; k < n_match; k++) {
		const int64_t o = matches[k].out_off, len = matches[k].len, dist = matches[k].dist;
		const uint8_t *from = out + o - dist;
		if (dist >= len) {
			for (int64_t i = threadIdx.x; i < len; i += 1024)
				out[o + i] = from[i];
		} else { // the copy overlaps itself: byte i repeats byte i mod dist of the dist bytes before the match
			for (int64_t i = threadIdx.x; i < len; i += 1024)
				out[o + i] = from[i % dist];
		}
		__syncthreads(); // the next match may read what this one wrote
	}
}

// ---- LZMA block decoder (LzmaDec.c, lc3 lp0 pb2), one thread per block -----------------------------------------
constexpr int kLc = 3, kPb = 2;
constexpr uint32_t kTop = 1u << 24, kBitTotal = 1u << 11;
constexpr int kMove = 5;
constexpr int kNumStates = 12, kPosStates = 1 << kPb;
constexpr int kLenLow = 8, kLenMid = 8, kLenHigh = 256;
// probability layout (uint16 each)
constexpr int oIsMatch = 0;                                   // [12][4]
constexpr int oIsRep = oIsMatch + kNumStates * kPosStates;    // [12]
constexpr int oIsRepG0 = oIsRep + kNumStates;
constexpr int oIsRepG1 = oIsRepG0 + kNumStates;
constexpr int oIsRepG2 = oIsRepG1 + kNumStates;
constexpr int oIsRep0Long = oIsRepG2 + kNumStates;            // [12][4]
constexpr int oPosSlot = oIsRep0Long + kNumStates * kPosStates; // [4][64]
constexpr int oSpecPos = oPosSlot + 4 * 64;                   // [128]
constexpr int oAlign = oSpecPos + 128;                        // [16]
constexpr int oLen = oAlign + 16;                             // choice, choice2, low[4][8], mid[4][8], high[256]
constexpr int kLenProbs = 2 + kPosStates * kLenLow + kPosStates * kLenMid + kLenHigh;
constexpr int oRepLen = oLen + kLenProbs;
constexpr int oLit = oRepLen + kLenProbs;                     // [0x300 << lc]
constexpr int kNumProbs = oLit + (0x300 << kLc);

struct RcDec {
	const uint8_t *p, *end;
	uint32_t range, code;
	int err;
};

__device__ __forceinline__ void rd_norm(RcDec &r)
{
	if (r.range < kTop) {
		r.range <<= 8;
		uint32_t b = 0;
		if (r.p < r.end)
			b = *r.p++;
		else
			r.err = 1;
		r.code = (r.code << 8) | b;
	}
}

__device__ __forceinline__ uint32_t rd_bit(RcDec &r, uint16_t *prob)
{
	rd_norm(r);
	const uint32_t p = *prob, bound = (r.range >> 11) * p;
	if (r.code < bound) {
		r.range = bound;
		*prob = (uint16_t)(p + ((kBitTotal - p) >> kMove));
		return 0;
	}
	r.range -= bound;
	r.code -= bound;
	*prob = (uint16_t)(p - (p >> kMove));
	return 1;
}

__device__ uint32_t rd_tree(RcDec &r, uint16_t *probs, int bits)
{
	uint32_t m = 1;
	for (int i = 0; i < bits; i++)
		m = (m << 1) | rd_bit(r, probs + m);
	return m - (1u << bits);
}

__device__ uint32_t rd_tree_rev(RcDec &r, uint16_t *probs, int bits)
{
	uint32_t m = 1, sym = 0;
	for (int i = 0; i < bits; i++) {
		const uint32_t b = rd_bit(r, probs + m);
		m = (m << 1) | b;
		sym |= b << i;
	}
	return sym;
}

__device__ uint32_t rd_len(RcDec &r, uint16_t *lp, uint32_t posState)
{
	if (!rd_bit(r, lp))
		return rd_tree(r, lp + 2 + posState * kLenLow, 3);
	if (!rd_bit(r, lp + 1))
		return kLenLow + rd_tree(r, lp + 2 + kPosStates * kLenLow + posState * kLenMid, 3);
	return kLenLow + kLenMid + rd_tree(r, lp + 2 + kPosStates * (kLenLow + kLenMid), 8);
}

__global__ void lzma_dec_kernel(LzmaDecJob *jobs, int njobs, uint16_t *prob_arena)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= njobs)
		return;
	LzmaDecJob &j = jobs[b];
	uint16_t *P = prob_arena + (size_t)b * kNumProbs;
	for (int i = 0; i < kNumProbs; i++)
		P[i] = kBitTotal >> 1;
	RcDec r;
	r.p = j.src;
	r.end = j.src + j.c_len;
	r.range = 0xFFFFFFFFu;
	r.code = 0;
	r.err = 0;
	if (j.c_len < 5 || j.src[0] != 0) {
		j.status = -1;
		return;
	}
	r.p++;
	for (int i = 0; i < 4; i++)
		r.code = (r.code << 8) | *r.p++;
	uint8_t *out = j.out;
	const int64_t n = j.u_len;
	uint32_t state = 0, rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0; // distances minus one
	int64_t pos = 0;
	int status = 0;
	while (pos < n) {
		const uint32_t posState = (uint32_t)pos & (kPosStates - 1);
		if (!rd_bit(r, P + oIsMatch + state * kPosStates + posState)) {
			const uint32_t prev = pos ? out[pos - 1] : 0;
			uint16_t *lp = P + oLit + 0x300 * (prev >> (8 - kLc)); // lp = 0: no position bits
			uint32_t sym = 1;
			if (state >= 7) {
				uint32_t mb = out[pos - rep0 - 1];
				while (sym < 0x100) {
					const uint32_t mbit = (mb >> 7) & 1;
					mb <<= 1;
					const uint32_t bit = rd_bit(r, lp + ((1 + mbit) << 8) + sym);
					sym = (sym << 1) | bit;
					if (mbit != bit)
						break;
				}
			}
			while (sym < 0x100)
				sym = (sym << 1) | rd_bit(r, lp + sym);
			out[pos++] = (uint8_t)sym;
			state = state < 4 ? 0 : (state < 10 ? state - 3 : state - 6);
			continue;
		}
		uint32_t len;
		if (rd_bit(r, P + oIsRep + state)) {
			if (pos == 0) {
				status = -2;
				break;
			}
			if (!rd_bit(r, P + oIsRepG0 + state)) {
				if (!rd_bit(r, P + oIsRep0Long + state * kPosStates + posState)) {
					state = state < 7 ? 9 : 11;
					out[pos] = out[pos - rep0 - 1];
					pos++;
					continue;
				}
			} else {
				uint32_t d;
				if (!rd_bit(r, P + oIsRepG1 + state))
					d = rep1;
				else {
					if (!rd_bit(r, P + oIsRepG2 + state))
						d = rep2;
					else {
						d = rep3;
						rep3 = rep2;
					}
					rep2 = rep1;
				}
				rep1 = rep0;
				rep0 = d;
			}
			len = rd_len(r, P + oRepLen, posState);
			state = state < 7 ? 8 : 11;
		} else {
			rep3 = rep2;
			rep2 = rep1;
			rep1 = rep0;
			len = rd_len(r, P + oLen, posState);
			state = state < 7 ? 7 : 10;
			const uint32_t slot = rd_tree(r, P + oPosSlot + (len < 4 ? len : 3) * 64, 6);
			if (slot < 4)
				rep0 = slot;
			else {
				const int nd = (int)(slot >> 1) - 1;
				rep0 = (2 | (slot & 1)) << nd;
				if (slot < 14)
					rep0 += rd_tree_rev(r, P + oSpecPos + rep0 - slot - 1, nd);
				else {
					uint32_t dbits = 0;
					for (int i = 0; i < nd - 4; i++) { // direct bits
						rd_norm(r);
						r.range >>= 1;
						const uint32_t t = (r.code - r.range) >> 31; // 1 when code < range
						r.code -= r.range & (t - 1);
						dbits = (dbits << 1) | (1 - t);
					}
					rep0 += dbits << 4;
					rep0 += rd_tree_rev(r, P + oAlign, 4);
				}
			}
			if ((int64_t)rep0 >= pos) {
				status = -2;
				break;
			}
		}
		len += 2;
		if (pos + len > n) {
			status = -3;
			break;
		}
		const uint8_t *from = out + pos - rep0 - 1;
		for (uint32_t i = 0; i < len; i++)
			out[pos + i] = from[i];
		pos += len;
	}
	rd_norm(r);
	if (!status && r.err)
		status = -4;
	j.status = status;
	j.produced = pos;
}

// One zstd frame per job (the payload of a CTYPE_ZSTD stream block), one thread per frame, one frame per CTA so that
// the frames of a chunk spread over the SMs (zstd_dec.cuh).
__global__ void zstd_dec_kernel(LzmaDecJob *jobs, int njobs, zd::Work *work)
{
	const int b = blockIdx.x;
	if (b >= njobs || threadIdx.x)
		return;
	LzmaDecJob &j = jobs[b];
	const int64_t r = zd::decode_frame(j.src, j.c_len, j.out, j.u_len, work + b);
	j.status = r < 0 ? (int32_t)r : 0;
	j.produced = r < 0 ? 0 : r;
}

} // namespace

#if !defined(LRZ_SIMT_HOST) // (the emulator calls the kernels directly)
int unrzip_parse_launch(const uint8_t *d_s0, int64_t s0_len, int cb, int64_t chunk_size, DecLit *d_lits, DecMatch *d_matches,
			int64_t cap, DecSummary *d_sum, cudaStream_t stream)
{
	s0_parse_kernel<<<1, 1, 0, stream>>>(d_s0, s0_len, cb, chunk_size, d_lits, d_matches, cap, d_sum);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int unrzip_replay_launch(const uint8_t *d_s1, int64_t s1_len, const DecLit *d_lits, int64_t n_lit, const DecMatch *d_matches,
			 int64_t n_match, uint8_t *d_out, int num_sms, cudaStream_t stream)
{
	if (s1_len > 0 && n_lit > 0) {
		int64_t grid = (s1_len + 256 * 16 - 1) / (256 * 16);
		if (grid > (int64_t)num_sms * 8)
			grid = (int64_t)num_sms * 8;
		lit_scatter_kernel<<<(unsigned)grid, 256, 0, stream>>>(d_s1, s1_len, d_lits, n_lit, d_out);
	}
	if (n_match > 0)
		match_replay_kernel<<<1, 1024, 0, stream>>>(d_matches, n_match, d_out);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

size_t lzma_dec_prob_bytes(int njobs) { return (size_t)njobs * kNumProbs * sizeof(uint16_t); }

int lzma_dec_launch(LzmaDecJob *d_jobs, int njobs, void *d_probs, cudaStream_t stream)
{
	if (njobs <= 0)
		return 0;
	lzma_dec_kernel<<<(unsigned)((njobs + 31) / 32), 32, 0, stream>>>(d_jobs, njobs, (uint16_t *)d_probs);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

size_t zstd_dec_work_bytes(int njobs) { return (size_t)njobs * sizeof(zd::Work); }

int zstd_dec_launch(LzmaDecJob *d_jobs, int njobs, void *d_work, cudaStream_t stream)
{
	if (njobs <= 0)
		return 0;
	zstd_dec_kernel<<<(unsigned)njobs, 32, 0, stream>>>(d_jobs, njobs, (zd::Work *)d_work);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int unrzip_preload()
{
	cudaFuncAttributes a;
	bool ok = cudaFuncGetAttributes(&a, s0_parse_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, lit_scatter_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, match_replay_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, lzma_dec_kernel) == cudaSuccess;
	ok = ok && cudaFuncGetAttributes(&a, zstd_dec_kernel) == cudaSuccess;
	return ok ? 0 : -1;
}
#endif // !LRZ_SIMT_HOST

} // namespace lrz
