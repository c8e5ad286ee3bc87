// lrzgpu.cu -- the C ABI (include/lrzgpu.h): context, device arenas, the per-chunk pipeline
//   H2D -> [K1 tag scan || K2 commit, segment-pipelined on two streams] || CRC-32 -> K4 emit ->
//   block plan -> backend -> framing -> D2H
// and whole-file orchestration (window loop src/rzip.c:1041-1186, magic src/lrzip.c:131-208, MD5).
#include "../../include/lrzgpu.h"

#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "backend.h"
#include "k2_commit.cuh"
#include "kernels.h"
#include "lrz_host.h"

using namespace lrz;

namespace {

constexpr int64_t kFrontPad = 256;
constexpr int64_t kSegment = 16ll << 20; // positions per K1/K2 segment (tile multiple)
constexpr int kCtypeNone = LRZGPU_CTYPE_NONE;

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t n)
	{
		if (n <= cap)
			return cudaSuccess;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		const size_t want = n + (n >> 4) + 4096;
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct lrzgpu_ctx {
	int device = 0, sms = 148;
	char err[512] = { 0 };
	cudaStream_t sA = nullptr, sB = nullptr, sC = nullptr, sD = nullptr, sE = nullptr;
	cudaEvent_t evK1[2] = { nullptr, nullptr }, evK2[2] = { nullptr, nullptr }, evInit = nullptr, evCrc = nullptr;
	DevBuf in, tab, state, cand[2], tc[2], recs, s0, s1, crc, w1, fside;
	ScanState *h_state = nullptr; // pinned, one per variant of the window being scanned
	size_t h_state_cap = 0;
	ScanState *h_snap = nullptr; // pinned, one per segment of a pipelined scan
	size_t h_snap_cap = 0;
	std::vector<cudaEvent_t> evSeg;
	cudaEvent_t evLit = nullptr;
	uint8_t *h_pin[2] = { nullptr, nullptr }; // pinned staging for the MD5 stream of device inputs
	size_t h_pin_cap = 0;
	BackendCtx *backend = nullptr;
	int64_t launches = 0;
	// window between lrzgpu_chunk_begin and lrzgpu_chunk_finish (rzip done, streams resident in s0 / s1)
	struct Pending {
		bool valid = false;
		lrzgpu_params p;
		lrzgpu_sizing_t sz;
		int64_t n = 0, s0_len = 0, s1_len = 0, n_rec = 0, rec_base = 0;
		int eof = 0, cb = 0;
		const uint8_t *d_chunk = nullptr;
		bool selected = false; // streams emitted (always true after lrzgpu_chunk_begin)
		bool pipelined = false; // lrzgpu_chunk_begin handed blocks to the backend during the scan
		int64_t blk1 = 0;
		std::vector<int64_t> v_s0, v_s1, v_nrec, v_base, v_out; // per variant, after lrzgpu_chunk_begin_all
		std::vector<ScanState> v_st;
	} pending;
};

namespace {

int fail(lrzgpu_ctx *c, int code, const char *fmt, ...)
{
	if (c) {
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(c->err, sizeof(c->err), fmt, ap);
		va_end(ap);
	}
	return code;
}

#define CU(c, call)                                                                                          \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess)                                                                       \
			return fail(c, e_ == cudaErrorMemoryAllocation ? LRZGPU_ENOMEM : LRZGPU_ECUDA, "%s: %s", #call, \
				    cudaGetErrorString(e_));                                                 \
	} while (0)

struct ChunkResult {
	int64_t s0_len = 0, s1_len = 0, n_rec = 0;
	int64_t rec_base = 0; // first record of this variant in c->recs
	ScanState st;
};

void account_rzip(lrzgpu_stats *stats, const ChunkResult &res)
{
	if (!stats)
		return;
	stats->matches += res.st.st_matches;
	stats->match_bytes += res.st.st_match_bytes;
	stats->literals += res.st.st_literals;
	stats->literal_bytes += res.st.st_literal_bytes;
	stats->tag_hits += res.st.st_tag_hits;
	stats->tag_misses += res.st.st_tag_misses;
	stats->inserts += res.st.st_inserts;
	stats->lookups += res.st.st_lookups;
	stats->chain_evictions += res.st.st_evictions;
	stats->sweeps += res.st.st_sweeps;
	stats->displacements += res.st.st_displacements;
	stats->hash_count = res.st.hash_count;
	stats->final_min_mask = res.st.min_mask;
	stats->final_tag_mask = res.st.tag_mask;
	stats->chunks += 1;
	stats->stream0_bytes += res.s0_len;
	stats->stream1_bytes += res.s1_len;
}

// The scan of one chunk that is resident in HBM at d_chunk (16-byte aligned, readable kFrontPad before and
// kInputPad after): CRC-32, K1 tag scan and K2 commit, segment-pipelined on two streams.  nvar == 1 runs the
// chunk from the given victim_round; nvar > 1 runs one commit per possible incoming value 0..nvar-1 of the
// reference's cross-window counter (all-values speculation), each with its own table, state and
// match records.  Leaves the match records in c->recs (variant v from res[v].rec_base).
// `progress` (single-variant scans only) is called on the host after every segment with the scan state as of the
// end of that segment -- match records [0, n_rec) and the stream lengths are final up to there -- while the
// device goes on with the following segments: this is where the backend gets its blocks from early.
using ScanProgress = std::function<int(const ScanState &)>;

int rzip_scan_device(lrzgpu_ctx *c, const uint8_t *d_chunk, int64_t n, int rzip_level, int cb, int64_t victim_round,
		     int nvar, std::vector<ChunkResult> &res, lrzgpu_stats *stats, const ScanProgress *progress = nullptr)
{
	const double t0 = now_ms();
	const RzipLevel &lv = kLevels[rzip_level];
	const size_t tab_bytes = (size_t)lv.mb_used << 20;
	const int64_t rec_cap = n / kMinMatch + 8;
	const int64_t seg = kSegment < ((n + kTile - 1) / kTile) * kTile ? kSegment : ((n + kTile - 1) / kTile) * kTile;
	CU(c, c->tab.ensure(tab_bytes * (size_t)nvar));
	CU(c, c->state.ensure(sizeof(ScanState) * (size_t)nvar));
	CU(c, c->recs.ensure((size_t)rec_cap * sizeof(MatchRec) * (size_t)nvar));
	CU(c, c->crc.ensure(16));
	for (int b = 0; b < 2; b++) {
		CU(c, c->cand[b].ensure((size_t)seg * sizeof(Cand)));
		CU(c, c->tc[b].ensure((size_t)(seg / kTile + 1) * sizeof(uint32_t)));
	}
	if ((size_t)nvar > c->h_state_cap) {
		if (c->h_state)
			cudaFreeHost(c->h_state);
		c->h_state = nullptr;
		c->h_state_cap = 0;
		CU(c, cudaHostAlloc((void **)&c->h_state, sizeof(ScanState) * (size_t)nvar, cudaHostAllocDefault));
		c->h_state_cap = (size_t)nvar;
	}
	ScanState *d_state = (ScanState *)c->state.p;
	for (int v = 0; v < nvar; v++) {
		k2_init_state(c->h_state + v, n, rzip_level, cb, nvar > 1 ? v : victim_round, rec_cap);
		if (const char *fl = getenv("LRZGPU_K2_FLAGS"))
			c->h_state[v].flags = atoi(fl);
	}
	CU(c, cudaMemcpyAsync(d_state, c->h_state, sizeof(ScanState) * (size_t)nvar, cudaMemcpyHostToDevice, c->sA));
	CU(c, cudaMemsetAsync(c->tab.p, 0, tab_bytes * (size_t)nvar, c->sA)); // src/rzip.c:599-600
	CU(c, cudaEventRecord(c->evInit, c->sA));
	CU(c, cudaStreamWaitEvent(c->sB, c->evInit, 0));
	CU(c, cudaStreamWaitEvent(c->sC, c->evInit, 0));
	if (crc32_launch(d_chunk, n, (uint32_t *)c->crc.p, c->sms, c->sC))
		return fail(c, LRZGPU_ECUDA, "crc32 launch: %s", cudaGetErrorString(cudaGetLastError()));
	CU(c, cudaEventRecord(c->evCrc, c->sC));
	c->launches += 1;

	const int64_t nseg = (n + seg - 1) / seg;
	const bool snap = progress && nvar == 1 && nseg > 1;
	if (snap) { // one pinned snapshot of the scan state and one event per segment
		if ((size_t)nseg > c->h_snap_cap) {
			if (c->h_snap)
				cudaFreeHost(c->h_snap);
			c->h_snap = nullptr;
			c->h_snap_cap = 0;
			CU(c, cudaHostAlloc((void **)&c->h_snap, sizeof(ScanState) * (size_t)nseg, cudaHostAllocDefault));
			c->h_snap_cap = (size_t)nseg;
		}
		while (c->evSeg.size() < (size_t)nseg) {
			cudaEvent_t e = nullptr;
			CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			c->evSeg.push_back(e);
		}
	}
	for (int64_t i = 0; i < nseg; i++) {
		const int b = (int)(i & 1);
		const int64_t lo = i * seg, hi = (lo + seg < n) ? lo + seg : n;
		if (i >= 2)
			CU(c, cudaStreamWaitEvent(c->sB, c->evK2[b], 0)); // cand[b] consumed, newer mask visible
		if (k1_launch(d_chunk, n, lo, hi, 0, d_state, nvar, (Cand *)c->cand[b].p, (uint32_t *)c->tc[b].p, c->sms, c->sB))
			return fail(c, LRZGPU_ECUDA, "k1 launch: %s", cudaGetErrorString(cudaGetLastError()));
		CU(c, cudaEventRecord(c->evK1[b], c->sB));
		CU(c, cudaStreamWaitEvent(c->sA, c->evK1[b], 0));
		if (k2_launch(d_chunk, d_state, (HEntry *)c->tab.p, (const Cand *)c->cand[b].p, (const uint32_t *)c->tc[b].p, lo,
			      hi, (MatchRec *)c->recs.p, i == nseg - 1, nvar, (int64_t)(tab_bytes / sizeof(HEntry)), rec_cap, c->sA))
			return fail(c, LRZGPU_ECUDA, "k2 launch: %s", cudaGetErrorString(cudaGetLastError()));
		CU(c, cudaEventRecord(c->evK2[b], c->sA));
		c->launches += 2;
		if (snap) {
			CU(c, cudaMemcpyAsync(c->h_snap + i, d_state, sizeof(ScanState), cudaMemcpyDeviceToHost, c->sA));
			CU(c, cudaEventRecord(c->evSeg[(size_t)i], c->sA));
		}
	}
	if (snap) {
		for (int64_t i = 0; i + 1 < nseg; i++) { // the last segment's state is handled by the caller as the final one
			CU(c, cudaEventSynchronize(c->evSeg[(size_t)i]));
			if (c->h_snap[i].status != kStatusRunning)
				break;
			const int prc = (*progress)(c->h_snap[i]);
			if (prc)
				return prc;
		}
	}
	CU(c, cudaMemcpyAsync(c->h_state, d_state, sizeof(ScanState) * (size_t)nvar, cudaMemcpyDeviceToHost, c->sA));
	CU(c, cudaStreamSynchronize(c->sA));
	CU(c, cudaStreamSynchronize(c->sB));
	res.assign((size_t)nvar, ChunkResult());
	for (int v = 0; v < nvar; v++) {
		ChunkResult &r = res[(size_t)v];
		r.st = c->h_state[v];
		if (getenv("LRZGPU_DEBUG")) {
			fprintf(stderr, "[lrzgpu] commit v%d: n=%lld lookups=%lld", v, (long long)n, (long long)r.st.st_lookups);
			for (int i = 0; i < 16; i++)
				fprintf(stderr, " d%d=%lld", i, (long long)r.st.dbg[i]);
			fprintf(stderr, "\n");
		}
		if (r.st.status != kStatusChunkDone)
			return fail(c, LRZGPU_EINTERNAL, "rzip commit ended with status %d at position %lld", r.st.status,
				    (long long)r.st.scan_pos);
		r.s0_len = r.st.s0_len;
		r.s1_len = r.st.s1_len;
		r.n_rec = r.st.n_rec;
		r.rec_base = (int64_t)v * rec_cap;
	}
	if (stats)
		stats->ms_rzip += now_ms() - t0;
	return LRZGPU_OK;
}

// K4: stream 0 / stream 1 of one scanned variant into c->s0 / c->s1.
int rzip_emit_device(lrzgpu_ctx *c, const uint8_t *d_chunk, int cb, const ChunkResult &res, lrzgpu_stats *stats,
		     int64_t s1_from = 0)
{
	const double t1 = now_ms();
	const MatchRec *recs = (const MatchRec *)c->recs.p + res.rec_base;
	CU(c, c->s0.ensure((size_t)res.s0_len + 64));
	if (!s1_from) // a pipelined chunk sized the buffer before the scan and already holds [0, s1_from)
		CU(c, c->s1.ensure((size_t)res.s1_len + 64));
	CU(c, cudaStreamWaitEvent(c->sA, c->evCrc, 0));
	if (k4_headers_launch(recs, res.n_rec, cb, (const uint32_t *)c->crc.p, (uint8_t *)c->s0.p, c->sA) ||
	    k4_literals_launch(d_chunk, recs, res.n_rec, s1_from, res.s1_len, (uint8_t *)c->s1.p, c->sms, c->sA))
		return fail(c, LRZGPU_ECUDA, "k4 launch: %s", cudaGetErrorString(cudaGetLastError()));
	c->launches += 2;
	uint32_t crc_acc = 0;
	CU(c, cudaMemcpyAsync(&crc_acc, c->crc.p, 4, cudaMemcpyDeviceToHost, c->sA));
	CU(c, cudaStreamSynchronize(c->sA));
	account_rzip(stats, res);
	if (stats) {
		stats->crc32 = crc_acc ^ 0xffffffffu;
		stats->ms_emit += now_ms() - t1;
	}
	return LRZGPU_OK;
}

// rzip of one chunk from a known victim_round: scan + emit.
int rzip_chunk_device(lrzgpu_ctx *c, const uint8_t *d_chunk, int64_t n, int rzip_level, int cb, int64_t victim_round,
		      ChunkResult &res, lrzgpu_stats *stats)
{
	std::vector<ChunkResult> all;
	int rc = rzip_scan_device(c, d_chunk, n, rzip_level, cb, victim_round, 1, all, stats);
	if (rc)
		return rc;
	res = all[0];
	return rzip_emit_device(c, d_chunk, cb, res, stats);
}

struct OutBuf {
	uint8_t *p = nullptr;
	int64_t len = 0, cap = 0;
	int reserve(int64_t extra)
	{
		if (len + extra <= cap)
			return 0;
		int64_t nc = cap ? cap : 65536;
		while (nc < len + extra)
			nc += nc / 2 + 4096;
		uint8_t *np = (uint8_t *)realloc(p, (size_t)nc);
		if (!np)
			return -1;
		p = np;
		cap = nc;
		return 0;
	}
};

// Global flush order of the chunk's stream blocks (src/stream.c:2198-2216, 2253-2259) as jobs over c->s0 / c->s1.
int plan_chunk_blocks(lrzgpu_ctx *c, const lrzgpu_sizing_t &sz, int cb, const ChunkResult &res, std::vector<BlockJob> &jobs)
{
	const int64_t nb0 = res.s0_len / sz.bufsize;
	std::vector<int64_t> w1((size_t)nb0 + 1, 0);
	if (nb0 > 0) {
		CU(c, c->w1.ensure((size_t)nb0 * 8));
		if (k4_flush_order_launch((const MatchRec *)c->recs.p + res.rec_base, res.n_rec, cb, sz.bufsize, nb0, (int64_t *)c->w1.p, c->sA))
			return fail(c, LRZGPU_ECUDA, "flush order launch failed");
		c->launches += 1;
		CU(c, cudaMemcpyAsync(w1.data(), c->w1.p, (size_t)nb0 * 8, cudaMemcpyDeviceToHost, c->sA));
		CU(c, cudaStreamSynchronize(c->sA));
	}
	std::vector<BlockPlan> plan;
	plan_blocks(res.s0_len, res.s1_len, sz.bufsize, w1.data(), plan);
	jobs.assign(plan.size(), BlockJob());
	for (size_t i = 0; i < plan.size(); i++) {
		jobs[i].d_src = (const uint8_t *)(plan[i].stream ? c->s1.p : c->s0.p) + plan[i].off;
		jobs[i].u_len = plan[i].u_len;
		jobs[i].stream = plan[i].stream;
		jobs[i].c_type = kCtypeNone;
		jobs[i].c_len = plan[i].u_len;
		jobs[i].d_payload = jobs[i].d_src;
	}
	return LRZGPU_OK;
}

// The chunk's blob appended to `out` (src/stream.c:1722-1821: chunk preamble, two initial stream headers, blocks
// in flush order with next_head patching); payloads come from the device.
int frame_chunk(lrzgpu_ctx *c, const lrzgpu_params &p, int64_t n, int eof, int cb, const std::vector<BlockJob> &jobs,
		OutBuf &out, lrzgpu_stats *stats)
{
	const int64_t hdr = 1 + 3 * cb;
	int64_t total = 2 + cb + 2 * hdr;
	for (auto &j : jobs)
		total += hdr + j.c_len;
	if (out.reserve(total))
		return fail(c, LRZGPU_ENOMEM, "out of host memory for %lld byte blob", (long long)total);
	uint8_t *w = out.p + out.len;
	const int64_t size_field = n < p.page_size ? p.page_size : n; // src/stream.c:1150-1152
	*w++ = (uint8_t)cb;
	*w++ = (uint8_t)eof;
	put_le(w, size_field, cb);
	w += cb;
	uint8_t *initial = w;
	int64_t cur_pos = 0, last_head[2];
	for (int s = 0; s < 2; s++) {
		last_head[s] = cur_pos + 1 + 2 * cb;
		*w++ = kCtypeNone;
		memset(w, 0, (size_t)(3 * cb));
		w += 3 * cb;
		cur_pos += hdr;
	}
	for (auto &j : jobs) {
		put_le(initial + last_head[j.stream], cur_pos, cb); // patch the previous header's next_head
		last_head[j.stream] = cur_pos + 1 + 2 * cb;
		*w++ = (uint8_t)j.c_type;
		put_le(w, j.c_len, cb);
		put_le(w + cb, j.u_len, cb);
		put_le(w + 2 * cb, 0, cb);
		w += 3 * cb;
		if (j.c_len)
			CU(c, cudaMemcpyAsync(w, j.d_payload, (size_t)j.c_len, cudaMemcpyDeviceToHost, c->sA));
		w += j.c_len;
		cur_pos += hdr + j.c_len;
		if (stats) {
			stats->blocks++;
			if (j.c_type == kCtypeNone)
				stats->blocks_stored++;
		}
	}
	CU(c, cudaStreamSynchronize(c->sA));
	out.len += total;
	return LRZGPU_OK;
}

// The pre-compression filter of the stream-1 blocks that make up c->s1[from, to), in place (src/stream.c:1587-1628: the
// first thing a compthread does to its block, whatever the backend).  c->fside must be sized already (Delta).
int filter_stream1(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t from, int64_t to, cudaStream_t st)
{
	if (!p.filter || to <= from)
		return LRZGPU_OK;
	if (filter_blocks_launch(p.filter, p.delta, (uint8_t *)c->s1.p, from, to, sz.bufsize, (uint8_t *)c->fside.p, st, &c->launches))
		return fail(c, LRZGPU_ECUDA, "filter launch failed: %s", cudaGetErrorString(cudaGetLastError()));
	CU(c, cudaStreamSynchronize(st));
	return LRZGPU_OK;
}

// Second half of a chunk: stream blocks -> backend -> framed blob, from the streams rzip left in c->s0 / c->s1.
int finish_chunk_device(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t n, int eof, int cb,
			const ChunkResult &res, OutBuf &out, lrzgpu_stats *stats)
{
	const double t0 = now_ms();
	std::vector<BlockJob> jobs;
	CU(c, c->fside.ensure(filter_side_bytes(p.filter, res.s1_len, sz.bufsize) + 16));
	int rc = filter_stream1(c, p, sz, 0, res.s1_len, c->sA);
	if (rc)
		return rc;
	rc = plan_chunk_blocks(c, sz, cb, res, jobs);
	if (rc)
		return rc;
	if (p.backend != LRZGPU_BACKEND_NONE) {
		rc = backend_encode_blocks(c->backend, p, sz, jobs, c->sms, c->sA, &c->launches, c->err, sizeof(c->err));
		if (rc)
			return rc;
	}
	const double t1 = now_ms();
	rc = frame_chunk(c, p, n, eof, cb, jobs, out, stats);
	if (rc)
		return rc;
	if (stats) {
		stats->ms_backend += t1 - t0;
		stats->ms_d2h += now_ms() - t1;
	}
	return LRZGPU_OK;
}

// One chunk with the LZMA backend running UNDER the rzip stage: every stream-1 block goes to the backend the
// moment the scan has produced its last byte (the reference does the same: write_stream() flushes a full buffer to
// a compthread, src/stream.c:2198-2216, 1836-1875), so that at the end of the scan only the blocks that were
// still open -- the last full one, the two tails and stream 0 -- remain to be encoded.
int pipelined_scan(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, const uint8_t *d_chunk, int64_t n, int cb,
		   int64_t *victim_round, ChunkResult &res_out, int64_t &blk1_out, lrzgpu_stats *stats)
{
	const int rzl = p.rzip_level ? p.rzip_level : p.level;
	const int64_t bs = sz.bufsize;
	// every device buffer is sized before the first block goes to the backend: cudaFree / cudaMalloc wait for the
	// whole device, i.e. for block encoders that run for tens of seconds
	const int64_t rec_max = n / kMinMatch + 8, pieces_max = n / 0xFFFF + rec_max;
	const int64_t s0_max = (6 + cb) * pieces_max + 64;
	CU(c, c->s1.ensure((size_t)n + 64));
	CU(c, c->s0.ensure((size_t)s0_max));
	CU(c, c->w1.ensure((size_t)(s0_max / bs + 2) * 8));
	CU(c, c->fside.ensure(filter_side_bytes(p.filter, n, bs) + 16));
	const int64_t max_blocks = n / bs + s0_max / bs + 4;
	int rc = backend_async_begin(c->backend, p, sz, bs < n ? bs : n, max_blocks, n + s0_max + (1 << 20), c->err, sizeof(c->err));
	if (rc)
		return rc;
	int64_t s1_done = 0, blk1 = 0;
	const ScanProgress progress = [&](const ScanState &st) -> int {
		const int64_t full = st.s1_len / bs;
		if (full <= blk1)
			return backend_async_poll(c->backend, &c->launches, c->err, sizeof(c->err));
		// stream-1 bytes [s1_done, st.s1_len) from the records known so far, on a stream of their own (sA holds
		// the rest of the scan)
		if (k4_literals_launch(d_chunk, (const MatchRec *)c->recs.p, st.n_rec, s1_done, st.s1_len, (uint8_t *)c->s1.p, c->sms, c->sE))
			return fail(c, LRZGPU_ECUDA, "k4 launch: %s", cudaGetErrorString(cudaGetLastError()));
		c->launches += 1;
		CU(c, cudaStreamSynchronize(c->sE)); // a millisecond; the backend's streams take no device-side waits
		if (int frc = filter_stream1(c, p, sz, blk1 * bs, full * bs, c->sE)) // the completed blocks, before anyone reads them
			return frc;
		std::vector<BlockJob> jobs((size_t)(full - blk1));
		for (int64_t k = blk1; k < full; k++) {
			BlockJob &j = jobs[(size_t)(k - blk1)];
			j.d_src = (const uint8_t *)c->s1.p + k * bs;
			j.u_len = bs;
			j.stream = 1;
			j.c_type = kCtypeNone;
			j.c_len = bs;
			j.d_payload = j.d_src;
		}
		int r = 0;
		if (bs >= 64) {
			r = backend_async_submit(c->backend, jobs.data(), (int)jobs.size(), nullptr, &c->launches, c->err, sizeof(c->err));
			if (r == 1) { // work space used up: wait for what is in flight (the scan goes on meanwhile), then go on
				r = backend_async_drain(c->backend, &c->launches, c->err, sizeof(c->err));
				if (!r)
					r = backend_async_submit(c->backend, jobs.data(), (int)jobs.size(), nullptr, &c->launches, c->err, sizeof(c->err));
				if (r == 1)
					r = fail(c, LRZGPU_EINTERNAL, "LZMA work space cannot hold %d blocks", (int)jobs.size());
			}
		}
		blk1 = full;
		s1_done = st.s1_len;
		return r;
	};
	std::vector<ChunkResult> all;
	rc = rzip_scan_device(c, d_chunk, n, rzl, cb, *victim_round, 1, all, stats, &progress);
	if (rc)
		return rc;
	res_out = all[0];
	*victim_round = res_out.st.victim_round;
	CU(c, cudaStreamSynchronize(c->sE));
	rc = rzip_emit_device(c, d_chunk, cb, res_out, stats, s1_done);
	if (rc)
		return rc;
	rc = filter_stream1(c, p, sz, blk1 * bs, res_out.s1_len, c->sA); // the blocks the scan did not complete
	if (rc)
		return rc;
	if (bs < 64)
		blk1 = 0; // nothing was handed over
	blk1_out = blk1;
	return LRZGPU_OK;
}

// Second half of a pipelined chunk: the blocks the scan could not hand over, drain, framing.
int pipelined_finish(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, int64_t n, int eof, int cb,
		     const ChunkResult &res, int64_t blk1, OutBuf &out, lrzgpu_stats *stats)
{
	const double t1 = now_ms();
	const int64_t bs = sz.bufsize;
	std::vector<BlockJob> jobs;
	int rc = plan_chunk_blocks(c, sz, cb, res, jobs);
	if (rc)
		return rc;
	// the blocks the scan could not hand over: whatever of stream 1 was not complete at the last look, stream 0
	std::vector<BlockJob> rest;
	std::vector<int> early((size_t)blk1, -1), late;
	for (size_t i = 0; i < jobs.size(); i++) {
		const BlockJob &j = jobs[i];
		const int64_t off = j.d_src - (const uint8_t *)(j.stream ? c->s1.p : c->s0.p);
		if (j.stream == 1 && off / bs < blk1 && j.u_len == bs)
			early[(size_t)(off / bs)] = (int)i;
		else if (j.u_len >= 64) { // src/stream.c:1633
			rest.push_back(j);
			late.push_back((int)i);
		}
	}
	const int n_early = backend_async_count(c->backend);
	if (n_early != blk1)
		return fail(c, LRZGPU_EINTERNAL, "pipelined blocks out of step");
	size_t at = 0;
	while (at < rest.size()) { // as many at a time as the work space takes
		size_t take = rest.size() - at;
		for (;;) {
			rc = backend_async_submit(c->backend, rest.data() + at, (int)take, nullptr, &c->launches, c->err, sizeof(c->err));
			if (rc != 1)
				break;
			if (take > 1)
				take = (take + 1) / 2;
			else if ((rc = backend_async_drain(c->backend, &c->launches, c->err, sizeof(c->err))))
				break;
		}
		if (rc)
			return rc;
		at += take;
	}
	rc = backend_async_drain(c->backend, &c->launches, c->err, sizeof(c->err));
	if (rc)
		return rc;
	auto take_result = [&](int job, int sub) -> int {
		const BlockJob *r = backend_async_result(c->backend, sub);
		if (!r || job < 0)
			return fail(c, LRZGPU_EINTERNAL, "missing result of a pipelined block");
		jobs[(size_t)job].c_type = r->c_type;
		jobs[(size_t)job].c_len = r->c_len;
		jobs[(size_t)job].d_payload = r->d_payload;
		return 0;
	};
	for (int k = 0; k < n_early; k++)
		if ((rc = take_result(early[(size_t)k], k)))
			return rc;
	for (size_t k = 0; k < late.size(); k++)
		if ((rc = take_result(late[k], n_early + (int)k)))
			return rc;
	const double t2 = now_ms();
	rc = frame_chunk(c, p, n, eof, cb, jobs, out, stats);
	if (rc)
		return rc;
	if (stats) {
		stats->ms_backend += t2 - t1;
		stats->ms_d2h += now_ms() - t2;
	}
	return LRZGPU_OK;
}

bool pipelined_ok(const lrzgpu_params &p, int64_t n)
{
	return (p.backend == LRZGPU_BACKEND_LZMA || p.backend == LRZGPU_BACKEND_ZSTD) && n > kSegment && !getenv("LRZGPU_NO_OVERLAP");
}

int compress_chunk_pipelined(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, const uint8_t *d_chunk,
			     int64_t n, int eof, int64_t *victim_round, OutBuf &out, lrzgpu_stats *stats)
{
	const int cb = chunk_bytes_for(n);
	ChunkResult res;
	int64_t blk1 = 0;
	int rc = pipelined_scan(c, p, sz, d_chunk, n, cb, victim_round, res, blk1, stats);
	if (rc)
		return rc;
	return pipelined_finish(c, p, sz, n, eof, cb, res, blk1, out, stats);
}

// One chunk: rzip on the device, then blocks -> backend -> framed blob appended to `out`.
int compress_chunk_device(lrzgpu_ctx *c, const lrzgpu_params &p, const lrzgpu_sizing_t &sz, const uint8_t *d_chunk,
			  int64_t n, int eof, int64_t *victim_round, OutBuf &out, lrzgpu_stats *stats)
{
	if (pipelined_ok(p, n))
		return compress_chunk_pipelined(c, p, sz, d_chunk, n, eof, victim_round, out, stats);
	const int rzl = p.rzip_level ? p.rzip_level : p.level;
	const int cb = chunk_bytes_for(n);
	ChunkResult res;
	int rc = rzip_chunk_device(c, d_chunk, n, rzl, cb, *victim_round, res, stats);
	if (rc)
		return rc;
	*victim_round = res.st.victim_round;
	return finish_chunk_device(c, p, sz, n, eof, cb, res, out, stats);
}

// Upload a host buffer behind a zeroed front pad and in front of a zeroed tail pad.
int upload(lrzgpu_ctx *c, const uint8_t *in, int64_t n, uint8_t **d_data)
{
	CU(c, c->in.ensure((size_t)(kFrontPad + n + kInputPad)));
	uint8_t *base = (uint8_t *)c->in.p;
	CU(c, cudaMemsetAsync(base, 0, kFrontPad, c->sA));
	CU(c, cudaMemsetAsync(base + kFrontPad + n, 0, kInputPad, c->sA));
	if (n)
		CU(c, cudaMemcpyAsync(base + kFrontPad, in, (size_t)n, cudaMemcpyHostToDevice, c->sA));
	*d_data = base + kFrontPad;
	return LRZGPU_OK;
}

int check_params(lrzgpu_ctx *c, const lrzgpu_params *p, int64_t n)
{
	if (!c || !p)
		return LRZGPU_EINVAL;
	if (n <= 0)
		return fail(c, LRZGPU_EINVAL, "empty input is not supported");
	return LRZGPU_OK;
}

// Whole file whose bytes are at d_in in HBM; md5 either given or produced by md5_thread.
int compress_resident(lrzgpu_ctx *c, const lrzgpu_params &p, const uint8_t *d_in, int64_t n, OutBuf &out,
		      lrzgpu_sizing_t &sz, lrzgpu_stats *stats)
{
	int rc = compute_sizing(p, n, sz);
	if (rc)
		return fail(c, rc, "unsupported parameters");
	if (out.reserve(21))
		return fail(c, LRZGPU_ENOMEM, "out of host memory");
	memset(out.p, 0, 21);
	out.len = 21;
	int64_t left = n, victim_round = 0;
	while (left > 0) { // src/rzip.c:1041
		const int64_t offset = n - left;
		const int64_t chunk = sz.max_chunk < left ? sz.max_chunk : left;
		rc = compress_chunk_device(c, p, sz, d_in + offset, chunk, chunk == left, &victim_round, out, stats);
		if (rc)
			return rc;
		left -= chunk;
	}
	return LRZGPU_OK;
}

} // namespace

extern "C" {

const char *lrzgpu_version(void) { return "lrzgpu 0.1 (lrzip-next 0.14 archive format, sm_100a)"; }

int lrzgpu_create(int device, lrzgpu_ctx **out)
{
	if (!out)
		return LRZGPU_EINVAL;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
		return LRZGPU_ENODEV; // no CPU fallback
	if (cudaSetDevice(device) != cudaSuccess)
		return LRZGPU_ECUDA;
	lrzgpu_ctx *c = new (std::nothrow) lrzgpu_ctx();
	if (!c)
		return LRZGPU_ENOMEM;
	c->device = device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
		c->sms = prop.multiProcessorCount;
	// the scan's streams outrank the backend's: while block encoders and match-finder kernels of earlier blocks
	// fill the GPU, the next segment's tag scan and commit kernel must not queue behind their pending CTAs
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	bool ok = cudaStreamCreateWithPriority(&c->sA, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
		  cudaStreamCreateWithPriority(&c->sB, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
		  cudaStreamCreateWithFlags(&c->sC, cudaStreamNonBlocking) == cudaSuccess &&
		  cudaStreamCreateWithFlags(&c->sD, cudaStreamNonBlocking) == cudaSuccess &&
		  cudaStreamCreateWithPriority(&c->sE, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
	for (int i = 0; i < 2 && ok; i++)
		ok = cudaEventCreateWithFlags(&c->evK1[i], cudaEventDisableTiming) == cudaSuccess &&
		     cudaEventCreateWithFlags(&c->evK2[i], cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->evLit, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&c->evInit, cudaEventDisableTiming) == cudaSuccess &&
	     cudaEventCreateWithFlags(&c->evCrc, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaHostAlloc((void **)&c->h_state, sizeof(ScanState), cudaHostAllocDefault) == cudaSuccess;
	if (ok)
		c->h_state_cap = 1;
	ok = ok && k1_init_tables() == 0 && k4_init_tables() == 0;
	ok = ok && k1_preload() == 0 && k2_preload() == 0 && k4_preload() == 0 && backend_preload() == 0 && unrzip_preload() == 0;
	ok = ok && filter_preload() == 0;
	if (ok) {
		c->backend = backend_create();
		ok = c->backend != nullptr;
	}
	if (!ok) {
		lrzgpu_destroy(c);
		return LRZGPU_ECUDA;
	}
	*out = c;
	return LRZGPU_OK;
}

void lrzgpu_destroy(lrzgpu_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	if (c->backend)
		backend_destroy(c->backend);
	DevBuf *bufs[] = { &c->fside, &c->in, &c->tab, &c->state, &c->cand[0], &c->cand[1], &c->tc[0], &c->tc[1], &c->recs, &c->s0, &c->s1,
			   &c->crc, &c->w1 };
	for (DevBuf *b : bufs)
		b->release();
	if (c->h_state)
		cudaFreeHost(c->h_state);
	if (c->h_snap)
		cudaFreeHost(c->h_snap);
	for (cudaEvent_t e : c->evSeg)
		cudaEventDestroy(e);
	if (c->evLit)
		cudaEventDestroy(c->evLit);
	for (int i = 0; i < 2; i++) {
		if (c->h_pin[i])
			cudaFreeHost(c->h_pin[i]);
		if (c->evK1[i])
			cudaEventDestroy(c->evK1[i]);
		if (c->evK2[i])
			cudaEventDestroy(c->evK2[i]);
	}
	if (c->evInit)
		cudaEventDestroy(c->evInit);
	if (c->evCrc)
		cudaEventDestroy(c->evCrc);
	cudaStream_t ss[] = { c->sA, c->sB, c->sC, c->sD, c->sE };
	for (cudaStream_t s : ss)
		if (s)
			cudaStreamDestroy(s);
	delete c;
}

const char *lrzgpu_last_error(const lrzgpu_ctx *c) { return c ? c->err : "no context"; }
void lrzgpu_free(void *p) { free(p); }
int lrzgpu_sm_count(const lrzgpu_ctx *c) { return c ? c->sms : 0; }

int lrzgpu_sizing(const lrzgpu_params *p, int64_t st_size, lrzgpu_sizing_t *out)
{
	if (!p || !out)
		return LRZGPU_EINVAL;
	return compute_sizing(*p, st_size, *out);
}

int lrzgpu_compress(lrzgpu_ctx *c, const lrzgpu_params *p, const uint8_t *in, int64_t n, uint8_t **out,
		    int64_t *out_len, lrzgpu_stats *stats)
{
	int rc = check_params(c, p, n);
	if (rc)
		return rc;
	if (!in || !out || !out_len)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	uint8_t md5[16];
	double md5_ms = 0;
	std::thread hasher([&] { // whole-file MD5 in file order, overlapped with the GPU (src/rzip.c:1195-1218)
		const double a = now_ms();
		Md5 m;
		for (int64_t o = 0; o < n; o += (64 << 20))
			m.update(in + o, (size_t)((n - o < (64 << 20)) ? n - o : (64 << 20)));
		m.final(md5);
		md5_ms = now_ms() - a;
	});
	uint8_t *d_in = nullptr;
	OutBuf ob;
	lrzgpu_sizing_t sz;
	rc = upload(c, in, n, &d_in);
	if (!rc) {
		cudaStreamSynchronize(c->sA);
		if (stats)
			stats->ms_h2d = now_ms() - t0;
		rc = compress_resident(c, *p, d_in, n, ob, sz, stats);
	}
	hasher.join();
	if (!rc && ob.reserve(16))
		rc = fail(c, LRZGPU_ENOMEM, "out of host memory");
	if (rc) {
		free(ob.p);
		return rc;
	}
	memcpy(ob.p + ob.len, md5, 16);
	ob.len += 16;
	make_magic(ob.p, *p, sz, n);
	*out = ob.p;
	*out_len = ob.len;
	if (stats) {
		stats->ms_md5 = md5_ms;
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_compress_device(lrzgpu_ctx *c, const lrzgpu_params *p, const void *d_in_v, int64_t n,
			   const uint8_t *md5_or_null, uint8_t **out, int64_t *out_len, lrzgpu_stats *stats)
{
	int rc = check_params(c, p, n);
	if (rc)
		return rc;
	if (!d_in_v || !out || !out_len || ((uintptr_t)d_in_v & 15))
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	const uint8_t *d_in = (const uint8_t *)d_in_v;
	uint8_t md5[16];
	double md5_ms = 0;
	int md5_rc = 0;
	std::thread hasher;
	if (md5_or_null)
		memcpy(md5, md5_or_null, 16);
	else {
		const size_t slice = 32u << 20;
		if (c->h_pin_cap < slice) {
			for (int i = 0; i < 2; i++) {
				if (c->h_pin[i])
					cudaFreeHost(c->h_pin[i]);
				c->h_pin[i] = nullptr;
				c->h_pin_cap = 0;
				if (cudaHostAlloc((void **)&c->h_pin[i], slice, cudaHostAllocDefault) != cudaSuccess) {
					c->h_pin[i] = nullptr;
					return fail(c, LRZGPU_ENOMEM, "pinned staging allocation failed");
				}
			}
			c->h_pin_cap = slice;
		}
		hasher = std::thread([&, slice] { // stream the input back to the host for the MD5, double buffered
			const double a = now_ms();
			cudaSetDevice(c->device);
			Md5 m;
			cudaEvent_t ev[2];
			cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
			cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
			const int64_t nsl = (n + (int64_t)slice - 1) / (int64_t)slice;
			auto issue = [&](int64_t i) {
				const int64_t o = i * (int64_t)slice;
				const size_t len = (size_t)((n - o < (int64_t)slice) ? n - o : (int64_t)slice);
				if (cudaMemcpyAsync(c->h_pin[i & 1], d_in + o, len, cudaMemcpyDeviceToHost, c->sD) != cudaSuccess)
					md5_rc = -1;
				cudaEventRecord(ev[i & 1], c->sD);
			};
			issue(0);
			for (int64_t i = 0; i < nsl; i++) {
				if (i + 1 < nsl)
					issue(i + 1);
				if (cudaEventSynchronize(ev[i & 1]) != cudaSuccess)
					md5_rc = -1;
				const int64_t o = i * (int64_t)slice;
				m.update(c->h_pin[i & 1], (size_t)((n - o < (int64_t)slice) ? n - o : (int64_t)slice));
			}
			m.final(md5);
			cudaEventDestroy(ev[0]);
			cudaEventDestroy(ev[1]);
			md5_ms = now_ms() - a;
		});
	}
	OutBuf ob;
	lrzgpu_sizing_t sz;
	rc = compress_resident(c, *p, d_in, n, ob, sz, stats);
	if (hasher.joinable())
		hasher.join();
	if (!rc && md5_rc)
		rc = fail(c, LRZGPU_ECUDA, "device->host MD5 stream failed");
	if (!rc && ob.reserve(16))
		rc = fail(c, LRZGPU_ENOMEM, "out of host memory");
	if (rc) {
		free(ob.p);
		return rc;
	}
	memcpy(ob.p + ob.len, md5, 16);
	ob.len += 16;
	make_magic(ob.p, *p, sz, n);
	*out = ob.p;
	*out_len = ob.len;
	if (stats) {
		stats->ms_md5 = md5_ms;
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_compress_file(lrzgpu_ctx *c, const lrzgpu_params *p, const char *in_path, const char *out_path,
			 lrzgpu_stats *stats)
{
	if (!c || !p || !in_path || !out_path)
		return LRZGPU_EINVAL;
	FILE *f = fopen(in_path, "rb");
	if (!f)
		return fail(c, LRZGPU_EIO, "cannot open %s: %s", in_path, strerror(errno));
	fseeko(f, 0, SEEK_END);
	const int64_t n = (int64_t)ftello(f);
	fseeko(f, 0, SEEK_SET);
	uint8_t *buf = nullptr;
	if (n > 0 && cudaHostAlloc((void **)&buf, (size_t)n, cudaHostAllocDefault) != cudaSuccess) {
		fclose(f);
		return fail(c, LRZGPU_ENOMEM, "cannot allocate %lld bytes of pinned memory", (long long)n);
	}
	const bool read_ok = n > 0 && fread(buf, 1, (size_t)n, f) == (size_t)n;
	fclose(f);
	if (!read_ok) {
		if (buf)
			cudaFreeHost(buf);
		return fail(c, n > 0 ? LRZGPU_EIO : LRZGPU_EINVAL, "cannot read %s", in_path);
	}
	uint8_t *out = nullptr;
	int64_t out_len = 0;
	int rc = lrzgpu_compress(c, p, buf, n, &out, &out_len, stats);
	cudaFreeHost(buf);
	if (rc)
		return rc;
	FILE *g = fopen(out_path, "wb");
	if (!g || fwrite(out, 1, (size_t)out_len, g) != (size_t)out_len) {
		if (g)
			fclose(g);
		free(out);
		return fail(c, LRZGPU_EIO, "cannot write %s: %s", out_path, strerror(errno));
	}
	fclose(g);
	free(out);
	return LRZGPU_OK;
}

int lrzgpu_compress_chunk(lrzgpu_ctx *c, const lrzgpu_params *p, const lrzgpu_sizing_t *sz, const uint8_t *in,
			  int64_t n, int eof, int64_t *victim_round, uint8_t **blob, int64_t *blob_len,
			  lrzgpu_stats *stats)
{
	int rc = check_params(c, p, n);
	if (rc)
		return rc;
	if (!sz || !in || !victim_round || !blob || !blob_len)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	uint8_t *d_in = nullptr;
	rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	CU(c, cudaStreamSynchronize(c->sA));
	if (stats)
		stats->ms_h2d = now_ms() - t0;
	OutBuf ob;
	rc = compress_chunk_device(c, *p, *sz, d_in, n, eof, victim_round, ob, stats);
	if (rc) {
		free(ob.p);
		return rc;
	}
	*blob = ob.p;
	*blob_len = ob.len;
	if (stats) {
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_chunk_begin(lrzgpu_ctx *c, const lrzgpu_params *p, const lrzgpu_sizing_t *sz, const uint8_t *in, int64_t n,
		       int eof, int64_t *victim_round, lrzgpu_stats *stats)
{
	int rc = check_params(c, p, n);
	if (rc)
		return rc;
	if (!sz || !in || !victim_round)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	c->pending.valid = false;
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	uint8_t *d_in = nullptr;
	rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	CU(c, cudaStreamSynchronize(c->sA));
	if (stats)
		stats->ms_h2d = now_ms() - t0;
	const int rzl = p->rzip_level ? p->rzip_level : p->level;
	const int cb = chunk_bytes_for(n);
	ChunkResult res;
	int64_t blk1 = 0;
	const bool pipe = pipelined_ok(*p, n);
	if (pipe)
		rc = pipelined_scan(c, *p, *sz, d_in, n, cb, victim_round, res, blk1, stats);
	else
		rc = rzip_chunk_device(c, d_in, n, rzl, cb, *victim_round, res, stats);
	if (rc)
		return rc;
	*victim_round = res.st.victim_round;
	lrzgpu_ctx::Pending &pd = c->pending;
	pd.valid = true;
	pd.selected = true;
	pd.pipelined = pipe;
	pd.blk1 = blk1;
	pd.p = *p;
	pd.sz = *sz;
	pd.n = n;
	pd.eof = eof;
	pd.cb = cb;
	pd.d_chunk = d_in;
	pd.s0_len = res.s0_len;
	pd.s1_len = res.s1_len;
	pd.n_rec = res.n_rec;
	pd.rec_base = res.rec_base;
	if (stats) {
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_victim_values(const lrzgpu_params *p)
{
	if (!p)
		return LRZGPU_EINVAL;
	const int rzl = p->rzip_level ? p->rzip_level : p->level;
	if (rzl < 0 || rzl > 9)
		return LRZGPU_EINVAL;
	return (int)kLevels[rzl].max_chain_len;
}

int lrzgpu_chunk_begin_all(lrzgpu_ctx *c, const lrzgpu_params *p, const lrzgpu_sizing_t *sz, const uint8_t *in, int64_t n,
			   int eof, int64_t *victim_out, int nvalues, lrzgpu_stats *stats)
{
	int rc = check_params(c, p, n);
	if (rc)
		return rc;
	if (!sz || !in || !victim_out)
		return LRZGPU_EINVAL;
	const int nvar = lrzgpu_victim_values(p);
	if (nvalues != nvar)
		return fail(c, LRZGPU_EINVAL, "victim_out must hold lrzgpu_victim_values() = %d entries", nvar);
	cudaSetDevice(c->device);
	c->pending.valid = false;
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	{ // nvar tables and record arrays must fit next to the window and the backend's work space
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		const int rzl0 = p->rzip_level ? p->rzip_level : p->level;
		const size_t need = (size_t)nvar * (((size_t)kLevels[rzl0].mb_used << 20) + (size_t)(n / kMinMatch + 8) * sizeof(MatchRec));
		const size_t have = free_b + c->tab.cap + c->recs.cap;
		if (need + (size_t)n * 4 + (8ull << 30) > have)
			return fail(c, LRZGPU_ENOMEM, "%d variants of a %lld byte window need %zu MiB of device memory", nvar,
				    (long long)n, need >> 20);
	}
	uint8_t *d_in = nullptr;
	rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	CU(c, cudaStreamSynchronize(c->sA));
	if (stats)
		stats->ms_h2d = now_ms() - t0;
	const int rzl = p->rzip_level ? p->rzip_level : p->level;
	const int cb = chunk_bytes_for(n);
	std::vector<ChunkResult> res;
	rc = rzip_scan_device(c, d_in, n, rzl, cb, 0, nvar, res, stats);
	if (rc)
		return rc;
	lrzgpu_ctx::Pending &pd = c->pending;
	pd.valid = true;
	pd.selected = false;
	pd.pipelined = false;
	pd.p = *p;
	pd.sz = *sz;
	pd.n = n;
	pd.eof = eof;
	pd.cb = cb;
	pd.d_chunk = d_in;
	pd.v_s0.clear();
	pd.v_s1.clear();
	pd.v_nrec.clear();
	pd.v_base.clear();
	pd.v_out.clear();
	pd.v_st.clear();
	for (int v = 0; v < nvar; v++) {
		pd.v_s0.push_back(res[(size_t)v].s0_len);
		pd.v_s1.push_back(res[(size_t)v].s1_len);
		pd.v_nrec.push_back(res[(size_t)v].n_rec);
		pd.v_base.push_back(res[(size_t)v].rec_base);
		pd.v_out.push_back(res[(size_t)v].st.victim_round);
		pd.v_st.push_back(res[(size_t)v].st);
		victim_out[v] = res[(size_t)v].st.victim_round;
	}
	if (stats) {
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_chunk_select(lrzgpu_ctx *c, int64_t victim_in, lrzgpu_stats *stats)
{
	if (!c)
		return LRZGPU_EINVAL;
	lrzgpu_ctx::Pending &pd = c->pending;
	if (!pd.valid || pd.selected)
		return fail(c, LRZGPU_EINVAL, "lrzgpu_chunk_select without a window from lrzgpu_chunk_begin_all");
	if (victim_in < 0 || victim_in >= (int64_t)pd.v_out.size())
		return fail(c, LRZGPU_EINVAL, "victim_round %lld out of range", (long long)victim_in);
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	ChunkResult res;
	const size_t v = (size_t)victim_in;
	res.s0_len = pd.v_s0[v];
	res.s1_len = pd.v_s1[v];
	res.n_rec = pd.v_nrec[v];
	res.rec_base = pd.v_base[v];
	res.st = pd.v_st[v];
	int rc = rzip_emit_device(c, pd.d_chunk, pd.cb, res, stats);
	if (rc)
		return rc;
	pd.selected = true;
	pd.s0_len = res.s0_len;
	pd.s1_len = res.s1_len;
	pd.n_rec = res.n_rec;
	pd.rec_base = res.rec_base;
	if (stats) {
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_chunk_finish(lrzgpu_ctx *c, uint8_t **blob, int64_t *blob_len, lrzgpu_stats *stats)
{
	if (!c || !blob || !blob_len)
		return LRZGPU_EINVAL;
	if (!c->pending.valid || !c->pending.selected)
		return fail(c, LRZGPU_EINVAL, "lrzgpu_chunk_finish without a window from lrzgpu_chunk_begin / lrzgpu_chunk_select");
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	const double t0 = now_ms();
	ChunkResult res;
	res.s0_len = c->pending.s0_len;
	res.s1_len = c->pending.s1_len;
	res.n_rec = c->pending.n_rec;
	res.rec_base = c->pending.rec_base;
	c->pending.valid = false;
	OutBuf ob;
	const int rc = c->pending.pipelined
			       ? pipelined_finish(c, c->pending.p, c->pending.sz, c->pending.n, c->pending.eof, c->pending.cb, res,
						  c->pending.blk1, ob, stats)
			       : finish_chunk_device(c, c->pending.p, c->pending.sz, c->pending.n, c->pending.eof, c->pending.cb, res, ob,
						     stats);
	if (rc) {
		free(ob.p);
		return rc;
	}
	*blob = ob.p;
	*blob_len = ob.len;
	if (stats) {
		stats->ms_total = now_ms() - t0;
		stats->kernel_launches = c->launches - launches0;
	}
	return LRZGPU_OK;
}

int lrzgpu_compress_multi(lrzgpu_ctx **ctxs, int nctx, const lrzgpu_params *p, const uint8_t *in, int64_t n, uint8_t **out,
			  int64_t *out_len, lrzgpu_stats *stats)
{
	if (!ctxs || nctx < 1 || !p || !in || n <= 0 || !out || !out_len)
		return LRZGPU_EINVAL;
	for (int i = 0; i < nctx; i++)
		if (!ctxs[i])
			return LRZGPU_EINVAL;
	lrzgpu_ctx *c0 = ctxs[0];
	if (nctx == 1)
		return lrzgpu_compress(c0, p, in, n, out, out_len, stats);
	lrzgpu_sizing_t sz;
	int rc = compute_sizing(*p, n, sz);
	if (rc)
		return fail(c0, rc, "unsupported parameters");
	const double t0 = now_ms();
	struct Win {
		int64_t off = 0, size = 0;
		int eof = 0;
		uint8_t *blob = nullptr;
		int64_t blob_len = 0;
		lrzgpu_stats st;
	};
	std::vector<Win> wins;
	for (int64_t off = 0; off < n; off += sz.max_chunk) { // src/rzip.c:1041: the window loop as a static plan
		Win w;
		w.off = off;
		w.size = sz.max_chunk < n - off ? sz.max_chunk : n - off;
		w.eof = off + w.size == n;
		memset(&w.st, 0, sizeof(w.st));
		wins.push_back(w);
	}
	const size_t nw = wins.size();
	// the true value of the cross-window counter after window i (src/rzip.c:308), handed from owner to owner
	std::vector<int64_t> v_out(nw, -1);
	std::mutex mu;
	std::condition_variable cv;
	std::vector<int> rcs((size_t)nctx, 0);
	auto wait_for = [&](size_t i) -> int64_t { // value after window i; -2 when its owner failed
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&] { return v_out[i] != -1; });
		return v_out[i];
	};
	auto publish = [&](size_t i, int64_t v) {
		{
			std::lock_guard<std::mutex> lk(mu);
			v_out[i] = v;
		}
		cv.notify_all();
	};
	auto worker = [&](int k) {
		lrzgpu_ctx *c = ctxs[k];
		for (size_t i = (size_t)k; i < nw; i += (size_t)nctx) { // window i belongs to context i mod nctx
			Win &w = wins[i];
			lrzgpu_stats a, b, s3;
			memset(&s3, 0, sizeof(s3));
			int r;
			int64_t v = 0;
			std::vector<int64_t> table;
			const int nval = lrzgpu_victim_values(p);
			if (i == 0)
				r = lrzgpu_chunk_begin(c, p, &sz, in + w.off, w.size, w.eof, &v, &a);
			else {
				table.resize((size_t)nval);
				r = lrzgpu_chunk_begin_all(c, p, &sz, in + w.off, w.size, w.eof, table.data(), nval, &a);
				if (r == LRZGPU_ENOMEM) { // the variants do not fit: wait for the predecessor instead
					table.clear();
					v = wait_for(i - 1);
					r = v < 0 ? LRZGPU_EINTERNAL : lrzgpu_chunk_begin(c, p, &sz, in + w.off, w.size, w.eof, &v, &a);
				} else if (!r) {
					const int64_t vin = wait_for(i - 1);
					r = vin < 0 ? LRZGPU_EINTERNAL : lrzgpu_chunk_select(c, vin, &s3);
					v = r ? 0 : table[(size_t)vin];
				}
			}
			publish(i, r ? -2 : v);
			if (!r)
				r = lrzgpu_chunk_finish(c, &w.blob, &w.blob_len, &b);
			if (r) {
				rcs[(size_t)k] = r;
				for (size_t j = i + (size_t)nctx; j < nw; j += (size_t)nctx)
					publish(j, -2); // nobody waits forever for a window that will not be scanned
				return;
			}
			// the window's counters: scan (begin or begin_all + select) and backend
			lrzgpu_stats &d = w.st;
			const lrzgpu_stats *parts[3] = { &a, &s3, &b };
			for (const lrzgpu_stats *q : parts) {
				d.matches += q->matches;
				d.match_bytes += q->match_bytes;
				d.literals += q->literals;
				d.literal_bytes += q->literal_bytes;
				d.tag_hits += q->tag_hits;
				d.tag_misses += q->tag_misses;
				d.inserts += q->inserts;
				d.lookups += q->lookups;
				d.chain_evictions += q->chain_evictions;
				d.sweeps += q->sweeps;
				d.displacements += q->displacements;
				d.chunks += q->chunks;
				d.blocks += q->blocks;
				d.blocks_stored += q->blocks_stored;
				d.stream0_bytes += q->stream0_bytes;
				d.stream1_bytes += q->stream1_bytes;
				d.ms_h2d += q->ms_h2d;
				d.ms_rzip += q->ms_rzip;
				d.ms_emit += q->ms_emit;
				d.ms_backend += q->ms_backend;
				d.ms_d2h += q->ms_d2h;
				d.kernel_launches += q->kernel_launches;
				if (q->crc32)
					d.crc32 = q->crc32;
				if (q->chunks) {
					d.hash_count = q->hash_count;
					d.final_min_mask = q->final_min_mask;
					d.final_tag_mask = q->final_tag_mask;
				}
			}
		}
	};
	uint8_t md5[16];
	double md5_ms = 0;
	std::thread hasher([&] { // whole-file MD5 in file order, overlapped with the GPUs (src/rzip.c:1195-1218)
		const double a = now_ms();
		Md5 m;
		for (int64_t o = 0; o < n; o += (64 << 20))
			m.update(in + o, (size_t)((n - o < (64 << 20)) ? n - o : (64 << 20)));
		m.final(md5);
		md5_ms = now_ms() - a;
	});
	std::vector<std::thread> th;
	for (int k = 0; k < nctx; k++)
		th.emplace_back(worker, k);
	for (auto &t : th)
		t.join();
	hasher.join();
	rc = 0;
	for (int k = 0; k < nctx && !rc; k++)
		if (rcs[(size_t)k]) {
			rc = rcs[(size_t)k];
			if (ctxs[k] != c0)
				snprintf(c0->err, sizeof(c0->err), "context %d: %s", k, ctxs[k]->err);
		}
	OutBuf ob;
	if (!rc) {
		int64_t total = 21 + 16;
		for (const Win &w : wins)
			total += w.blob_len;
		if (ob.reserve(total))
			rc = fail(c0, LRZGPU_ENOMEM, "out of host memory");
	}
	if (!rc) {
		memset(ob.p, 0, 21);
		ob.len = 21;
		for (const Win &w : wins) {
			memcpy(ob.p + ob.len, w.blob, (size_t)w.blob_len);
			ob.len += w.blob_len;
		}
		memcpy(ob.p + ob.len, md5, 16);
		ob.len += 16;
		make_magic(ob.p, *p, sz, n);
		*out = ob.p;
		*out_len = ob.len;
	} else
		free(ob.p);
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		for (const Win &w : wins) {
			stats->matches += w.st.matches;
			stats->match_bytes += w.st.match_bytes;
			stats->literals += w.st.literals;
			stats->literal_bytes += w.st.literal_bytes;
			stats->tag_hits += w.st.tag_hits;
			stats->tag_misses += w.st.tag_misses;
			stats->inserts += w.st.inserts;
			stats->lookups += w.st.lookups;
			stats->chain_evictions += w.st.chain_evictions;
			stats->sweeps += w.st.sweeps;
			stats->displacements += w.st.displacements;
			stats->chunks += w.st.chunks;
			stats->blocks += w.st.blocks;
			stats->blocks_stored += w.st.blocks_stored;
			stats->stream0_bytes += w.st.stream0_bytes;
			stats->stream1_bytes += w.st.stream1_bytes;
			stats->kernel_launches += w.st.kernel_launches;
			stats->ms_h2d += w.st.ms_h2d;
			stats->ms_d2h += w.st.ms_d2h;
			if (w.st.ms_rzip > stats->ms_rzip)
				stats->ms_rzip = w.st.ms_rzip;
			if (w.st.ms_backend > stats->ms_backend)
				stats->ms_backend = w.st.ms_backend;
			stats->crc32 = w.st.crc32;
			stats->hash_count = w.st.hash_count;
			stats->final_min_mask = w.st.final_min_mask;
			stats->final_tag_mask = w.st.final_tag_mask;
		}
		stats->ms_md5 = md5_ms;
		stats->ms_total = now_ms() - t0;
	}
	for (Win &w : wins)
		free(w.blob);
	return rc;
}

// ---- decode: runzip_fd / runzip_chunk, src/runzip.c:261-470, on the device ---------------
namespace {
int64_t get_le(const uint8_t *p, int width)
{
	int64_t v = 0;
	for (int i = 0; i < width; i++)
		v |= (int64_t)p[i] << (8 * i);
	return v;
}
struct ArcBlock {
	int stream, ctype;
	int64_t c_len, u_len, payload; // payload: offset in the archive
};
} // namespace

int lrzgpu_decompress(lrzgpu_ctx *c, const uint8_t *arc, int64_t arc_len, uint8_t **out, int64_t *out_len)
{
	if (!c || !arc || !out || !out_len)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	// magic header, src/lrzip.c:131-208 (written) / 227-330 (read)
	if (arc_len < 21 + 16 || memcmp(arc, "LRZI", 4) || arc[4] != 0)
		return fail(c, LRZGPU_EINVAL, "not an lrzip-next archive");
	if (arc[15] != 0)
		return fail(c, LRZGPU_EUNSUPPORTED, "encrypted archives are not supported");
	// filter byte (src/lrzip.c:323-340): 128 + coded distance = Delta, else the flag; undone per stream-1 block after
	// the block is decompressed (src/stream.c:2082-2130)
	int flt_id = arc[16], flt_delta = 0;
	if (arc[16] & 128) {
		const int i = arc[16] & 127;
		flt_id = LRZGPU_FILTER_DELTA;
		flt_delta = i <= 16 ? i : (i - 15) * 16;
	}
	if (flt_id > LRZGPU_FILTER_RISCV && flt_id != LRZGPU_FILTER_DELTA)
		return fail(c, LRZGPU_EUNSUPPORTED, "filter %d is not supported", flt_id);
	if (arc[14] != 1)
		return fail(c, LRZGPU_EUNSUPPORTED, "only MD5 archives are supported (hash code %d)", arc[14]);
	const int64_t st_size = get_le(arc + 6, 8);
	int64_t pos = 21 + arc[20];
	const int64_t end = arc_len - 16;
	uint8_t *res = (uint8_t *)malloc((size_t)(st_size > 0 ? st_size : 1));
	if (!res)
		return fail(c, LRZGPU_ENOMEM, "out of host memory");
	struct Guard { // every error path below returns through the CU macro or bail(): the buffer goes with it
		uint8_t *p;
		~Guard() { free(p); }
	} guard{ res };
	int64_t done = 0;
	int rc = LRZGPU_OK;
	auto bail = [&](int code, const char *msg) { return fail(c, code, "%s", msg); };
	for (bool last = false; !last;) {
		if (pos + 2 > end)
			return bail(LRZGPU_EINVAL, "truncated archive (chunk header)");
		const int cb = arc[pos], eof = arc[pos + 1];
		if (cb < 1 || cb > 8 || pos + 2 + cb + 2 * (1 + 3 * cb) > end)
			return bail(LRZGPU_EINVAL, "bad chunk header");
		pos += 2 + cb; // the stored chunk size is not needed: the terminator ends the chunk
		const int64_t initial = pos, hdr = 1 + 3 * cb;
		std::vector<ArcBlock> blocks;
		int64_t chunk_end = initial + 2 * hdr, total_u[2] = { 0, 0 };
		for (int s = 0; s < 2; s++) { // follow the stream's chain of block headers (src/stream.c:1883-2016)
			int64_t head = get_le(arc + initial + s * hdr + 1 + 2 * cb, cb);
			while (head) {
				const int64_t at = initial + head;
				if (at < initial || at + hdr > end)
					return bail(LRZGPU_EINVAL, "block header outside the archive");
				ArcBlock b;
				b.stream = s;
				b.ctype = arc[at];
				b.c_len = get_le(arc + at + 1, cb);
				b.u_len = get_le(arc + at + 1 + cb, cb);
				b.payload = at + hdr;
				if (b.c_len < 0 || b.u_len < 0 || b.payload + b.c_len > end)
					return bail(LRZGPU_EINVAL, "block payload outside the archive");
				if (b.ctype != LRZGPU_CTYPE_NONE && b.ctype != LRZGPU_CTYPE_LZMA && b.ctype != LRZGPU_CTYPE_ZSTD)
					return bail(LRZGPU_EUNSUPPORTED, "only stored, LZMA and zstd blocks can be decoded on the device");
				blocks.push_back(b);
				total_u[s] += b.u_len;
				if (b.payload + b.c_len > chunk_end)
					chunk_end = b.payload + b.c_len;
				head = get_le(arc + at + 1 + 2 * cb, cb);
			}
		}
		// stream bytes on the device: stored blocks are copied in place, LZMA blocks decoded there
		CU(c, c->s0.ensure((size_t)total_u[0] + 64));
		CU(c, c->s1.ensure((size_t)total_u[1] + 64));
		int64_t comp_bytes = 0;
		std::vector<LzmaDecJob> jobs, zjobs;
		for (const ArcBlock &b : blocks)
			if (b.ctype != LRZGPU_CTYPE_NONE)
				comp_bytes += (b.c_len + 15) & ~(int64_t)15;
		CU(c, c->in.ensure((size_t)comp_bytes + 64));
		int64_t so[2] = { 0, 0 }, co = 0;
		for (const ArcBlock &b : blocks) {
			uint8_t *dst = (uint8_t *)(b.stream ? c->s1.p : c->s0.p) + so[b.stream];
			if (b.ctype == LRZGPU_CTYPE_NONE) {
				if (b.c_len != b.u_len)
					return bail(LRZGPU_EINVAL, "stored block with c_len != u_len");
				if (b.u_len)
					CU(c, cudaMemcpyAsync(dst, arc + b.payload, (size_t)b.u_len, cudaMemcpyHostToDevice, c->sA));
			} else {
				uint8_t *src = (uint8_t *)c->in.p + co;
				CU(c, cudaMemcpyAsync(src, arc + b.payload, (size_t)b.c_len, cudaMemcpyHostToDevice, c->sA));
				LzmaDecJob j;
				memset(&j, 0, sizeof(j));
				j.src = src;
				j.c_len = b.c_len;
				j.out = dst;
				j.u_len = b.u_len;
				(b.ctype == LRZGPU_CTYPE_LZMA ? jobs : zjobs).push_back(j);
				co += (b.c_len + 15) & ~(int64_t)15;
			}
			so[b.stream] += b.u_len;
		}
		if (!jobs.empty()) {
			const size_t jb = jobs.size() * sizeof(LzmaDecJob);
			CU(c, c->w1.ensure(jb));
			CU(c, c->tab.ensure(lzma_dec_prob_bytes((int)jobs.size())));
			CU(c, cudaMemcpyAsync(c->w1.p, jobs.data(), jb, cudaMemcpyHostToDevice, c->sA));
			if (lzma_dec_launch((LzmaDecJob *)c->w1.p, (int)jobs.size(), c->tab.p, c->sA))
				return bail(LRZGPU_ECUDA, "LZMA decoder launch failed");
			c->launches++;
			CU(c, cudaMemcpyAsync(jobs.data(), c->w1.p, jb, cudaMemcpyDeviceToHost, c->sA));
			CU(c, cudaStreamSynchronize(c->sA));
			for (const LzmaDecJob &j : jobs)
				if (j.status || j.produced != j.u_len)
					return bail(LRZGPU_EINVAL, "corrupt LZMA block");
		}
		if (!zjobs.empty()) { // zstd frames: one thread per frame, frames side by side
			const size_t jb = zjobs.size() * sizeof(LzmaDecJob);
			CU(c, c->w1.ensure(jb));
			CU(c, c->tab.ensure(zstd_dec_work_bytes((int)zjobs.size())));
			CU(c, cudaMemcpyAsync(c->w1.p, zjobs.data(), jb, cudaMemcpyHostToDevice, c->sA));
			if (zstd_dec_launch((LzmaDecJob *)c->w1.p, (int)zjobs.size(), c->tab.p, c->sA))
				return bail(LRZGPU_ECUDA, "zstd decoder launch failed");
			c->launches++;
			CU(c, cudaMemcpyAsync(zjobs.data(), c->w1.p, jb, cudaMemcpyDeviceToHost, c->sA));
			CU(c, cudaStreamSynchronize(c->sA));
			for (const LzmaDecJob &j : zjobs)
				if (j.status || j.produced != j.u_len)
					return bail(LRZGPU_EINVAL, "corrupt zstd block");
		}
		if (flt_id) { // every stream-1 block was filtered on its own, from its position 0: all but the last are equally long
			int64_t bs1 = 0, seen = 0;
			bool uniform = true;
			for (const ArcBlock &b : blocks)
				if (b.stream == 1) {
					if (!bs1)
						bs1 = b.u_len;
					else if (seen % bs1)
						uniform = false; // a short block that is not the last one
					seen += b.u_len;
				}
			if (!uniform)
				return bail(LRZGPU_EUNSUPPORTED, "filtered archive with irregular stream-1 blocks");
			if (bs1 && filter_blocks_launch(flt_id, flt_delta, (uint8_t *)c->s1.p, 0, total_u[1], bs1, nullptr, c->sA, &c->launches, false))
				return bail(LRZGPU_ECUDA, "unfilter launch failed");
		}
		// stream 0 -> records, then the replay into the chunk's bytes
		const int64_t cap = total_u[0] / 3 + 2;
		CU(c, c->recs.ensure((size_t)cap * (sizeof(DecLit) + sizeof(DecMatch)) + sizeof(DecSummary) + 64));
		DecLit *d_lits = (DecLit *)c->recs.p;
		DecMatch *d_matches = (DecMatch *)(d_lits + cap);
		DecSummary *d_sum = (DecSummary *)(d_matches + cap);
		const int64_t room = st_size - done;
		if (unrzip_parse_launch((const uint8_t *)c->s0.p, total_u[0], cb, room, d_lits, d_matches, cap, d_sum, c->sA))
			return bail(LRZGPU_ECUDA, "stream parse launch failed");
		DecSummary sum;
		CU(c, cudaMemcpyAsync(&sum, d_sum, sizeof(sum), cudaMemcpyDeviceToHost, c->sA));
		CU(c, cudaStreamSynchronize(c->sA));
		c->launches++;
		if (sum.status || sum.lit_len != total_u[1] || sum.out_len > room)
			return bail(LRZGPU_EINVAL, "corrupt rzip stream");
		CU(c, c->cand[0].ensure((size_t)sum.out_len + kFrontPad + kInputPad));
		uint8_t *d_out = (uint8_t *)c->cand[0].p + kFrontPad;
		if (unrzip_replay_launch((const uint8_t *)c->s1.p, total_u[1], d_lits, sum.n_lit, d_matches, sum.n_match, d_out, c->sms, c->sA))
			return bail(LRZGPU_ECUDA, "replay launch failed");
		c->launches += 2;
		// the chunk's CRC-32 (src/runzip.c:346-357)
		CU(c, c->crc.ensure(16));
		uint32_t crc_acc = 0;
		if (sum.out_len > 0) {
			if (crc32_launch(d_out, sum.out_len, (uint32_t *)c->crc.p, c->sms, c->sA))
				return bail(LRZGPU_ECUDA, "crc32 launch failed");
			c->launches++;
			CU(c, cudaMemcpyAsync(&crc_acc, c->crc.p, 4, cudaMemcpyDeviceToHost, c->sA));
		}
		if (sum.out_len)
			CU(c, cudaMemcpyAsync(res + done, d_out, (size_t)sum.out_len, cudaMemcpyDeviceToHost, c->sA));
		CU(c, cudaStreamSynchronize(c->sA));
		if (sum.out_len > 0 && (crc_acc ^ 0xffffffffu) != sum.crc)
			return bail(LRZGPU_EINVAL, "chunk CRC mismatch");
		done += sum.out_len;
		pos = chunk_end;
		last = eof != 0;
	}
	if (done != st_size || pos != end)
		return bail(LRZGPU_EINVAL, "archive size does not match its header");
	uint8_t md5[16];
	Md5 m;
	m.update(res, (size_t)done);
	m.final(md5);
	if (memcmp(md5, arc + end, 16))
		return bail(LRZGPU_EINVAL, "MD5 mismatch");
	guard.p = nullptr;
	*out = res;
	*out_len = done;
	return rc;
}

int lrzgpu_rzip_chunk(lrzgpu_ctx *c, const uint8_t *in, int64_t n, int rzip_level, int chunk_bytes,
		      int64_t *victim_round, uint8_t **s0, int64_t *s0_len, uint8_t **s1, int64_t *s1_len,
		      lrzgpu_stats *stats)
{
	if (!c || !in || n <= 0 || rzip_level < 0 || rzip_level > 9 || chunk_bytes < 1 || chunk_bytes > 8 || !s0 || !s0_len ||
	    !s1 || !s1_len)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	if (stats)
		memset(stats, 0, sizeof(*stats));
	const int64_t launches0 = c->launches;
	uint8_t *d_in = nullptr;
	int rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	ChunkResult res;
	rc = rzip_chunk_device(c, d_in, n, rzip_level, chunk_bytes, victim_round ? *victim_round : 0, res, stats);
	if (rc)
		return rc;
	if (victim_round)
		*victim_round = res.st.victim_round;
	uint8_t *h0 = (uint8_t *)malloc((size_t)res.s0_len + 1), *h1 = (uint8_t *)malloc((size_t)res.s1_len + 1);
	if (!h0 || !h1) {
		free(h0);
		free(h1);
		return fail(c, LRZGPU_ENOMEM, "out of host memory");
	}
	CU(c, cudaMemcpy(h0, c->s0.p, (size_t)res.s0_len, cudaMemcpyDeviceToHost));
	if (res.s1_len)
		CU(c, cudaMemcpy(h1, c->s1.p, (size_t)res.s1_len, cudaMemcpyDeviceToHost));
	*s0 = h0;
	*s0_len = res.s0_len;
	*s1 = h1;
	*s1_len = res.s1_len;
	if (stats)
		stats->kernel_launches = c->launches - launches0;
	return LRZGPU_OK;
}

int lrzgpu_tag_scan(lrzgpu_ctx *c, const uint8_t *in, int64_t n, int64_t pos_lo, int64_t pos_hi, int64_t mask,
		    int64_t *out_pos, int64_t *out_tag, int64_t cap, int64_t *count)
{
	if (!c || !in || n <= 0 || pos_lo < 0 || pos_hi > n || !count)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	uint8_t *d_in = nullptr;
	int rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	const int64_t seg = 1 << 20;
	CU(c, c->cand[0].ensure((size_t)seg * sizeof(Cand)));
	CU(c, c->tc[0].ensure((size_t)(seg / kTile + 1) * 4));
	std::vector<Cand> hc((size_t)seg);
	std::vector<uint32_t> ht((size_t)(seg / kTile + 1));
	int64_t total = 0;
	for (int64_t lo = pos_lo - pos_lo % kTile; lo < pos_hi; lo += seg) {
		const int64_t a = lo < pos_lo ? pos_lo : lo, b = lo + seg < pos_hi ? lo + seg : pos_hi;
		// pos_lo of a launch must be tile aligned for the tile-strided layout: scan from `lo`, filter to [a,b)
		if (k1_launch(d_in, n, lo, b, mask, nullptr, 0, (Cand *)c->cand[0].p, (uint32_t *)c->tc[0].p, c->sms, c->sA))
			return fail(c, LRZGPU_ECUDA, "k1 launch failed");
		c->launches++;
		const int64_t ntiles = (b - 1) / kTile - lo / kTile + 1;
		CU(c, cudaMemcpyAsync(ht.data(), c->tc[0].p, (size_t)ntiles * 4, cudaMemcpyDeviceToHost, c->sA));
		CU(c, cudaMemcpyAsync(hc.data(), c->cand[0].p, (size_t)ntiles * kTile * sizeof(Cand), cudaMemcpyDeviceToHost, c->sA));
		CU(c, cudaStreamSynchronize(c->sA));
		for (int64_t t = 0; t < ntiles; t++)
			for (uint32_t i = 0; i < ht[(size_t)t]; i++) {
				const Cand &cd = hc[(size_t)(t * kTile + i)];
				if (cd.pos < a || cd.pos >= b)
					continue;
				if (total < cap && out_pos && out_tag) {
					out_pos[total] = cd.pos;
					out_tag[total] = cd.tag;
				}
				total++;
			}
	}
	*count = total;
	return LRZGPU_OK;
}

int lrzgpu_crc32(lrzgpu_ctx *c, const uint8_t *in, int64_t n, uint32_t *crc)
{
	if (!c || !in || n <= 0 || !crc)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	uint8_t *d_in = nullptr;
	int rc = upload(c, in, n, &d_in);
	if (rc)
		return rc;
	CU(c, c->crc.ensure(16));
	if (crc32_launch(d_in, n, (uint32_t *)c->crc.p, c->sms, c->sA))
		return fail(c, LRZGPU_ECUDA, "crc32 launch failed");
	c->launches++;
	uint32_t acc = 0;
	CU(c, cudaMemcpyAsync(&acc, c->crc.p, 4, cudaMemcpyDeviceToHost, c->sA));
	CU(c, cudaStreamSynchronize(c->sA));
	*crc = acc ^ 0xffffffffu;
	return LRZGPU_OK;
}

int lrzgpu_block_compress(lrzgpu_ctx *c, const lrzgpu_params *p, uint32_t dict_size, const uint8_t *in, int64_t u_len,
			  uint8_t **out, int64_t *c_len, int *c_type)
{
	if (!c || !p || !in || u_len <= 0 || !out || !c_len || !c_type)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	uint8_t *d_in = nullptr;
	int rc = upload(c, in, u_len, &d_in);
	if (rc)
		return rc;
	lrzgpu_sizing_t sz;
	memset(&sz, 0, sizeof(sz));
	sz.dict_size = dict_size;
	sz.bufsize = u_len;
	sz.threads = p->threads;
	std::vector<BlockJob> jobs(1);
	jobs[0].d_src = d_in;
	jobs[0].u_len = u_len;
	jobs[0].stream = 1;
	jobs[0].c_type = kCtypeNone;
	jobs[0].c_len = u_len;
	jobs[0].d_payload = d_in;
	if (p->backend != LRZGPU_BACKEND_NONE && u_len >= 64) {
		rc = backend_encode_blocks(c->backend, *p, sz, jobs, c->sms, c->sA, &c->launches, c->err, sizeof(c->err));
		if (rc)
			return rc;
	}
	uint8_t *h = (uint8_t *)malloc((size_t)jobs[0].c_len + 1);
	if (!h)
		return fail(c, LRZGPU_ENOMEM, "out of host memory");
	CU(c, cudaMemcpy(h, jobs[0].d_payload, (size_t)jobs[0].c_len, cudaMemcpyDeviceToHost));
	*out = h;
	*c_len = jobs[0].c_len;
	*c_type = jobs[0].c_type;
	return LRZGPU_OK;
}

int lrzgpu_lz4_gate(lrzgpu_ctx *c, const uint8_t *in, int64_t len, int threshold, int *compressible)
{
	if (!c || !in || len <= 0 || !compressible)
		return LRZGPU_EINVAL;
	cudaSetDevice(c->device);
	uint8_t *d_in = nullptr;
	int rc = upload(c, in, len, &d_in);
	if (rc)
		return rc;
	rc = backend_lz4_gate(c->backend, d_in, len, threshold, compressible, c->sA, &c->launches);
	if (rc)
		return fail(c, rc, "lz4 gate failed");
	return LRZGPU_OK;
}

int lrzgpu_k1_launch(lrzgpu_ctx *c, const void *d_buf, int64_t n, int64_t mask, void *d_cand, void *d_tile_count,
		     void *stream)
{
	if (!c || !d_buf || n <= 0 || !d_cand || !d_tile_count)
		return LRZGPU_EINVAL;
	if (k1_launch((const uint8_t *)d_buf, n, 0, n, mask, nullptr, 0, (Cand *)d_cand, (uint32_t *)d_tile_count, c->sms,
		      (cudaStream_t)stream))
		return fail(c, LRZGPU_ECUDA, "k1 launch: %s", cudaGetErrorString(cudaGetLastError()));
	c->launches++;
	return LRZGPU_OK;
}

int lrzgpu_crc32_launch(lrzgpu_ctx *c, const void *d_buf, int64_t n, void *d_crc, void *stream)
{
	if (!c || !d_buf || n <= 0 || !d_crc)
		return LRZGPU_EINVAL;
	if (crc32_launch((const uint8_t *)d_buf, n, (uint32_t *)d_crc, c->sms, (cudaStream_t)stream))
		return fail(c, LRZGPU_ECUDA, "crc32 launch failed");
	c->launches++;
	return LRZGPU_OK;
}

} // extern "C"
