"""Window-sharded compression across the GPUs of one box (one process per GPU).

lrzip-next windows ("chunks") are independent in the format and in the algorithm (hash table reset
per chunk src/rzip.c:599-600, chunk-relative offsets, position-independent blobs src/stream.c:1021-1036),
so chunk i goes to rank i mod world and no collective is needed on the data path.  Two couplings are
carried explicitly:

  * the ``eof`` byte is 1 only in the last chunk (src/rzip.c:1173-1174) -- known up front;
  * the reference's function-static ``victim_round`` (src/rzip.c:308) leaks from one chunk into the
    next.  Every rank first compresses its chunks assuming the value 0 and reports, per chunk,
    (chain_evictions, victim_round_out).  A chunk without chain evictions never consulted the counter
    and passes its input through unchanged; walking the chunks in order gives every chunk's true
    incoming value, and only chunks that both had evictions and assumed the wrong value are redone.

The only collective is the final gather of the finished blobs to rank 0 (``gather_blobs``): NCCL over
NVLink when the tensors are on the GPUs, gloo in the CPU tests.  Rank 0 appends the blobs in chunk
order between the 21-byte magic and the whole-file MD5 (src/lrzip.c:131-208, src/rzip.c:1195-1218).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from .api import BACKEND_LZMA, BACKEND_ZSTD, Params, Sizing


@dataclass
class ChunkPlan:
    index: int
    offset: int
    size: int
    eof: bool
    rank: int


def plan_chunks(st_size: int, max_chunk: int, world: int) -> list[ChunkPlan]:
    """The reference's window loop (src/rzip.c:1041-1186) as a static plan, round-robin over ranks."""
    plans, off, i = [], 0, 0
    while off < st_size:
        size = min(max_chunk, st_size - off)
        plans.append(ChunkPlan(i, off, size, off + size == st_size, i % world))
        off += size
        i += 1
    return plans


def resolve_victim_rounds(reports: list[tuple[int, int, int]]) -> list[int]:
    """reports[i] = (assumed_in, chain_evictions, victim_round_out) of chunk i, in chunk order.
    Returns the indices of the chunks that must be redone, together fixing up nothing else: the
    caller re-runs those with the value returned by ``true_incoming``."""
    redo, cur = [], 0
    for i, (assumed, evictions, vr_out) in enumerate(reports):
        if evictions and assumed != cur:
            redo.append(i)
            return redo  # later chunks depend on this one's true output: resolve one at a time
        cur = vr_out if evictions else cur
    return redo


def true_incoming(reports: list[tuple[int, int, int]], upto: int) -> int:
    cur = 0
    for assumed, evictions, vr_out in reports[:upto]:
        cur = vr_out if evictions else cur
    return cur


def make_magic(params: Params, sz: Sizing, st_size: int) -> bytes:
    """21-byte header of src/lrzip.c:131-208 (file -> file, MD5, no encryption / filter / comment)."""
    m = bytearray(21)
    m[0:4] = b"LRZI"
    m[4], m[5] = 0, 14
    m[6:14] = int(st_size).to_bytes(8, "little")
    m[14] = 1
    if params.backend == BACKEND_LZMA:
        m[17] = 1
        prop = 0
        while prop <= 40 and sz.dict_size > (0xFFFFFFFF if prop == 40 else ((2 | (prop & 1)) << (prop // 2 + 11))):
            prop += 1
        m[18] = prop
    elif params.backend == BACKEND_ZSTD:
        m[17] = (params.level << 4) | 4
        m[18] = [-1, 2, 4, 5, 7, 12, 15, 17, 18, 22][params.level] & 0xFF
    rzl = params.rzip_level or params.level
    m[19] = (rzl << 4) + params.level
    return bytes(m)


def gather_blobs(local: dict[int, bytes], device: torch.device | None = None) -> dict[int, bytes] | None:
    """Gather {chunk index: blob} from every rank to rank 0 (returns None elsewhere).
    One all_gather of the per-rank byte counts, then one gather of the padded payloads."""
    world, rank = dist.get_world_size(), dist.get_rank()
    device = device or torch.device("cpu")
    idx = sorted(local)
    payload = b"".join(local[i] for i in idx)
    meta = torch.tensor([len(idx), len(payload)], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_chunks = max(int(m[0]) for m in metas)
    max_bytes = max(int(m[1]) for m in metas)
    table = torch.full((max(max_chunks, 1), 2), -1, dtype=torch.int64, device=device)
    for k, i in enumerate(idx):
        table[k, 0], table[k, 1] = i, len(local[i])
    buf = torch.zeros(max(max_bytes, 1), dtype=torch.uint8, device=device)
    if payload:
        buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    tables = [torch.zeros_like(table) for _ in range(world)] if rank == 0 else None
    bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(table, tables, dst=0)
    dist.gather(buf, bufs, dst=0)
    if rank != 0:
        return None
    out: dict[int, bytes] = {}
    for r in range(world):
        t = tables[r].cpu().numpy()
        b = bufs[r].cpu().numpy()
        off = 0
        for i, ln in t:
            if i < 0:
                continue
            out[int(i)] = b[off:off + int(ln)].tobytes()
            off += int(ln)
    return out


def assemble(params: Params, sz: Sizing, st_size: int, blobs: dict[int, bytes], md5: bytes) -> bytes:
    return make_magic(params, sz, st_size) + b"".join(blobs[i] for i in sorted(blobs)) + md5


def _resolve_md5(whole_md5) -> bytes:
    """The trailing MD5 is over the whole file (src/rzip.c:1195-1218) and MD5 states do not combine, so rank 0
    must be given it: either the 16 bytes or a callable producing them (e.g. the join of a hashing thread)."""
    md5 = whole_md5() if callable(whole_md5) else whole_md5
    if not isinstance(md5, (bytes, bytearray)) or len(md5) != 16:
        raise ValueError("rank 0 needs the whole file's 16-byte MD5 (bytes or a callable returning them)")
    return bytes(md5)


def compress_sharded(ctx, params: Params, sz: Sizing, shards: dict[int, np.ndarray], plans: list[ChunkPlan],
                     whole_md5: bytes | None = None, device: torch.device | None = None):
    """Compress this rank's chunks (``shards``: chunk index -> bytes), settle victim_round, gather.
    Returns (archive or None, per-rank stats list)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = [p for p in plans if p.rank == rank]
    if rank == 0 and whole_md5 is None:
        raise ValueError("rank 0 needs the whole file's MD5")
    blobs, reports, stats_by = {}, {}, {}
    for p in mine:
        blob, vr_out, st = ctx.compress_chunk(shards[p.index], params, sz, p.eof, 0)
        blobs[p.index], reports[p.index] = blob, (0, st["chain_evictions"], vr_out)
        stats_by[p.index] = st
    # settle the victim_round chain (a few integers per chunk; gloo/nccl all_gather_object)
    while True:
        allrep: list = [None] * world
        dist.all_gather_object(allrep, reports)
        merged = {}
        for r in allrep:
            merged.update(r)
        ordered = [merged[i] for i in range(len(plans))]
        redo = resolve_victim_rounds(ordered)
        if not redo:
            break
        i = redo[0]
        if plans[i].rank == rank:
            vin = true_incoming(ordered, i)
            blob, vr_out, st = ctx.compress_chunk(shards[i], params, sz, plans[i].eof, vin)
            blobs[i], reports[i] = blob, (vin, st["chain_evictions"], vr_out)
            stats_by[i] = st  # the redone chunk's counters replace the speculative ones
    stats = [stats_by[i] for i in sorted(stats_by)]
    got = gather_blobs(blobs, device)
    if rank != 0:
        return None, stats
    return assemble(params, sz, sum(p.size for p in plans), got, _resolve_md5(whole_md5)), stats


_chain_calls = 0


def compress_chained(ctx, params: Params, sz: Sizing, shards: dict[int, np.ndarray], plans: list[ChunkPlan],
                     whole_md5: bytes | None = None, device: torch.device | None = None):
    """Like compress_sharded, for data that consults ``victim_round`` all the time (text: the XOR tags collide
    and fill equal-tag chains), where speculation would re-run nearly every chunk.  The chunks form a chain
    through that one counter, but only through the rzip stage: the owner of chunk i receives the counter from
    the owner of chunk i-1, runs ``chunk_begin`` (rzip), passes the new counter on at once and only then runs
    ``chunk_finish`` (blocks -> backend -> framing), so the backend phases of all chunks overlap and the
    critical path is  sum(rzip) + one backend  instead of  sum(rzip + backend).
    Returns (archive or None, per-rank stats list)."""
    global _chain_calls
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0 and whole_md5 is None:
        raise ValueError("rank 0 needs the whole file's MD5")
    device = device or torch.device("cpu")
    # The counter (one small integer per chunk) travels through the process group's key-value store: plain
    # TCP, the same under nccl and gloo, no device work and no pairing rules to get wrong.  Every rank makes
    # the same number of calls, so the call index keeps the keys of successive archives apart.
    from datetime import timedelta
    store = dist.distributed_c10d._get_default_store()
    store.set_timeout(timedelta(hours=6))
    _chain_calls += 1
    tag = f"lrzgpu/victim_round/{_chain_calls}"
    blobs, stats = {}, []
    last_out = 0
    for p in plans:
        if p.rank != rank:
            continue
        vin = 0
        if p.index > 0:
            vin = last_out if plans[p.index - 1].rank == rank else int(store.get(f"{tag}/{p.index - 1}"))
        vr_out, st_a = ctx.chunk_begin(shards[p.index], params, sz, p.eof, vin)
        last_out = vr_out
        if p.index + 1 < len(plans) and plans[p.index + 1].rank != rank:
            store.set(f"{tag}/{p.index}", str(vr_out))
        blob, st_b = ctx.chunk_finish()
        blobs[p.index] = blob
        merged = dict(st_b)
        for k, v in st_a.items():  # begin holds the rzip counters and timings, finish the backend ones
            if isinstance(v, (int, float)) and k not in ("crc32",):
                merged[k] = merged.get(k, 0) + v
        merged["crc32"] = st_a.get("crc32", 0)
        stats.append(merged)
    got = gather_blobs(blobs, device)
    if rank != 0:
        return None, stats
    return assemble(params, sz, sum(p.size for p in plans), got, _resolve_md5(whole_md5)), stats


def _merge_stats(parts: list[dict]) -> dict:
    """Sum the numeric fields of the stats of the calls that made up one window (begin / select / finish)."""
    merged: dict = {}
    for st in parts:
        for k, v in st.items():
            if isinstance(v, (int, float)) and k != "crc32":
                merged[k] = merged.get(k, 0) + v
    for st in parts:
        if st.get("crc32"):
            merged["crc32"] = st["crc32"]
    merged.setdefault("crc32", 0)
    return merged


_spec_calls = 0


def compress_speculated(ctx, params: Params, sz: Sizing, shards: dict[int, np.ndarray], plans: list[ChunkPlan],
                        whole_md5=None, device: torch.device | None = None):
    """All-values speculation of ``victim_round``: every window runs its rzip stage at once, on its own GPU, for
    EVERY value the reference's cross-window counter can take (``chunk_begin_all``: one commit CTA, one table and
    one record array per value -- the commit stage needs one SM of 148).  Only the counter's true value then
    travels from window to window (one small integer through the process group's store, as in
    ``compress_chained``), each owner picks the matching variant (``chunk_select``) and runs its backend.  The
    rzip stages of all windows overlap, so the critical path is one rzip stage + one backend for any number of
    windows, and the archive is byte-identical to the serial one.  Window 0 starts from the known value 0.
    Falls back to waiting for the predecessor (the chained form) when the variants do not fit device memory.
    Returns (archive or None, per-window stats list of this rank)."""
    global _spec_calls
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0 and whole_md5 is None:
        raise ValueError("rank 0 needs the whole file's MD5")
    device = device or torch.device("cpu")
    from datetime import timedelta
    store = dist.distributed_c10d._get_default_store()
    store.set_timeout(timedelta(hours=6))
    _spec_calls += 1
    tag = f"lrzgpu/victim_round_true/{_spec_calls}"
    blobs, stats = {}, []
    last_out = 0

    def incoming(p):
        return last_out if plans[p.index - 1].rank == rank else int(store.get(f"{tag}/{p.index - 1}"))

    for p in plans:
        if p.rank != rank:
            continue
        parts = []
        if p.index == 0:
            vr_out, st_a = ctx.chunk_begin(shards[p.index], params, sz, p.eof, 0)
            parts.append(st_a)
        else:
            try:
                table, st_a = ctx.chunk_begin_all(shards[p.index], params, sz, p.eof)
            except Exception as e:  # LrzGpuError ENOMEM: the variants do not fit -> wait for the predecessor
                if getattr(e, "code", None) != -2:
                    raise
                table = None
            if table is None:
                vr_out, st_a = ctx.chunk_begin(shards[p.index], params, sz, p.eof, incoming(p))
                parts.append(st_a)
            else:
                vin = incoming(p)
                parts += [st_a, ctx.chunk_select(vin)]
                vr_out = int(table[vin])
        last_out = vr_out
        if p.index + 1 < len(plans) and plans[p.index + 1].rank != rank:
            store.set(f"{tag}/{p.index}", str(vr_out))
        blob, st_b = ctx.chunk_finish()
        blobs[p.index] = blob
        stats.append(_merge_stats(parts + [st_b]))
    got = gather_blobs(blobs, device)
    if rank != 0:
        return None, stats
    return assemble(params, sz, sum(p.size for p in plans), got, _resolve_md5(whole_md5)), stats
