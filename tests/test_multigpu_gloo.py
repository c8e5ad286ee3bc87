"""world_size-2 gloo tests (CPU) of the window-sharded path: chunk planning, victim_round settlement,
blob gather and archive assembly.  The per-chunk compressor is played by the oracle here (the GPU
library cannot run on this box); on GPU boxes `bench.py --gpus N` runs the same paths with the real one."""
import hashlib
import os
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle
from lrzip_next_b200 import api, datagen, make_params, multigpu, sizing


class OracleChunkCtx:
    """compress_chunk() with the C ABI's contract, backed by the CPU oracle (tests only)."""

    def compress_chunk(self, data, params, sz, eof, victim_round=0):
        d = np.ascontiguousarray(data, dtype=np.uint8)
        p = oracle.make_params(level=params.level, backend=0, threads=1)
        s0, s1, st, vr = oracle.rzip_chunk(d, params.level, victim_round=victim_round)
        arc, _ = oracle.compress(d, p)  # single-chunk archive: body = blob with eof = 1
        blob = bytearray(arc[21:-16])
        blob[1] = 1 if eof else 0
        if victim_round:
            # the single-chunk oracle archive always starts from victim_round 0; rebuild stream bytes
            raise AssertionError("test data must not depend on victim_round")
        return bytes(blob), vr, {"chain_evictions": st["chain_evictions"]}


def _worker(rank, world, port, path, n, window):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = datagen.gen_rep(n, block=1 << 16)
    params = make_params(backend=0, threads=1, window=window, ramsize=100 * 1048576 * 3)
    sz = sizing(params, n)
    plans = multigpu.plan_chunks(n, sz.max_chunk, world)
    shards = {p.index: d[p.offset:p.offset + p.size] for p in plans if p.rank == rank}
    arc, _ = multigpu.compress_sharded(OracleChunkCtx(), params, sz, shards, plans, hashlib.md5(d.tobytes()).digest())
    if rank == 0:
        with open(path, "wb") as fh:
            fh.write(arc)
    dist.destroy_process_group()


def test_sharded_archive_equals_single_process(tmp_path):
    n, window = 250 << 20, 1   # 3 windows of 100 MiB / 100 MiB / 50 MiB over 2 ranks
    path = str(tmp_path / "out.lrz")
    mp.spawn(_worker, args=(2, 29731, path, n, window), nprocs=2, join=True)
    d = datagen.gen_rep(n, block=1 << 16)
    want, _ = oracle.compress(d, oracle.make_params(backend=0, threads=1, window=window, ramsize=100 * 1048576 * 3))
    with open(path, "rb") as fh:
        assert fh.read() == want


def test_plan_chunks_round_robin():
    plans = multigpu.plan_chunks(1000, 300, 2)
    assert [(p.offset, p.size, p.eof, p.rank) for p in plans] == [(0, 300, False, 0), (300, 300, False, 1),
                                                                    (600, 300, False, 0), (900, 100, True, 1)]


def test_victim_round_settlement():
    # chunk 1 had evictions and ends at 5; chunk 2 had evictions but assumed 0 => must be redone with 5
    rep = [(0, 0, 0), (0, 3, 5), (0, 2, 7), (0, 0, 0)]
    assert multigpu.resolve_victim_rounds(rep) == [2]
    assert multigpu.true_incoming(rep, 2) == 5
    rep[2] = (5, 2, 9)
    assert multigpu.resolve_victim_rounds(rep) == []
    assert multigpu.true_incoming(rep, 4) == 9


def test_magic_matches_reference_layout():
    p = make_params(backend=api.BACKEND_LZMA, threads=8)
    m = multigpu.make_magic(p, sizing(p, 123456789), 123456789)
    assert m[:6] == b"LRZI\x00\x0e" and int.from_bytes(m[6:14], "little") == 123456789
    assert (m[14], m[17], m[18], m[19]) == (1, 1, 0x1a, 0x77)   # SURVEY.md Appendix A probe: 01 1a 77
    pz = make_params(backend=api.BACKEND_ZSTD, threads=8)
    mz = multigpu.make_magic(pz, sizing(pz, 1000), 1000)
    assert (mz[17], mz[18], mz[19]) == (0x74, 0x11, 0x77)


class ChainFakeCtx:
    """chunk_begin / chunk_finish with a made-up but victim_round-DEPENDENT result, to test the chained
    orchestration (who waits for whom, what is passed on, gather order) without a GPU."""

    def __init__(self):
        self.pending = None

    @staticmethod
    def _vr_out(data, vin):
        return (vin * 5 + int(np.asarray(data[:64], dtype=np.int64).sum()) + 3) % 16

    def chunk_begin(self, data, params, sz, eof, victim_round=0):
        assert self.pending is None
        d = np.ascontiguousarray(data, dtype=np.uint8)
        self.pending = (d, bool(eof), victim_round)
        return self._vr_out(d, victim_round), {"chain_evictions": 1, "lookups": int(d.size), "crc32": 7}

    def chunk_finish(self):
        d, eof, vin = self.pending
        self.pending = None
        blob = bytes([4, 1 if eof else 0, vin]) + hashlib.sha1(d.tobytes()).digest() + d[:1000].tobytes()
        return blob, {"blocks": 2, "lookups": 0, "crc32": 0}


def _chain_worker(rank, world, port, path, n, chunk):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = datagen.gen_text(n)
    params = make_params(backend=0, threads=1)
    sz = sizing(params, n)
    plans = multigpu.plan_chunks(n, chunk, world)
    shards = {p.index: d[p.offset:p.offset + p.size] for p in plans if p.rank == rank}
    arc, sts = multigpu.compress_chained(ChainFakeCtx(), params, sz, shards, plans, hashlib.md5(d.tobytes()).digest())
    if rank == 0:
        with open(path, "wb") as fh:
            fh.write(arc)
        assert all(s["lookups"] > 0 and s["blocks"] == 2 for s in sts)
    dist.destroy_process_group()


def test_chained_windows_pass_victim_round_along(tmp_path):
    n, chunk = 700_000, 100_000  # 7 windows over 2 ranks: ranks alternate, each waits for its predecessor
    path = str(tmp_path / "chain.lrz")
    mp.spawn(_chain_worker, args=(2, 29741, path, n, chunk), nprocs=2, join=True)
    d = datagen.gen_text(n)
    params = make_params(backend=0, threads=1)
    sz = sizing(params, n)
    fake, vr, blobs = ChainFakeCtx(), 0, {}
    for p in multigpu.plan_chunks(n, chunk, 2):  # the serial meaning: one process, chunks in order
        vr, _ = fake.chunk_begin(d[p.offset:p.offset + p.size], params, sz, p.eof, vr)
        blobs[p.index], _ = fake.chunk_finish()
    want = multigpu.assemble(params, sz, n, blobs, hashlib.md5(d.tobytes()).digest())
    with open(path, "rb") as fh:
        assert fh.read() == want


class SpecFakeCtx(ChainFakeCtx):
    """Adds chunk_begin_all / chunk_select (all-values speculation) to the fake: the table of outgoing counter
    values for every incoming one, and the pick of one variant."""

    def chunk_begin_all(self, data, params, sz, eof):
        assert self.pending is None
        d = np.ascontiguousarray(data, dtype=np.uint8)
        self.pending = (d, bool(eof), None)
        return [self._vr_out(d, v) for v in range(16)], {"chain_evictions": 1, "lookups": int(d.size), "crc32": 0}

    def chunk_select(self, victim_in):
        d, eof, v = self.pending
        assert v is None
        self.pending = (d, eof, victim_in)
        return {"lookups": 0, "crc32": 7}


def _spec_worker(rank, world, port, path, n, chunk):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = datagen.gen_text(n)
    params = make_params(backend=0, threads=1)
    sz = sizing(params, n)
    plans = multigpu.plan_chunks(n, chunk, world)
    shards = {p.index: d[p.offset:p.offset + p.size] for p in plans if p.rank == rank}
    md5 = hashlib.md5(d.tobytes()).digest()
    arc, sts = multigpu.compress_speculated(SpecFakeCtx(), params, sz, shards, plans, (lambda: md5) if rank == 0 else None)
    if rank == 0:
        with open(path, "wb") as fh:
            fh.write(arc)
        assert all(s["lookups"] > 0 and s["blocks"] == 2 and s["crc32"] == 7 for s in sts)
    dist.destroy_process_group()


def test_speculated_windows_equal_the_serial_chain(tmp_path):
    n, chunk = 700_000, 100_000  # 7 windows over 2 ranks
    path = str(tmp_path / "spec.lrz")
    mp.spawn(_spec_worker, args=(2, 29751, path, n, chunk), nprocs=2, join=True)
    d = datagen.gen_text(n)
    params = make_params(backend=0, threads=1)
    sz = sizing(params, n)
    fake, vr, blobs = ChainFakeCtx(), 0, {}
    for p in multigpu.plan_chunks(n, chunk, 2):
        vr, _ = fake.chunk_begin(d[p.offset:p.offset + p.size], params, sz, p.eof, vr)
        blobs[p.index], _ = fake.chunk_finish()
    want = multigpu.assemble(params, sz, n, blobs, hashlib.md5(d.tobytes()).digest())
    with open(path, "rb") as fh:
        assert fh.read() == want


def test_missing_md5_is_an_error():
    with pytest.raises(ValueError):
        multigpu._resolve_md5(None)
    with pytest.raises(ValueError):
        multigpu._resolve_md5(b"short")
    assert multigpu._resolve_md5(lambda: b"0123456789abcdef") == b"0123456789abcdef"
