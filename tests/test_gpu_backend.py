"""GPU parity tests for the block backends through the C ABI: lz4 gate, LZMA block encoder, zstd frames,
and whole archives against the unmodified reference binary (oracle/_ref, built from /root/reference by
oracle/Makefile and shipped to the GPU box)."""
import os
import time

import numpy as np
import pytest

import oracle
from lrzip_next_b200 import BACKEND_LZMA, BACKEND_ZSTD, make_params, sizing
from lrzip_next_b200 import datagen

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(11)


def _inputs():
    return {
        "text": datagen.gen_text(300_000).tobytes(),
        "random": RNG.integers(0, 256, 200_000, dtype=np.uint8).tobytes(),
        "zeros": bytes(400_000),
        "lowent": RNG.integers(0, 3, 150_000, dtype=np.uint8).tobytes(),
        "skew": (RNG.integers(0, 256, 300_000, dtype=np.uint8) & RNG.integers(0, 256, 300_000, dtype=np.uint8)).astype(np.uint8).tobytes(),
        "tiny": b"hello hello hello hello hello hello hello hello hello hello hello!",
    }


@pytest.mark.skipif(not oracle.have_lz4(), reason="system liblz4 not present")
@pytest.mark.parametrize("threshold", [100, 50, 10])
def test_lz4_gate_matches_liblz4(ctx, threshold):
    for name, d in _inputs().items():
        assert ctx.lz4_gate(d, threshold) == oracle.ref_lz4_gate(d, threshold), name


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("level,dict_size", [(7, 1 << 25), (5, 1 << 24), (9, 1 << 27), (3, 1 << 22), (1, 1 << 18)])
def test_lzma_block_bit_exact(ctx, level, dict_size):
    p = make_params(level=level, backend=BACKEND_LZMA, threads=8, threshold=0)
    for name, d in _inputs().items():
        if len(d) < 64:
            continue
        want = oracle.ref_lzma_block(d, level, dict_size, 2)
        t = time.time()
        got, ctype = ctx.block_compress(d, p, dict_size)
        dt = time.time() - t
        print(f"lzma L{level} {name}: {len(d)} -> {len(got)} in {dt:.2f}s ({len(d) / dt / 1e3:.0f} KB/s)")
        if want is None:
            assert ctype == 3 and got == d, name
        else:
            assert ctype == 6 and got == want, name


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,n,kw", [
    ("text", 1_500_000, dict(threads=8)),
    ("mix", 2_000_000, dict(threads=8)),
    ("trees", 3_000_000, dict(threads=2, level=5)),
    ("text", 2_000_000, dict(threads=4, level=3)),
    # several LZMA blocks per stream (block waves, flush order, next_head patching over compressed blocks):
    # -p8 on 45 MiB gives the 10 MiB minimum block => 5 stream-1 blocks
    ("text", 45 << 20, dict(threads=8)),
])
def test_lzma_archive_bit_identical_to_reference(ctx, kind, n, kw):
    d = datagen.generate(kind, n)
    # the reference binary sizes threads / dictionary from the core count of the machine it runs on
    # (PROCESSORS in open_stream_out, src/stream.c:1180): -p8 -m100 gives dict 32 MiB on a 16-core box and
    # 24 MiB from 20 cores up, so both sides must be told the cores of THIS box
    kw = dict(kw, processors=os.cpu_count() or 8)
    p = make_params(backend=BACKEND_LZMA, **kw)
    op = oracle.make_params(backend=oracle.BACKEND_LZMA, **kw)
    want = oracle.ref_compress(d, op)
    got = ctx.compress(d, p)
    assert got == want
    assert oracle.ref_test(got)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_zstd_archive_decodes_with_reference(ctx):
    # zstd payload parity is unpinned (libzstd is not vendored): the archive must be valid for the
    # reference decoder and keep the reference's stored / compressed decisions for the easy cases
    d = np.concatenate([RNG.integers(0, 256, 1 << 20, dtype=np.uint8), np.zeros(3 << 20, dtype=np.uint8),
                        datagen.gen_text(1 << 20)])
    p = make_params(backend=BACKEND_ZSTD, threads=8)
    got, st = ctx.compress(d, p, want_stats=True)
    assert got[17] == (7 << 4) | 4 and got[18] == 17
    assert oracle.ref_test(got)
    assert oracle.ref_decompress(got) == d.tobytes()
    assert len(got) < d.size - (2 << 20)  # the zero run went through RLE blocks


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_pipelined_chunk_equals_plain_and_reference(ctx, monkeypatch):
    """Chunks above one scan segment run the LZMA backend UNDER the rzip stage (blocks submitted as they fill,
    gate verdicts read on the device).  Same bytes as the plain order, and as the reference: text blocks that
    compress, random blocks the lz4 gate leaves stored, zero blocks, several blocks per stream."""
    d = np.concatenate([datagen.generate("text", 18 << 20), RNG.integers(0, 256, 25 << 20, dtype=np.uint8),
                        np.zeros(5 << 20, dtype=np.uint8), datagen.generate("text", 3 << 20)])
    kw = dict(threads=8, processors=os.cpu_count() or 8)
    p = make_params(backend=BACKEND_LZMA, **kw)
    got, st = ctx.compress(d, p, want_stats=True)
    assert st["blocks"] >= 5 and 0 < st["blocks_stored"] < st["blocks"]
    monkeypatch.setenv("LRZGPU_NO_OVERLAP", "1")
    plain = ctx.compress(d, p)
    monkeypatch.delenv("LRZGPU_NO_OVERLAP")
    assert got == plain
    assert got == oracle.ref_compress(d, oracle.make_params(backend=oracle.BACKEND_LZMA, **kw))


def _libzstd():
    import ctypes as C
    try:
        Z = C.CDLL("libzstd.so.1")
    except OSError:
        pytest.skip("system libzstd not present")
    Z.ZSTD_decompress.restype = C.c_size_t
    Z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    Z.ZSTD_isError.restype = C.c_uint
    Z.ZSTD_compress.restype = C.c_size_t
    Z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    return Z


def test_zstd_blocks_are_compressed_frames_libzstd_decodes(ctx):
    """zstd backend, block level.  PARITY UNPINNED for payload bytes (libzstd 1.5.5 is not vendored and its source
    is absent): what is pinned is (a) every payload is a Zstandard frame libzstd decodes back to the block,
    (b) the stored / compressed decision equals the reference's (lz4 gate + ZSTD_compress(17) + "not smaller =>
    stored", src/stream.c:167-229), (c) the size delta against ZSTD_compress(level 17) is reported."""
    import ctypes as C
    Z = _libzstd()
    p = make_params(level=7, backend=BACKEND_ZSTD, threads=8)
    report = {}
    for name, d in _inputs().items():
        if len(d) < 64:
            continue
        got, ctype = ctx.block_compress(d, p)
        n = len(d)
        ref = C.create_string_buffer(n + n // 8 + 1024)
        rs = Z.ZSTD_compress(ref, len(ref), d, n, 17)
        ref_kept = bool(oracle.ref_lz4_gate(d, 100)) and not Z.ZSTD_isError(rs) and rs < n
        if ctype == 10:
            back = C.create_string_buffer(n + 16)
            r = Z.ZSTD_decompress(back, len(back), got, len(got))
            assert not Z.ZSTD_isError(r) and r == n and back.raw[:n] == d, name
            assert len(got) < n
        else:
            assert ctype == 3 and got == d, name
        assert (ctype == 10) == ref_kept, (name, ctype, rs, n)
        report[name] = {"n": n, "ours": len(got), "zstd17": int(rs), "delta_pct": round(100.0 * (len(got) - rs) / rs, 1)}
    print("zstd payload parity: UNPINNED; sizes vs ZSTD_compress(17):", report)
    assert report["text"]["ours"] < 0.55 * report["text"]["n"]  # text is compressed for real now


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_zstd_c4_shaped_block_decisions_match_reference(ctx):
    """C4 shape (random || zeros, -Z, gate on), scaled: per block, the same c_type and u_len as the reference's
    archive (random blocks stay stored behind the lz4 gate, zero blocks become zstd frames), and the reference
    decodes our archive."""
    d = np.concatenate([RNG.integers(0, 256, 24 << 20, dtype=np.uint8), np.zeros(24 << 20, dtype=np.uint8)])
    kw = dict(threads=8, processors=os.cpu_count() or 8)
    got = ctx.compress(d, make_params(backend=BACKEND_ZSTD, **kw))
    want = oracle.ref_compress(d, oracle.make_params(backend=oracle.BACKEND_ZSTD, **kw))
    ours = [(c, s, t, u) for c, s, t, cl, u in oracle.walk_blocks(got)]
    theirs = [(c, s, t, u) for c, s, t, cl, u in oracle.walk_blocks(want)]
    assert ours == theirs
    assert any(t == 3 for _, _, t, _ in ours) and any(t == 10 for _, _, t, _ in ours)
    assert got[:21] == want[:21] and got[-16:] == want[-16:]
    assert oracle.ref_decompress(got) == d.tobytes()


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_device_decode_round_trips_and_reads_reference_archives(ctx):
    """lrzgpu_decompress (runzip_fd on the device): our own archives and the reference's, stored and
    LZMA, one and several chunks / blocks, come back byte-identical; a flipped byte is rejected (CRC / MD5 / stream
    checks) instead of returned."""
    cases = [
        (datagen.generate("text", 3_000_000), dict(backend=BACKEND_LZMA, threads=8)),
        (np.concatenate([datagen.generate("rep", 6 << 20, block=1 << 18), datagen.generate("text", 2 << 20),
                         np.zeros(3 << 20, dtype=np.uint8)]), dict(backend=BACKEND_LZMA, threads=8)),
        (datagen.generate("text", 45 << 20), dict(backend=0, threads=1, ramsize=30 * 1048576)),   # 3 chunks x 2 blocks
        (datagen.generate("vm", 12 << 20), dict(backend=0, threads=1)),
        (datagen.generate("text", 40), dict(backend=0, threads=1)),
    ]
    for d, kw in cases:
        kw = dict(kw, processors=os.cpu_count() or 8)
        ours = ctx.compress(d, make_params(**kw))
        assert ctx.decompress(ours) == d.tobytes()
        if kw.get("ramsize", 0) == 0:  # -m is in units of 100 MiB on the reference's command line
            theirs = oracle.ref_compress(d, oracle.make_params(**kw))
            assert ctx.decompress(theirs) == d.tobytes()
    bad = bytearray(ours)
    bad[len(bad) // 2] ^= 0x40
    with pytest.raises(Exception):
        ctx.decompress(bytes(bad))
    # zstd blocks (one RFC 8878 frame per stream block, csrc/zstd_dec.cuh): our own frames and libzstd's level-17 ones
    # written by the reference binary, compressed and stored blocks mixed (C4 shape)
    for d in (datagen.generate("text", 2_500_000), datagen.generate("randzero", 24 << 20),
              np.concatenate([datagen.generate("rep", 3 << 20, block=1 << 17), datagen.generate("text", 1 << 20)])):
        kw = dict(backend=BACKEND_ZSTD, threads=8, processors=os.cpu_count() or 8)
        z = ctx.compress(d, make_params(**kw))
        assert ctx.decompress(z) == d.tobytes()
        theirs = oracle.ref_compress(d, oracle.make_params(**kw))
        assert ctx.decompress(theirs) == d.tobytes()
        bad = bytearray(theirs)
        bad[len(bad) // 3] ^= 0x10
        with pytest.raises(Exception):
            ctx.decompress(bytes(bad))


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flt,delta,flag,backend", [
    (1, 0, "--x86", BACKEND_LZMA), (2, 0, "--arm", 0), (7, 0, "--arm64", BACKEND_LZMA), (4, 0, "--ppc", 0),
    (5, 0, "--sparc", 0), (3, 0, "--armt", 0), (6, 0, "--ia64", 0),
    (128, 1, "--delta=1", 0), (128, 4, "--delta=4", BACKEND_LZMA), (128, 48, "--delta=18", 0),
])
def test_filtered_archives_bit_identical_to_reference(ctx, flt, delta, flag, backend):
    """Pre-compression filters (src/stream.c:1587-1628): stream-1 blocks are converted on the device
    before the gate / backend; archives (filter byte in the magic, filtered stored blocks, LZMA over filtered bytes,
    several blocks so that every block restarts the converter at 0) equal the reference binary's."""
    rng = np.random.default_rng(flt + delta)
    n = (25 if backend else 80) << 20  # LZMA: 3 blocks of one chunk; stored at -m1: 2 chunks x 3 blocks
    d = rng.integers(0, 256, n, dtype=np.uint8)
    # code-like: plant E8 calls and 4-byte aligned branch opcodes of several architectures, plus a compressible part
    pos = (rng.integers(0, n // 4 - 2, n // 24) * 4).astype(np.int64)
    d[pos + 3] = rng.choice(np.array([0xEB, 0x94, 0x97, 0x90], dtype=np.uint8), pos.size)
    d[pos] = rng.choice(np.array([0x48, 0x4B, 0x40, 0x7F, 0xE8, 0x10, 0x16], dtype=np.uint8), pos.size)
    d[pos + 1] = np.where(rng.integers(0, 3, pos.size) == 0, 0xF0 | (d[pos + 1] & 7), d[pos + 1]).astype(np.uint8)  # Thumb BL high half
    d[5 << 20:9 << 20] = datagen.generate("text", 4 << 20)
    d[12 << 20:13 << 20] = (np.arange(1 << 20) // 3 % 251).astype(np.uint8)  # a ramp: what Delta is for
    kw = dict(threads=2 if backend else 1, processors=os.cpu_count() or 8)
    if not backend:
        kw["ramsize"] = 100 * 1048576  # -m1: two chunks, three blocks each
    p = make_params(backend=backend, filter=flt, delta=delta, **kw)
    got = ctx.compress(d, p)
    op = oracle.make_params(backend=oracle.BACKEND_LZMA if backend else 0, **kw)
    want = oracle.ref_compress(d, op, extra=(flag,))
    assert got[16] == want[16] and got[16] != 0
    assert got == want
    assert ctx.decompress(want) == d.tobytes()  # the decode side undoes the filter per stream-1 block


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_riscv_filtered_archive_bit_identical_to_reference(ctx):
    """--riscv (z7_BranchConv_RISCV_Enc, src/lzma/C/Bra.c:423-720): JAL ra / t0 and AUIPC pairs planted on 2-byte
    boundaries, two chunks of three stored blocks; archive equal to the reference binary's, decoded back on the device."""
    rng = np.random.default_rng(88)
    n = 80 << 20
    d = rng.integers(0, 256, n, dtype=np.uint8)
    pos = (rng.integers(0, n // 2 - 8, n // 40) * 2).astype(np.int64)
    k = pos.size // 2
    d[pos[:k]] = 0xEF
    d[pos[:k] + 1] = (d[pos[:k] + 1] & 0xF0) | rng.choice(np.array([0, 2], dtype=np.uint8), k)
    rd = rng.integers(1, 32, pos.size - k).astype(np.uint8)
    d[pos[k:]] = 0x17 | ((rd & 1) << 7)
    d[pos[k:] + 1] = (d[pos[k:] + 1] & 0xF0) | (rd >> 1)
    d[pos[k:] + 4] |= 3                                             # a 32-bit second instruction ...
    d[pos[k:] + 5] = (d[pos[k:] + 5] & 0x7F) | ((rd & 1) << 7)      # ... whose rs1 is the AUIPC's rd
    d[pos[k:] + 6] = (d[pos[k:] + 6] & 0xF0) | (rd >> 1)
    d[5 << 20:9 << 20] = datagen.generate("text", 4 << 20)
    kw = dict(threads=1, processors=os.cpu_count() or 8, ramsize=100 * 1048576)
    got = ctx.compress(d, make_params(backend=0, filter=8, **kw))
    want = oracle.ref_compress(d, oracle.make_params(backend=0, **kw), extra=("--riscv",))
    assert got[16] == want[16] == 8
    assert got == want
    plain = oracle.ref_compress(d, oracle.make_params(backend=0, **kw))
    assert sum(a != b for a, b in zip(want[:4 << 20], plain[:4 << 20])) > 1000  # the converter did convert
    assert ctx.decompress(want) == d.tobytes()


def test_unbuilt_filters_are_rejected(ctx):
    for flt in (9, 77):
        with pytest.raises(Exception):
            ctx.compress(np.zeros(1000, dtype=np.uint8), make_params(filter=flt))
    with pytest.raises(Exception):
        ctx.compress(np.zeros(1000, dtype=np.uint8), make_params(filter=128, delta=17))
