"""GPU parity tests for the rzip path (K1 tag scan, K2 commit, K4 emit, CRC) through the C ABI.

Every comparison is bit-exact against the CPU oracle (oracle/liboracle.so), which is itself pinned to
the unmodified reference (tests/test_oracle.py, tests/golden/)."""
import hashlib
import zlib

import numpy as np
import pytest

import oracle
from lrzip_next_b200 import BACKEND_NONE, make_params
from lrzip_next_b200 import datagen

pytestmark = pytest.mark.gpu


def _first_diff(a: bytes, b: bytes):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    nz = np.flatnonzero(x)
    return int(nz[0]) if nz.size else (n if len(a) != len(b) else None)


@pytest.mark.parametrize("n", [1, 31, 32, 4095, 4096, 4097, 100_000, (1 << 20) + 77])
def test_crc32_matches_zlib(ctx, n):
    d = np.random.default_rng(n).integers(0, 256, size=n, dtype=np.uint8)
    assert ctx.crc32(d) == zlib.crc32(d.tobytes())


@pytest.mark.parametrize("kind,n,mask", [("text", 300_000, 1), ("rep", 1 << 20, 1), ("text", 70_000, 15),
                                         ("randzero", 200_000, 3), ("text", 40, 1), ("text", 31, 1), ("text", 5000, 0)])
def test_tag_scan_matches_full_tag(ctx, kind, n, mask):
    d = datagen.generate(kind, n)
    pos, tag = ctx.tag_scan(d, mask)
    hi = oracle.hash_index()
    # reference tags by prefix XOR (src/rzip.c:405-416)
    x = np.concatenate(([0], np.bitwise_xor.accumulate(hi[d])))
    end = n - 31
    if end < 1:
        assert pos.size == 0
        return
    p = np.arange(1, end + 1)
    t = x[p + 31] ^ x[p]
    keep = (t & mask) == mask
    assert np.array_equal(pos, p[keep])
    assert np.array_equal(tag, t[keep])


CASES = [
    ("rep", 8 << 20, 7), ("text", 6 << 20, 7), ("text", 3 << 20, 3), ("mix", 8 << 20, 7), ("vm", 8 << 20, 7),
    ("trees", 8 << 20, 9), ("randzero", 4 << 20, 1), ("text", 1000, 7), ("text", 20, 7), ("text", 31, 7),
    ("text", 32, 7), ("text", 24 << 20, 1), ("trees", 40 << 20, 7), ("rep", 33 << 20, 5),
]


@pytest.mark.parametrize("kind,n,level", CASES)
def test_rzip_chunk_streams_bit_exact(ctx, kind, n, level):
    d = datagen.generate(kind, n)
    o0, o1, ost, ovr = oracle.rzip_chunk(d, level)
    s0, s1, st, vr = ctx.rzip_chunk(d, level)
    for k in ("inserts", "lookups", "tag_hits", "tag_misses", "chain_evictions", "sweeps", "hash_count",
              "final_min_mask", "final_tag_mask", "matches", "match_bytes", "literals", "literal_bytes"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert st["crc32"] == ost["crc32"]
    assert len(s0) == len(o0) and _first_diff(s0, o0) is None, _first_diff(s0, o0)
    assert len(s1) == len(o1) and _first_diff(s1, o1) is None, _first_diff(s1, o1)
    assert vr == ovr


def test_rzip_victim_round_carried(ctx):
    # a periodic input whose tag passes the mask builds equal-tag chains that hit max_chain_len
    blk = np.frombuffer(b"abcdefg" * 5, dtype=np.uint8)
    d = np.tile(blk, (2 << 20) // blk.size + 1)[:2 << 20].copy()
    rnd = np.random.default_rng(5).integers(0, 256, size=d.size // 8, dtype=np.uint8)
    d[::8] ^= (rnd & 1)  # break most long matches, keep many equal windows
    for vr_in in (0, 3):
        o0, o1, ost, ovr = oracle.rzip_chunk(d, 7, victim_round=vr_in)
        s0, s1, st, vr = ctx.rzip_chunk(d, 7, victim_round=vr_in)
        assert (s0, s1, vr) == (o0, o1, ovr)
        assert st["chain_evictions"] == ost["chain_evictions"]


@pytest.mark.parametrize("kind,n,kw", [
    ("rep", 12 << 20, {}), ("text", 5 << 20, {}), ("mix", 8 << 20, {}),
    ("text", 30 << 20, dict(window=1, ramsize=3 * 100 * 1048576)),   # -w1 = 100 MiB window: still ONE chunk
    ("vm", 25 << 20, dict(level=9)),
    # several chunks (victim_round carried between them, eof only on the last) and several blocks per stream
    # (flush order, next_head patching); each has an oracle-vs-reference-binary twin in tests/test_oracle.py
    ("text", 250 << 20, dict(window=1)),                              # 3 chunks of 100 / 100 / 50 MiB
    ("text", 80 << 20, dict(ramsize=100 * 1048576)),                  # 2 chunks x 3 blocks of 33.3 MiB
    ("mix", 70 << 20, dict(ramsize=30 * 1048576)),                    # 4 chunks of 20 MiB x 2 blocks of 10 MiB
])
def test_archive_stored_bit_identical_to_oracle(ctx, kind, n, kw):
    d = datagen.generate(kind, n)
    p = make_params(backend=BACKEND_NONE, threads=1, **kw)
    op = oracle.make_params(backend=oracle.BACKEND_NONE, threads=1, **kw)
    want, _ = oracle.compress(d, op)
    got = ctx.compress(d, p)
    assert len(got) == len(want) and _first_diff(got, want) is None, _first_diff(got, want)
    assert got[-16:] == hashlib.md5(d.tobytes()).digest()


def test_rzip_chunk_bytes_5_forced(ctx):
    """chunk_bytes = 5 (files >= 2^32, C3/C4/C5; src/rzip.c:1129-1133) forced on a small chunk: 5-byte match
    distances in stream 0."""
    d = np.concatenate([datagen.generate("text", 600_000), datagen.generate("rep", 400_000, block=1 << 15)])
    o0, o1, ost, _ = oracle.rzip_chunk(d, 7, chunk_bytes=5)
    s0, s1, st, _ = ctx.rzip_chunk(d, 7, chunk_bytes=5)
    assert (s0, s1) == (o0, o1)
    o4, _, _, _ = oracle.rzip_chunk(d, 7, chunk_bytes=4)
    assert len(s0) == len(o4) + st["matches"]  # one more distance byte per match record


def test_c1_full_size_appendix_f_facts(ctx):
    """BASELINE config C1 at its stated size: 100 MiB of a repeated 1 MiB block, -n -p1.  SURVEY.md Appendix F
    records what the reference produces: 1,059,858 bytes, 1585 matches, MD5 535d79..."""
    d = datagen.gen_rep(100 << 20)
    got, st = ctx.compress(d, make_params(backend=BACKEND_NONE, threads=1), want_stats=True)
    assert len(got) == 1_059_858
    assert (st["matches"], st["match_bytes"], st["literals"], st["literal_bytes"], st["inserts"]) == \
           (1585, 103_808_993, 19, 1_048_607, 524_840)
    assert got[-16:].hex() == "535d792a42ccc7c552714d2a97bf158f"
    want, _ = oracle.compress(d, oracle.make_params(backend=oracle.BACKEND_NONE, threads=1))
    assert got == want


def test_compress_file_equals_compress(ctx, tmp_path):
    """lrzgpu_compress_file (the B1 seam: rzip_fd + write_magic, file -> file) writes the bytes lrzgpu_compress returns."""
    d = np.concatenate([datagen.generate("text", 2_500_000), datagen.generate("rep", 1_500_000, block=1 << 16)])
    src, dst = tmp_path / "in.bin", tmp_path / "out.lrz"
    d.tofile(src)
    p = make_params(backend=BACKEND_NONE, threads=1)
    st = ctx.compress_file(str(src), str(dst), p)
    got = dst.read_bytes()
    assert got == ctx.compress(d, p)
    assert got == oracle.compress(d, oracle.make_params(backend=oracle.BACKEND_NONE, threads=1))[0]
    assert st["chunks"] == 1 and st["stream1_bytes"] > 0
    with pytest.raises(Exception):
        ctx.compress_file(str(tmp_path / "missing.bin"), str(dst), p)


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_archive_verifies_with_reference_binary(ctx):
    d = datagen.generate("trees", 16 << 20)
    p = make_params(backend=BACKEND_NONE, threads=1)
    got = ctx.compress(d, p)
    assert oracle.ref_test(got)
    assert oracle.ref_decompress(got) == d.tobytes()


def test_two_phase_chunk_equals_one_call(ctx):
    """lrzgpu_chunk_begin + lrzgpu_chunk_finish (the chained multi-GPU form) == lrzgpu_compress_chunk."""
    from lrzip_next_b200 import BACKEND_LZMA, sizing
    d = np.concatenate([datagen.generate("text", 1_200_000), datagen.generate("rep", 800_000, block=1 << 16)])
    for p in (make_params(backend=BACKEND_NONE, threads=1), make_params(backend=BACKEND_LZMA, threads=8)):
        sz = sizing(p, d.size)
        for vin in (0, 5):
            blob, vr, st = ctx.compress_chunk(d, p, sz, True, vin)
            vr2, st_a = ctx.chunk_begin(d, p, sz, True, vin)
            blob2, st_b = ctx.chunk_finish()
            assert (blob2, vr2) == (blob, vr)
            assert st_a["lookups"] == st["lookups"] and st_b["blocks"] == st["blocks"]


def test_all_values_speculation_equals_chunk_from_known_value(ctx):
    """lrzgpu_chunk_begin_all + _select(v) + _finish == lrzgpu_compress_chunk(victim_round = v) for every tested v,
    and the reported table of outgoing counter values agrees with the oracle's."""
    from lrzip_next_b200 import sizing
    from lrzip_next_b200.api import victim_values
    blk = np.frombuffer(b"abcdefg" * 5, dtype=np.uint8)
    d = np.tile(blk, (3 << 20) // blk.size + 1)[:3 << 20].copy()
    rnd = np.random.default_rng(5).integers(0, 256, size=d.size // 8, dtype=np.uint8)
    d[::8] ^= (rnd & 1)  # many equal 31-byte windows => equal-tag chains at max_chain_len => the counter matters
    d = np.concatenate([d, datagen.generate("text", 2 << 20)])
    p = make_params(backend=BACKEND_NONE, threads=1)
    assert victim_values(p) == 16
    sz = sizing(p, d.size)
    table, st_a = ctx.chunk_begin_all(d, p, sz, True)
    assert len(table) == 16
    blobs = {}
    for v in (0, 3, 15):
        if v:
            table2, _ = ctx.chunk_begin_all(d, p, sz, True)
            assert table2 == table
        st_s = ctx.chunk_select(v)
        blob, _ = ctx.chunk_finish()
        want, vr_out, st = ctx.compress_chunk(d, p, sz, True, v)
        assert blob == want and table[v] == vr_out
        assert st_s["chain_evictions"] == st["chain_evictions"] > 0
        _, _, _, ovr = oracle.rzip_chunk(d, 7, victim_round=v)
        assert table[v] == ovr
        blobs[v] = blob
    assert len(set(blobs.values())) > 1, "the input was meant to depend on the incoming counter"
    with pytest.raises(Exception):
        ctx.chunk_select(0)  # nothing pending any more


def test_compress_multi_equals_single_context(ctx):
    """lrzgpu_compress_multi (one process, windows dealt to several contexts, all-values speculation of the
    cross-window counter) == lrzgpu_compress, on text whose windows DO consult the counter.  Two contexts on the
    same device here; on a multi-GPU box they would sit on different devices."""
    from lrzip_next_b200 import Context
    from lrzip_next_b200.api import compress_multi
    d = datagen.generate("text", 230 << 20)   # 3 windows of 100 / 100 / 30 MiB
    p = make_params(backend=BACKEND_NONE, threads=1, window=1)
    want, st1 = ctx.compress(d, p, want_stats=True)
    with Context(0) as c2:
        got, st = compress_multi([ctx, c2], d, p, want_stats=True)
    assert got == want
    assert st["chunks"] == 3 and st["chain_evictions"] == st1["chain_evictions"] > 0
    assert st["lookups"] == st1["lookups"] and st["matches"] == st1["matches"]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_stdin_mode_archive_bit_identical_to_reference_pipe_run(ctx):
    """`cat f | lrzip-next -o out` (STDIN mode: chunks of one mmap buffer, the last one shrunk at EOF, block size from the
    first chunk) against lrzgpu_params.stdin_mode = 1: three chunks, byte-identical; and the small-file
    case where pipe and file runs coincide."""
    d = datagen.generate("text", 80 << 20)
    kw = dict(backend=0, threads=1, ramsize=100 * 1048576)
    want = oracle.ref_compress(d, oracle.make_params(**kw), via_stdin=True)
    got = ctx.compress(d, make_params(stdin_mode=1, **kw))
    assert got == want
    assert got != ctx.compress(d, make_params(**kw))  # the file -> file archive has two larger chunks
    small = d[:20 << 20]
    assert ctx.compress(small, make_params(stdin_mode=1, **kw)) == oracle.ref_compress(small, oracle.make_params(**kw), via_stdin=True)
