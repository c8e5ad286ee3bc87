"""ctypes loader for the CPU oracle (oracle/liboracle.so) and a runner for the compiled
reference (oracle/_ref/lrzip-next).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "lrzip-next")
REF_LZMA = os.path.join(ORACLE_DIR, "_ref", "liblzmaref.so")

BACKEND_NONE, BACKEND_LZMA, BACKEND_ZSTD = 0, 1, 4


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "matches", "match_bytes", "literals", "literal_bytes", "tag_hits", "tag_misses", "inserts",
        "lookups", "chain_evictions", "sweeps",
        "displacements", "max_depth", "insert_probes", "lookup_probes", "max_probe", "probes_ge32", "hash_count", "final_min_mask", "final_tag_mask")] + [("crc32", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Params(C.Structure):
    _fields_ = [("level", C.c_int), ("rzip_level", C.c_int), ("backend", C.c_int), ("threads", C.c_int),
                ("window", C.c_int), ("unlimited", C.c_int), ("ramsize", C.c_int64), ("page_size", C.c_int),
                ("processors", C.c_int), ("threshold", C.c_int), ("nobemt", C.c_int)]


class Sizing(C.Structure):
    _fields_ = [("threads", C.c_int), ("dict_size", C.c_uint32), ("overhead", C.c_int64),
                ("bufsize", C.c_int64), ("max_chunk", C.c_int64)]


BLOCK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_int64, C.c_int,
                       C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int))

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
        L = C.CDLL(path)
        L.rzo_rzip_chunk.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Stats)]
        L.rzo_rzip_chunk.restype = C.c_int
        L.rzo_free.argtypes = [C.c_void_p]
        L.rzo_hash_index.argtypes = [C.c_void_p]
        L.rzo_full_tag.argtypes = [C.c_void_p, C.c_int64]
        L.rzo_full_tag.restype = C.c_int64
        L.rzo_sizing_compute.restype = C.c_int
        L.rzo_sizing_compute.argtypes = [C.POINTER(Params), C.c_int64, C.POINTER(Sizing)]
        L.rzo_compress.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int64, BLOCK_FN, C.c_void_p,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Stats)]
        L.rzo_compress.restype = C.c_int
        L.rzo_md5.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def hash_index() -> np.ndarray:
    hi = np.zeros(256, dtype=np.int64)
    lib().rzo_hash_index(hi.ctypes.data)
    return hi


def _take(ptr: C.c_void_p, n: int) -> bytes:
    data = C.string_at(ptr, n) if n else b""
    lib().rzo_free(ptr)
    return data


def rzip_chunk(data: np.ndarray, rzip_level: int = 7, chunk_bytes: int | None = None, victim_round: int = 0):
    """Oracle rzip of one chunk -> (stream0 bytes, stream1 bytes, stats dict, victim_round out)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = int(data.size)
    if chunk_bytes is None:
        chunk_bytes = chunk_bytes_for(n)
    vr = C.c_int64(victim_round)
    s0, s1 = C.c_void_p(), C.c_void_p()
    l0, l1 = C.c_int64(), C.c_int64()
    st = Stats()
    rc = lib().rzo_rzip_chunk(data.ctypes.data, n, rzip_level, chunk_bytes, C.byref(vr),
                              C.byref(s0), C.byref(l0), C.byref(s1), C.byref(l1), C.byref(st))
    if rc:
        raise RuntimeError(f"rzo_rzip_chunk rc={rc}")
    return _take(s0, l0.value), _take(s1, l1.value), st.as_dict(), vr.value


def chunk_bytes_for(n: int) -> int:
    bits = 8
    while n >> bits > 0:
        bits += 1
    return bits // 8 + (1 if bits % 8 else 0)


def make_params(level=7, rzip_level=0, backend=BACKEND_NONE, threads=1, window=0, unlimited=0,
                ramsize=100 * 100 * 1048576, page_size=4096, processors=8, threshold=100, nobemt=0) -> Params:
    return Params(level, rzip_level, backend, threads, window, unlimited, ramsize, page_size,
                  processors, threshold, nobemt)


def sizing(params: Params, st_size: int) -> Sizing:
    s = Sizing()
    if lib().rzo_sizing_compute(C.byref(params), st_size, C.byref(s)):
        raise ValueError("no encoder fits the RAM budget (the reference would crash)")
    return s


def ref_lzma_block(data: bytes, level: int, dict_size: int, threads: int = 2):
    """LzmaCompress of the vendored SDK (src/lzma/C/LzmaLib.c:12) as called by
    lzma_compress_buf (src/stream.c:443-456). Returns payload bytes or None (stored)."""
    L = C.CDLL(REF_LZMA)
    n = len(data)
    dlen = int(n * 1.02)
    dlen += (-dlen) % 4096
    dst = C.create_string_buffer(max(dlen, 1))
    dl = C.c_size_t(dlen)
    props = C.create_string_buffer(5)
    ps = C.c_size_t(5)
    L.LzmaCompress.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_void_p,
                               C.POINTER(C.c_size_t), C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    rc = L.LzmaCompress(dst, C.byref(dl), data, n, props, C.byref(ps), level, dict_size, 3, 0, 2,
                        32 if level < 7 else 64, threads)
    if rc != 0 or dl.value >= n:
        return None
    return dst.raw[:dl.value]


def compress(data: np.ndarray, params: Params, block_fn=None):
    """Oracle whole-archive compress (stored blocks unless block_fn given)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out, ol = C.c_void_p(), C.c_int64()
    st = Stats()
    cb = BLOCK_FN(block_fn) if block_fn else C.cast(None, BLOCK_FN)
    rc = lib().rzo_compress(C.byref(params), data.ctypes.data, int(data.size), cb, None,
                            C.byref(out), C.byref(ol), C.byref(st))
    if rc:
        raise RuntimeError(f"rzo_compress rc={rc}")
    return _take(out, ol.value), st.as_dict()


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def ref_flags(params: Params):
    f = ["-Q", "-f", f"-L{params.level}", f"-p{params.threads}", f"-m{params.ramsize // (100 * 1048576)}"]
    if params.rzip_level:
        f.append(f"-R{params.rzip_level}")
    if params.backend == BACKEND_NONE:
        f.append("-n")
    elif params.backend == BACKEND_ZSTD:
        f.append("-Z")
    if params.window:
        f.append(f"-w{params.window}")
    if params.unlimited:
        f.append("-U")
    if params.threshold == 0:
        f.append("-T")
    elif params.threshold != 100:
        f.append(f"-T{params.threshold}")
    if params.nobemt:
        f.append("--nobemt")
    return f


def ref_compress(data, params: Params, workdir: str | None = None, extra=(), via_stdin: bool = False) -> bytes:
    """Run the unmodified reference binary (oracle/_ref/lrzip-next) file -> file, or pipe -> file (via_stdin: the
    reference's STDIN mode, where the input's size is unknown while it is read)."""
    env = dict(os.environ, LRZIP="NOCONFIG")
    with tempfile.TemporaryDirectory(dir=workdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)) as d:
        src = os.path.join(d, "in.bin")
        dst = os.path.join(d, "out.lrz")
        if isinstance(data, (bytes, bytearray)):
            with open(src, "wb") as fh:
                fh.write(data)
        else:
            np.ascontiguousarray(data, dtype=np.uint8).tofile(src)
        if via_stdin:
            with open(src, "rb") as fin:
                subprocess.run([REF_BIN, *ref_flags(params), *extra, "-o", dst], check=True, env=env, stdin=fin,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        else:
            subprocess.run([REF_BIN, *ref_flags(params), *extra, "-o", dst, src], check=True, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(dst, "rb") as fh:
            return fh.read()


def ref_test(archive: bytes) -> bool:
    """`lrzip-next -t` (structural walk + CRC + MD5) on an archive; True when it verifies."""
    env = dict(os.environ, LRZIP="NOCONFIG")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        a = os.path.join(d, "a.lrz")
        with open(a, "wb") as fh:
            fh.write(archive)
        r = subprocess.run([REF_BIN, "-t", "-Q", a], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return r.returncode == 0


def ref_decompress(archive: bytes) -> bytes:
    env = dict(os.environ, LRZIP="NOCONFIG")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        a = os.path.join(d, "a.lrz")
        o = os.path.join(d, "a.out")
        with open(a, "wb") as fh:
            fh.write(archive)
        subprocess.run([REF_BIN, "-d", "-Q", "-f", "-o", o, a], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(o, "rb") as fh:
            return fh.read()


# ---- lz4 gate reference: lz4_compresses() of src/stream.c:2325-2380 over the system liblz4 ----------
_lz4 = None


def _liblz4():
    global _lz4
    if _lz4 is None:
        for name in ("liblz4.so.1", "/lib/x86_64-linux-gnu/liblz4.so.1"):
            try:
                _lz4 = C.CDLL(name)
                break
            except OSError:
                continue
        if _lz4 is None:
            raise RuntimeError("liblz4.so.1 not found")
        _lz4.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    return _lz4


def have_lz4() -> bool:
    try:
        _liblz4()
        return True
    except RuntimeError:
        return False


def ref_lz4_gate(data: bytes, threshold: int = 100) -> int:
    L = _liblz4()
    s_len = len(data)
    test_len = s_len
    in_len = min(test_len, 100 * 1048576)
    buftest = in_len
    pct = 101.0
    dst = C.create_string_buffer(in_len + 2)
    while test_len > 0:
        ret = L.LZ4_compress_default(data, dst, in_len, in_len + 1)
        if ret > 0:
            pct = 100 * (ret / in_len)
            if ret < in_len * (threshold / 100):
                break
        test_len -= in_len
        if test_len > 0:
            buftest += in_len
            if buftest < 10 * 1048576:
                buftest <<= 1
            in_len = min(test_len, buftest)
            dst = C.create_string_buffer(in_len + 2)
    return 0 if pct > threshold else (int(pct + 1) if pct < 1 else int(pct))


def lzma_block_fn(level: int, dict_size: int, threshold: int = 100):
    """Block callback for oracle.compress(): lz4 gate + the reference's LzmaCompress."""
    def fn(user, src, u_len, stream, out, out_cap, c_len, c_type):
        data = C.string_at(src, u_len)
        if threshold and have_lz4() and not ref_lz4_gate(data, threshold):
            return 0
        payload = ref_lzma_block(data, level, dict_size, 2)
        if payload is not None:
            C.memmove(out, payload, len(payload))
            c_len[0] = len(payload)
            c_type[0] = 6
        return 0
    return fn


def walk_blocks(archive: bytes):
    """Block headers of a v0.14 archive in file order: [(chunk, ctype, c_len, u_len)] (SURVEY.md Appendix A:
    21-byte magic, per chunk [cb][eof][chunk_size:cb], two initial stream headers, then blocks
    [ctype][c_len:cb][u_len:cb][next_head:cb] + payload; 16-byte MD5 at the end)."""
    out, pos, chunk = [], 21 + archive[20], 0
    end = len(archive) - 16
    while pos < end:
        cb, eof = archive[pos], archive[pos + 1]
        pos += 2 + cb
        initial = pos
        hdr = 1 + 3 * cb
        pos += 2 * hdr
        # blocks follow back to back until the chunk ends: the last block of stream 1 is physically last
        # (src/lrzip.c:1276), so walk until both streams have seen a header with next_head == 0
        open_streams = 2
        heads = {}
        for s in range(2):
            nh = int.from_bytes(archive[initial + s * hdr + 1 + 2 * cb: initial + s * hdr + hdr], "little")
            heads[initial + nh] = s
        while open_streams and pos < end:
            ctype = archive[pos]
            c_len = int.from_bytes(archive[pos + 1:pos + 1 + cb], "little")
            u_len = int.from_bytes(archive[pos + 1 + cb:pos + 1 + 2 * cb], "little")
            nh = int.from_bytes(archive[pos + 1 + 2 * cb:pos + hdr], "little")
            s = heads.pop(pos)
            out.append((chunk, s, ctype, c_len, u_len))
            if nh:
                heads[initial + nh] = s
            else:
                open_streams -= 1
            pos += hdr + c_len
        chunk += 1
        if eof:
            break
    return out
