"""ctypes binding of include/lrzgpu.h (the reference-facing C ABI).

Names follow the reference's own vocabulary: ``compress`` stands for ``rzip_fd`` + ``write_magic``
(src/rzip.c:922, src/lrzip.c:131), ``rzip_chunk`` for ``hash_search`` (src/rzip.c:586),
``block_compress`` for ``lzma_compress_buf`` / ``zstd_compress_buf`` (src/stream.c:429/167),
``lz4_gate`` for ``lz4_compresses`` (src/stream.c:2325).  Errors raise LrzGpuError carrying the
library's code and message; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

BACKEND_NONE, BACKEND_LZMA, BACKEND_ZSTD = 0, 1, 4
CTYPE_NONE, CTYPE_LZMA, CTYPE_ZSTD = 3, 6, 10

_HERE = os.path.dirname(os.path.abspath(__file__))


class LrzGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lrzgpu error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("level", C.c_int), ("rzip_level", C.c_int), ("backend", C.c_int), ("threads", C.c_int),
                ("window", C.c_int), ("unlimited", C.c_int), ("ramsize", C.c_int64), ("page_size", C.c_int),
                ("processors", C.c_int), ("threshold", C.c_int), ("nobemt", C.c_int),
                ("filter", C.c_int), ("delta", C.c_int), ("stdin_mode", C.c_int)]


class Sizing(C.Structure):
    _fields_ = [("threads", C.c_int), ("dict_size", C.c_uint32), ("overhead", C.c_int64),
                ("bufsize", C.c_int64), ("max_chunk", C.c_int64)]


class Stats(C.Structure):
    _fields_ = ([(n, C.c_int64) for n in (
        "matches", "match_bytes", "literals", "literal_bytes", "tag_hits", "tag_misses", "inserts", "lookups",
        "chain_evictions", "sweeps", "displacements", "hash_count", "final_min_mask", "final_tag_mask",
        "chunks", "blocks", "blocks_stored", "stream0_bytes", "stream1_bytes")]
        + [("crc32", C.c_uint32), ("pad", C.c_uint32)]
        + [(n, C.c_double) for n in ("ms_h2d", "ms_rzip", "ms_emit", "ms_backend", "ms_d2h", "ms_md5", "ms_total")]
        + [("kernel_launches", C.c_int64)])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "pad"}


def make_params(level=7, rzip_level=0, backend=BACKEND_NONE, threads=1, window=0, unlimited=0,
                ramsize=100 * 100 * 1048576, page_size=4096, processors=8, threshold=100, nobemt=0,
                filter=0, delta=0, stdin_mode=0) -> Params:
    return Params(level, rzip_level, backend, threads, window, unlimited, ramsize, page_size, processors,
                  threshold, nobemt, filter, delta, stdin_mode)


class ArchiveInfo(C.Structure):
    """lrzgpu_archive_info (include/lrzgpu.h): what `lrzip-next -i` reports."""
    _fields_ = [("major", C.c_int), ("minor", C.c_int), ("expected_size", C.c_int64), ("hash_type", C.c_int),
                ("encrypted", C.c_int), ("filter", C.c_int), ("delta", C.c_int), ("backend_code", C.c_int),
                ("backend_prop", C.c_int), ("lzma_dict_size", C.c_uint32), ("rzip_level", C.c_int), ("level", C.c_int),
                ("chunks", C.c_int64), ("blocks", C.c_int64), ("stream_c_bytes", C.c_int64 * 2),
                ("stream_u_bytes", C.c_int64 * 2), ("blocks_by_ctype", C.c_int64 * 16), ("archive_bytes", C.c_int64),
                ("md5", C.c_uint8 * 16)]


class BlockInfo(C.Structure):
    _fields_ = [("chunk", C.c_int64), ("stream", C.c_int), ("ctype", C.c_int), ("c_len", C.c_int64), ("u_len", C.c_int64),
                ("offset", C.c_int64), ("next_head", C.c_int64)]


def archive_info(archive: bytes):
    """get_fileinfo (src/lrzip.c:1069) on an archive in memory: (ArchiveInfo, [BlockInfo]).  Host only."""
    L = load_library()
    L.lrzgpu_info.argtypes = [C.c_char_p, C.c_int64, C.POINTER(ArchiveInfo), C.POINTER(BlockInfo), C.c_int64,
                              C.POINTER(C.c_int64)]
    info, nb = ArchiveInfo(), C.c_int64()
    rc = L.lrzgpu_info(archive, len(archive), C.byref(info), None, 0, C.byref(nb))
    if rc:
        raise LrzGpuError(f"lrzgpu_info: error {rc}")
    arr = (BlockInfo * max(1, nb.value))()
    rc = L.lrzgpu_info(archive, len(archive), C.byref(info), arr, nb.value, C.byref(nb))
    if rc:
        raise LrzGpuError(f"lrzgpu_info: error {rc}")
    return info, list(arr[:nb.value])


def lib_path() -> str:
    return os.path.join(_HERE, "liblrzgpu.so")


def build_library() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)


_lib = None

EXPORTS = [
    "lrzgpu_create", "lrzgpu_destroy", "lrzgpu_last_error", "lrzgpu_free", "lrzgpu_version", "lrzgpu_sizing",
    "lrzgpu_compress", "lrzgpu_compress_file", "lrzgpu_compress_device", "lrzgpu_compress_multi", "lrzgpu_compress_chunk",
    "lrzgpu_chunk_begin", "lrzgpu_chunk_finish", "lrzgpu_victim_values", "lrzgpu_chunk_begin_all", "lrzgpu_chunk_select",
    "lrzgpu_decompress", "lrzgpu_info", "lrzgpu_rzip_chunk", "lrzgpu_tag_scan", "lrzgpu_crc32", "lrzgpu_block_compress", "lrzgpu_lz4_gate",
    "lrzgpu_k1_launch", "lrzgpu_crc32_launch", "lrzgpu_sm_count",
]


def load_library():
    """Load liblrzgpu.so (must have been built: __graft_entry__.build() or make -C csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    # more hardware work queues than the default 8: the backend keeps one stream per group of blocks in flight
    # (only read by the CUDA runtime when it creates its context, so it must be set before the first CUDA call)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    path = lib_path()
    if not os.path.exists(path):
        raise LrzGpuError(-4, f"{path} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i64, pi64, pvp = C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_void_p)
    L.lrzgpu_create.argtypes = [C.c_int, pvp]
    L.lrzgpu_destroy.argtypes = [vp]
    L.lrzgpu_destroy.restype = None
    L.lrzgpu_last_error.argtypes = [vp]
    L.lrzgpu_last_error.restype = C.c_char_p
    L.lrzgpu_free.argtypes = [vp]
    L.lrzgpu_free.restype = None
    L.lrzgpu_version.restype = C.c_char_p
    L.lrzgpu_sizing.argtypes = [C.POINTER(Params), i64, C.POINTER(Sizing)]
    L.lrzgpu_compress.argtypes = [vp, C.POINTER(Params), vp, i64, pvp, pi64, C.POINTER(Stats)]
    L.lrzgpu_compress_file.argtypes = [vp, C.POINTER(Params), C.c_char_p, C.c_char_p, C.POINTER(Stats)]
    L.lrzgpu_compress_device.argtypes = [vp, C.POINTER(Params), vp, i64, vp, pvp, pi64, C.POINTER(Stats)]
    L.lrzgpu_compress_multi.argtypes = [pvp, C.c_int, C.POINTER(Params), vp, i64, pvp, pi64, C.POINTER(Stats)]
    L.lrzgpu_compress_chunk.argtypes = [vp, C.POINTER(Params), C.POINTER(Sizing), vp, i64, C.c_int, pi64, pvp, pi64,
                                        C.POINTER(Stats)]
    L.lrzgpu_chunk_begin.argtypes = [vp, C.POINTER(Params), C.POINTER(Sizing), vp, i64, C.c_int, pi64, C.POINTER(Stats)]
    L.lrzgpu_chunk_finish.argtypes = [vp, pvp, pi64, C.POINTER(Stats)]
    L.lrzgpu_victim_values.argtypes = [C.POINTER(Params)]
    L.lrzgpu_chunk_begin_all.argtypes = [vp, C.POINTER(Params), C.POINTER(Sizing), vp, i64, C.c_int, pi64, C.c_int,
                                         C.POINTER(Stats)]
    L.lrzgpu_chunk_select.argtypes = [vp, i64, C.POINTER(Stats)]
    L.lrzgpu_decompress.argtypes = [vp, vp, i64, pvp, pi64]
    L.lrzgpu_rzip_chunk.argtypes = [vp, vp, i64, C.c_int, C.c_int, pi64, pvp, pi64, pvp, pi64, C.POINTER(Stats)]
    L.lrzgpu_tag_scan.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, i64, pi64]
    L.lrzgpu_crc32.argtypes = [vp, vp, i64, C.POINTER(C.c_uint32)]
    L.lrzgpu_block_compress.argtypes = [vp, C.POINTER(Params), C.c_uint32, vp, i64, pvp, pi64, C.POINTER(C.c_int)]
    L.lrzgpu_lz4_gate.argtypes = [vp, vp, i64, C.c_int, C.POINTER(C.c_int)]
    L.lrzgpu_k1_launch.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.lrzgpu_crc32_launch.argtypes = [vp, vp, i64, vp, vp]
    L.lrzgpu_sm_count.argtypes = [vp]
    _lib = L
    return L


def sizing(params: Params, st_size: int) -> Sizing:
    s = Sizing()
    rc = load_library().lrzgpu_sizing(C.byref(params), st_size, C.byref(s))
    if rc:
        raise LrzGpuError(rc, "lrzgpu_sizing rejected the parameters")
    return s


def victim_values(params: Params) -> int:
    """Number of values the reference's cross-window counter can take (max_chain_len of the rzip level)."""
    return load_library().lrzgpu_victim_values(C.byref(params))


def compress_multi(contexts, data, params: Params, want_stats: bool = False):
    """lrzgpu_compress_multi: one process, several contexts (GPUs); windows dealt round-robin."""
    L = load_library()
    addr, n, keep = _ptr(data)
    arr = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    out, ol, st = C.c_void_p(), C.c_int64(), Stats()
    rc = L.lrzgpu_compress_multi(arr, len(contexts), C.byref(params), addr, n, C.byref(out), C.byref(ol), C.byref(st))
    if rc:
        raise LrzGpuError(rc, (L.lrzgpu_last_error(contexts[0].handle) or b"").decode())
    res = C.string_at(out, ol.value)
    L.lrzgpu_free(out)
    return (res, st.as_dict()) if want_stats else res


def chunk_bytes_for(n: int) -> int:
    bits = 8
    while n >> bits > 0:
        bits += 1
    return bits // 8 + (1 if bits % 8 else 0)


def _ptr(data) -> tuple[int, int, object]:
    """(address, nbytes, keepalive) of a host buffer: bytes, numpy uint8 array or a pinned torch tensor."""
    if isinstance(data, (bytes, bytearray)):
        a = np.frombuffer(data, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    if hasattr(data, "data_ptr"):  # torch tensor on the host
        return data.data_ptr(), data.numel() * data.element_size(), data
    raise TypeError(type(data))


class Context:
    """One lrzgpu context == one GPU (lrzgpu_create)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.lrzgpu_create(device, C.byref(h))
        if rc:
            raise LrzGpuError(rc, "lrzgpu_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.lrzgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def _check(self, rc: int):
        if rc:
            raise LrzGpuError(rc, (self._L.lrzgpu_last_error(self._h) or b"").decode())

    def _take(self, p: C.c_void_p, n: int) -> bytes:
        data = C.string_at(p, n) if n else b""
        self._L.lrzgpu_free(p)
        return data

    @property
    def sm_count(self) -> int:
        return self._L.lrzgpu_sm_count(self._h)

    # ---- whole archive (rzip_fd + write_magic) ---------------------------------------------------
    def compress(self, data, params: Params, want_stats: bool = False):
        addr, n, keep = _ptr(data)
        out, ol, st = C.c_void_p(), C.c_int64(), Stats()
        self._check(self._L.lrzgpu_compress(self._h, C.byref(params), addr, n, C.byref(out), C.byref(ol), C.byref(st)))
        res = self._take(out, ol.value)
        return (res, st.as_dict()) if want_stats else res

    def compress_raw(self, addr: int, n: int, params: Params):
        """Like compress() but returns (pointer, length, Stats) without copying the archive (bench)."""
        out, ol, st = C.c_void_p(), C.c_int64(), Stats()
        self._check(self._L.lrzgpu_compress(self._h, C.byref(params), addr, n, C.byref(out), C.byref(ol), C.byref(st)))
        return out, ol.value, st

    def compress_device_raw(self, d_addr: int, n: int, params: Params, md5: bytes | None = None):
        out, ol, st = C.c_void_p(), C.c_int64(), Stats()
        self._check(self._L.lrzgpu_compress_device(self._h, C.byref(params), d_addr, n, md5, C.byref(out), C.byref(ol),
                                                   C.byref(st)))
        return out, ol.value, st

    def free(self, p):
        self._L.lrzgpu_free(p)

    def decompress(self, archive) -> bytes:
        """runzip_fd on the device: archive bytes -> original bytes (CRC and MD5 verified)."""
        addr, n, keep = _ptr(archive)
        out, ol = C.c_void_p(), C.c_int64()
        self._check(self._L.lrzgpu_decompress(self._h, addr, n, C.byref(out), C.byref(ol)))
        return self._take(out, ol.value)

    def compress_file(self, src: str, dst: str, params: Params) -> dict:
        st = Stats()
        self._check(self._L.lrzgpu_compress_file(self._h, C.byref(params), src.encode(), dst.encode(), C.byref(st)))
        return st.as_dict()

    def compress_chunk(self, data, params: Params, sz: Sizing, eof: bool, victim_round: int = 0):
        """One window -> (blob, victim_round_out, stats): the unit sharded across GPUs."""
        addr, n, keep = _ptr(data)
        out, ol, st, vr = C.c_void_p(), C.c_int64(), Stats(), C.c_int64(victim_round)
        self._check(self._L.lrzgpu_compress_chunk(self._h, C.byref(params), C.byref(sz), addr, n, int(eof), C.byref(vr),
                                                  C.byref(out), C.byref(ol), C.byref(st)))
        return self._take(out, ol.value), vr.value, st.as_dict()

    def chunk_begin(self, data, params: Params, sz: Sizing, eof: bool, victim_round: int = 0):
        """rzip stage of one window -> (victim_round_out, stats); the window stays pending for chunk_finish()."""
        addr, n, keep = _ptr(data)
        st, vr = Stats(), C.c_int64(victim_round)
        self._check(self._L.lrzgpu_chunk_begin(self._h, C.byref(params), C.byref(sz), addr, n, int(eof), C.byref(vr),
                                               C.byref(st)))
        return vr.value, st.as_dict()

    def chunk_begin_all(self, data, params: Params, sz: Sizing, eof: bool):
        """rzip stage of one window for EVERY incoming victim_round value at once -> (victim_out list, stats):
        victim_out[v] is the counter after the window when it started at v.  chunk_select(v) then picks one."""
        addr, n, keep = _ptr(data)
        nv = victim_values(params)
        out = (C.c_int64 * nv)()
        st = Stats()
        self._check(self._L.lrzgpu_chunk_begin_all(self._h, C.byref(params), C.byref(sz), addr, n, int(eof), out, nv,
                                                   C.byref(st)))
        return list(out), st.as_dict()

    def chunk_select(self, victim_in: int) -> dict:
        """Emit the streams of the variant that started from `victim_in`; chunk_finish() follows."""
        st = Stats()
        self._check(self._L.lrzgpu_chunk_select(self._h, victim_in, C.byref(st)))
        return st.as_dict()

    def chunk_finish(self):
        """backend + framing of the pending window -> (blob, stats)."""
        out, ol, st = C.c_void_p(), C.c_int64(), Stats()
        self._check(self._L.lrzgpu_chunk_finish(self._h, C.byref(out), C.byref(ol), C.byref(st)))
        return self._take(out, ol.value), st.as_dict()

    # ---- scan primitives -------------------------------------------------------------------------
    def rzip_chunk(self, data, rzip_level: int = 7, chunk_bytes: int | None = None, victim_round: int = 0):
        addr, n, keep = _ptr(data)
        if chunk_bytes is None:
            chunk_bytes = chunk_bytes_for(n)
        s0, s1, l0, l1 = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        vr, st = C.c_int64(victim_round), Stats()
        self._check(self._L.lrzgpu_rzip_chunk(self._h, addr, n, rzip_level, chunk_bytes, C.byref(vr), C.byref(s0),
                                              C.byref(l0), C.byref(s1), C.byref(l1), C.byref(st)))
        return self._take(s0, l0.value), self._take(s1, l1.value), st.as_dict(), vr.value

    def tag_scan(self, data, mask: int, pos_lo: int = 0, pos_hi: int | None = None):
        addr, n, keep = _ptr(data)
        pos_hi = n if pos_hi is None else pos_hi
        cap = max(1, pos_hi - pos_lo)
        pos = np.empty(cap, dtype=np.int64)
        tag = np.empty(cap, dtype=np.int64)
        cnt = C.c_int64()
        self._check(self._L.lrzgpu_tag_scan(self._h, addr, n, pos_lo, pos_hi, mask, pos.ctypes.data, tag.ctypes.data, cap,
                                            C.byref(cnt)))
        return pos[:cnt.value], tag[:cnt.value]

    def crc32(self, data) -> int:
        addr, n, keep = _ptr(data)
        v = C.c_uint32()
        self._check(self._L.lrzgpu_crc32(self._h, addr, n, C.byref(v)))
        return v.value

    # ---- backends --------------------------------------------------------------------------------
    def block_compress(self, data, params: Params, dict_size: int = 0):
        addr, n, keep = _ptr(data)
        out, cl, ct = C.c_void_p(), C.c_int64(), C.c_int()
        self._check(self._L.lrzgpu_block_compress(self._h, C.byref(params), dict_size, addr, n, C.byref(out), C.byref(cl),
                                                  C.byref(ct)))
        return self._take(out, cl.value), ct.value

    def lz4_gate(self, data, threshold: int = 100) -> int:
        addr, n, keep = _ptr(data)
        r = C.c_int()
        self._check(self._L.lrzgpu_lz4_gate(self._h, addr, n, threshold, C.byref(r)))
        return r.value

    # ---- measurement hooks -----------------------------------------------------------------------
    def k1_launch(self, d_buf: int, n: int, mask: int, d_cand: int, d_tile_count: int, stream: int = 0):
        self._check(self._L.lrzgpu_k1_launch(self._h, d_buf, n, mask, d_cand, d_tile_count, stream))

    def crc32_launch(self, d_buf: int, n: int, d_crc: int, stream: int = 0):
        self._check(self._L.lrzgpu_crc32_launch(self._h, d_buf, n, d_crc, stream))
