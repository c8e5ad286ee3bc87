#!/usr/bin/env python
"""Randomised CPU check of the batched commit kernel (k2_commit.cu under tests/hostsim/simt.h) against the scalar
control logic: adversarial byte streams (zero runs, short periods, duplicated and edited blocks), shrunken hash tables,
every rzip level, incoming victim_round values, segment sizes, development switches.

  python tools/k2_simt_fuzz.py [--cases N] [--jobs J] [--seed S] [--max-kib K]
"""
import argparse
import ctypes as C
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HS = os.path.join(ROOT, "tests", "hostsim")
_S = None


def lib():
    global _S
    if _S is None:
        _S = C.CDLL(os.path.join(HS, "libk2simt.so"))
        _S.simt_rzip_chunk.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int64, C.c_int,
                                       C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    return _S


def run(d, level, seg, bits, flags, mode, vr):
    S = lib()
    v, s0, s1, l0, l1 = C.c_int64(vr), C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
    st, dbg = (C.c_int64 * 10)(), (C.c_int64 * 16)()
    rc = S.simt_rzip_chunk(d.ctypes.data, d.size, level, 4, C.byref(v), seg, bits, flags, mode, C.byref(s0), C.byref(l0),
                           C.byref(s1), C.byref(l1), st, dbg)
    if rc:
        return ("rc", rc), list(dbg)
    a, b = C.string_at(s0, l0.value), C.string_at(s1, l1.value)
    S.simt_free(s0)
    S.simt_free(s1)
    return (a, b, v.value, list(st)), list(dbg)


def make_data(rng, n):
    """A concatenation of pieces of different character."""
    from lrzip_next_b200 import datagen
    out = []
    left = n
    pool = []
    while left > 0:
        k = int(min(left, rng.integers(2000, max(2001, n // 3))))
        kind = rng.integers(0, 9)
        if kind == 0:
            piece = np.zeros(k, dtype=np.uint8)
        elif kind == 1:
            piece = rng.integers(0, 256, size=k, dtype=np.uint8)
        elif kind == 2:  # short period
            per = int(rng.integers(1, 70))
            piece = np.resize(rng.integers(0, 256, size=per, dtype=np.uint8), k)
        elif kind == 3:
            piece = datagen.gen_text(k, seed=int(rng.integers(0, 1 << 30)))
        elif kind in (4, 5) and pool:  # an earlier piece again, with edits
            src = pool[int(rng.integers(0, len(pool)))]
            piece = np.resize(src, k).copy()
            ne = int(k * rng.choice([0.0, 0.0005, 0.005, 0.03]))
            if ne:
                piece[rng.integers(0, k, size=ne)] = rng.integers(0, 256, size=ne, dtype=np.uint8)
        elif kind == 6:  # few symbols
            piece = rng.integers(0, int(rng.integers(2, 5)), size=k, dtype=np.uint8)
        elif kind == 7:  # long period
            per = int(rng.integers(100, 5000))
            piece = np.resize(rng.integers(0, 256, size=per, dtype=np.uint8), k)
        else:
            piece = datagen.gen_text(k, seed=int(rng.integers(0, 1 << 30)))
        pool.append(piece)
        out.append(piece)
        left -= k
    return np.ascontiguousarray(np.concatenate(out)[:n])


def one(args):
    seed, max_kib = args
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, max_kib * 1024))
    if rng.integers(0, 3):
        d = make_data(rng, n)
    else:  # plain text: the sweeps, gate tightening and same-tag runs of the headline workload
        from lrzip_next_b200 import datagen
        d = datagen.gen_text(n, seed=int(rng.integers(0, 1 << 30)))
    level = int(rng.choice([1, 3, 5, 6, 7, 7, 7, 8, 9]))
    bits = int(rng.integers(8, 16))
    seg = int(rng.choice([4096, 12288, 1 << 16, 1 << 18, 1 << 22]))
    vr = int(rng.integers(0, [1, 2, 2, 2, 3, 4, 6, 16, 32, 128][level]))  # the counter stays below max_chain_len
    flags = int(rng.choice([0, 0, 0, 1, 2, 4, 8, 12]))
    want, _ = run(d, level, seg, bits, 0, 0, vr)
    got, dbg = run(d, level, seg, bits, flags, 1, vr)
    ok = want == got
    return seed, ok, dict(n=n, level=level, bits=bits, seg=seg, vr=vr, flags=flags), dbg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=64)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-kib", type=int, default=768)
    a = ap.parse_args()
    subprocess.run(["make", "-s", "-C", HS, "libk2simt.so"], check=True)
    bad = 0
    tot = np.zeros(16, dtype=np.int64)
    with mp.Pool(a.jobs) as pool:
        for seed, ok, cfg, dbg in pool.imap_unordered(one, [(a.seed * 100003 + i, a.max_kib) for i in range(a.cases)]):
            tot += np.array(dbg, dtype=np.int64)
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, cfg, flush=True)
    print(f"{a.cases - bad}/{a.cases} equal; rounds fresh {tot[0]} resumed {tot[7]} committed {tot[1]} serial "
          f"{tot[2] + tot[3] + tot[4]} tails {tot[15]} conflicts {tot[5]}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
