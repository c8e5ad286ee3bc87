"""Repeat the whole LZMA pipeline on small inputs and localise any difference from the reference (development aid)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
from lrzip_next_b200 import BACKEND_LZMA, Context, datagen, make_params

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = Context(0)
for kind, n, kw in (("text", 1_500_000, dict(threads=8)), ("mix", 2_000_000, dict(threads=8))):
    d = datagen.generate(kind, n)
    want = oracle.ref_compress(d, oracle.make_params(backend=oracle.BACKEND_LZMA, **kw))
    o0, o1, _, _ = oracle.rzip_chunk(d, 7)
    p = make_params(backend=BACKEND_LZMA, **kw)
    bad = 0
    for i in range(reps):
        got = ctx.compress(d, p)
        s0, s1, st, vr = ctx.rzip_chunk(d, 7)
        ok_r = (s0 == o0 and s1 == o1)
        if got != want or not ok_r:
            bad += 1
            m = min(len(got), len(want))
            a = np.frombuffer(got[:m], dtype=np.uint8) != np.frombuffer(want[:m], dtype=np.uint8)
            first = int(np.flatnonzero(a)[0]) if a.any() else m
            # block-level: each stream's blocks through the single-block ABI
            blk = []
            for name, s in (("s0", o0), ("s1", o1)):
                ref = oracle.ref_lzma_block(s, 7, 1 << 25, 2)
                for j in range(3):
                    g, ct = ctx.block_compress(s, p, 1 << 25)
                    blk.append((name, j, (g == ref) if ref is not None else (ct == 3)))
            print(kind, "iter", i, "archive_ok", got == want, "rzip_ok", ok_r, "len", len(got), len(want), "first_diff", first, "blocks", blk, flush=True)
    print(kind, "bad", bad, "of", reps, flush=True)
