"""Development aid: the commit kernel against the oracle with its evaluation switches (LRZGPU_K2_FLAGS:
1 = no twin evaluation, 2 = exact validation of every lane) to localise a divergence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from lrzip_next_b200 import Context, datagen
ctx = Context(0)
cases = [("text", 6 << 20, 7), ("text", 24 << 20, 1), ("trees", 40 << 20, 7)]
for kind, n, level in cases:
    d = datagen.generate(kind, n)
    o0, o1, ost, ovr = oracle.rzip_chunk(d, level)
    for flags in (0, 1, 2, 3):
        os.environ["LRZGPU_K2_FLAGS"] = str(flags)
        s0, s1, st, vr = ctx.rzip_chunk(d, level)
        bad = [k for k in ("inserts", "lookups", "tag_hits", "tag_misses", "chain_evictions", "sweeps", "hash_count", "matches")
               if st[k] != ost[k]]
        print(kind, n, level, "flags", flags, "streams_equal", (s0, s1) == (o0, o1), "vr", vr == ovr, "bad", bad,
              {k: (st[k], ost[k]) for k in bad}, flush=True)
