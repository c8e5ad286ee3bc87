"""K2 commit A/B on the GPU (development aid): the same window scanned with different LRZGPU_K2_FLAGS.

  python tools/k2_ab_probe.py kind:MiB[:flags,flags...] ...      e.g.  trees:600:12,0 text:400:12,0
Prints ms_rzip, the rzip statistics (must not depend on the flags) and the kernel's counters (LRZGPU_DEBUG)."""
import json, os, sys, time
os.environ["LRZGPU_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrzip_next_b200 import BACKEND_NONE, Context, datagen, make_params, sizing

ctx = Context(0)
for spec in sys.argv[1:]:
    f = spec.split(":")
    kind, mb = f[0], int(f[1])
    flags = [int(x) for x in f[2].split(",")] if len(f) > 2 else [12, 0]
    n = mb << 20
    h = {"text": lambda: datagen.gen_text_blocks(n), "trees": lambda: datagen.gen_trees(n, seed=3),
         "rep": lambda: datagen.gen_rep(n), "randzero": lambda: datagen.gen_randzero(n), "vm": lambda: datagen.gen_vm(n)}[kind]()
    pin = torch.from_numpy(h).pin_memory()
    p = make_params(backend=BACKEND_NONE, threads=1)
    sz = sizing(p, n)
    ref = None
    for fl in flags:
        os.environ["LRZGPU_K2_FLAGS"] = str(fl)
        t = time.time()
        vr, st = ctx.chunk_begin(pin, p, sz, True, 0)
        dt = time.time() - t
        ctx.chunk_finish()
        key = {k: st[k] for k in ("matches", "match_bytes", "literals", "literal_bytes", "inserts", "lookups", "tag_hits",
                                  "tag_misses", "chain_evictions", "sweeps")}
        key["vr_out"] = vr
        ref = ref or key
        print(json.dumps({"kind": kind, "mb": mb, "flags": fl, "ms_rzip": st["ms_rzip"], "wall_s": round(dt, 2),
                          "same_as_first": key == ref, **key}), flush=True)
