"""K2 commit timing / round statistics on text (development aid, not the bench).

  python tools/k2_probe.py [MiB ...] [--all]      (--all also times the all-values form, chunk_begin_all)
Prints ms_rzip and the commit kernel's bookkeeping counters (ScanState.dbg, lrz_common.h) via LRZGPU_DEBUG."""
import json, os, sys, time
os.environ["LRZGPU_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrzip_next_b200 import BACKEND_NONE, Context, datagen, make_params, sizing

args = [a for a in sys.argv[1:] if not a.startswith("--")]
do_all = "--all" in sys.argv
ctx = Context(0)
for mb in [int(x) for x in (args or ["256"])]:
    n = mb << 20
    h = datagen.gen_text_blocks(n)
    pin = torch.from_numpy(h).pin_memory()
    p = make_params(backend=BACKEND_NONE, threads=1)
    sz = sizing(p, n)
    t = time.time()
    vr, st = ctx.chunk_begin(pin, p, sz, True, 0)
    dt = time.time() - t
    ctx.chunk_finish()
    print(json.dumps({"k": "chunk_begin", "mb": mb, "wall_s": dt, "MBps": n / dt / 1e6, "ms_rzip": st["ms_rzip"],
                      "lookups": st["lookups"], "evictions": st["chain_evictions"], "vr_out": vr}), flush=True)
    if do_all:
        t = time.time()
        table, st = ctx.chunk_begin_all(pin, p, sz, True)
        dt = time.time() - t
        ctx.chunk_select(0)
        ctx.chunk_finish()
        print(json.dumps({"k": "chunk_begin_all", "mb": mb, "wall_s": dt, "MBps": n / dt / 1e6, "ms_rzip": st["ms_rzip"],
                          "table": table, "agrees_with_single": table[0] == vr}), flush=True)
