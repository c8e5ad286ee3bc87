"""Device LZMA block encoder speed on one text block (development aid)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrzip_next_b200 import Context, datagen, make_params, BACKEND_LZMA
ctx = Context(0)
p = make_params(level=7, backend=BACKEND_LZMA, threads=8, threshold=0)
for kb in [int(x) for x in (sys.argv[1:] or ["512", "2048"])]:
    d = datagen.gen_text_blocks(kb << 10)
    t = time.time()
    got, ct = ctx.block_compress(d, p, 1 << 25)
    dt = time.time() - t
    print(json.dumps({"k": "lzma_block", "kb": kb, "out": len(got), "ctype": ct, "s": dt, "KBps": kb / dt}), flush=True)
