"""Backend timing probe (development aid): one text window through lrzgpu_compress with the LZMA backend, with and
without the backend running under the scan, plus a single 10 MiB block alone."""
import json, os, sys, time
os.environ["LRZGPU_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrzip_next_b200 import BACKEND_LZMA, Context, datagen, make_params

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = Context(0)
h = datagen.gen_text_blocks(mb << 20)
pin = torch.from_numpy(h).pin_memory()
p = make_params(level=7, backend=BACKEND_LZMA, threads=160, ramsize=600 * 100 * 1048576, processors=os.cpu_count() or 16)
blk = bytes(h[:10 << 20])
t = time.time()
got, ct = ctx.block_compress(blk, make_params(level=7, backend=BACKEND_LZMA, threads=8), 1 << 25)
print(json.dumps({"k": "one 10 MiB block alone", "s": time.time() - t, "out": len(got)}), flush=True)
for mode in ("overlap", "plain", "overlap"):
    if mode == "plain":
        os.environ["LRZGPU_NO_OVERLAP"] = "1"
    else:
        os.environ.pop("LRZGPU_NO_OVERLAP", None)
    t = time.time()
    out, ol, st = ctx.compress_raw(pin.data_ptr(), pin.numel(), p)
    dt = time.time() - t
    ctx.free(out)
    s = st.as_dict()
    print(json.dumps({"k": mode, "mb": mb, "wall_s": dt, "MBps": pin.numel() / dt / 1e6, "blocks": s["blocks"],
                      **{k: round(s[k]) for k in ("ms_rzip", "ms_emit", "ms_backend", "ms_d2h", "ms_total")}}), flush=True)
