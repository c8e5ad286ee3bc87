"""Quick device-side timings (development aid, not the bench): K1 roofline, CRC, rzip stage."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lrzip_next_b200 import Context, datagen, make_params, BACKEND_NONE

def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

ctx = Context(0)
sizes = [int(x) for x in (sys.argv[1:] or ["256"])]
for mb in sizes:
    n = mb << 20
    h = datagen.gen_text_blocks(n)
    d = torch.zeros(n + 8192 + 256, dtype=torch.uint8, device="cuda")
    d[256:256 + n] = torch.from_numpy(h).cuda()
    base = d.data_ptr() + 256
    cand = torch.empty(((n + 511) // 512) * 512 * 2, dtype=torch.int64, device="cuda")
    tc = torch.empty((n + 511) // 512, dtype=torch.int32, device="cuda")
    crc = torch.zeros(4, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for mask in (1, 15, 255):
        best, avg = ev_time(lambda: ctx.k1_launch(base, n, mask, cand.data_ptr(), tc.data_ptr(), st))
        ncand = int(tc.sum().item())
        alg = n + 16 * ncand
        print(json.dumps({"k": "k1", "mb": mb, "mask": mask, "ms_best": best, "ms_avg": avg, "cands": ncand,
                          "alg_GBps": alg / best / 1e6, "input_GBps": n / best / 1e6}))
    best, avg = ev_time(lambda: ctx.crc32_launch(base, n, crc.data_ptr(), st))
    print(json.dumps({"k": "crc32", "mb": mb, "ms_best": best, "GBps": n / best / 1e6}))
    del cand
    for kind, hh in (("text", h), ("rep", datagen.gen_rep(min(n, 100 << 20)))):
        p = make_params(backend=BACKEND_NONE, threads=1)
        pin = torch.from_numpy(hh).pin_memory()
        for it in range(3):
            t = time.time()
            out, ol, stt = ctx.compress_raw(pin.data_ptr(), pin.numel(), p)
            dt = time.time() - t
            ctx.free(out)
        s = stt.as_dict()
        print(json.dumps({"k": "compress -n", "kind": kind, "mb": pin.numel() >> 20, "wall_s": dt, "MBps": pin.numel() / dt / 1e6,
                          **{k: s[k] for k in ("ms_h2d", "ms_rzip", "ms_emit", "ms_backend", "ms_d2h", "ms_md5", "ms_total", "lookups", "inserts", "matches", "kernel_launches")}}))
