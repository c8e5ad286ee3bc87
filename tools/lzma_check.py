"""Device LZMA block encoder: bit-exactness against the reference's LzmaCompress on a spread of inputs, then
speed on one text block (development aid; the tests proper are tests/test_gpu_backend.py).
usage: python tools/lzma_check.py [speed_kb ...]"""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from lrzip_next_b200 import Context, datagen, make_params, BACKEND_LZMA

ctx = Context(0)
rng = np.random.default_rng(5)
inputs = {
    "text": datagen.gen_text_blocks(700_000).tobytes(),
    "text2": datagen.gen_text(300_000).tobytes(),
    "trees": datagen.generate("trees", 600_000).tobytes(),
    "mix": datagen.generate("mix", 500_000).tobytes(),
    "vm": datagen.generate("vm", 500_000).tobytes(),
    "lowent": rng.integers(0, 3, 150_000, dtype=np.uint8).tobytes(),
    "skew": (rng.integers(0, 256, 200_000, dtype=np.uint8) & rng.integers(0, 256, 200_000, dtype=np.uint8)).astype(np.uint8).tobytes(),
    "zeros": bytes(300_000),
    "rep": (bytes(rng.integers(0, 256, 5000, dtype=np.uint8)) * 40),
    "tiny": b"hello hello hello hello hello hello hello hello hello hello hello!",
}
bad = 0
for level, dict_size in [(7, 1 << 25), (5, 1 << 24), (9, 1 << 27)]:
    p = make_params(level=level, backend=BACKEND_LZMA, threads=8, threshold=0)
    for name, d in inputs.items():
        if level != 7 and name in ("vm", "skew", "text2"):
            continue
        want = oracle.ref_lzma_block(d, level, dict_size, 2)
        got, ctype = ctx.block_compress(d, p, dict_size)
        ok = (ctype == 3 and got == d) if want is None else (ctype == 6 and got == want)
        if not ok:
            bad += 1
            print(f"MISMATCH L{level} {name}: got {len(got)} want {None if want is None else len(want)}", flush=True)
print("parity:", "OK" if bad == 0 else f"{bad} MISMATCHES", flush=True)
p = make_params(level=7, backend=BACKEND_LZMA, threads=8, threshold=0)
for kb in [int(x) for x in (sys.argv[1:] or ["1024"])]:
    d = datagen.gen_text_blocks(kb << 10)
    t = time.time()
    got, ct = ctx.block_compress(d, p, 1 << 25)
    dt = time.time() - t
    print(json.dumps({"k": "lzma_block", "kb": kb, "out": len(got), "ctype": ct, "s": dt, "KBps": kb / dt}), flush=True)
