"""Regenerates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Each case pins, for a seeded synthetic input and a pinned option tuple (-L/-p/-m/-w/backend), the
SHA-256 and size of the reference's .lrz archive, the rzip stream sizes, and for small cases the archive
bytes themselves (hex).  LZMA block vectors pin LzmaCompress() (src/lzma/C/LzmaLib.c:12) directly."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from lrzip_next_b200 import datagen  # noqa: E402

ARCHIVES = [
    # name, generator, size, params
    ("rep_n", "rep", 3 << 20, dict(backend=0, threads=1), dict(block=1 << 18)),
    ("text_n", "text", 1 << 20, dict(backend=0, threads=1), {}),
    ("text_n_l3", "text", 2 << 20, dict(backend=0, threads=1, level=3), {}),
    ("mix_n_w", "mix", 6 << 20, dict(backend=0, threads=1, window=1, ramsize=1 * 100 * 1048576), {}),
    ("vm_n_l9", "vm", 4 << 20, dict(backend=0, threads=1, level=9), {}),
    ("text_lzma", "text", 600_000, dict(backend=1, threads=8), {}),
    ("trees_lzma_l5", "trees", 1 << 20, dict(backend=1, threads=2, level=5), {}),
    ("tiny_n", "text", 100, dict(backend=0, threads=1), {}),
]
LZMA_BLOCKS = [
    ("text", 200_000, 7, 1 << 25), ("text", 200_000, 5, 1 << 24), ("vm", 1 << 20, 7, 1 << 25), ("mix", 400_000, 9, 1 << 27),
]


def main():
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"reference": "pete4abw/lrzip-next v0.14.0 (a465d8c), gcc -O2, libzstd 1.5.5 / liblz4 1.9.4 runtime",
           "archives": {}, "lzma_blocks": []}
    for name, kind, n, pk, gk in ARCHIVES:
        d = datagen.generate(kind, n, **gk)
        p = oracle.make_params(**pk)
        arc = oracle.ref_compress(d, p)
        s0, s1, st, _ = oracle.rzip_chunk(d, pk.get("level", 7)) if not pk.get("window") else (b"", b"", {}, 0)
        e = {"kind": kind, "n": n, "gen": gk, "params": pk, "archive_len": len(arc),
             "archive_sha256": hashlib.sha256(arc).hexdigest(), "input_sha256": hashlib.sha256(d.tobytes()).hexdigest()}
        if len(arc) <= 4096:
            e["archive_hex"] = arc.hex()
        out["archives"][name] = e
        print(name, len(arc))
    for kind, n, level, dic in LZMA_BLOCKS:
        d = datagen.generate(kind, n).tobytes()
        pay = oracle.ref_lzma_block(d, level, dic, 2)
        out["lzma_blocks"].append({"kind": kind, "n": n, "level": level, "dict": dic,
                                   "len": None if pay is None else len(pay),
                                   "sha256": None if pay is None else hashlib.sha256(pay).hexdigest()})
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
