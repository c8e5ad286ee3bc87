"""CPU tests of the checker itself: the C restatement (oracle/liboracle.so) against the committed golden
vectors of the reference (tests/golden/golden.json) and, where the compiled reference is present
(oracle/_ref, this container and the GPU box), against the reference binary directly."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from lrzip_next_b200 import datagen

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as fh:
    GOLDEN = json.load(fh)


@pytest.mark.parametrize("name", sorted(GOLDEN["archives"]))
def test_oracle_archive_matches_golden(name):
    g = GOLDEN["archives"][name]
    d = datagen.generate(g["kind"], g["n"], **g["gen"])
    assert hashlib.sha256(d.tobytes()).hexdigest() == g["input_sha256"], "generator drifted"
    p = oracle.make_params(**g["params"])
    if g["params"]["backend"] == oracle.BACKEND_LZMA:
        if not oracle.have_ref():
            pytest.skip("LZMA payloads need oracle/_ref/liblzmaref.so")
        sz = oracle.sizing(p, g["n"])
        arc, _ = oracle.compress(d, p, oracle.lzma_block_fn(p.level, sz.dict_size, p.threshold))
    else:
        arc, _ = oracle.compress(d, p)
    assert len(arc) == g["archive_len"]
    assert hashlib.sha256(arc).hexdigest() == g["archive_sha256"]
    if "archive_hex" in g:
        assert arc.hex() == g["archive_hex"]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,n,kw", [
    ("text", 3 << 20, dict(level=3)), ("trees", 5 << 20, dict(level=9)), ("randzero", 2 << 20, dict(level=1)),
    ("text", 9 << 20, dict(window=1, ramsize=100 * 1048576 * 2)), ("rep", 70_000, {}), ("text", 31, {}), ("text", 4096, {}),
    # twins of the multi-chunk / multi-block GPU cases (tests/test_gpu_rzip.py): victim_round carried between
    # chunks, eof on the last chunk only, flush order and next_head patching over several blocks per stream
    ("text", 250 << 20, dict(window=1)),              # 3 chunks of 100 / 100 / 50 MiB
    ("text", 80 << 20, dict(ramsize=100 * 1048576)),  # 2 chunks x 3 blocks
])
def test_oracle_equals_reference_binary_stored(kind, n, kw):
    d = datagen.generate(kind, n)
    p = oracle.make_params(backend=oracle.BACKEND_NONE, threads=1, **kw)
    assert oracle.compress(d, p)[0] == oracle.ref_compress(d, p)


def test_hash_index_constants():
    hi = oracle.hash_index()
    # glibc random() with the default seed: first four entries and the OR of all
    assert [hex(int(v)) for v in hi[:4]] == ["0x6b8b771c23c6", "0x643cfe5a4873", "0x74b0c5185cff", "0x2ae8f61f58ec"]
    assert int(np.bitwise_or.reduce(hi)) == 0x7FFFFFFFFFFF


def test_appendix_f_rep100_facts():
    # 100 MiB of a repeated 1 MiB block, -n -p1 => 1,059,858 bytes, 1585 matches
    d = datagen.gen_rep(100 << 20)
    arc, st = oracle.compress(d, oracle.make_params(backend=oracle.BACKEND_NONE, threads=1))
    assert len(arc) == 1_059_858
    assert (st["matches"], st["match_bytes"], st["literals"], st["literal_bytes"], st["inserts"]) == \
           (1585, 103_808_993, 19, 1_048_607, 524_840)
    assert arc[-16:].hex() == "535d792a42ccc7c552714d2a97bf158f"
