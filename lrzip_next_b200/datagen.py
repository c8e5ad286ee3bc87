"""Seeded synthetic inputs for the BASELINE.json configs.

Nothing ships with the reference (its test/speedtest.sh needs a user-supplied file), so the
workloads are generated: every generator is a pure function of (size, seed) and returns a
numpy uint8 array. The shapes follow SURVEY.md 8(d):

  rep      C1  one random block repeated (long-range redundancy at a fixed distance)
  text     C2  enwik-style Zipf text over a 50k-word vocabulary
  trees    C3  tar-like stream of N copies of a "source tree" with 0.5 % per-copy byte edits
  randzero C4  random bytes followed by zeros
  vm       C5  VM-image-like mixture of zero runs / repeated 4 KiB blocks / text / random
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_SEPS = [b" ", b" ", b" ", b" ", b" ", b", ", b". ", b"\n", b" [[", b"]] ", b" &quot;"]


def gen_rep(size: int, block: int = 1 << 20, seed: int = 1234) -> np.ndarray:
    """C1: `block` random bytes repeated until `size` bytes."""
    rng = np.random.default_rng(seed)
    blk = rng.integers(0, 256, size=block, dtype=np.uint8)
    reps = -(-size // block)
    return np.tile(blk, reps)[:size].copy()


def _vocab(rng: np.random.Generator, nwords: int):
    lens = np.clip(rng.geometric(0.22, size=nwords) + 1, 2, 14)
    p = (np.arange(1, len(_LETTERS) + 1, dtype=np.float64)) ** -0.8
    p /= p.sum()
    total = int(lens.sum())
    letters = _LETTERS[rng.choice(len(_LETTERS), size=total, p=p)]
    offs = np.concatenate(([0], np.cumsum(lens)))
    return letters, offs.astype(np.int64), lens.astype(np.int64)


def gen_text(size: int, seed: int = 7, nwords: int = 50000, zipf_a: float = 1.05) -> np.ndarray:
    """C2: Zipf(1.05) word draws from a 50k-word vocabulary joined by wiki-ish separators."""
    rng = np.random.default_rng(seed)
    letters, offs, lens = _vocab(rng, nwords)
    seps = [np.frombuffer(s, dtype=np.uint8) for s in _SEPS]
    sep_lens = np.array([len(s) for s in seps], dtype=np.int64)
    sep_cat = np.concatenate(seps)
    sep_offs = np.concatenate(([0], np.cumsum(sep_lens)))
    ranks = np.arange(1, nwords + 1, dtype=np.float64) ** -zipf_a
    cdf = np.cumsum(ranks / ranks.sum())
    out = np.empty(size + 64, dtype=np.uint8)
    pos = 0
    batch = 1 << 20
    while pos < size:
        w = np.searchsorted(cdf, rng.random(batch)).clip(0, nwords - 1)
        s = rng.integers(0, len(seps), size=batch)
        wl = lens[w]
        sl = sep_lens[s]
        tl = wl + sl
        ends = np.cumsum(tl)
        starts = ends - tl
        total = int(ends[-1])
        # per-byte token index and offset inside the token, then gather letters / separator bytes
        tok = np.repeat(np.arange(batch, dtype=np.int64), tl)
        within = np.arange(total, dtype=np.int64) - np.repeat(starts, tl)
        wl_b = np.repeat(wl, tl)
        is_word = within < wl_b
        src = np.where(is_word, np.repeat(offs[w], tl) + within, np.repeat(sep_offs[s], tl) + (within - wl_b))
        buf = np.where(is_word, letters[np.minimum(src, len(letters) - 1)], sep_cat[np.minimum(src, len(sep_cat) - 1)])
        del tok
        n = min(total, size - pos)
        out[pos:pos + n] = buf[:n]
        pos += n
    return out[:size].copy()


def gen_trees(size: int, copies: int = 8, seed: int = 3, edit_rate: float = 0.005) -> np.ndarray:
    """C3: `copies` copies of one synthetic source tree (text files behind 512-byte tar-like
    headers), each copy with `edit_rate` random byte edits."""
    rng = np.random.default_rng(seed)
    tree = -(-size // copies)
    base = gen_text(tree, seed=seed + 100)
    # sprinkle tar-like 512-byte headers every ~16 KiB
    hdr_every = 16384
    for o in range(0, tree - 512, hdr_every):
        base[o:o + 512] = 0
        name = (b"src/dir%03d/file%06d.c" % ((o // hdr_every) % 512, o // hdr_every))
        base[o:o + len(name)] = np.frombuffer(name, dtype=np.uint8)
        base[o + 100:o + 108] = np.frombuffer(b"0000644\0", dtype=np.uint8)
    out = np.empty(tree * copies, dtype=np.uint8)
    for c in range(copies):
        seg = out[c * tree:(c + 1) * tree]
        seg[:] = base
        if c:
            ne = int(tree * edit_rate)
            where = rng.integers(0, tree, size=ne)
            seg[where] = rng.integers(0, 256, size=ne, dtype=np.uint8)
    return out[:size].copy()


def gen_randzero(size: int, seed: int = 4) -> np.ndarray:
    """C4: first half random bytes, second half zeros."""
    rng = np.random.default_rng(seed)
    out = np.zeros(size, dtype=np.uint8)
    half = size // 2
    out[:half] = rng.integers(0, 256, size=half, dtype=np.uint8)
    return out


def gen_mix(size: int, seed: int = 11) -> np.ndarray:
    """random | zeros | the same random again (Appendix F 'mix' shape)."""
    rng = np.random.default_rng(seed)
    q = size // 4
    r = rng.integers(0, 256, size=q, dtype=np.uint8)
    out = np.zeros(size, dtype=np.uint8)
    out[:q] = r
    out[size - q:] = r
    return out


def gen_vm(size: int, seed: int = 5) -> np.ndarray:
    """C5: VM-image-like mixture: 40 % zero runs, 35 % repeated 4 KiB blocks from a pool,
    15 % text, 10 % random; laid out in 64 KiB extents."""
    rng = np.random.default_rng(seed)
    ext = 65536
    next_ = -(-size // ext)
    out = np.zeros(next_ * ext, dtype=np.uint8)
    pool_blocks = max(16, min(1 << 19, (size // 16) // 4096))
    pool = rng.integers(0, 256, size=(pool_blocks, 4096), dtype=np.uint8)
    text = gen_text(max(ext * 4, size // 8), seed=seed + 1)
    kinds = rng.choice(4, size=next_, p=[0.40, 0.35, 0.15, 0.10])
    tpos = 0
    for e in range(next_):
        k = kinds[e]
        seg = out[e * ext:(e + 1) * ext]
        if k == 1:
            ids = rng.integers(0, pool_blocks, size=ext // 4096)
            seg[:] = pool[ids].reshape(-1)
        elif k == 2:
            if tpos + ext > len(text):
                tpos = 0
            seg[:] = text[tpos:tpos + ext]
            tpos += ext
        elif k == 3:
            seg[:] = rng.integers(0, 256, size=ext, dtype=np.uint8)
    return out[:size].copy()


def _text_block(args):
    size, seed = args
    return gen_text(size, seed=seed)


def gen_text_blocks(size: int, seed: int = 7, block: int = 64 << 20, workers: int | None = None) -> np.ndarray:
    """C2 at full size: `size` bytes of enwik-style text as independent `block`-byte pieces (seed,
    seed+1, ...) generated in parallel worker processes -- same bytes for any worker count."""
    import multiprocessing as mp
    import os
    nb = -(-size // block)
    jobs = [(min(block, size - i * block), seed + i) for i in range(nb)]
    workers = workers or min(nb, max(1, (os.cpu_count() or 2) - 1), 32)
    if workers <= 1 or nb == 1:
        parts = [_text_block(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_text_block, jobs)
    return np.concatenate(parts)


GENERATORS = {
    "rep": gen_rep,
    "text": gen_text,
    "trees": gen_trees,
    "randzero": gen_randzero,
    "mix": gen_mix,
    "vm": gen_vm,
}


def generate(kind: str, size: int, **kw) -> np.ndarray:
    return GENERATORS[kind](size, **kw)
