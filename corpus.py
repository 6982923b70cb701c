"""Deterministic synthetic corpus: file i is seeded with PCG64(0x504E4100 + i); content in
64 KiB runs drawn 50% text-like (Zipf 1.1 over a 4096-word vocabulary, space/newline separated), 25% structured
64-byte records with 10% mutated bytes and a little-endian u32 counter, 25% uniform random.  Fully vectorised
(numpy) so that a 4 GiB shard is generated in about a minute on a few cores."""
from __future__ import annotations

import numpy as np

_VOCAB = None


def _vocab():
    """flat byte array of all words, each followed by one separator slot; start offsets; lengths incl. separator"""
    global _VOCAB
    if _VOCAB is None:
        r = np.random.Generator(np.random.PCG64(0x504E41))
        lens = r.integers(2, 11, 4096)
        tot = int(lens.sum() + 4096)
        flat = np.empty(tot, dtype=np.uint8)
        starts = np.zeros(4096, dtype=np.int64)
        pos = 0
        for i, n in enumerate(lens):
            starts[i] = pos
            flat[pos:pos + n] = r.integers(97, 123, int(n), dtype=np.uint8)
            flat[pos + n] = 32
            pos += int(n) + 1
        _VOCAB = (flat, starts, (lens + 1).astype(np.int64))
    return _VOCAB


def _text(r, n):
    flat, starts, lens = _vocab()
    nw = n // 3 + 8
    ranks = np.minimum(r.zipf(1.1, nw) - 1, 4095)
    wl = lens[ranks]
    ends = np.cumsum(wl)
    k = int(np.searchsorted(ends, n)) + 1
    ranks, wl, ends = ranks[:k], wl[:k], ends[:k]
    total = int(ends[-1])
    begin = ends - wl
    src = np.repeat(starts[ranks] - begin, wl) + np.arange(total, dtype=np.int64)
    out = flat[src]
    nl = r.integers(0, 12, k) == 0
    out[ends[nl] - 1] = 10
    return out[:n].tobytes()


def _records(r, n):
    tmpl = r.integers(0, 256, 64, dtype=np.uint8)
    cnt = (n + 63) // 64
    a = np.tile(tmpl, cnt).reshape(cnt, 64).copy()
    mut = r.random((cnt, 64)) < 0.10
    a[mut] = r.integers(0, 256, int(mut.sum()), dtype=np.uint8)
    a[:, :4] = np.arange(cnt, dtype=np.uint32).view(np.uint8).reshape(cnt, 4)
    return a.tobytes()[:n]


def make_file(i: int, size: int) -> bytes:
    r = np.random.Generator(np.random.PCG64(0x504E4100 + i))
    out = bytearray()
    while len(out) < size:
        n = min(65536, size - len(out))
        k = r.integers(0, 4)
        if k < 2:
            out += _text(r, n)
        elif k == 2:
            out += _records(r, n)
        else:
            out += r.integers(0, 256, n, dtype=np.uint8).tobytes()
    return bytes(out)


def lognormal_sizes(count: int, total: int, seed: int = 1) -> np.ndarray:
    """cfg1 sizes: log-normal (mu = ln 64 KiB, sigma = 1) rescaled to sum to `total`."""
    r = np.random.Generator(np.random.PCG64(seed))
    s = r.lognormal(np.log(65536.0), 1.0, count)
    s = np.maximum(1, (s * (total / s.sum())).astype(np.int64))
    s[-1] += total - int(s.sum())
    return s
