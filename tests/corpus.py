"""Deterministic synthetic corpus: file i is seeded with PCG64(0x504E4100 + i); content in
64 KiB runs drawn 50% text-like (Zipf 1.1 over a 4096-word vocabulary), 25% structured 64-byte records with 10%
mutated bytes, 25% uniform random."""
from __future__ import annotations

import numpy as np

_VOCAB = None


def _vocab():
    global _VOCAB
    if _VOCAB is None:
        r = np.random.Generator(np.random.PCG64(0x504E41))
        lens = r.integers(2, 11, 4096)
        _VOCAB = [bytes(r.integers(97, 123, int(n), dtype=np.uint8)) for n in lens]
    return _VOCAB


def _text(r, n):
    v = _vocab()
    # Zipf(1.1) ranks clipped to the vocabulary
    ranks = np.minimum(r.zipf(1.1, n // 4 + 16) - 1, 4095)
    seps = r.integers(0, 12, ranks.size)
    out = bytearray()
    for k, s in zip(ranks, seps):
        out += v[int(k)]
        out += b"\n" if s == 0 else b" "
        if len(out) >= n:
            break
    while len(out) < n:
        out += b" "
    return bytes(out[:n])


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
