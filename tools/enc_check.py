"""Encode a few shapes with the GPU deflate / zstd writers and check them with zlib / the oracle -- development aid for the
block writer (run it under compute-sanitizer when a build misbehaves)."""
import importlib, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
pna = importlib.import_module("portable-network-archive_b200")
ctx = pna.Context(0)
SIZES = [0, 1, 3, 4, 15, 16, 17, 31, 32, 33, 255, 4096, 32767, 32768, 32769, 65536, 100_000, 300_001]
def plain(i, n):
    k = i % 4
    if k == 0: return corpus.make_file(500 + i, n)
    if k == 1: return bytes(n)
    if k == 2: return bytes((j * 7919 + (j >> 8) * 31) & 255 for j in range(n))
    return (b"abcdefgh" * (n // 8 + 1))[:n]
ents = [{"plain": plain(i, n), "compression": 1, "level": -1} for i, n in enumerate(SIZES)]
streams, _, st = ctx.encode_batch(ents)
bad = []
for e, s in zip(ents, streams):
    try:
        ok = zlib.decompress(s.tobytes()) == e["plain"]
    except Exception as ex:
        ok = False
    if not ok:
        bad.append(len(e["plain"]))
print("status", st, "bad", bad)
