import importlib, os, sys, zlib, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import corpus
if os.environ.get("EMU"):
    _ffi = importlib.import_module("portable-network-archive_b200._ffi")
    _ffi.LIB_PATH = os.path.join(ROOT, "tests/emu/_gen/libpna_cuda.so")
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
res = []
for e, s in zip(ents, streams):
    b = s.tobytes()
    try:
        ok = zlib.decompress(b) == e["plain"]
    except Exception as ex:
        ok = str(ex)
    res.append((len(e["plain"]), len(b), ok))
print(st); print(res)
pickle.dump([s.tobytes() for s in streams], open(sys.argv[1], "wb"))
