"""xz decode rate of the GPU path (one warp per stream): N streams of S KiB (text-like corpus, preset 6) through a decode plan,
kernel-only (stage 'inflate' holds the xz kernel), against liblzma on one host core for the same streams.
    python tools/xz_rate.py [n_streams] [kib_per_stream]
"""
import importlib, lzma, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import corpus  # noqa: E402

pna = importlib.import_module("portable-network-archive_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
kib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
base = [corpus.make_file(i, kib * 1024) for i in range(16)]
comp = [lzma.compress(b, preset=6) for b in base]
t0 = time.perf_counter()
for c in comp:
    lzma.decompress(c)
cpu = sum(len(b) for b in base) / (time.perf_counter() - t0) / 1e9
ctx = pna.Context()
entries = [{"bodies": [comp[i % 16]], "compression": 4, "encryption": 0, "cipher_mode": 0, "key": None, "raw_size_hint": kib * 1024}
           for i in range(n)]
plan = ctx.decode_plan(entries)
for _ in range(2):
    plan.run()
ms = plan.stage_ms()
total = n * kib * 1024
print({"streams": n, "kib": kib, "ratio": round(sum(map(len, base)) / sum(map(len, comp)), 2), "xz_ms": ms.get("inflate"),
       "gpu_GBps": round(total / ms["inflate"] / 1e6, 3), "per_stream_MBps": round(kib * 1024 / ms["inflate"] / 1e3, 2),
       "liblzma_1core_GBps": round(cpu, 3)})
