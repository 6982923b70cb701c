"""BASELINE config 1 alone (1 GiB, 10k files, zstd 3, no encryption: create + extract) with the host layer's trace on stderr.
    PNA_HOST_TRACE=1 python tools/cfg1_run.py [workers] [group_mib]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import benchlib, corpus  # noqa: E402

pna = importlib.import_module("portable-network-archive_b200")
host = importlib.import_module("portable-network-archive_b200._host")
workers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gmib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = pna.Context(0)
files = benchlib.gen_files(range(256), os.cpu_count() or 8)
corpus_np = np.concatenate([np.frombuffer(f, dtype=np.uint8) for f in files])
n, total = 10000, 1 << 30
sizes = corpus.lognormal_sizes(n, total)
offs = np.zeros(n + 1, dtype=np.int64); np.cumsum(sizes, out=offs[1:])
plain = ctx.pinned(total); plain[:] = corpus_np[:total]
names = [f"c/{i:06d}" for i in range(n)]
views = [plain[int(offs[i]):int(offs[i + 1])] for i in range(n)]
arch = ctx.pinned(int(total * 1.05) + (8 << 20))
fl = list(zip(names, views))
for it in range(4):
    t0 = time.perf_counter()
    blob = host.create_archive((names, plain, offs) if os.environ.get('PACKED', '1') == '1' else fl, compression=2, level=3, max_chunk_size=0, device=0, workers=workers, group_bytes=gmib << 20, out=arch)
    dt = time.perf_counter() - t0
    print(json.dumps({"create_ms": round(dt * 1e3, 1), "GBps": round(total / dt / 1e9, 2), "archive": int(blob.size), "workers": workers, "group_mib": gmib}), flush=True)
    print("----", file=sys.stderr, flush=True)
out = ctx.pinned(total + 16 * n + 64)
for it in range(3):
    t0 = time.perf_counter()
    ha = host.HostArchive(blob)
    _, xo, st = ha.extract_files(out=out, device=0, workers=4, group_bytes=64 << 20, verify=True)
    ha.close()
    dt = time.perf_counter() - t0
    print(json.dumps({"extract_ms": round(dt * 1e3, 1), "GBps": round(total / dt / 1e9, 2)}), flush=True)
    print("----", file=sys.stderr, flush=True)
assert st == [0] * n
