"""BASELINE config 3 alone (many small files, zlib 6 + Camellia-256-CBC, extract) with the host layer's trace on stderr.
    PNA_HOST_TRACE=1 python tools/cfg3_run.py [scale] [workers] [group_mib]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import benchlib, benchcfg  # noqa: E402

pna = importlib.import_module("portable-network-archive_b200")
host = importlib.import_module("portable-network-archive_b200._host")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gmib = int(sys.argv[3]) if len(sys.argv) > 3 else 64
threads = os.cpu_count() or 8
ctx = pna.Context(0)
n = max(1024, int((1 << 20) * scale))
rng = np.random.Generator(np.random.PCG64(3))
sizes = rng.integers(1, 16385, n).astype(np.int64)
offs = np.zeros(n + 1, dtype=np.int64); np.cumsum(sizes, out=offs[1:])
U = int(offs[-1])
files = benchlib.gen_files(range((U >> 22) + 2), threads)
corpus_np = np.concatenate([np.frombuffer(f, dtype=np.uint8) for f in files])
opts = pna.WriteOptions(compression=1, encryption=2, cipher_mode=0, password=b"pw", kdf_params={"i": 1000})
streams, s_offs, _ = benchlib.oracle_encode(corpus_np, offs, 1, 6, 2, 0, benchcfg.KEY, threads)
buf = benchlib.frame_archive(streams, s_offs, sizes, bytes([0, 0, 0, 1, 2, 0]), opts.phsf, 16, "s/%07d", ctx.pinned, threads)
out = ctx.pinned(U + 16 * n + 64)
for it in range(3):
    t0 = time.perf_counter()
    ha = host.HostArchive(buf)
    t1 = time.perf_counter()
    ha.set_key(opts.phsf, benchcfg.KEY)
    _, xo, st = ha.extract_files(out=out, device=0, workers=workers, group_bytes=gmib << 20, verify=True)
    t2 = time.perf_counter()
    ha.close()
    dt = time.perf_counter() - t0
    print(json.dumps({"n": n, "bytes": U, "ms": round(dt * 1e3, 1), "index_ms": round((t1 - t0) * 1e3, 1), "extract_ms": round((t2 - t1) * 1e3, 1),
                      "close_ms": round((time.perf_counter() - t2) * 1e3, 1), "GBps": round(U / dt / 1e9, 2), "workers": workers, "group_mib": gmib}), flush=True)
    print("----", file=sys.stderr, flush=True)
assert st == [0] * n
