"""Development aid: the create_xz configuration of benchcfg.py alone (N x 4 MiB), with a per-stage diagnosis when the read-back fails."""
import importlib, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import corpus, benchcfg
pna = importlib.import_module("portable-network-archive_b200")
host = importlib.import_module("portable-network-archive_b200._host")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = pna.Context(0)
files = [corpus.make_file(1000 + i, 4 << 20) for i in range(n)]
plain = ctx.pinned(n * (4 << 20))
offs = np.arange(n + 1, dtype=np.int64) * (4 << 20)
for i, f in enumerate(files):
    plain[int(offs[i]):int(offs[i + 1])] = np.frombuffer(f, dtype=np.uint8)
try:
    t0 = time.perf_counter()
    r = benchcfg.create_xz(pna, host, ctx, files, plain, offs, len(os.sched_getaffinity(0)))
    print({k: r[k] for k in ("value", "kernel_ms", "stage_ms", "e2e", "ratio", "c_gpu_over_c_ref", "cpu_baseline")}, round(time.perf_counter() - t0, 1), "s")
except Exception:
    traceback.print_exc()
