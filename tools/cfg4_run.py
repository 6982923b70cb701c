"""BASELINE config 4 at quarter size (1024 x 4 MiB create, zstd and deflate) -- development tool.
    python tools/cfg4_run.py [deflate|zstd|both]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import benchlib, benchcfg  # noqa: E402

pna = importlib.import_module("portable-network-archive_b200"); host = importlib.import_module("portable-network-archive_b200._host")
which = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = pna.Context(0)
threads = os.cpu_count() or 8
files = benchlib.gen_files(range(1024), threads)
plain = ctx.pinned(sum(len(f) for f in files)); p_offs = np.zeros(1025, dtype=np.int64); pos = 0
for i, f in enumerate(files):
    plain[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8); pos += len(f); p_offs[i + 1] = pos
for comp, lv, nm in ((1, 6, "deflate"), (2, 3, "zstd")):
    if which not in (nm, "both"):
        continue
    r = benchcfg.cfg4(pna, host, ctx, files, plain, p_offs, threads, 4, comp, lv, nm)
    print(nm, {k: r[k] for k in ("value", "kernel_ms", "stage_ms", "ratio", "c_gpu_over_c_ref")}, "e2e", r["e2e"]["value"], flush=True)
