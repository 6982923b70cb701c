#!/usr/bin/env python
"""Sweep the end-to-end extract pipeline's knobs (worker threads, entry-group size) on the cfg2 shard; one JSON line per point.
Development tool: bench.py carries the chosen defaults."""
import argparse
import importlib
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entries", type=int, default=1024)
    ap.add_argument("--workers", default="2,3,4,6")
    ap.add_argument("--group-mib", default="32,64,128")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    pna = importlib.import_module("portable-network-archive_b200")
    host = importlib.import_module("portable-network-archive_b200._host")
    ctx = pna.Context(0)
    files, streams, key = bench.make_shard(0, a.entries, os.cpu_count() or 8, 1)
    sizes = [len(f) for f in files]
    U = sum(sizes)
    opts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=1, password=bench.PASSWORD, kdf_params={"i": 1000})
    buf = bench.build_archive(streams, sizes, opts.phsf, into=ctx.pinned)
    out = ctx.pinned(U + 16 * a.entries + 64)
    print("mem", os.popen("free -g | head -2 | tail -1").read().strip(), flush=True)
    for w in [int(x) for x in a.workers.split(",")]:
        for g in [int(x) for x in a.group_mib.split(",")]:
            ts = []
            for it in range(a.reps + 1):
                t0 = time.perf_counter()
                ha = host.HostArchive(buf)
                ha.set_key(opts.phsf, key)
                _, offs, st = ha.extract_files(out=out, device=0, workers=w, group_bytes=g << 20, verify=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                ha.close()
                if it:
                    ts.append(dt)
            assert st == [0] * a.entries
            print(json.dumps({"workers": w, "group_mib": g, "ms": [round(t * 1e3, 1) for t in ts], "GBps": U / statistics.median(ts) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
