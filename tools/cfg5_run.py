#!/usr/bin/env python
"""Development tool: BASELINE config 5 alone (solid archive, one reference-written zstd frame) at a chosen number of 4 MiB files."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import benchcfg  # noqa: E402
import benchlib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pna = importlib.import_module("portable-network-archive_b200")
host = importlib.import_module("portable-network-archive_b200._host")
ctx = pna.Context(0)
files = benchlib.gen_files(range(n), os.cpu_count() or 8)
print(json.dumps(benchcfg.cfg5(pna, host, ctx, files, os.cpu_count() or 8, 4)), flush=True)
