#!/usr/bin/env python
"""Development check: run benchcfg.run_all at a tiny scale against the SIMT emulator build (tests/emu) on a box without a GPU,
so that the bench plumbing (input preparation, framing, oracle legs, parity checks) is debugged before GPU minutes are spent."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "emu")):
    sys.path.insert(0, p)
import build_emu  # noqa: E402

gen = build_emu.build()
ffi = importlib.import_module("portable-network-archive_b200._ffi")
host = importlib.import_module("portable-network-archive_b200._host")
ffi.LIB_PATH = os.path.join(gen, "libpna_cuda.so")
host.LIB_PATH = os.path.join(gen, "libpna_host.so")
pna = importlib.import_module("portable-network-archive_b200")
import benchcfg  # noqa: E402
import benchlib  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1 / 512
ctx = pna.Context(0)
files = benchlib.gen_files(range(max(8, int(4096 * scale))), 8)
res = benchcfg.run_all(pna, host, ctx, files, 8, scale, 2)
print(json.dumps(res, indent=1)[:6000])
assert not any("error" in v for v in res.values() if isinstance(v, dict)), "a configuration failed"
