"""xz create path on the GPU: N x 4 MiB bench-corpus files, GPU LZMA2 (one chunk per 32 KiB segment) + AES-256-CTR + CRC, timed per
stage with CUDA events; liblzma (the reference's encoder, preset 6, all host cores) beside it on a bounded sample; every sampled
stream read back by liblzma, all of them by our own decoder."""
import argparse, importlib, json, lzma, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import corpus
pna = importlib.import_module("portable-network-archive_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--entries", type=int, default=256)
ap.add_argument("--ref-entries", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--level", type=int, default=6)
ap.add_argument("--serial", type=int, default=0, help="also decode liblzma-written streams of the same files (serial decoder; takes ~1 min of host time to prepare)")
a = ap.parse_args()
ctx = pna.Context(0)
FILE = 4 << 20
files = [corpus.make_file(1000 + i, FILE) for i in range(a.entries)]
KEY = bytes(range(32))
ents = [{"plain": f, "compression": 4, "level": a.level, "encryption": 1, "cipher_mode": 1, "key": KEY, "iv": os.urandom(16), "max_chunk_size": 0} for f in files]
plan = ctx.encode_plan(ents)
plan.run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = torch.cuda.current_stream()
ts = []
for _ in range(a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter(); plan.run(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
stage = plan.stage_ms()
streams, crcs, st = plan.fetch()
assert st == [0] * len(ents)
c_gpu = sum(int(s.size) for s in streams)
import pna_oracle as O
for k in range(0, a.entries, max(1, a.entries // 8)):
    assert O.decode_stream(streams[k].tobytes(), 4, 1, 1, KEY, None) == files[k]
back, st2, _ = ctx.decode_batch([{"bodies": [s], "compression": 4, "encryption": 1, "cipher_mode": 1, "key": KEY, "raw_size_hint": FILE} for s in streams])
assert st2 == [0] * len(ents) and all(b.tobytes() == f for b, f in zip(back, files))
# extract of what we wrote: chunk-parallel xz decode (a warp per 32 KiB window), kernel-only through a decode plan
dplan = ctx.decode_plan([{"bodies": [s], "compression": 4, "encryption": 1, "cipher_mode": 1, "key": KEY, "raw_size_hint": FILE} for s in streams])
for _ in range(2):
    dplan.run()
dms = dplan.stage_ms()
# the same files written by liblzma (one dictionary for the whole stream: the serial decoder, a warp per stream)
nser = min(a.entries, 64)
ser = [lzma.compress(f, preset=a.level) for f in files[:nser]] if a.serial else []
sms = None
if ser:
    splan = ctx.decode_plan([{"bodies": [c], "compression": 4, "encryption": 0, "cipher_mode": 0, "key": None, "raw_size_hint": FILE} for c in ser])
    for _ in range(2):
        splan.run()
    sms = splan.stage_ms()
nref = min(a.ref_entries, a.entries)
cores = len(os.sched_getaffinity(0))
t0 = time.perf_counter()
with ThreadPoolExecutor(cores) as ex:
    ref = list(ex.map(lambda f: len(lzma.compress(f, preset=a.level)), files[:nref]))
t_ref = time.perf_counter() - t0
c_ref = sum(ref); c_gpu_s = sum(int(s.size) - 16 for s in streams[:nref])
ms = min(ts)
print(json.dumps({"workload": f"create {a.entries} x 4 MiB, GPU xz (LZMA2 chunk per 32 KiB) + aes-256-ctr + crc32", "ms_per_step": ms,
                  "GBps": a.entries * FILE / ms / 1e6, "stage_ms": stage, "ratio": a.entries * FILE / c_gpu, "c_gpu_over_c_ref": c_gpu_s / c_ref,
                  "cpu_liblzma": {"GBps": nref * FILE / t_ref / 1e9, "cores": cores, "preset": a.level, "sample": f"{nref} x 4 MiB"},
                  "extract_back": {"stage_ms": dms, "xz_GBps": a.entries * FILE / dms["inflate"] / 1e6, "how": "decode plan over the streams written above: chunk-parallel pass + container check"},
                  "extract_liblzma_written": None if sms is None else {"streams": nser, "xz_ms": sms["inflate"], "xz_GBps": nser * FILE / sms["inflate"] / 1e6},
                  "checked": "sampled streams by liblzma + OpenSSL, all streams by our xz decode kernel"}))
