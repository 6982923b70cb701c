#!/usr/bin/env python
"""bench.py -- extract GB/s (uncompressed) of the PNA data-chunk hot path on B200, vs the host-CPU path.

Workload (BASELINE.json configs[1], weak-scaled): each GPU extracts one shard of `--entries` x 4 MiB
files, zstd level 3 + AES-256-CTR, layout FHED,fSIZ,PHSF,FDAT(16),FDAT(C),FEND.  A "step" = one pass of
the whole hot path over the shard: CRC-32 check of every chunk, AES-256-CTR decrypt, zstd decode.

  value  kernel-only: archive already resident in HBM, CUDA events on the library's stream, max over ranks
  e2e    the same through the C ABI with HOST buffers: pinned archive -> H2D -> kernels -> D2H pinned outputs
  roofline      dominant stage, algorithmic bytes / event time vs the measured HBM copy peak
  cpu_baseline  the oracle (reference dataflow restated on libzstd/OpenSSL/zlib) on the host cores

`--impl reference` times that CPU path alone (the reference is Rust; there is no cargo in this image).
Inputs are synthetic (corpus.py), produced outside every timed region by the oracle's encoder, i.e. the
same libzstd level-3 streaming frames the reference writes.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import struct
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import corpus  # noqa: E402

FILE_SIZE = 4 << 20
PASSWORD = b"bench-password"


def _gen(i):
    return corpus.make_file(i, FILE_SIZE)


def make_shard(rank: int, entries: int, threads: int, world: int = 1):
    """Plain files + oracle-encoded streams (zstd 3 + AES-256-CTR) of this rank's shard.  Not timed.
    The global corpus has world x entries files; the rank's share comes from the same LPT partition by entry that the
    extract path uses (portable-network-archive_b200/shard.py) -- no collective, every rank derives it alone."""
    import multiprocessing as mp
    import pna_oracle as O
    shard = importlib.import_module("portable-network-archive_b200.shard")
    idx = shard.rank_entries([FILE_SIZE] * (entries * world), rank, world)
    with mp.get_context("fork").Pool(max(1, min(threads, 64))) as pool:
        files = pool.map(_gen, idx, chunksize=4)
    key = bytes(range(32))
    rng = np.random.Generator(np.random.PCG64(1234 + rank))
    L = O.lib()
    jobs = (O.EncJob * entries)()
    outs = []
    for j, f in enumerate(files):
        cap = L.pna_oracle_encode_bound(2, len(f))
        o = C.create_string_buffer(cap)
        outs.append(o)
        jobs[j].plain = C.cast(C.c_char_p(f), C.c_void_p)
        jobs[j].len = len(f)
        jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode, jobs[j].level = 2, 1, 1, 3
        C.memmove(jobs[j].key, key, 32)
        C.memmove(jobs[j].iv, rng.bytes(16), 16)
        jobs[j].out = C.cast(o, C.c_void_p)
        jobs[j].cap = cap
    L.pna_oracle_encode_batch_mt(jobs, entries, threads, None)
    streams = []
    for j in range(entries):
        assert jobs[j].status == 0
        streams.append(outs[j].raw[:jobs[j].out_len])
    return files, streams, key


def make_shard_encode_only(files, key, threads: int):
    """The reference's create dataflow on the host (oracle): `threads` workers, one entry each."""
    import pna_oracle as O
    L = O.lib()
    n = len(files)
    jobs = (O.EncJob * n)()
    outs = []
    for j, f in enumerate(files):
        cap = L.pna_oracle_encode_bound(2, len(f))
        o = C.create_string_buffer(cap)
        outs.append(o)
        jobs[j].plain = C.cast(C.c_char_p(f), C.c_void_p)
        jobs[j].len = len(f)
        jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode, jobs[j].level = 2, 1, 1, 3
        C.memmove(jobs[j].key, key, 32)
        jobs[j].out = C.cast(o, C.c_void_p)
        jobs[j].cap = cap
    L.pna_oracle_encode_batch_mt(jobs, n, threads, None)
    assert all(jobs[j].status == 0 for j in range(n))


def build_archive(streams, sizes, phsf: str, into=None):
    """PNA container bytes (signature, AHED, entries, AEND); chunk CRCs via zlib (input preparation)."""
    parts = [b"\x89PNA\r\n\x1a\n"]

    def chunk(ty, data):
        parts.append(struct.pack(">I", len(data)) + ty)
        parts.append(data)
        parts.append(struct.pack(">I", zlib.crc32(data, zlib.crc32(ty))))
    chunk(b"AHED", bytes(8))
    for i, (s, n) in enumerate(zip(streams, sizes)):
        chunk(b"FHED", bytes([0, 0, 0, 2, 1, 1]) + f"corpus/{i:07d}.bin".encode())
        chunk(b"fSIZ", int(n).to_bytes(8, "big").lstrip(b"\0") or b"\0")
        chunk(b"PHSF", phsf.encode())
        chunk(b"FDAT", s[:16])
        chunk(b"FDAT", s[16:])
        chunk(b"FEND", b"")
    chunk(b"AEND", b"")
    total = sum(len(p) for p in parts)
    buf = into(total) if into else np.empty(total, dtype=np.uint8)
    pos = 0
    for p in parts:
        buf[pos:pos + len(p)] = np.frombuffer(p, dtype=np.uint8)
        pos += len(p)
    return buf


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_decode(streams, sizes, key, threads: int, repeat: int = 1):
    """The reference's extract dataflow on the host (oracle): one thread CRCs every FDAT, `threads` workers decode
    one entry each.  Returns seconds per pass."""
    import pna_oracle as O
    L = O.lib()
    n = len(streams)
    jobs = (O.Job * n)()
    outs = []
    for j, (s, u) in enumerate(zip(streams, sizes)):
        o = C.create_string_buffer(int(u))
        outs.append(o)
        jobs[j].stream = C.cast(C.c_char_p(s), C.c_void_p)
        jobs[j].len = len(s)
        jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode = 2, 1, 1
        C.memmove(jobs[j].key, key, 32)
        jobs[j].out = C.cast(o, C.c_void_p)
        jobs[j].cap = int(u)
    crc = (C.c_uint32 * n)()
    t0 = time.perf_counter()
    for _ in range(repeat):
        L.pna_oracle_decode_batch_mt(jobs, n, threads, 1, crc)
    dt = (time.perf_counter() - t0) / repeat
    assert all(jobs[j].status == 0 and jobs[j].out_len == sizes[j] for j in range(n))
    return dt, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--entries", type=int, default=1024, help="4 MiB files per GPU (cfg2: 8192 over 8 GPUs)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--workers", type=int, default=3, help="host worker threads / contexts of the end-to-end path")
    ap.add_argument("--group-mib", type=int, default=128, help="compressed MiB per pipelined entry group (end-to-end path)")
    ap.add_argument("--create-workers", type=int, default=4)
    ap.add_argument("--create-group-mib", type=int, default=256)
    ap.add_argument("--create", type=int, default=1, help="also measure the create path (GPU zstd + AES-CTR + CRC) on the same files")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncpu = os.cpu_count() or 1
    threads = max(1, ncpu // max(world, 1))
    E = args.entries
    cfg = {"workload": f"cfg2 shard: extract {E} x 4 MiB files per GPU, zstd level 3 + AES-256-CTR, chunk CRC-32 check "
                       f"(32 GiB / 8 GPUs at the full config)", "entries_per_gpu": E, "file_bytes": FILE_SIZE,
           "codec": "zstd-3", "cipher": "aes-256-ctr", "parallelism": f"entry-sharded x{world}, no collective",
           "l2": "inputs (C+U per step ~5.7 GiB) far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(E, 256)
        files, streams, key = make_shard(0, sample, ncpu)
        sizes = [len(f) for f in files]
        U = sum(sizes)
        for _ in range(max(args.warmup, 1)):
            cpu_decode(streams, sizes, key, ncpu)
        tot = 0.0
        for _ in range(args.steps):   # buffer set-up is outside the timed region (inputs/outputs resident in RAM)
            dt, outs = cpu_decode(streams, sizes, key, ncpu)
            tot += dt
        assert outs[0].raw == files[0]
        v = U * args.steps / tot / 1e9
        line = {"impl": "reference", "metric": "extract_uncompressed_GBps", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": "GB/s", "cores": ncpu, "kind": "port",
                                 "sample": f"{sample} x 4 MiB entries per step (oracle: libzstd + OpenSSL AES-NI + zlib crc32, "
                                           f"1 CRC thread + {ncpu} workers, reference dataflow)"},
                "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pna = importlib.import_module("portable-network-archive_b200")
    ctx = pna.Context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs (not timed)
    files, streams, key = make_shard(rank, E, threads, world)
    sizes = [len(f) for f in files]
    U = sum(sizes)
    opts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=1, password=PASSWORD, kdf_params={"i": 1000})
    # the shard's streams were encrypted with `key`; record it as the options' derived key (KDF is host work)
    archive_buf = build_archive(streams, sizes, opts.phsf, into=ctx.pinned)
    ro = pna.ReadOptions.with_password(PASSWORD)
    ro._keys[opts.phsf] = key
    Cbytes = sum(len(s) for s in streams)

    archive = pna.Archive.read_header(archive_buf, ctx, verify=False)   # index pass (host); CRC runs inside the plan
    plan, ents = archive.extract_plan(ro)
    assert len(ents) == E
    stream = torch.cuda.ExternalStream(ctx.stream)
    # ---- kernel-only: W warm-ups, K timed steps, CUDA events on the library's stream
    for _ in range(max(args.warmup, 1)):
        plan.run()
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    stage_acc = {}
    for _ in range(args.steps):
        plan.run()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if sampler else None
    ms_step = max_over_ranks(ms_total / args.steps)
    counts = plan.counts()
    stage_ms = plan.stage_ms()                     # last timed step, CUDA events between stages on the same stream
    # parity spot-check of what was just timed (all entries, SHA-free exact compare)
    outs, st, _ = plan.fetch(sizes)
    crcs, broken = plan.crc_results()
    assert st == [0] * E and broken == 0, "decode failed"
    for k in range(0, E, max(1, E // 16)):
        assert outs[k].tobytes() == files[k], "GPU output differs from the source file"
    del outs
    plan.close()

    # ---- end to end through the reference-facing host API (C++ pna::Archive over the C ABI) with HOST buffers:
    # index pass over the pinned archive bytes, then entry groups pipelined over `--workers` contexts (H2D + chunk CRC
    # check + decrypt + decode + D2H into pinned output), everything inside the timed region
    host = importlib.import_module("portable-network-archive_b200._host")
    out_pinned = ctx.pinned(U + 16 * E + 64)
    e2e_times = []
    for it in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        ha = host.HostArchive(archive_buf)
        ha.set_key(opts.phsf, key)
        _, offs, stv = ha.extract_files(out=out_pinned, device=local_rank, workers=args.workers, group_bytes=args.group_mib << 20, verify=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ha.close()
        if it > 0:
            e2e_times.append(dt)
    if e2e_times:
        assert stv == [0] * E
        for k in range(0, E, max(1, E // 16)):
            assert out_pinned[int(offs[k]):int(offs[k]) + sizes[k]].tobytes() == files[k], "e2e output differs from the source file"
    e2e_s = max_over_ranks(statistics.median(e2e_times)) if e2e_times else float('nan')
    # PCIe yardstick for the end-to-end number: pinned <-> HBM copies of 1 GiB on this box (CUDA events); the transfer
    # floor of a step is the slower direction (full duplex), everything else of e2e is pipeline fill and host work
    pcie = None
    if e2e_times and rank == 0:
        nprobe = 1 << 30
        hbuf = torch.empty(nprobe, dtype=torch.uint8, pin_memory=True)
        dbuf = torch.empty(nprobe, dtype=torch.uint8, device="cuda")
        pcie = {}
        for name, dst, src in (("h2d_GBps", dbuf, hbuf), ("d2h_GBps", hbuf, dbuf)):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pa.record()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            pb.record()
            torch.cuda.synchronize()
            pcie[name] = 3 * nprobe / (pa.elapsed_time(pb) * 1e-3) / 1e9
        pcie["transfer_floor_ms"] = max(archive_buf.size / pcie["h2d_GBps"], U / pcie["d2h_GBps"]) / 1e6
        del hbuf, dbuf

    # ---- create path (BASELINE config 4 shape: GPU zstd encode + AES-256-CTR + FDAT CRC-32) on the same files
    create = None
    if args.create:
        import pna_oracle as O
        rng = np.random.Generator(np.random.PCG64(99 + rank))
        plain_pinned = ctx.pinned(U)
        pos = 0
        views = []
        for f in files:
            plain_pinned[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
            views.append(plain_pinned[pos:pos + len(f)])
            pos += len(f)
        ents = [{"plain": v, "compression": 2, "level": 3, "encryption": 1, "cipher_mode": 1, "key": key, "iv": rng.bytes(16),
                 "max_chunk_size": 0} for v in views]
        eplan = ctx.encode_plan(ents)
        for _ in range(max(args.warmup, 1)):
            eplan.run()
        torch.cuda.synchronize()
        l1 = ctx.launch_count
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(args.steps):
            eplan.run()
        c1.record(stream)
        barrier()
        c_ms = max_over_ranks(c0.elapsed_time(c1) / args.steps)
        c_launches = ctx.launch_count - l1
        c_stage = eplan.stage_ms()
        outp = ctx.pinned(sum(eplan.bounds))
        streams_gpu, crcs_gpu, cst = eplan.fetch(into=outp)
        assert cst == [0] * E
        c_gpu = sum(int(s.size) for s in streams_gpu)
        for k in range(0, E, max(1, E // 8)):   # the reference pipeline must read what we wrote, CRCs must match
            s = streams_gpu[k].tobytes()
            assert O.decode_stream(s, 2, 1, 1, key, None) == files[k], "GPU-created stream is not reference-readable"
            assert int(crcs_gpu[k][0]) == O.chunk_crc(b"FDAT", s[16:])
        eplan.close()
        ce2e = []
        names = [f"corpus/{i:07d}.bin" for i in range(E)]
        arch_out = ctx.pinned(int(U * 1.02) + (64 << 20))
        ivs = rng.bytes(16 * E)
        for it in range(args.e2e_steps + 1):
            barrier()
            t0 = time.perf_counter()
            blob = host.create_archive(list(zip(names, views)), compression=2, level=3, encryption=1, cipher_mode=1, key=key, phsf=opts.phsf,
                                       ivs=ivs, max_chunk_size=0, device=local_rank, workers=args.create_workers, group_bytes=args.create_group_mib << 20,
                                       out=arch_out)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it > 0:
                ce2e.append(dt)
        if ce2e:   # the created archive must extract bit-exactly with the reference reader (oracle) -- sample -- and with ours
            got = O.extract_all(blob[:min(blob.size, 64 << 20)].tobytes() if False else blob.tobytes(), PASSWORD, _keys={opts.phsf: key}) if E <= 64 else None
            hb = host.HostArchive(blob)
            hb.set_key(opts.phsf, key)
            o2, of2, st2 = hb.extract_files(device=local_rank, workers=args.workers)
            assert st2 == [0] * E
            for k in range(0, E, max(1, E // 16)):
                assert o2[int(of2[k]):int(of2[k]) + sizes[k]].tobytes() == files[k]
            if got is not None:
                assert [d for _, d in got] == files
            del o2
            hb.close()
        ce2e_s = max_over_ranks(statistics.median(ce2e)) if ce2e else float("nan")
        create = {"metric": "create_uncompressed_GBps", "value": world * U / (c_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": c_ms,
                  "e2e": {"value": world * U / ce2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(U) * world, "d2h_bytes_per_step": int(c_gpu) * world,
                          "ms_per_step": ce2e_s * 1e3},
                  "codec": "gpu zstd (32 KiB blocks, predefined-FSE sequences, Huffman literals where the alphabet allows the direct weight form) + aes-256-ctr + crc32", "stage_ms": c_stage,
                  "gpu_launches": c_launches, "ratio": U / c_gpu, "c_gpu_over_c_ref": c_gpu / Cbytes,
                  "checked": "sampled streams decoded by the oracle (libzstd + OpenSSL) == source files; FDAT CRCs == zlib crc32"}
        del outp, plain_pinned

    # ---- CPU baseline on this box's cores (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    cpu_create = None
    if rank == 0 and world == 1:
        sample = min(E, 256)
        cpu_decode(streams[:sample], sizes[:sample], key, ncpu)
        dt, _ = cpu_decode(streams[:sample], sizes[:sample], key, ncpu, repeat=3)
        if args.create:
            t0 = time.perf_counter()
            make_shard_encode_only(files[:sample], key, ncpu)
            dtc = time.perf_counter() - t0
            cpu_create = {"value": sum(sizes[:sample]) / dtc / 1e9, "unit": "GB/s", "cores": ncpu, "kind": "port",
                          "sample": f"{sample} x 4 MiB entries, oracle encode (libzstd level 3 streaming + OpenSSL AES-256-CTR), {ncpu} threads"}
        cpu = {"value": sum(sizes[:sample]) / dt / 1e9, "unit": "GB/s", "cores": ncpu, "kind": "port",
               "sample": f"{sample} x 4 MiB entries x3 passes (oracle: libzstd + OpenSSL AES-NI + zlib crc32; 1 CRC thread + "
                         f"{ncpu} worker threads, reference extract dataflow)"}

    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant stage of the step and its algorithmic bytes (DESIGN.md "algorithmic bytes")
    # per-stage algorithmic bytes: seq = bitstreams in (<= C) + 8 B records out; lz = records + literals in, U out
    alg = {"crc": Cbytes, "cipher": 2 * Cbytes, "zstd_scan": 0, "zstd_seq": Cbytes + 8 * counts["sequences"],
           "zstd_lit": Cbytes + counts["literal_bytes"], "zstd_prefix": 0,
           "zstd_lz": 8 * counts["sequences"] + counts["literal_bytes"] + U, "inflate": 0, "store": 0}
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    step_alg = Cbytes + U                           # fully fused accounting: ciphertext read once, plaintext written once
    dom_ms = stage_ms[dom]
    roof = {"bound": "hbm", "kernel": dom, "achieved": alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None, "peak": peak,
            "unit": "GB/s", "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
            "traffic": None, "stage_ms": stage_ms, "stage_share": {k: v / sum(stage_ms.values()) for k, v in stage_ms.items()},
            "step_algorithmic_GBps": step_alg / (ms_step * 1e-3) / 1e9, "step_frac": step_alg / (ms_step * 1e-3) / 1e9 / peak}
    roof["frac"] = roof["achieved"] / peak if roof["achieved"] else None
    roof["algorithmic_bytes"] = alg[dom]
    try:   # DRAM bytes of the dominant kernel per launch from the committed ncu --set full capture of this same configuration
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        k = tr["kernels"].get(dom + "_kernel")
        if k and E == 1024:
            roof["traffic"] = k["dram_read_bytes"] + k["dram_write_bytes"]
            roof["traffic_source"] = "profiles/r1_traffic.json (ncu --set full, one launch, same --entries)"
    except Exception:
        pass
    value = world * U / (ms_step * 1e-3) / 1e9
    line = {"metric": "extract_uncompressed_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": dict(cfg, compressed_bytes_per_gpu=Cbytes, plain_bytes_per_gpu=U,
                                                             ratio=U / Cbytes),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * U / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(archive_buf.size) * world,
                    "d2h_bytes_per_step": int(U) * world, "ms_per_step": e2e_s * 1e3, "pcie": pcie,
                    "path": f"pna::Archive::read_header_from_slice + extract_files (C++ host layer, {args.workers} worker threads x 2 contexts, {args.group_mib} MiB entry groups software-pipelined create->run->fetch): index pass, H2D, chunk CRC check, decrypt, decode, D2H to pinned buffers; host clock"},
            "roofline": roof}
    if cpu:
        line["cpu_baseline"] = cpu
    if create:
        if cpu_create:
            create["cpu_baseline"] = cpu_create
        line["create"] = create
    print(json.dumps(line))


if __name__ == "__main__":
    main()
