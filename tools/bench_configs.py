#!/usr/bin/env python
"""Secondary BASELINE.json configurations, measured with the same machinery as bench.py (not the driver's bench line):
  cfg3  many small files (uniform 1..16384 B), zlib level 6 + Camellia-256-CBC          -> inflate + CBC + index-pass stress
  cfg5  N x 4 MiB files in ONE solid zstd entry (inner STORE entries, 32 KiB SDATs)       -> single-stream decode vs per-entry
  cfg1  log-normal file sizes, zstd 3, no encryption: create + extract                    -> the CPU-runnable case
Prints one JSON line per configuration.  Inputs come from the oracle's encoders (reference dataflow); outputs are compared
with the source files."""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import struct
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import corpus  # noqa: E402
import bench  # noqa: E402


def _small(i_n):
    i, n = i_n
    return corpus.make_file(10_000_000 + i, n)


def _big(i):
    return corpus.make_file(20_000_000 + i, 4 << 20)


def oracle_encode(files, comp, level, enc, mode, key, threads, seed=7):
    import pna_oracle as O
    L = O.lib()
    n = len(files)
    jobs = (O.EncJob * n)()
    outs = []
    rng = np.random.Generator(np.random.PCG64(seed))
    for j, f in enumerate(files):
        cap = L.pna_oracle_encode_bound(comp, len(f)) + 64
        o = C.create_string_buffer(cap)
        outs.append(o)
        jobs[j].plain = C.cast(C.c_char_p(f), C.c_void_p)
        jobs[j].len = len(f)
        jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode, jobs[j].level = comp, enc, mode, level
        C.memmove(jobs[j].key, key, 32)
        C.memmove(jobs[j].iv, rng.bytes(16), 16)
        jobs[j].out = C.cast(o, C.c_void_p)
        jobs[j].cap = cap
    L.pna_oracle_encode_batch_mt(jobs, n, threads, None)
    assert all(jobs[j].status == 0 for j in range(n))
    return [outs[j].raw[:jobs[j].out_len] for j in range(n)]


def cpu_decode_time(streams, sizes, comp, enc, mode, key, threads):
    import pna_oracle as O
    L = O.lib()
    n = len(streams)
    jobs = (O.Job * n)()
    outs = []
    for j, (s, u) in enumerate(zip(streams, sizes)):
        o = C.create_string_buffer(max(int(u), 1))
        outs.append(o)
        jobs[j].stream = C.cast(C.c_char_p(s), C.c_void_p)
        jobs[j].len = len(s)
        jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode = comp, enc, mode
        C.memmove(jobs[j].key, key, 32)
        jobs[j].out = C.cast(o, C.c_void_p)
        jobs[j].cap = int(u)
    crc = (C.c_uint32 * n)()
    L.pna_oracle_decode_batch_mt(jobs, n, threads, 1, crc)
    t0 = time.perf_counter()
    L.pna_oracle_decode_batch_mt(jobs, n, threads, 1, crc)
    return time.perf_counter() - t0


def chunk(parts, ty, data):
    parts.append(struct.pack(">I", len(data)) + ty)
    parts.append(data)
    parts.append(struct.pack(">I", zlib.crc32(data, zlib.crc32(ty))))


def archive_of(entries, phsf, into):
    """entries: (name, header6, raw_size, stream, iv_len) -> FHED,fSIZ,[PHSF,FDAT(iv)],FDAT(body),FEND"""
    parts = [b"\x89PNA\r\n\x1a\n"]
    chunk(parts, b"AHED", bytes(8))
    for name, hdr, n, s, iv_len in entries:
        chunk(parts, b"FHED", hdr + name.encode())
        chunk(parts, b"fSIZ", int(n).to_bytes(8, "big").lstrip(b"\0"))
        if iv_len:
            chunk(parts, b"PHSF", phsf.encode())
            chunk(parts, b"FDAT", s[:iv_len])
        if len(s) > iv_len:
            chunk(parts, b"FDAT", s[iv_len:])
        chunk(parts, b"FEND", b"")
    chunk(parts, b"AEND", b"")
    total = sum(len(p) for p in parts)
    buf = into(total)
    pos = 0
    for p in parts:
        buf[pos:pos + len(p)] = np.frombuffer(p, dtype=np.uint8)
        pos += len(p)
    return buf


def timed_extract(host, ctx, archive_buf, phsf, key, U, nfiles, reps=3, workers=2, group_mib=64):
    import torch
    out = ctx.pinned(U + 16 * nfiles + 64)
    ts, t_index = [], []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        ha = host.HostArchive(archive_buf)
        t1 = time.perf_counter()
        if phsf:
            ha.set_key(phsf, key)
        _, offs, st = ha.extract_files(out=out, device=0, workers=workers, group_bytes=group_mib << 20, verify=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        t_index.append(t1 - t0)
        files = ha.files() if nfiles <= 4096 else None
        ha.close()
    return min(ts[1:]), min(t_index[1:]), out, offs, st, files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg3-files", type=int, default=131072)
    ap.add_argument("--cfg5-files", type=int, default=32)
    ap.add_argument("--cfg1-files", type=int, default=2500)
    ap.add_argument("--io-files", type=int, default=256)
    ap.add_argument("--gcm-files", type=int, default=256)
    ap.add_argument("--own-files", type=int, default=256)
    ap.add_argument("--only", default="")
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--group-mib", type=int, default=64)
    args = ap.parse_args()
    import multiprocessing as mp
    pna = importlib.import_module("portable-network-archive_b200")
    host = importlib.import_module("portable-network-archive_b200._host")
    ctx = pna.Context(0)
    ncpu = os.cpu_count() or 1
    key = bytes(range(32))
    opts = pna.WriteOptions(compression=1, encryption=2, cipher_mode=0, password=b"pw", kdf_params={"i": 1000})

    if args.only in ("", "cfg3"):
        n = args.cfg3_files
        rng = np.random.Generator(np.random.PCG64(3))
        sizes = [int(x) for x in rng.integers(1, 16385, n)]
        with mp.get_context("fork").Pool(min(ncpu, 64)) as pool:
            files = pool.map(_small, list(enumerate(sizes)), chunksize=256)
        streams = oracle_encode(files, 1, 6, 2, 0, key, ncpu)
        U, Cb = sum(sizes), sum(len(s) for s in streams)
        buf = archive_of([(f"s/{i:07d}", bytes([0, 0, 0, 1, 2, 0]), sizes[i], streams[i], 16) for i in range(n)], opts.phsf, ctx.pinned)
        # kernel-only through one plan
        ha = pna.Archive.read_header(buf, ctx, verify=False) if n <= 20000 else None
        dt, t_index, out, offs, st, flist = timed_extract(host, ctx, buf, opts.phsf, key, U, n, workers=args.workers, group_mib=args.group_mib)
        assert st == [0] * n
        for k in range(0, n, max(1, n // 64)):
            assert out[int(offs[k]):int(offs[k]) + sizes[k]].tobytes() == files[k]
        cpu_dt = cpu_decode_time(streams[:min(n, 65536)], sizes[:min(n, 65536)], 1, 2, 0, key, ncpu)
        print(json.dumps({"config": "cfg3", "files": n, "plain_bytes": U, "stream_bytes": Cb, "chunks": 6 * n + 2,
                          "codec": "zlib-6 + camellia-256-cbc", "e2e_GBps": U / dt / 1e9, "e2e_ms": dt * 1e3, "index_pass_ms": t_index * 1e3,
                          "path": "pna::Archive (C++ host) index + extract_files, pinned buffers, CRC verified on GPU",
                          "cpu_baseline_GBps": sum(sizes[:min(n, 65536)]) / cpu_dt / 1e9, "cpu_cores": ncpu}), flush=True)
        del out, buf, files, streams

    if args.only in ("", "cfg5"):
        import pna_oracle as O
        n = args.cfg5_files
        files = [corpus.make_file(i, 4 << 20) for i in range(n)]
        inner = []
        for i, f in enumerate(files):
            chunk(inner, b"FHED", bytes([0, 0, 0, 0, 0, 0]) + f"solid/{i:05d}.bin".encode())
            chunk(inner, b"fSIZ", len(f).to_bytes(8, "big").lstrip(b"\0"))
            for o in range(0, len(f), 1 << 20):
                chunk(inner, b"FDAT", f[o:o + (1 << 20)])
            chunk(inner, b"FEND", b"")
        inner = b"".join(inner)
        stream = O.compress(2, inner, 3)          # one zstd frame over the whole inner archive (window 2 MiB)
        parts = [b"\x89PNA\r\n\x1a\n"]
        chunk(parts, b"AHED", bytes(8))
        chunk(parts, b"SHED", bytes([0, 0, 2, 0, 0]))
        for o in range(0, len(stream), 32768):
            chunk(parts, b"SDAT", stream[o:o + 32768])
        chunk(parts, b"SEND", b"")
        chunk(parts, b"AEND", b"")
        blob = b"".join(parts)
        buf = ctx.pinned(len(blob))
        buf[:] = np.frombuffer(blob, dtype=np.uint8)
        U = sum(len(f) for f in files)
        dt, t_index, out, offs, st, flist = timed_extract(host, ctx, buf, None, key, U, n, reps=2)
        assert st == [0] * n and [nm for nm, _, _ in flist] == [f"solid/{i:05d}.bin" for i in range(n)]
        for k in range(n):
            assert out[int(offs[k]):int(offs[k]) + len(files[k])].tobytes() == files[k]
        # the same corpus as per-entry archive
        streams = oracle_encode(files, 2, 3, 0, 0, key, ncpu)
        buf2 = archive_of([(f"e/{i:05d}", bytes([0, 0, 0, 2, 0, 0]), len(files[i]), streams[i], 0) for i in range(n)], "", ctx.pinned)
        dt2, _, out2, offs2, st2, _ = timed_extract(host, ctx, buf2, None, key, U, n, reps=2)
        assert st2 == [0] * n
        t0 = time.perf_counter()
        assert O.decompress(2, stream) == inner
        cpu_solid = time.perf_counter() - t0
        # the same inner archive compressed by THIS library's writer: one frame per MiB (reference-readable), so the
        # LZ stage runs one unit per frame instead of one CTA for the whole stream
        # ... through the C++ solid writer (pna::create_solid_archive_into: inner chunk CRCs, encode, outer chunk CRCs on the GPU)
        sfiles = [(f"solid/{i:05d}.bin", np.frombuffer(f, dtype=np.uint8)) for i, f in enumerate(files)]
        buf3 = ctx.pinned(len(inner) + (16 << 20))
        cts3 = []
        for _ in range(3):
            t0 = time.perf_counter()
            blob3 = host.create_solid_archive(sfiles, compression=2, level=3, max_chunk_size=32768, out=buf3)
            cts3.append(time.perf_counter() - t0)
        assert [d for _, d in O.extract_all(blob3.tobytes(), None)] == files       # the reference reader extracts it
        arch3_len = int(blob3.size)
        buf3 = blob3
        dt3, _, out3, offs3, st3, _ = timed_extract(host, ctx, buf3, None, key, U, n, reps=2)
        assert st3 == [0] * n
        for k in range(n):
            assert out3[int(offs3[k]):int(offs3[k]) + len(files[k])].tobytes() == files[k]
        print(json.dumps({"config": "cfg5", "files": n, "plain_bytes": U, "solid_stream_bytes": len(stream), "sdat_chunks": (len(stream) + 32767) // 32768,
                          "solid_e2e_GBps": U / dt / 1e9, "solid_e2e_ms": dt * 1e3, "per_entry_e2e_GBps": U / dt2 / 1e9, "per_entry_e2e_ms": dt2 * 1e3,
                          "note": "a solid entry is ONE zstd frame: the sequence stage is block-parallel (lane per block), the LZ stage runs "
                                  "the frame on one CTA of 16 warps; the stream is decoded once, stays in HBM for the inner chunk CRC check and the range copies of the STORE entries",
                          "cpu_baseline_solid_GBps": U / cpu_solid / 1e9, "cpu_cores_solid": 1,
                          "gpu_written_solid_e2e_GBps": U / dt3 / 1e9, "gpu_written_solid_e2e_ms": dt3 * 1e3, "gpu_written_archive_bytes": arch3_len,
                          "gpu_written_frames": (len(inner) + (1 << 20) - 1) >> 20, "gpu_solid_create_e2e_GBps": U / min(cts3[1:]) / 1e9,
                          "gpu_solid_create_e2e_ms": min(cts3[1:]) * 1e3}), flush=True)
        del out, out2, out3

    if args.only in ("", "io"):
        # file-system side: create_from_files -> archive file -> extract_to_dir, on tmpfs (/dev/shm) so that
        # the numbers show the data path and not the box's disk
        import shutil
        import tempfile
        n = args.io_files
        root = tempfile.mkdtemp(prefix="pna_io_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            files = [corpus.make_file(50_000 + i, 4 << 20) for i in range(n)]
            src = os.path.join(root, "src")
            os.makedirs(src)
            pairs = []
            for i, f in enumerate(files):
                pth = os.path.join(src, f"f{i:05d}.bin")
                with open(pth, "wb") as fh:
                    fh.write(f)
                pairs.append((f"corpus/f{i:05d}.bin", pth))
            U = sum(len(f) for f in files)
            wopts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=1, password=b"pw", kdf_params={"i": 1000})
            arch = os.path.join(root, "a.pna")
            best_c = best_x = None
            for _ in range(3):
                if os.path.exists(arch):
                    os.remove(arch)      # freeing the old file's tmpfs pages is not part of the create path
                sc = host.create_from_files(pairs, arch, compression=2, level=3, encryption=1, cipher_mode=1, key=wopts.key, phsf=wopts.phsf,
                                            io_threads=16)
                best_c = sc if best_c is None or sc["total_ms"] < best_c["total_ms"] else best_c
            for _ in range(3):
                out_dir = os.path.join(root, "out")
                shutil.rmtree(out_dir, ignore_errors=True)
                ha = host.HostArchive.open_file(arch)
                ha.set_key(wopts.phsf, wopts.key)
                t0 = time.perf_counter()
                sx, stx = ha.extract_to_dir(out_dir, window_bytes=1 << 30, io_threads=16)
                sx["wall_ms"] = (time.perf_counter() - t0) * 1e3
                ha.close()
                assert stx == [0] * n
                best_x = sx if best_x is None or sx["wall_ms"] < best_x["wall_ms"] else best_x
            for i in range(0, n, max(1, n // 8)):
                with open(os.path.join(root, "out", "corpus", f"f{i:05d}.bin"), "rb") as fh:
                    assert fh.read() == files[i]
            print(json.dumps({"config": "io (tmpfs)", "files": n, "plain_bytes": U, "archive_bytes": os.path.getsize(arch),
                              "create_from_files_GBps": U / best_c["total_ms"] / 1e6, "create_ms": best_c,
                              "extract_to_dir_GBps": U / best_x["wall_ms"] / 1e6, "extract_ms": best_x,
                              "path": "files on tmpfs -> pinned -> GPU zstd+AES-CTR -> archive file; archive file (mmap) -> GPU -> pinned windows -> files"}), flush=True)
        finally:
            shutil.rmtree(root, ignore_errors=True)

    if args.only in ("", "gcm"):
        # cfg2's entry shape under the GCM STREAM cipher mode: 4 MiB files, zstd 3 + AES-256-GCM, 1 MiB segments, one stream key
        # per entry (aead.rs:188).  Kernel-only through one plan (stage times from CUDA events) and the CPU port beside it.
        import pna_oracle as O
        from concurrent.futures import ThreadPoolExecutor
        n = args.gcm_files
        with mp.get_context("fork").Pool(min(ncpu, 64)) as pool:
            files = pool.map(_big, list(range(n)), chunksize=4)
        keys = [bytes(np.random.Generator(np.random.PCG64(900 + i)).bytes(32)) for i in range(n)]

        def enc_one(i):
            hdr = bytes(np.random.Generator(np.random.PCG64(500 + i)).bytes(39)) + struct.pack(">I", 1 << 20) + bytes(32)
            return O.gcm_encrypt_stream(1, keys[i], hdr, O.compress(2, files[i], 3))
        with ThreadPoolExecutor(ncpu) as ex:
            streams = list(ex.map(enc_one, range(n)))
        U, Cb = sum(len(f) for f in files), sum(len(x) for x in streams)
        img = ctx.pinned(Cb)
        ents, pos = [], 0
        for i, x in enumerate(streams):
            img[pos:pos + len(x)] = np.frombuffer(x, dtype=np.uint8)
            ents.append({"bodies": [img[pos:pos + len(x)]], "compression": 2, "encryption": 1, "cipher_mode": 2, "key": keys[i],
                         "raw_size_hint": len(files[i])})
            pos += len(x)
        plan = ctx.decode_plan(ents)
        for _ in range(4):
            plan.run()
        stage = plan.stage_ms()               # CUDA events of the last run, on the library's stream
        dt = sum(stage.values()) * 1e-3
        outs, st, _ = plan.fetch([len(f) for f in files])
        assert list(st) == [0] * n
        for k in range(0, n, max(1, n // 16)):
            assert outs[k].tobytes() == files[k]
        plan.close()
        L = O.lib()
        jobs = (O.Job * n)()
        keep = []
        for j in range(n):
            o = C.create_string_buffer(len(files[j]))
            keep.append(o)
            jobs[j].stream = C.cast(C.c_char_p(streams[j]), C.c_void_p)
            jobs[j].len = len(streams[j])
            jobs[j].compression, jobs[j].encryption, jobs[j].cipher_mode = 2, 1, 2
            C.memmove(jobs[j].key, keys[j], 32)
            jobs[j].out = C.cast(o, C.c_void_p)
            jobs[j].cap = len(files[j])
        crc = (C.c_uint32 * n)()
        L.pna_oracle_decode_batch_mt(jobs, n, ncpu, 1, crc)
        t0 = time.perf_counter()
        L.pna_oracle_decode_batch_mt(jobs, n, ncpu, 1, crc)
        cpu_dt = time.perf_counter() - t0
        assert all(jobs[j].status == 0 for j in range(n))
        # create leg: GPU zstd + AES-256-GCM, kernel-only through one plan; the oracle must read what was written
        hdrs = [bytes(np.random.Generator(np.random.PCG64(700 + i)).bytes(39)) + struct.pack(">I", 1 << 20) + bytes(32) for i in range(n)]
        pl = ctx.pinned(U)
        views, pos = [], 0
        for f in files:
            pl[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
            views.append(pl[pos:pos + len(f)])
            pos += len(f)
        eplan = ctx.encode_plan([{"plain": v, "compression": 2, "level": 3, "encryption": 1, "cipher_mode": 2, "key": keys[i],
                                  "stream_header": hdrs[i], "max_chunk_size": 0} for i, v in enumerate(views)])
        for _ in range(4):
            eplan.run()
        c_stage = eplan.stage_ms()
        outp = ctx.pinned(sum(eplan.bounds))
        sg, _, cst = eplan.fetch(into=outp)
        assert cst == [0] * n
        for k in range(0, n, max(1, n // 8)):
            assert O.decompress(2, O.gcm_decrypt_stream(1, keys[k], sg[k].tobytes()), len(files[k])) == files[k]
        c_gpu = sum(int(x.size) for x in sg)
        eplan.close()
        # end to end through the C++ host layer (pna::create_archive_into / pna::Archive::extract_files), pinned buffers:
        # stream headers, key confirmation and per-entry stream keys on the host, everything else on the GPU
        import torch
        gopts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=2, password=b"pw", kdf_params={"i": 1000})
        names = [f"gcm/{i:05d}.bin" for i in range(n)]
        arch = ctx.pinned(int(U * 1.05) + (8 << 20))
        cts = []
        for _ in range(3):
            t0 = time.perf_counter()
            blob = host.create_archive(list(zip(names, views)), compression=2, level=3, encryption=1, cipher_mode=2, key=gopts.key,
                                       phsf=gopts.phsf, max_chunk_size=0, device=0, workers=4, group_bytes=256 << 20, out=arch)
            torch.cuda.synchronize()
            cts.append(time.perf_counter() - t0)
        probe = pna.Archive.read_header(blob, ctx, verify=False)
        e0 = next(iter(probe.entries()))
        assert O.decode_stream(e0.stream_bytes() if hasattr(e0, "stream_bytes") else b"".join(bytes(b) for b in e0.bodies), 2, 1, 2,
                               O.gcm_derive_stream_key(gopts.key, b"".join(bytes(b) for b in e0.bodies)[:75], b"FHED", e0.header_bytes,
                                                       gopts.phsf.encode())) == files[0]
        edt, _, eout, eoffs, est, _ = timed_extract(host, ctx, blob, gopts.phsf, gopts.key, U, n, workers=args.workers, group_mib=args.group_mib)
        assert est == [0] * n
        for k in range(0, n, max(1, n // 16)):
            assert eout[int(eoffs[k]):int(eoffs[k]) + len(files[k])].tobytes() == files[k]
        print(json.dumps({"config": "gcm", "files": n, "plain_bytes": U, "stream_bytes": Cb, "codec": "zstd-3 + aes-256-gcm (1 MiB segments)",
                          "kernel_only_GBps": U / dt / 1e9, "kernel_only_ms": dt * 1e3, "stage_ms": stage,
                          "gcm_stage_GBps_of_ciphertext": Cb / (stage.get("cipher", 0) * 1e-3 + 1e-12) / 1e9,
                          "cpu_baseline_GBps": U / cpu_dt / 1e9, "cpu_cores": ncpu, "cpu_kind": "port (OpenSSL AES-256-GCM + libzstd)",
                          "extract_e2e_GBps": U / edt / 1e9, "extract_e2e_ms": edt * 1e3, "create_e2e_GBps": U / min(cts[1:]) / 1e9,
                          "create_e2e_ms": min(cts[1:]) * 1e3,
                          "create_kernel_only_GBps": U / (sum(c_stage.values()) * 1e-3) / 1e9, "create_stage_ms": c_stage,
                          "create_gcm_stage_GBps_of_ciphertext": c_gpu / (c_stage.get("cipher", 0) * 1e-3 + 1e-12) / 1e9}), flush=True)
        del files, streams, img

    if args.only in ("", "own"):
        # an archive WRITTEN BY THIS LIBRARY (cfg2's shape: 4 MiB files, zstd + AES-256-CTR; 32 KiB blocks, one frame per MiB) read back:
        # kernel-only through one plan and end to end through the C++ host layer
        import torch
        n = args.own_files
        with mp.get_context("fork").Pool(min(ncpu, 64)) as pool:
            files = pool.map(_big, list(range(n)), chunksize=4)
        U = sum(len(f) for f in files)
        pl = ctx.pinned(U)
        views, pos = [], 0
        for f in files:
            pl[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
            views.append(pl[pos:pos + len(f)])
            pos += len(f)
        oopts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=1, password=b"pw", kdf_params={"i": 1000})
        arch = ctx.pinned(int(U * 1.05) + (8 << 20))
        blob = host.create_archive([(f"own/{i:05d}.bin", v) for i, v in enumerate(views)], compression=2, level=3, encryption=1, cipher_mode=1,
                                   key=oopts.key, phsf=oopts.phsf, max_chunk_size=0, device=0, workers=4, group_bytes=256 << 20, out=arch)
        torch.cuda.synchronize()
        a = pna.Archive.read_header(blob, ctx, verify=False)
        ro = pna.ReadOptions.with_password(b"pw")
        ro._keys[oopts.phsf] = oopts.key
        plan, ents = a.extract_plan(ro)
        for _ in range(4):
            plan.run()
        stage = plan.stage_ms()
        cnt = plan.counts()
        outs, st, _ = plan.fetch([len(f) for f in files])
        assert list(st) == [0] * n and all(outs[k].tobytes() == files[k] for k in range(0, n, max(1, n // 16)))
        plan.close()
        edt, _, eout, eoffs, est, _ = timed_extract(host, ctx, blob, oopts.phsf, oopts.key, U, n, workers=args.workers, group_mib=args.group_mib)
        assert est == [0] * n
        print(json.dumps({"config": "own", "files": n, "plain_bytes": U, "archive_bytes": int(blob.size), "codec": "this library's zstd writer + aes-256-ctr",
                          "kernel_only_GBps": U / (sum(stage.values()) * 1e-3) / 1e9, "kernel_only_ms": sum(stage.values()), "stage_ms": stage,
                          "counts": cnt, "extract_e2e_GBps": U / edt / 1e9, "extract_e2e_ms": edt * 1e3}), flush=True)
        # the same archive as a split archive (split_parts.rs writer, read.rs:105-165 reader): 64 MiB parts
        t0 = time.perf_counter()
        parts = host.split_archive(blob, 64 << 20)
        t_split = time.perf_counter() - t0
        mts = []
        for _ in range(3):
            t0 = time.perf_counter()
            hm = host.HostArchive.open_multipart(parts, pinned_device=0)
            t_join = time.perf_counter() - t0
            hm.set_key(oopts.phsf, oopts.key)
            _, moffs, mst = hm.extract_files(out=eout, device=0, workers=args.workers, group_bytes=args.group_mib << 20, verify=True)
            torch.cuda.synchronize()
            mts.append(time.perf_counter() - t0)
            hm.close()
        assert mst == [0] * n and all(eout[int(moffs[k]):int(moffs[k]) + len(files[k])].tobytes() == files[k] for k in range(0, n, max(1, n // 16)))
        print(json.dumps({"config": "own_split", "parts": len(parts), "part_bytes": 64 << 20, "split_ms": t_split * 1e3,
                          "split_GBps": int(blob.size) / t_split / 1e9, "join_ms": t_join * 1e3, "extract_e2e_ms": min(mts[1:]) * 1e3,
                          "extract_e2e_GBps": U / min(mts[1:]) / 1e9, "note": "split: host copy + one GPU CRC batch of the re-cut chunks; "
                          "read: parts joined into one owned chunk stream (pinned pool), then the single-archive path"}), flush=True)
        del files, pl, arch

    if args.only in ("", "cfg1"):
        n = args.cfg1_files
        sizes = [int(s) for s in corpus.lognormal_sizes(n, n * 107374)]     # 10,000 files sum to 1 GiB in the full config
        with mp.get_context("fork").Pool(min(ncpu, 64)) as pool:
            files = pool.map(_small, list(enumerate(sizes)), chunksize=32)
        U = sum(sizes)
        names = [f"c/{i:06d}" for i in range(n)]
        plain = ctx.pinned(U + 16)
        views, pos = [], 0
        for f in files:
            plain[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
            views.append(plain[pos:pos + len(f)])
            pos += len(f)
        arch = ctx.pinned(int(U * 1.05) + (8 << 20))
        import torch
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            blob = host.create_archive(list(zip(names, views)), compression=2, level=3, max_chunk_size=0, device=0, workers=1,
                                       group_bytes=4096 << 20, out=arch)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        import pna_oracle as O
        got = list(O.extract_all(blob.tobytes(), None))
        assert [d for _, d in got] == files, "GPU-created archive must extract bit-exactly with the reference reader"
        dt, t_index, out, offs, st, _ = timed_extract(host, ctx, blob, None, key, U, n)
        assert st == [0] * n
        t0 = time.perf_counter()
        ref_streams = oracle_encode(files, 2, 3, 0, 0, key, ncpu)
        cpu_create = time.perf_counter() - t0
        cpu_dt = cpu_decode_time(ref_streams, sizes, 2, 0, 0, key, ncpu)
        c_ref = sum(len(s) for s in ref_streams)
        print(json.dumps({"config": "cfg1", "files": n, "plain_bytes": U, "create_e2e_GBps": U / min(ts[1:]) / 1e9, "extract_e2e_GBps": U / dt / 1e9,
                          "archive_bytes": int(blob.size), "c_gpu_over_c_ref": float(blob.size) / c_ref,
                          "checked": "oracle.extract_all(GPU-created archive) == source files",
                          "cpu_create_GBps": U / cpu_create / 1e9, "cpu_extract_GBps": U / cpu_dt / 1e9, "cpu_cores": ncpu}), flush=True)


if __name__ == "__main__":
    main()
