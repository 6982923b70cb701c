"""benchcfg.py -- BASELINE.json configs 1, 3, 4, 5 at their stated sizes, for bench.py's `configs` object (N = 1).

Every configuration reports: value (kernel-only where a single plan expresses it), e2e (through the C++ host layer over the C
ABI with pinned host buffers, host<->device copies inside the timed region), cpu_baseline (the oracle port on the box's cores,
bounded sample) and a parity check against the source files; create configurations add c_gpu_over_c_ref (archive bytes against
the reference's encoder at the same level: libzstd level 3 / zlib level 6).  Content comes from the same corpus generator as
the headline shard: 4 MiB corpus files, and -- for the many-small-files shapes -- consecutive slices of their concatenation."""
from __future__ import annotations

import os
import statistics
import time

import numpy as np

import benchlib
import corpus

KEY = bytes(range(32))


def _timed(f, reps):
    ts = []
    res = None
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        res = f()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts[1:]) if reps > 1 else ts[-1], res


def _extract_e2e(host, ctx, buf, phsf, n, U, workers, group_mib, reps=2):
    out = ctx.pinned(U + 16 * n + 64)
    state = {}

    def once():
        ha = host.HostArchive(buf)
        t1 = time.perf_counter()
        if phsf:
            ha.set_key(phsf, KEY)
        _, offs, st = ha.extract_files(out=out, device=ctx.device, workers=workers, group_bytes=group_mib << 20, verify=True)
        state.update(offs=offs, st=st, t_index=t1)
        ha.close()
    t0s = []

    def wrapped():
        t0 = time.perf_counter()
        once()
        t0s.append(state["t_index"] - t0)
    dt, _ = _timed(wrapped, reps)
    return dt, min(t0s), out, state["offs"], state["st"]


def cfg1(pna, host, ctx, corpus_np, threads, scale, workers):
    """config 1: create + extract, 1 GiB corpus of 10k files (log-normal sizes), zstd level 3, no encryption"""
    import pna_oracle as O
    n = max(16, int(10000 * scale))
    total = min(int((1 << 30) * scale), corpus_np.size)
    sizes = corpus.lognormal_sizes(n, total)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    plain = ctx.pinned(total)
    plain[:] = corpus_np[:total]
    names = [f"c/{i:06d}" for i in range(n)]
    views = [plain[int(offs[i]):int(offs[i + 1])] for i in range(n)]
    arch = ctx.pinned(int(total * 1.05) + (8 << 20))
    state = {}

    def create():
        state["blob"] = host.create_archive((names, plain, offs), compression=2, level=3, max_chunk_size=0, device=ctx.device, workers=2,
                                            group_bytes=256 << 20, out=arch)
    c_dt, _ = _timed(create, 2)
    blob = state["blob"]
    got = list(O.extract_all(blob.tobytes(), None))
    assert len(got) == n and all(d == views[i].tobytes() for i, (_, d) in enumerate(got)), "GPU-created archive must extract bit-exactly with the reference reader"
    x_dt, t_index, out, xoffs, st = _extract_e2e(host, ctx, blob, None, n, total, workers, 64)
    assert st == [0] * n
    for k in range(0, n, max(1, n // 64)):
        assert out[int(xoffs[k]):int(xoffs[k]) + int(sizes[k])].tobytes() == views[k].tobytes()
    ctx.pinned_free(out)
    del out
    ref_streams, r_offs, cpu_c = benchlib.oracle_encode(corpus_np, offs, 2, 3, 0, 0, KEY, threads)
    cpu_x, _, _ = benchlib.oracle_decode_time(ref_streams, r_offs, sizes, 2, 0, 0, KEY, threads, crc_impl=2, passes=2)
    res_blob = int(blob.size)
    del blob, views, state
    ctx.pinned_free(arch)
    ctx.pinned_free(plain)
    return {"shape": f"{n} files, log-normal sizes, {total} bytes, zstd level 3, no encryption", "plain_bytes": int(total),
            "create": {"e2e": {"value": total / c_dt / 1e9, "unit": "GB/s", "ms": c_dt * 1e3, "h2d_bytes": int(total), "d2h_bytes": res_blob},
                       "archive_bytes": res_blob, "c_gpu_over_c_ref": float(res_blob) / float(r_offs[-1] + 60 * n),
                       "cpu_baseline": {"value": total / cpu_c / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "sample": "whole corpus, libzstd level 3"}},
            "extract": {"e2e": {"value": total / x_dt / 1e9, "unit": "GB/s", "ms": x_dt * 1e3, "index_pass_ms": t_index * 1e3,
                                "h2d_bytes": res_blob, "d2h_bytes": int(total)},
                        "cpu_baseline": {"value": total / cpu_x / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                         "sample": "whole corpus (reference-written streams), 1 CRC thread + workers"}},
            "checked": "oracle.extract_all(GPU-created archive) == every source file; our extract == sampled source files"}


def cfg3(pna, host, ctx, corpus_np, threads, scale, workers):
    """config 3: extract 8 GiB archive of 1M small (<= 16 KiB) files, deflate (zlib 6) + Camellia-256-CBC"""
    n = max(1024, int((1 << 20) * scale))
    rng = np.random.Generator(np.random.PCG64(3))
    sizes = rng.integers(1, 16385, n).astype(np.int64)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=offs[1:])
    U = int(offs[-1])
    assert U <= corpus_np.size
    opts = pna.WriteOptions(compression=1, encryption=2, cipher_mode=0, password=b"pw", kdf_params={"i": 1000})
    streams, s_offs, enc_dt = benchlib.oracle_encode(corpus_np, offs, 1, 6, 2, 0, KEY, threads)
    buf = benchlib.frame_archive(streams, s_offs, sizes, bytes([0, 0, 0, 1, 2, 0]), opts.phsf, 16, "s/%07d", ctx.pinned, threads)
    dt, t_index, out, xoffs, st = _extract_e2e(host, ctx, buf, opts.phsf, n, U, max(2, workers), 64, reps=3)
    assert st == [0] * n
    for k in range(0, n, max(1, n // 256)):
        assert out[int(xoffs[k]):int(xoffs[k]) + int(sizes[k])].tobytes() == corpus_np[offs[k]:offs[k + 1]].tobytes()
    ns = min(n, 1 << 17)
    cpu_dt, _, _ = benchlib.oracle_decode_time(streams, s_offs[:ns + 1], sizes[:ns], 1, 2, 0, KEY, threads, crc_impl=2)
    buf_size = int(buf.size)
    ctx.pinned_free(out)
    ctx.pinned_free(buf)
    del out, buf
    return {"shape": f"{n} files, sizes uniform 1..16384 ({U} bytes), zlib level 6 + Camellia-256-CBC, {6 * n + 2} chunks",
            "plain_bytes": U, "stream_bytes": int(s_offs[-1]), "archive_bytes": buf_size,
            "e2e": {"value": U / dt / 1e9, "unit": "GB/s", "ms": dt * 1e3, "index_pass_ms": t_index * 1e3, "h2d_bytes": buf_size, "d2h_bytes": U,
                    "path": "pna::Archive index pass (6.3 M chunk frames) + extract_files, all chunk CRCs checked on the GPU"},
            "cpu_baseline": {"value": float(sizes[:ns].sum()) / cpu_dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"first {ns} entries: zlib inflate + OpenSSL Camellia-256-CBC + folding CRC, 1 CRC thread + {threads} workers"},
            "checked": "every 4096th file == source bytes; all statuses OK"}


def cfg4(pna, host, ctx, files, plain_pinned, p_offs, threads, workers, compression, level, label):
    """config 4: create 16 GiB archive, GPU zstd / deflate encode + AES-256-CTR + CRC-32, round trip checked by the reference reader"""
    import pna_oracle as O
    n = len(files)
    U = int(p_offs[-1])
    opts = pna.WriteOptions(compression=compression, encryption=1, cipher_mode=1, password=b"pw", kdf_params={"i": 1000})
    names = [f"corpus/{i:07d}.bin" for i in range(n)]
    views = [plain_pinned[int(p_offs[i]):int(p_offs[i + 1])] for i in range(n)]
    arch = ctx.pinned(int(U * 1.03) + (64 << 20))
    rng = np.random.Generator(np.random.PCG64(4))
    ivs = rng.bytes(16 * n)
    state = {}

    def create():
        state["blob"] = host.create_archive(list(zip(names, views)), compression=compression, level=level, encryption=1, cipher_mode=1, key=KEY,
                                            phsf=opts.phsf, ivs=ivs, max_chunk_size=0, device=ctx.device, workers=4, group_bytes=256 << 20, out=arch)
    c_dt, _ = _timed(create, 2)
    blob = state["blob"]
    # kernel-only: the whole corpus as one encode plan, inputs resident in HBM
    step = max(1, n // 1024)
    ents = [{"plain": v, "compression": compression, "level": level, "encryption": 1, "cipher_mode": 1, "key": KEY, "iv": ivs[16 * i:16 * i + 16],
             "max_chunk_size": 0} for i, v in enumerate(views)]
    eplan = ctx.encode_plan(ents)
    eplan.run()
    eplan.run()
    stage = eplan.stage_ms()
    k_ms = sum(stage.values())
    eplan.close()
    # round trip: the reference reader (oracle) extracts a sample of entries from the archive we wrote; ours extracts all of it
    a = pna.Archive.read_header(blob, ctx, verify=False) if n <= 8192 else None
    checked = 0
    if a is not None:
        for i, e in enumerate(a.entries()):
            if i % max(1, n // 32):
                continue
            s = b"".join(bytes(b) for b in e.bodies)
            assert O.decode_stream(s, compression, 1, 1, KEY, None) == files[i], "GPU-created entry is not reference-readable"
            checked += 1
    # (the deflate streams this library writes are one zlib stream per file: their bit-serial entropy stage runs one LANE per
    #  stream, so reading 4 MiB streams back is slow -- sampled through the reference pipeline above instead of all through ours)
    x_dt = None
    if compression == 2:
        x_dt, _, out, xoffs, st = _extract_e2e(host, ctx, blob, opts.phsf, n, U, workers, 128, reps=1)
        assert st == [0] * n
        for k in range(0, n, max(1, n // 64)):
            assert out[int(xoffs[k]):int(xoffs[k]) + len(files[k])].tobytes() == files[k]
        ctx.pinned_free(out)
        del out
    # reference encoder at the same level: archive bytes and host-core rate (deflate: zlib level 6 is slow -- an eighth of the corpus)
    ns = n if compression == 2 else max(1, n // 8)
    plain_np = np.frombuffer(plain_pinned, dtype=np.uint8)
    _, r_offs, cpu_dt = benchlib.oracle_encode(plain_np, p_offs[:ns + 1], compression, level, 1, 1, KEY, threads)
    c_ref = float(r_offs[-1])
    gpu_stream_bytes = float(blob.size) - n * 150.0   # container framing (FHED, fSIZ, PHSF, FDAT headers, FEND) is not codec output
    c_gpu_sample = gpu_stream_bytes * (ns / n)
    blob_size = int(blob.size)
    del blob, state, a
    ctx.pinned_free(arch)
    return {"shape": f"{n} x 4 MiB files ({U} bytes), GPU {label} + AES-256-CTR + chunk CRC-32", "plain_bytes": U, "archive_bytes": blob_size,
            "value": U / (k_ms * 1e-3) / 1e9, "unit": "GB/s", "kernel_ms": k_ms, "stage_ms": stage,
            "e2e": {"value": U / c_dt / 1e9, "unit": "GB/s", "ms": c_dt * 1e3, "h2d_bytes": U, "d2h_bytes": blob_size},
            "ratio": U / float(blob_size), "c_gpu_over_c_ref": c_gpu_sample / c_ref,
            "c_ref": f"oracle: {'libzstd level 3' if compression == 2 else 'zlib level 6'} on {ns} of the {n} files",
            "cpu_baseline": {"value": float(p_offs[ns]) / cpu_dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"{ns} x 4 MiB entries, reference create dataflow (compress + AES-256-CTR per worker)"},
            "extract_back_e2e_GBps": (U / x_dt / 1e9) if x_dt else None,
            "checked": f"{checked} entries decoded by the reference pipeline (oracle: libzstd/zlib + OpenSSL) == source files; " + ("the whole archive extracted by this library == sampled source files" if x_dt else "")}


def create_xz(pna, host, ctx, files, plain_pinned, p_offs, threads):
    """create with Compress::XZ (entry/write.rs:263) on a quarter of the cfg2 shard: GPU LZMA2 (one chunk per 32 KiB segment) +
    AES-256-CTR + CRC-32; liblzma preset 6 (the reference's encoder and level) on a bounded sample beside it"""
    import lzma
    from concurrent.futures import ThreadPoolExecutor
    import pna_oracle as O
    n = len(files)
    U = int(p_offs[n])
    views = [plain_pinned[int(p_offs[i]):int(p_offs[i + 1])] for i in range(n)]
    names = [f"corpus/{i:07d}.bin" for i in range(n)]
    opts = pna.WriteOptions(compression=4, encryption=1, cipher_mode=1, password=b"pw", kdf_params={"i": 1000})
    rng = np.random.Generator(np.random.PCG64(44))
    ivs = rng.bytes(16 * n)
    arch = ctx.pinned(int(U * 1.03) + (64 << 20))
    state = {}

    def create():
        state["blob"] = host.create_archive(list(zip(names, views)), compression=4, level=6, encryption=1, cipher_mode=1, key=opts.key,
                                            phsf=opts.phsf, ivs=ivs, max_chunk_size=0, device=ctx.device, workers=4, group_bytes=256 << 20, out=arch)
    c_dt, _ = _timed(create, 2)
    blob = state["blob"]
    ents = [{"plain": v, "compression": 4, "level": 6, "encryption": 1, "cipher_mode": 1, "key": opts.key, "iv": ivs[16 * i:16 * i + 16],
             "max_chunk_size": 0} for i, v in enumerate(views)]
    eplan = ctx.encode_plan(ents)
    eplan.run()
    eplan.run()
    stage = eplan.stage_ms()
    k_ms = sum(stage.values())
    eplan.close()
    a = pna.Archive.read_header(blob, ctx, verify=False)
    checked = 0
    for i, e in enumerate(a.entries()):
        if i % max(1, n // 16):
            continue
        s = b"".join(bytes(b) for b in e.bodies)
        assert O.decode_stream(s, 4, 1, 1, opts.key, None) == files[i], "GPU-created xz entry is not reference-readable"
        checked += 1
    back = dict((e.name, d) for e, d in a.read_all(pna.ReadOptions.with_password(b"pw")))   # and all of it by our own xz decoder
    assert all(bytes(back[nm]) == f for nm, f in zip(names, files))
    ns = max(1, min(n, threads))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        ref = list(ex.map(lambda f: len(lzma.compress(f, preset=6)), files[:ns]))
    cpu_dt = time.perf_counter() - t0
    blob_size = int(blob.size)
    c_gpu_sample = (float(blob_size) - n * 150.0) * (ns / n)
    del blob, state, a, back
    ctx.pinned_free(arch)
    return {"shape": f"{n} x 4 MiB files ({U} bytes), GPU xz (LZMA2, one chunk per 32 KiB segment, level 6 request) + AES-256-CTR + chunk CRC-32",
            "plain_bytes": U, "archive_bytes": blob_size, "value": U / (k_ms * 1e-3) / 1e9, "unit": "GB/s", "kernel_ms": k_ms, "stage_ms": stage,
            "e2e": {"value": U / c_dt / 1e9, "unit": "GB/s", "ms": c_dt * 1e3, "h2d_bytes": U, "d2h_bytes": blob_size},
            "ratio": U / float(blob_size), "c_gpu_over_c_ref": c_gpu_sample / float(sum(ref)),
            "c_ref": f"liblzma preset 6 (python lzma = the C library the reference links) on {ns} of the {n} files",
            "cpu_baseline": {"value": ns * len(files[0]) / cpu_dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "reference",
                             "sample": f"{ns} x 4 MiB entries, liblzma preset 6, one stream per thread (compression only, no cipher)"},
            "checked": f"{checked} entries decoded by liblzma + OpenSSL == source files; the whole archive extracted by this library's xz decoder == source files"}


def cfg5(pna, host, ctx, files, threads, workers):
    """config 5: solid-mode zstd archive (ONE entry, one reference-written frame), block-parallel decode on 1 GPU vs per-entry mode"""
    import pna_oracle as O
    n = len(files)
    U = sum(len(f) for f in files)
    names = [f"solid/{i:05d}.bin" for i in range(n)]
    inner = benchlib.inner_store_archive(files, names)
    t0 = time.perf_counter()
    stream = O.compress(2, inner, 3)            # one zstd frame over the whole inner archive, as the reference's solid writer produces
    t_comp = time.perf_counter() - t0
    buf = benchlib.frame_solid(stream, 32768, ctx.pinned)
    dt, _, out, xoffs, st = _extract_e2e(host, ctx, buf, None, n, U, workers, 128, reps=2)
    assert st == [0] * n
    for k in range(0, n, max(1, n // 128)):
        assert out[int(xoffs[k]):int(xoffs[k]) + len(files[k])].tobytes() == files[k]
    buf_size = int(buf.size)
    ctx.pinned_free(out)
    ctx.pinned_free(buf)
    del out, buf
    # kernel-only: the solid stream as one decode plan resident in HBM
    sarr = np.frombuffer(stream, dtype=np.uint8)
    plan = ctx.decode_plan([{"bodies": [sarr], "compression": 2, "encryption": 0, "cipher_mode": 0, "key": None, "raw_size_hint": None}])
    plan.run()
    plan.run()
    stage = plan.stage_ms()
    k_ms = sum(stage.values())
    lens, pst = plan.lengths()
    assert pst == [0] and lens[0] == len(inner)
    plan.close()
    # the same files as a per-entry archive (zstd 3, no cipher)
    plain, p_offs = benchlib.pack(files)
    streams, s_offs, _ = benchlib.oracle_encode(plain, p_offs, 2, 3, 0, 0, KEY, threads)
    buf2 = benchlib.frame_archive(streams, s_offs, [len(f) for f in files], bytes([0, 0, 0, 2, 0, 0]), "", 0, "e/%05d", ctx.pinned, threads)
    dt2, _, out2, xo2, st2 = _extract_e2e(host, ctx, buf2, None, n, U, workers, 128, reps=2)
    assert st2 == [0] * n
    ctx.pinned_free(out2)
    ctx.pinned_free(buf2)
    del out2, buf2
    t0 = time.perf_counter()
    got = O.decompress(2, stream, len(inner))
    cpu_dt = time.perf_counter() - t0
    assert got == inner
    return {"shape": f"{n} x 4 MiB files in ONE solid entry (inner STORE entries), zstd level 3, one reference-written frame of {len(stream)} bytes in "
                     f"{(len(stream) + 32767) // 32768} SDAT chunks of 32 KiB", "plain_bytes": U, "inner_stream_bytes": len(inner),
            "value": len(inner) / (k_ms * 1e-3) / 1e9, "unit": "GB/s", "kernel_ms": k_ms, "stage_ms": stage,
            "e2e": {"value": U / dt / 1e9, "unit": "GB/s", "ms": dt * 1e3, "h2d_bytes": buf_size, "d2h_bytes": 2 * U,
                    "path": "pna::Archive: SDAT chunk CRCs + decode on the GPU, decoded stream kept in HBM for the inner chunk CRC check, host copy for the inner index pass, range copies of the STORE entries"},
            "per_entry_mode_e2e": {"value": U / dt2 / 1e9, "unit": "GB/s", "ms": dt2 * 1e3},
            "cpu_baseline": {"value": U / cpu_dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                             "sample": "the whole solid stream, libzstd on one thread (the reference decodes a solid entry on one thread: lib/src/entry.rs:567-583)"},
            "input_prep_s": t_comp, "checked": "every 8th inner file == source bytes; inner chunk CRCs verified on the GPU"}


def run_all(pna, host, ctx, all_files, threads, scale, workers):
    res = {}
    t_all = time.perf_counter()

    def guarded(name, f):
        t0 = time.perf_counter()
        try:
            res[name] = f()
        except Exception as e:   # a failing configuration must not hide the headline line
            res[name] = {"error": f"{type(e).__name__}: {e}"}
        res[name]["wall_s"] = round(time.perf_counter() - t0, 1)
    n5 = max(4, int(1024 * scale))
    guarded("cfg5", lambda: cfg5(pna, host, ctx, all_files[:n5], threads, workers))
    n4 = len(all_files)
    plain_pinned = ctx.pinned(sum(len(f) for f in all_files))
    p_offs = np.zeros(n4 + 1, dtype=np.int64)
    pos = 0
    for i, f in enumerate(all_files):
        plain_pinned[pos:pos + len(f)] = np.frombuffer(f, dtype=np.uint8)
        pos += len(f)
        p_offs[i + 1] = pos
    guarded("cfg4_zstd", lambda: cfg4(pna, host, ctx, all_files, plain_pinned, p_offs, threads, workers, 2, 3, "zstd (level 3 request)"))
    guarded("cfg4_deflate", lambda: cfg4(pna, host, ctx, all_files, plain_pinned, p_offs, threads, workers, 1, 6, "deflate (level 6 request)"))
    nx = max(4, int(256 * scale))
    guarded("create_xz", lambda: create_xz(pna, host, ctx, all_files[:nx], plain_pinned, p_offs, threads))
    corpus_np = np.frombuffer(plain_pinned, dtype=np.uint8)
    guarded("cfg3", lambda: cfg3(pna, host, ctx, corpus_np, threads, scale, workers))
    guarded("cfg1", lambda: cfg1(pna, host, ctx, corpus_np, threads, scale, workers))
    del corpus_np
    ctx.pinned_free(plain_pinned)
    res["total_wall_s"] = round(time.perf_counter() - t_all, 1)
    return res
