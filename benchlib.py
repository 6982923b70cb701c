"""benchlib.py -- input preparation and CPU-baseline plumbing shared by bench.py's workloads (BASELINE.json configs 1-5).

Everything here is OUTSIDE the timed GPU regions: synthetic corpora (corpus.py), reference-dataflow encoding of the inputs
by the oracle (libzstd / zlib / OpenSSL -- the streams the reference would have written), PNA container framing, and the CPU
baseline legs (the oracle on the box's host cores).  Job arrays for the oracle's thread pools are filled with numpy so that a
million entries do not cost a million Python iterations."""
from __future__ import annotations

import ctypes as C
import multiprocessing as mp
import os
import struct
import time
import zlib

import numpy as np

import corpus

FILE_SIZE = 4 << 20
SIG = b"\x89PNA\r\n\x1a\n"

JOB_DT = np.dtype([("stream", "<u8"), ("len", "<u8"), ("compression", "u1"), ("encryption", "u1"), ("cipher_mode", "u1"), ("_pad", "u1"),
                   ("key", "u1", 32), ("out", "<u8"), ("cap", "<u8"), ("out_len", "<u8"), ("status", "<i4")], align=True)
ENC_DT = np.dtype([("plain", "<u8"), ("len", "<u8"), ("compression", "u1"), ("encryption", "u1"), ("cipher_mode", "u1"), ("_pad", "u1"),
                   ("level", "<i4"), ("key", "u1", 32), ("iv", "u1", 16), ("out", "<u8"), ("cap", "<u8"), ("out_len", "<u8"),
                   ("status", "<i4")], align=True)


def _oracle():
    import pna_oracle as O
    assert JOB_DT.itemsize == C.sizeof(O.Job) and ENC_DT.itemsize == C.sizeof(O.EncJob), "job layouts drifted from oracle/pna_oracle.c"
    return O, O.lib()


def _gen(i):
    return corpus.make_file(i, FILE_SIZE)


def gen_files(indices, threads):
    """corpus files (4 MiB each) by index, generated on `threads` processes"""
    with mp.get_context("fork").Pool(max(1, min(threads, 64))) as pool:
        return pool.map(_gen, list(indices), chunksize=4)


def pack(blobs, into=None):
    """concatenate bytes-likes into one uint8 array; returns (array, offsets[n+1])"""
    lens = np.fromiter((len(b) for b in blobs), dtype=np.int64, count=len(blobs))
    offs = np.zeros(len(blobs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    buf = into(int(offs[-1])) if into else np.empty(int(offs[-1]), dtype=np.uint8)
    for b, o in zip(blobs, offs):
        buf[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf, offs


def oracle_encode(plain: np.ndarray, offs, comp, level, enc, mode, key: bytes, threads, seed=7):
    """The reference's create dataflow on the host: every entry plain[offs[i]:offs[i+1]] -> [IV ||] cipher(compress(.)).
    Returns (streams array, stream offsets[n+1], seconds)."""
    O, L = _oracle()
    n = len(offs) - 1
    lens = np.diff(offs).astype(np.uint64)
    # per-entry bound: raw-block fallbacks bound zstd / deflate expansion; a little slack for the IV and CBC padding
    caps = (lens + (lens >> np.uint64(7)) + np.uint64(1024)).astype(np.uint64)
    out_offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(caps.astype(np.int64), out=out_offs[1:])
    out = np.empty(int(out_offs[-1]), dtype=np.uint8)
    jobs = np.zeros(n, dtype=ENC_DT)
    jobs["plain"] = plain.ctypes.data + np.asarray(offs[:-1], dtype=np.uint64)
    jobs["len"] = lens
    jobs["compression"], jobs["encryption"], jobs["cipher_mode"], jobs["level"] = comp, enc, mode, level
    jobs["key"] = np.frombuffer(key, dtype=np.uint8)
    jobs["iv"] = np.random.Generator(np.random.PCG64(seed)).integers(0, 256, (n, 16), dtype=np.uint8)
    jobs["out"] = out.ctypes.data + out_offs[:-1].astype(np.uint64)
    jobs["cap"] = caps
    t0 = time.perf_counter()
    L.pna_oracle_encode_batch_mt(C.cast(jobs.ctypes.data, C.POINTER(O.EncJob)), n, threads, None)
    dt = time.perf_counter() - t0
    assert (jobs["status"] == 0).all(), "oracle encode failed"
    slens = jobs["out_len"].astype(np.int64)
    s_offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(slens, out=s_offs[1:])
    streams = np.empty(int(s_offs[-1]), dtype=np.uint8)
    for i in range(n) if n <= 8192 else ():
        streams[s_offs[i]:s_offs[i + 1]] = out[out_offs[i]:out_offs[i] + slens[i]]
    if n > 8192:   # many small entries: gather with one fancy index instead of a Python loop
        idx = np.repeat(out_offs[:-1] - s_offs[:-1], slens) + np.arange(int(s_offs[-1]), dtype=np.int64)
        streams[:] = out[idx]
    return streams, s_offs, dt


def oracle_decode_time(streams: np.ndarray, s_offs, sizes, comp, enc, mode, key: bytes, threads, crc_impl=2, passes=1):
    """The reference's extract dataflow on the host (cli/src/command/extract.rs:868-1019): ONE thread walks the archive and
    checks every chunk CRC (crc_impl 2 = the PCLMULQDQ folding CRC crc32fast runs, 1 = zlib's table CRC), `threads` workers
    decrypt + decompress one entry each.  Returns (seconds per pass, output array, output offsets)."""
    O, L = _oracle()
    n = len(s_offs) - 1
    sizes = np.asarray(sizes, dtype=np.int64)
    o_offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes, out=o_offs[1:])
    out = np.empty(max(int(o_offs[-1]), 1), dtype=np.uint8)
    jobs = np.zeros(n, dtype=JOB_DT)
    jobs["stream"] = streams.ctypes.data + np.asarray(s_offs[:-1], dtype=np.uint64)
    jobs["len"] = np.diff(s_offs).astype(np.uint64)
    jobs["compression"], jobs["encryption"], jobs["cipher_mode"] = comp, enc, mode
    jobs["key"] = np.frombuffer(key, dtype=np.uint8)
    jobs["out"] = out.ctypes.data + o_offs[:-1].astype(np.uint64)
    jobs["cap"] = sizes.astype(np.uint64)
    crc = np.zeros(max(n, 1), dtype=np.uint32)
    t0 = time.perf_counter()
    for _ in range(passes):
        L.pna_oracle_decode_batch_mt(C.cast(jobs.ctypes.data, C.POINTER(O.Job)), n, threads, crc_impl, crc.ctypes.data)
    dt = (time.perf_counter() - t0) / passes
    assert (jobs["status"] == 0).all() and (jobs["out_len"].astype(np.int64) == sizes).all(), "oracle decode failed"
    return dt, out, o_offs


def single_core_rates(sample: bytes, key: bytes):
    """Single-core rates of the reference's stages on this box: chunk CRC both ways, AES-256-CTR,
    zstd level 3 encode / decode -- what one rayon worker / the one iterating thread of the reference can do."""
    O, L = _oracle()
    L.pna_oracle_crc32_fold.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
    L.pna_oracle_crc32_fold.restype = C.c_uint32
    n = len(sample)
    res = {"sample_bytes": n, "pclmul": bool(L.pna_oracle_have_pclmul())}

    def rate(f, nbytes, reps=3):
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            f()
            best = min(best, time.perf_counter() - t0)
        return nbytes / best / 1e9
    res["crc32_zlib_GBps"] = rate(lambda: zlib.crc32(sample), n)
    res["crc32_fold_pclmul_GBps"] = rate(lambda: L.pna_oracle_crc32_fold(0, sample, n), n)
    res["aes256_ctr_GBps"] = rate(lambda: O.ctr(1, key, bytes(16), sample), n)
    comp = O.compress(2, sample, 3)
    res["zstd3_encode_GBps"] = rate(lambda: O.compress(2, sample, 3), n, reps=2)
    res["zstd3_decode_GBps"] = rate(lambda: O.decompress(2, comp, n), n)
    res["zstd3_ratio"] = n / len(comp)
    return res


# ---------------------------------------------------------------------------------------------- container framing
def _chunk(parts, ty, data):
    parts.append(struct.pack(">I", len(data)) + ty)
    parts.append(data)
    parts.append(struct.pack(">I", zlib.crc32(data, zlib.crc32(ty))))


_FRAME = {}


def _frame_range(r):
    lo, hi = r
    g = _FRAME
    streams, s_offs, sizes, hdr6, phsf, iv_len, name_fmt = g["streams"], g["s_offs"], g["sizes"], g["hdr6"], g["phsf"], g["iv_len"], g["name_fmt"]
    parts = []
    for i in range(lo, hi):
        s = streams[s_offs[i]:s_offs[i + 1]].tobytes()
        _chunk(parts, b"FHED", hdr6 + (name_fmt % i).encode())
        _chunk(parts, b"fSIZ", int(sizes[i]).to_bytes(8, "big").lstrip(b"\0") or b"\0")
        if iv_len:
            _chunk(parts, b"PHSF", phsf)
            _chunk(parts, b"FDAT", s[:iv_len])
        if len(s) > iv_len:
            _chunk(parts, b"FDAT", s[iv_len:])
        _chunk(parts, b"FEND", b"")
    return b"".join(parts)


def frame_archive(streams: np.ndarray, s_offs, sizes, hdr6: bytes, phsf: str, iv_len: int, name_fmt: str, into, threads):
    """PNA container around the entries' streams: signature, AHED, per entry FHED,fSIZ,[PHSF,FDAT(iv)],FDAT(body),FEND, AEND
    (wire order lib/src/entry.rs:895-912; chunk CRCs by zlib -- input preparation).  Returns the archive in `into(nbytes)`."""
    n = len(s_offs) - 1
    _FRAME.update(streams=streams, s_offs=s_offs, sizes=sizes, hdr6=hdr6, phsf=phsf.encode(), iv_len=iv_len, name_fmt=name_fmt)
    step = max(1, min(65536, (n + 4 * threads - 1) // (4 * threads)))
    ranges = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
    if n >= 4096 and threads > 1:
        with mp.get_context("fork").Pool(min(threads, 32)) as pool:
            blobs = pool.map(_frame_range, ranges)
    else:
        blobs = [_frame_range(r) for r in ranges]
    head, tail = [], []
    head.append(SIG)
    _chunk(head, b"AHED", bytes(8))
    _chunk(tail, b"AEND", b"")
    head, tail = b"".join(head), b"".join(tail)
    total = len(head) + sum(len(b) for b in blobs) + len(tail)
    buf = into(total)
    pos = 0
    for b in [head] + blobs + [tail]:
        buf[pos:pos + len(b)] = np.frombuffer(b, dtype=np.uint8)
        pos += len(b)
    _FRAME.clear()
    return buf


def frame_solid(inner_stream: bytes, sdat: int, into):
    """one solid entry: SHED(zstd, no cipher), SDAT bodies of `sdat` bytes, SEND (lib/src/entry.rs:471-483)"""
    parts = [SIG]
    _chunk(parts, b"AHED", bytes(8))
    _chunk(parts, b"SHED", bytes([0, 0, 2, 0, 0]))
    mv = memoryview(inner_stream)
    for o in range(0, len(inner_stream), sdat):
        _chunk(parts, b"SDAT", bytes(mv[o:o + sdat]))
    _chunk(parts, b"SEND", b"")
    _chunk(parts, b"AEND", b"")
    total = sum(len(p) for p in parts)
    buf = into(total)
    pos = 0
    for p in parts:
        buf[pos:pos + len(p)] = np.frombuffer(p, dtype=np.uint8)
        pos += len(p)
    return buf


def inner_store_archive(files, names, fdat: int = 1 << 20) -> bytes:
    """the chunk stream a solid entry carries: STORE normal entries FHED,fSIZ,FDAT..,FEND (lib/src/entry.rs:401-423)"""
    parts = []
    for f, nm in zip(files, names):
        _chunk(parts, b"FHED", bytes(6) + nm.encode())
        _chunk(parts, b"fSIZ", len(f).to_bytes(8, "big").lstrip(b"\0") or b"\0")
        for o in range(0, len(f), fdat):
            _chunk(parts, b"FDAT", f[o:o + fdat])
        _chunk(parts, b"FEND", b"")
    return b"".join(parts)


def cpu_affinity_for_rank(rank: int, world: int):
    """Give every rank of a multi-process run its own slice of the host cores (and keep its worker threads there): eight
    ranks x (workers + index threads) on one box otherwise migrate over each other's caches.  Returns the cpu list."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    if world <= 1 or len(cpus) < world:
        return cpus
    k = len(cpus) // world
    mine = cpus[rank * k:(rank + 1) * k]
    os.sched_setaffinity(0, mine)
    return mine
