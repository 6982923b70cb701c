"""ctypes binding of libpna_host.so -- the C++ host layer (include/pna_host.hpp) above the C ABI: index pass, entry
grouping, pipelined multi-context extract, archive writer.  KDF stays here in Python (host work in the reference too)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpna_host.so")
EXPORTS = ["pnah_open", "pnah_close", "pnah_entry_count", "pnah_entry_get", "pnah_chunk_count", "pnah_set_key", "pnah_prepare",
           "pnah_file_count", "pnah_file_get", "pnah_file_sizes", "pnah_extract_files", "pnah_create", "pnah_create_bound", "pnah_create_solid", "pnah_create_solid_bound",
           "pnah_open_file", "pnah_extract_to_dir", "pnah_create_from_files", "pnah_open_multipart", "pnah_split",
           "pnah_extract_range", "pnah_extract_files_on", "pnah_create_on"]


class EntryInfo(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("data_kind", C.c_uint8), ("compression", C.c_uint8), ("encryption", C.c_uint8),
                ("cipher_mode", C.c_uint8), ("n_bodies", C.c_uint32), ("compressed_size", C.c_uint64),
                ("raw_file_size", C.c_uint64), ("name", C.c_char_p), ("phsf", C.c_char_p)]


class IoStats(C.Structure):
    _fields_ = [("files", C.c_uint64), ("dirs", C.c_uint64), ("skipped", C.c_uint64), ("bytes", C.c_uint64),
                ("index_ms", C.c_double), ("gpu_ms", C.c_double), ("io_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        _ffi.lib()   # libpna_cuda.so first (rpath $ORIGIN resolves it as well)
        L = C.CDLL(LIB_PATH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.pnah_open.argtypes = [vp, u64, C.POINTER(vp), C.c_char_p, u64]
        L.pnah_close.argtypes = [vp]
        L.pnah_close.restype = None
        L.pnah_entry_count.argtypes = [vp]
        L.pnah_entry_count.restype = u32
        L.pnah_chunk_count.argtypes = [vp]
        L.pnah_chunk_count.restype = u32
        L.pnah_entry_get.argtypes = [vp, u32, C.POINTER(EntryInfo)]
        L.pnah_set_key.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.pnah_prepare.argtypes = [vp, C.c_int, C.c_char_p, u64]
        L.pnah_file_count.argtypes = [vp]
        L.pnah_file_count.restype = u32
        L.pnah_file_get.argtypes = [vp, u32, C.POINTER(C.c_char_p), C.POINTER(u64)]
        L.pnah_file_sizes.argtypes = [vp, C.POINTER(u64), C.POINTER(C.c_int32)]
        L.pnah_extract_files.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(C.c_int32), C.c_int, C.c_int, u64, C.c_int, C.c_char_p, u64]
        L.pnah_extract_range.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(C.c_int32), C.c_int, C.c_int, u64, C.c_int, u64, u64, C.c_char_p, u64]
        L.pnah_extract_files_on.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(C.c_int32), C.POINTER(C.c_int), u32, C.c_int, u64, C.c_int,
                                            C.c_char_p, u64]
        L.pnah_create_on.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(u64), C.c_char_p, C.c_uint8, C.c_int32,
                                     C.c_uint8, C.c_uint8, C.c_char_p, C.c_char_p, u32, C.POINTER(C.c_int), u32, C.c_int, u64, vp, u64,
                                     C.POINTER(u64), C.c_char_p, u64]
        L.pnah_create.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(u64), C.c_char_p, C.c_uint8, C.c_int32,
                                  C.c_uint8, C.c_uint8, C.c_char_p, C.c_char_p, u32, C.c_int, C.c_int, u64, vp, u64, C.POINTER(u64),
                                  C.c_char_p, u64]
        L.pnah_create_bound.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(u64), C.c_uint8, C.c_uint8, C.c_char_p, u32]
        L.pnah_create_bound.restype = u64
        L.pnah_create_solid_bound.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(u64), C.c_uint8, C.c_uint8, C.c_uint8, C.c_char_p, u32]
        L.pnah_create_solid_bound.restype = u64
        L.pnah_create_solid.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(u64), C.c_uint8, C.c_int32, C.c_uint8, C.c_uint8,
                                        C.c_char_p, C.c_char_p, u32, C.c_int, vp, u64, C.POINTER(u64), C.c_char_p, u64]
        L.pnah_open_file.argtypes = [C.c_char_p, C.POINTER(vp), C.c_char_p, u64]
        L.pnah_split.argtypes = [vp, u64, u64, C.c_int, vp, u64, C.POINTER(u64), C.POINTER(u64), u32, C.POINTER(u32), C.c_char_p, u64]
        L.pnah_open_multipart.argtypes = [C.POINTER(vp), C.POINTER(u64), u32, C.c_int, C.POINTER(vp), C.c_char_p, u64]
        L.pnah_extract_to_dir.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, u64, u64, C.c_int, C.c_int, C.POINTER(IoStats),
                                          C.POINTER(C.c_int32), C.c_char_p, u64]
        L.pnah_create_from_files.argtypes = [u32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint8, C.c_int32, C.c_uint8, C.c_uint8,
                                             C.c_char_p, C.c_char_p, u32, C.c_char_p, C.c_int, C.c_int, u64, C.c_int, C.POINTER(IoStats),
                                             C.c_char_p, u64]
        _lib = L
    return _lib


class HostError(Exception):
    def __init__(self, kind, msg):
        super().__init__(msg)
        self.kind = kind


class HostArchive:
    """pna::Archive (C++): read_header_from_slice + batched, pipelined extract of every FILE entry."""

    def __init__(self, data):
        self.L = lib()
        self.buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.L.pnah_open(self.buf.ctypes.data, self.buf.size, C.byref(h), err, 512)
        if rc:
            raise HostError(rc, err.value.decode())
        self.h = h

    @classmethod
    def open_multipart(cls, parts, pinned_device=-1):
        """Split archive (archive/read.rs:105-165): `parts` in order, bytes-like or uint8 arrays; the handle owns a joined copy
        (pinned_device >= 0: in pinned memory of that device's pool, so uploads are DMA)."""
        self = cls.__new__(cls)
        self.L = lib()
        bufs = [p if isinstance(p, np.ndarray) else np.frombuffer(p, dtype=np.uint8) for p in parts]
        self.buf = None
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        lens = (C.c_uint64 * len(bufs))(*[b.size for b in bufs])
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.L.pnah_open_multipart(ptrs, lens, len(bufs), pinned_device, C.byref(h), err, 512)
        if rc:
            raise HostError(rc, err.value.decode())
        self.h = h
        return self

    @classmethod
    def open_file(cls, path: str):
        """Archive over an mmap of `path` (cli/src/utils/mmap.rs:36-45); the handle owns the mapping."""
        self = cls.__new__(cls)
        self.L = lib()
        self.buf = None
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.L.pnah_open_file(os.fsencode(path), C.byref(h), err, 512)
        if rc:
            raise HostError(rc, err.value.decode())
        self.h = h
        return self

    def extract_to_dir(self, out_dir: str, device=0, workers=3, group_bytes=128 << 20, window_bytes=2 << 30, io_threads=8, verify=True):
        """`pna extract` data path: GPU decode in pinned windows, files written by io_threads writers.  Returns (stats, statuses)."""
        self.prepare(device)
        nf = int(self.L.pnah_file_count(self.h))
        st = (C.c_int32 * max(nf, 1))()
        stats = IoStats()
        err = C.create_string_buffer(512)
        rc = self.L.pnah_extract_to_dir(self.h, os.fsencode(out_dir), device, workers, group_bytes, window_bytes, io_threads, int(verify),
                                        C.byref(stats), st, err, 512)
        if rc:
            raise HostError(rc, err.value.decode())
        return stats.as_dict(), list(st)[:nf]

    def close(self):
        if getattr(self, "h", None):
            self.L.pnah_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_chunks(self):
        return int(self.L.pnah_chunk_count(self.h))

    def entries(self):
        out = []
        for i in range(self.L.pnah_entry_count(self.h)):
            e = EntryInfo()
            self.L.pnah_entry_get(self.h, i, C.byref(e))
            out.append({"kind": e.kind, "data_kind": e.data_kind, "compression": e.compression, "encryption": e.encryption,
                        "cipher_mode": e.cipher_mode, "n_bodies": e.n_bodies, "compressed_size": e.compressed_size,
                        "raw_file_size": None if e.raw_file_size == _ffi.UINT64_MAX else e.raw_file_size,
                        "name": e.name.decode() if e.name else "", "phsf": e.phsf.decode() if e.phsf else None})
        return out

    def set_password(self, password: bytes, key_cache: dict | None = None):
        """ReadOptions::with_password: derive one key per distinct PHSF (KDF = host work, lib/src/hash.rs:45-85)."""
        from .archive import derive_key
        cache = key_cache if key_cache is not None else {}
        for e in self.entries():
            if e["phsf"] and e["phsf"] not in cache:
                cache[e["phsf"]] = derive_key(e["phsf"], password)
        for phsf, key in cache.items():
            self.L.pnah_set_key(self.h, phsf.encode(), key)

    def set_key(self, phsf: str, key: bytes):
        self.L.pnah_set_key(self.h, phsf.encode(), key)

    def prepare(self, device=0):
        err = C.create_string_buffer(512)
        rc = self.L.pnah_prepare(self.h, device, err, 512)
        if rc:
            raise HostError(rc, err.value.decode())

    def files(self):
        out = []
        for i in range(self.L.pnah_file_count(self.h)):
            name, size = C.c_char_p(), C.c_uint64()
            st = self.L.pnah_file_get(self.h, i, C.byref(name), C.byref(size))
            out.append((name.value.decode(), int(size.value), st))
        return out

    def extract_files(self, out: np.ndarray | None = None, device=0, workers=3, group_bytes=256 << 20, verify=True, devices=None):
        """Returns (out buffer, offsets, statuses).  `out`: optional (pinned) uint8 array of sum(sizes) bytes.
        devices: list of GPUs to partition the entries over (workers threads each); default: the single `device`."""
        self.prepare(device if devices is None else devices[0])
        nf = int(self.L.pnah_file_count(self.h))
        sizes = np.zeros(max(nf, 1), dtype=np.uint64)
        self.L.pnah_file_sizes(self.h, sizes.ctypes.data_as(C.POINTER(C.c_uint64)), None)
        offs = np.zeros(nf + 1, dtype=np.uint64)
        # checked sum (sizes are derived from untrusted archive fields), vectorised: a million entries must not mean a million
        # interpreter steps.  The float sum bounds the exact one, so the uint64 prefix sums below cannot wrap.
        if nf and (int(sizes[:nf].max()) >= 1 << 59 or float(sizes[:nf].astype(np.float64).sum()) + 16.0 * nf >= float(1 << 60)):
            raise HostError(_ffi.E_OOM, "decoded sizes overflow the buffer layout")
        if nf:
            np.cumsum((sizes[:nf] + np.uint64(15)) // np.uint64(16) * np.uint64(16), out=offs[1:])
        total = int(offs[nf])
        if out is None:
            out = np.empty(total + 16, dtype=np.uint8)
        elif out.size < total:
            raise HostError(_ffi.E_NOSPACE, f"output buffer of {out.size} bytes, {total} needed")
        st = (C.c_int32 * max(nf, 1))()
        err = C.create_string_buffer(512)
        if devices is None:
            rc = self.L.pnah_extract_files(self.h, out.ctypes.data, offs.ctypes.data_as(C.POINTER(C.c_uint64)), st, device, workers,
                                           group_bytes, int(verify), err, 512)
        else:
            dv = (C.c_int * len(devices))(*devices)
            rc = self.L.pnah_extract_files_on(self.h, out.ctypes.data, offs.ctypes.data_as(C.POINTER(C.c_uint64)), st, dv, len(devices),
                                              workers, group_bytes, int(verify), err, 512)
        if rc:
            raise HostError(rc, err.value.decode())
        return out, offs, np.ctypeslib.as_array(st)[:nf].tolist()

    def read_all(self, **kw):
        """[(name, status, bytes | None)] of every FILE entry.  The length of each result is the DECODED length the stream
        produced (files() after the extraction), not the archive's fSIZ hint; an entry that decodes to more than its hint said
        (PNA_E_NOSPACE on the first pass) is taken again on its own with the length the first pass reported."""
        out, offs, st = self.extract_files(**kw)
        files = self.files()
        res = []
        for i, (name, size, _) in enumerate(files):
            if st[i] == _ffi.E_NOSPACE:
                one = np.empty((size + 15) // 16 * 16 + 16, dtype=np.uint8)
                off2 = np.zeros(len(files) + 1, dtype=np.uint64)
                off2[i + 1] = (size + 15) // 16 * 16
                st2 = (C.c_int32 * len(files))()
                err = C.create_string_buffer(512)
                rc = self.L.pnah_extract_range(self.h, one.ctypes.data, off2.ctypes.data_as(C.POINTER(C.c_uint64)), st2, kw.get("device", 0), 1,
                                               kw.get("group_bytes", 256 << 20), 0, i, i + 1, err, 512)
                if rc:
                    raise HostError(rc, err.value.decode())
                size2 = self.files()[i][1]
                res.append((name, st2[i], one[:size2].tobytes() if st2[i] == 0 else None))
                continue
            res.append((name, st[i], out[int(offs[i]):int(offs[i]) + size].tobytes() if st[i] == 0 else None))
        return res


def split_archive(archive, max_part_bytes: int, device=0):
    """Split writer (lib/src/archive/split_parts.rs): a finished archive cut into parts of at most max_part_bytes.  Returns the
    parts as uint8 arrays."""
    L = lib()
    buf = archive if isinstance(archive, np.ndarray) else np.frombuffer(archive, dtype=np.uint8)
    err = C.create_string_buffer(512)
    total, n_parts = C.c_uint64(0), C.c_uint32(0)
    guess = buf.size // max(1, max_part_bytes - 52) + 2            # one call when the guess holds, a sizing pass otherwise
    out, lens = np.empty(buf.size + 128 * guess, dtype=np.uint8), (C.c_uint64 * (2 * guess))()
    for _ in range(2):
        rc = L.pnah_split(buf.ctypes.data, buf.size, max_part_bytes, device, out.ctypes.data, out.size, C.byref(total), lens, len(lens),
                          C.byref(n_parts), err, 512)
        if rc != _ffi.E_NOSPACE:
            break
        out, lens = np.empty(total.value, dtype=np.uint8), (C.c_uint64 * n_parts.value)()
    if rc:
        raise HostError(rc, err.value.decode())
    parts, at = [], 0
    for k in range(n_parts.value):
        parts.append(out[at:at + lens[k]])
        at += lens[k]
    return parts


def split_layout(archive, max_part_bytes: int, max_parts: int = 1 << 16):
    """Part lengths the split writer would produce (sizing call of pnah_split: budget arithmetic only, no copy, no GPU work)."""
    L = lib()
    buf = archive if isinstance(archive, np.ndarray) else np.frombuffer(archive, dtype=np.uint8)
    err = C.create_string_buffer(512)
    total, n_parts = C.c_uint64(0), C.c_uint32(0)
    lens = (C.c_uint64 * max_parts)()
    rc = L.pnah_split(buf.ctypes.data, buf.size, max_part_bytes, 0, None, 0, C.byref(total), lens, max_parts, C.byref(n_parts), err, 512)
    if rc != _ffi.E_NOSPACE:
        raise HostError(rc, err.value.decode())
    return [int(lens[k]) for k in range(min(n_parts.value, max_parts))], int(total.value)


def create_archive(files, compression=0, level=-1, encryption=0, cipher_mode=1, key=None, phsf=None, ivs=None, max_chunk_size=0,
                   device=0, workers=3, group_bytes=256 << 20, out=None, devices=None):
    """files: list of (name, bytes-like), or the packed form (names, uint8 buffer, offsets[n + 1]).  Returns the archive bytes (numpy view of `out` when given).
    ivs: 16 bytes per file, or None: the writer draws a fresh IV per entry from the OS (entry/write.rs:108-111).
    devices: list of GPUs to partition the files over (`workers` threads each); default: the single `device`."""
    L = lib()
    if isinstance(files, tuple) and len(files) == 3:
        # packed form (names, buffer, offsets[n + 1]): file i is buffer[offsets[i]:offsets[i + 1]] -- no per-file Python objects
        fnames, fbuf, foffs = files
        n = len(fnames)
        foffs = np.ascontiguousarray(foffs, dtype=np.uint64)
        assert foffs.size == n + 1 and int(foffs[-1]) <= fbuf.size
        names = (C.c_char_p * max(n, 1))(*[x.encode() for x in fnames])
        lens_np = np.ascontiguousarray(np.diff(foffs), dtype=np.uint64)
        ptrs_np = np.ascontiguousarray(foffs[:-1] + np.uint64(fbuf.ctypes.data), dtype=np.uint64)
        ptrs = (C.c_void_p * max(n, 1)).from_buffer(ptrs_np) if n else (C.c_void_p * 1)()
        lens = (C.c_uint64 * max(n, 1)).from_buffer(lens_np) if n else (C.c_uint64 * 1)()
    else:
        n = len(files)
        arrs = [f[1] if isinstance(f[1], np.ndarray) else np.frombuffer(f[1], dtype=np.uint8) for f in files]
        names = (C.c_char_p * max(n, 1))(*[f[0].encode() for f in files])
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
        lens = (C.c_uint64 * max(n, 1))(*[a.size for a in arrs])
    bound = int(L.pnah_create_bound(n, names, lens, compression, encryption, (phsf or "").encode(), max_chunk_size))
    if out is None:
        out = np.empty(bound, dtype=np.uint8)
    olen = C.c_uint64(0)
    err = C.create_string_buffer(512)
    if devices is None:
        rc = L.pnah_create(n, names, ptrs, lens, ivs, compression, level, encryption, cipher_mode, key or bytes(32), (phsf or "").encode(),
                           max_chunk_size, device, workers, group_bytes, out.ctypes.data, out.size, C.byref(olen), err, 512)
    else:
        dv = (C.c_int * len(devices))(*devices)
        rc = L.pnah_create_on(n, names, ptrs, lens, ivs, compression, level, encryption, cipher_mode, key or bytes(32), (phsf or "").encode(),
                              max_chunk_size, dv, len(devices), workers, group_bytes, out.ctypes.data, out.size, C.byref(olen), err, 512)
    if rc:
        raise HostError(rc, err.value.decode())
    return out[:olen.value]


def create_solid_archive(files, compression=0, level=-1, encryption=0, cipher_mode=1, key=None, phsf=None, max_chunk_size=32 * 1024,
                         device=0, out=None):
    """Solid mode (archive/write.rs:438-471): every file a STORE entry inside one compressed (+ encrypted) stream of SDAT bodies.
    files: list of (name, bytes-like).  Returns the archive bytes."""
    L = lib()
    if isinstance(files, tuple) and len(files) == 3:
        # packed form (names, buffer, offsets[n + 1]): file i is buffer[offsets[i]:offsets[i + 1]] -- no per-file Python objects
        fnames, fbuf, foffs = files
        n = len(fnames)
        foffs = np.ascontiguousarray(foffs, dtype=np.uint64)
        assert foffs.size == n + 1 and int(foffs[-1]) <= fbuf.size
        names = (C.c_char_p * max(n, 1))(*[x.encode() for x in fnames])
        lens_np = np.ascontiguousarray(np.diff(foffs), dtype=np.uint64)
        ptrs_np = np.ascontiguousarray(foffs[:-1] + np.uint64(fbuf.ctypes.data), dtype=np.uint64)
        ptrs = (C.c_void_p * max(n, 1)).from_buffer(ptrs_np) if n else (C.c_void_p * 1)()
        lens = (C.c_uint64 * max(n, 1)).from_buffer(lens_np) if n else (C.c_uint64 * 1)()
    else:
        n = len(files)
        arrs = [f[1] if isinstance(f[1], np.ndarray) else np.frombuffer(f[1], dtype=np.uint8) for f in files]
        names = (C.c_char_p * max(n, 1))(*[f[0].encode() for f in files])
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
        lens = (C.c_uint64 * max(n, 1))(*[a.size for a in arrs])
    bound = int(L.pnah_create_solid_bound(n, names, lens, compression, encryption, cipher_mode, (phsf or "").encode(), max_chunk_size))
    if out is None:
        out = np.empty(bound, dtype=np.uint8)
    olen = C.c_uint64(0)
    err = C.create_string_buffer(512)
    rc = L.pnah_create_solid(n, names, ptrs, lens, compression, level, encryption, cipher_mode, key or bytes(32), (phsf or "").encode(),
                             max_chunk_size, device, out.ctypes.data, out.size, C.byref(olen), err, 512)
    if rc:
        raise HostError(rc, err.value.decode())
    return out[:olen.value]


def create_from_files(names_and_paths, archive_path, compression=0, level=-1, encryption=0, cipher_mode=1, key=None, phsf=None,
                      max_chunk_size=0, device=0, workers=4, group_bytes=256 << 20, io_threads=8):
    """`pna create` data path: files read into pinned memory by io_threads readers, one GPU encode pass, archive written to
    `archive_path`.  names_and_paths: list of (entry name, file path).  Returns the stats dict."""
    L = lib()
    n = len(names_and_paths)
    names = (C.c_char_p * max(n, 1))(*[a.encode() for a, _ in names_and_paths])
    paths = (C.c_char_p * max(n, 1))(*[os.fsencode(b) for _, b in names_and_paths])
    stats = IoStats()
    err = C.create_string_buffer(512)
    rc = L.pnah_create_from_files(n, names, paths, compression, level, encryption, cipher_mode, key or bytes(32), (phsf or "").encode(),
                                  max_chunk_size, os.fsencode(archive_path), device, workers, group_bytes, io_threads, C.byref(stats),
                                  err, 512)
    if rc:
        raise HostError(rc, err.value.decode())
    return stats.as_dict()
