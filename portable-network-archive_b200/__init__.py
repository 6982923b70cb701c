"""pna-b200: B200-native data-chunk pipeline behind the `libpna` API seam.

Host-side mirror of the reference's public interface for the hot path (names and argument meaning
follow /root/reference/lib/src: `Archive`, `NormalEntry.reader`, `SolidEntry.entries`, `ReadOptions`,
`WriteOptions`, `FileEntryBuilder`), with the three internal seams (chunk CRC, decode, encode) routed to
`libpna_cuda.so` through the C ABI in include/pna_cuda.h.  All arithmetic on entry data happens in
CUDA kernels; this module only parses chunk framing, derives keys (KDF is host work in the reference
too, lib/src/hash.rs) and batches calls.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import (E_BAD_ARG, E_CUDA, E_INTERNAL, E_INVALID_DATA, E_INVALID_INPUT, E_NOSPACE, E_OOM,
                   E_UNEXPECTED_EOF, E_UNSUPPORTED, OK, UINT64_MAX, PnaCudaError)
from .archive import (Archive, ChunkType, CipherMode, Compression, DataKind, Encryption, EntryBuilder,
                      FileEntryBuilder, NormalEntry, PnaError, ReadOptions, SolidEntry, SolidEntryBuilder, WriteOptions,
                      derive_key, index_archive)

__all__ = ["Context", "Archive", "NormalEntry", "SolidEntry", "ReadOptions", "WriteOptions", "FileEntryBuilder",
           "EntryBuilder", "SolidEntryBuilder", "Compression", "Encryption", "CipherMode", "DataKind", "ChunkType",
           "PnaError", "PnaCudaError", "derive_key", "index_archive", "default_context"]


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
        return a if a.flags.c_contiguous else np.ascontiguousarray(a)
    return np.frombuffer(buf, dtype=np.uint8)


class DecodePlan:
    """A decode batch resident in HBM (pna_plan).  run() launches the kernels on Context.stream."""

    def __init__(self, ctx: "Context", handle, n: int, keep):
        self.ctx, self.h, self.n, self._keep = ctx, handle, n, keep

    def run(self):
        self.ctx._ck(self.ctx.L.pna_cuda_decode_plan_run(self.h), "decode_plan_run")

    def fetch(self, caps):
        """Returns (list of numpy outputs, statuses, lens)."""
        outs = [np.empty(max(int(c), 1), dtype=np.uint8) for c in caps]
        bufs = (_ffi.Buf * self.n)()
        for i, o in enumerate(outs):
            bufs[i].ptr = o.ctypes.data
            bufs[i].cap = int(caps[i])
        st = (C.c_int32 * self.n)()
        self.ctx._ck(self.ctx.L.pna_cuda_decode_plan_fetch(self.h, bufs, st), "decode_plan_fetch")
        return [o[:bufs[i].len] if st[i] == OK else o[:0] for i, o in enumerate(outs)], list(st), [b.len for b in bufs]

    def lengths(self):
        """(decoded length, status) of every entry after run(): E_NOSPACE + the required length when its size hint was too small."""
        lens = (C.c_uint64 * max(self.n, 1))()
        st = (C.c_int32 * max(self.n, 1))()
        self.ctx._ck(self.ctx.L.pna_cuda_decode_plan_lengths(self.h, lens, st), "decode_plan_lengths")
        return [int(x) for x in lens][:self.n], list(st)[:self.n]

    def fetch_into(self, bufs, st):
        self.ctx._ck(self.ctx.L.pna_cuda_decode_plan_fetch(self.h, bufs, st), "decode_plan_fetch")

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.ctx.L.pna_cuda_plan_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"stream_bytes": a.value, "plain_bytes": b.value, "launches_per_run": c.value}

    def counts(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.ctx.L.pna_cuda_plan_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"blocks": a.value, "sequences": b.value, "literal_bytes": c.value}

    def crc_results(self):
        """(computed CRC per registered chunk span, number of mismatches) of the last run."""
        n = getattr(self, "n_crc", 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        broken = C.c_uint32(0)
        self.ctx._ck(self.ctx.L.pna_cuda_plan_crc_results(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(broken)),
                     "plan_crc_results")
        return out[:n], int(broken.value)

    def stage_ms(self) -> dict:
        ms = (C.c_float * 16)()
        n = self.ctx.L.pna_cuda_plan_stage_ms(self.h, ms, 16)
        return {self.ctx.L.pna_cuda_stage_name(i).decode(): float(ms[i]) for i in range(max(n, 0))}

    def close(self):
        if self.h:
            self.ctx.L.pna_cuda_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EncodePlan:
    """A create batch resident in HBM (pna_plan, encode kind)."""

    def __init__(self, ctx, handle, n, bounds, ncrc):
        self.ctx, self.h, self.n, self.bounds, self.ncrc = ctx, handle, n, bounds, ncrc

    def run(self):
        self.ctx._ck(self.ctx.L.pna_cuda_encode_plan_run(self.h), "encode_plan_run")

    def fetch(self, into=None):
        """Returns (streams, crc lists, statuses).  `into`: optional pinned uint8 array that receives the streams."""
        n = self.n
        bufs = (_ffi.Buf * max(n, 1))()
        if into is None:
            outs = [np.empty(b, dtype=np.uint8) for b in self.bounds]
        else:
            outs, pos = [], 0
            for b in self.bounds:
                outs.append(into[pos:pos + b])
                pos += b
        for i, o in enumerate(outs):
            bufs[i].ptr = o.ctypes.data
            bufs[i].cap = self.bounds[i]
        crcs = np.zeros(sum(self.ncrc) + 1, dtype=np.uint32)
        cnt = np.zeros(max(n, 1), dtype=np.uint32)
        st = (C.c_int32 * max(n, 1))()
        self.ctx._ck(self.ctx.L.pna_cuda_encode_plan_fetch(self.h, bufs, crcs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                         cnt.ctypes.data_as(C.POINTER(C.c_uint32)), st), "encode_plan_fetch")
        res, crc_lists, pos = [], [], 0
        for i in range(n):
            res.append(outs[i][:bufs[i].len])
            crc_lists.append(crcs[pos:pos + int(cnt[i])].copy())
            pos += int(cnt[i])
        return res, crc_lists, list(st)[:n]

    def stage_ms(self) -> dict:
        ms = (C.c_float * 16)()
        n = self.ctx.L.pna_cuda_plan_stage_ms(self.h, ms, 16)
        return {self.ctx.L.pna_cuda_encode_stage_name(i).decode(): float(ms[i]) for i in range(min(max(n, 0), 5))}

    def close(self):
        if self.h:
            self.ctx.L.pna_cuda_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """A pna_ctx over one GPU (`device`) or several (`devices=[...]`): with several, every batch / plan call is sharded by
    entry across them inside the library (no collective; entries are independent).  One process per GPU (one Context per
    rank) works just as well -- bench.py does that, because the driver launches it with torchrun."""

    def __init__(self, device: int = 0, devices=None):
        self.L = _ffi.lib()
        h = C.c_void_p()
        ids = [int(device)] if devices is None else [int(d) for d in devices]
        arr = (C.c_int * len(ids))(*ids)
        rc = self.L.pna_cuda_init(C.byref(h), arr, len(ids))
        if rc != OK:
            raise PnaCudaError(rc, "pna_cuda_init failed: no usable sm_100 device (there is no CPU fallback)")
        self.h = h
        self.device = ids[0]
        self.devices = ids

    def close(self):
        if getattr(self, "h", None):
            self.L.pna_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != OK:
            raise PnaCudaError(rc, f"{what}: {self.L.pna_cuda_strerror(rc).decode()} / {self.L.pna_cuda_last_error(self.h).decode()}")

    @property
    def stream(self) -> int:
        return int(self.L.pna_cuda_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.pna_cuda_launch_count(self.h))

    def pinned(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a numpy array (lives until pinned_free or the end of the process)."""
        p = self.L.pna_cuda_host_alloc(self.h, nbytes)
        if not p:
            raise PnaCudaError(E_OOM, "pna_cuda_host_alloc")
        if not hasattr(self, "_pinned"):
            self._pinned = set()
        self._pinned.add(int(p))
        return np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p))[:nbytes]

    def pinned_free(self, arr: np.ndarray):
        """Give a buffer from pinned() back (the caller must drop every view of it)."""
        p = int(arr.ctypes.data)
        if p in getattr(self, "_pinned", ()):
            self._pinned.discard(p)
            self.L.pna_cuda_host_free(self.h, p)

    def transfer_probe(self, h2d_src: np.ndarray, d2h_dst: np.ndarray):
        """pinned -> HBM copy of h2d_src and HBM -> pinned copy into d2h_dst: milliseconds alone and both at once (CUDA events)."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self._ck(self.L.pna_cuda_transfer_probe(self.h, h2d_src.ctypes.data, h2d_src.size, d2h_dst.ctypes.data, d2h_dst.size,
                                                C.byref(a), C.byref(b), C.byref(c)), "transfer_probe")
        return {"h2d_ms": a.value, "d2h_ms": b.value, "both_ms": c.value}

    # ---- seam 1
    def crc32(self, spans) -> np.ndarray:
        """CRC-32 of each span (bytes-like); chunk CRC = crc32(type || data) (format/chunk.rs:7-12)."""
        arrs = [_as_u8(s) for s in spans]
        n = len(arrs)
        sp = (_ffi.Span * max(n, 1))()
        for i, a in enumerate(arrs):
            sp[i].ptr = a.ctypes.data if a.size else None
            sp[i].len = a.size
        out = np.zeros(n, dtype=np.uint32)
        self._ck(self.L.pna_cuda_crc32(self.h, sp, n, out.ctypes.data_as(C.POINTER(C.c_uint32))), "crc32")
        return out

    def crc32_image(self, image, offs, lens) -> np.ndarray:
        img = _as_u8(image)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        out = np.zeros(len(offs), dtype=np.uint32)
        self._ck(self.L.pna_cuda_crc32_image(self.h, img.ctypes.data, img.size, offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             lens.ctypes.data_as(C.POINTER(C.c_uint64)), len(offs),
                                             out.ctypes.data_as(C.POINTER(C.c_uint32))), "crc32_image")
        return out

    # ---- seam 2
    def _descs(self, entries):
        """entries: iterable of dicts {bodies:[array..], compression, encryption, cipher_mode, key, raw_size_hint}."""
        n = len(entries)
        descs = (_ffi.DecodeDesc * max(n, 1))()
        keep = []
        for i, e in enumerate(entries):
            bodies = [_as_u8(b) for b in e["bodies"]]
            sp = (_ffi.Span * max(len(bodies), 1))()
            for j, b in enumerate(bodies):
                sp[j].ptr = b.ctypes.data if b.size else None
                sp[j].len = b.size
            keep.append((bodies, sp))
            d = descs[i]
            d.bodies = sp
            d.n_bodies = len(bodies)
            d.compression = e.get("compression", 0)
            d.encryption = e.get("encryption", 0)
            d.cipher_mode = e.get("cipher_mode", 0)
            key = e.get("key") or bytes(32)
            if len(key) != 32:
                raise PnaCudaError(E_INVALID_INPUT, "key must be 32 bytes")
            C.memmove(d.key, key, 32)
            h = e.get("raw_size_hint")
            d.raw_size_hint = UINT64_MAX if h is None else int(h)
        return descs, keep

    def decode_plan(self, entries, crc=None) -> DecodePlan:
        """Upload a batch once (pna_plan).  crc = (ptrs u64[], lens u64[], expect u32[], entry_of i32[]) fuses the
        chunk-CRC check (seam 1) into the plan: spans are host addresses of type||data inside the archive buffer."""
        descs, keep = self._descs(entries)
        h = C.c_void_p()
        if crc is None:
            self._ck(self.L.pna_cuda_decode_plan_create(self.h, descs, len(entries), C.byref(h)), "decode_plan_create")
            return DecodePlan(self, h, len(entries), None)
        ptrs, lens, expect, entry_of = crc
        n = len(ptrs)
        sp = np.zeros(n, dtype=[("ptr", "<u8"), ("len", "<u8")])
        sp["ptr"], sp["len"] = ptrs, lens
        expect = np.ascontiguousarray(expect, dtype=np.uint32)
        entry_of = np.ascontiguousarray(entry_of, dtype=np.int32)
        self._ck(self.L.pna_cuda_decode_plan_create_crc(
            self.h, descs, len(entries), sp.ctypes.data_as(C.POINTER(_ffi.Span)), expect.ctypes.data_as(C.POINTER(C.c_uint32)),
            entry_of.ctypes.data_as(C.POINTER(C.c_int32)), n, C.byref(h)), "decode_plan_create_crc")
        p = DecodePlan(self, h, len(entries), None)
        p.n_crc = n
        return p

    def decode_batch(self, entries, caps=None):
        """One-shot decode through pna_cuda_decode_batch.  Returns (outputs, statuses, lens).  When caps is None
        a sizing call (cap 0 -> E_NOSPACE with the required length) precedes the real one."""
        n = len(entries)
        if n == 0:
            return [], [], []
        descs, keep = self._descs(entries)
        st = (C.c_int32 * n)()
        bufs = (_ffi.Buf * n)()
        if caps is None:
            self._ck(self.L.pna_cuda_decode_batch(self.h, descs, n, bufs, st), "decode_batch(size)")
            caps = [bufs[i].len if st[i] in (OK, E_NOSPACE) else 0 for i in range(n)]
        outs = [np.empty(max(int(c), 1), dtype=np.uint8) for c in caps]
        for i, o in enumerate(outs):
            bufs[i].ptr = o.ctypes.data
            bufs[i].cap = int(caps[i])
            bufs[i].len = 0
        self._ck(self.L.pna_cuda_decode_batch(self.h, descs, n, bufs, st), "decode_batch")
        del keep
        return [o[:bufs[i].len] if st[i] == OK else o[:0] for i, o in enumerate(outs)], list(st), [b.len for b in bufs]

    # ---- seam 3
    def _enc_descs(self, entries):
        n = len(entries)
        descs = (_ffi.EncodeDesc * max(n, 1))()
        keep = []
        for i, e in enumerate(entries):
            p = _as_u8(e["plain"])
            keep.append(p)
            d = descs[i]
            d.plain.ptr = p.ctypes.data if p.size else None
            d.plain.len = p.size
            d.compression = e.get("compression", 0)
            d.encryption = e.get("encryption", 0)
            d.cipher_mode = e.get("cipher_mode", 0)
            d.level = e.get("level", -1)
            C.memmove(d.key, e.get("key") or bytes(32), 32)
            C.memmove(d.iv, e.get("iv") or bytes(16), 16)
            d.max_chunk_size = e.get("max_chunk_size", 0)
            sh = e.get("stream_header")     # GCM: the 75-byte stream header (archive.gcm_stream_header)
            if sh is not None:
                hb = C.create_string_buffer(bytes(sh), len(sh))
                keep.append(hb)
                d.stream_header = C.cast(hb, C.c_void_p)
        return descs, keep

    def encode_plan(self, entries) -> "EncodePlan":
        """Upload a create batch once (pna_plan); run() launches match -> block writers -> layout -> cipher -> CRC."""
        descs, keep = self._enc_descs(entries)
        h = C.c_void_p()
        self._ck(self.L.pna_cuda_encode_plan_create(self.h, descs, len(entries), C.byref(h)), "encode_plan_create")
        bounds = [int(self.L.pna_cuda_encode_bound(C.byref(descs[i]))) for i in range(len(entries))]
        ncrc = [int(self.L.pna_cuda_encode_crc_count(C.byref(descs[i]))) for i in range(len(entries))]
        return EncodePlan(self, h, len(entries), bounds, ncrc)

    def encode_batch(self, entries):
        """entries: dicts {plain, compression, level, encryption, cipher_mode, key, iv, max_chunk_size}.
        Returns (streams, fdat_crcs per entry, statuses)."""
        n = len(entries)
        if n == 0:
            return [], [], []
        descs, keep = self._enc_descs(entries)
        bounds = [int(self.L.pna_cuda_encode_bound(C.byref(descs[i]))) for i in range(n)]
        ncrc = [int(self.L.pna_cuda_encode_crc_count(C.byref(descs[i]))) for i in range(n)]
        outs = [np.empty(b, dtype=np.uint8) for b in bounds]
        bufs = (_ffi.Buf * n)()
        for i, o in enumerate(outs):
            bufs[i].ptr = o.ctypes.data
            bufs[i].cap = bounds[i]
        crcs = np.zeros(sum(ncrc) + 1, dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        st = (C.c_int32 * n)()
        self._ck(self.L.pna_cuda_encode_batch(self.h, descs, n, bufs, crcs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              cnt.ctypes.data_as(C.POINTER(C.c_uint32)), st), "encode_batch")
        res, crc_lists, pos = [], [], 0
        for i in range(n):
            res.append(outs[i][:bufs[i].len])
            crc_lists.append(crcs[pos:pos + int(cnt[i])].copy())
            pos += int(cnt[i])
        return res, crc_lists, list(st)

    # ---- test hook
    def ecb(self, encryption: int, encrypt: bool, key: bytes, data: bytes) -> bytes:
        src = _as_u8(data)
        out = np.zeros(src.size, dtype=np.uint8)
        self._ck(self.L.pna_cuda_ecb(self.h, encryption, int(encrypt), key, src.ctypes.data, src.size, out.ctypes.data), "ecb")
        return out.tobytes()[: src.size & ~15]


_default = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0)
    return _default
