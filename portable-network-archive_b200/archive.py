"""Host-side mirror of the `libpna` container API for the data-chunk hot path.

Only chunk framing, entry grouping and option plumbing live here (citations are relative to
/root/reference).  Every byte of entry data is checked (CRC), decrypted, decompressed, compressed and
encrypted by libpna_cuda.so.

* signature, chunk framing   lib/src/format/signature.rs:6, lib/src/io.rs:117-197, lib/src/bytes.rs:39-111
* archive header / gather    lib/src/archive/header.rs:27-56, lib/src/archive/read.rs:22-73
* NormalEntry / SolidEntry   lib/src/entry.rs:457-483, :567-583, :757-912, :1150
* headers                    lib/src/entry/header.rs:123-162, :274-296
* options                    lib/src/entry/options.rs:237-247, :483-491, :596-604, :840-851, :1035, :1344
* builders                   lib/src/entry/builder.rs:45-69, :171-189; builder/file.rs:65-140; builder/solid.rs:85-320
* writer                     lib/src/archive/write.rs:92, :368, :545; lib/src/io.rs:183-197
"""
from __future__ import annotations

import base64
import ctypes as C
import hashlib
import os
import struct

import numpy as np

from . import _ffi

SIGNATURE = b"\x89PNA\r\n\x1a\n"
MIN_CHUNK_BYTES_SIZE = 12


class PnaError(Exception):
    """io::Error mirror: .kind is one of the _ffi status codes (== reference io::ErrorKind class)."""

    def __init__(self, kind: int, msg: str):
        super().__init__(msg)
        self.kind = kind


class ChunkType:
    AHED, AEND, ANXT = b"AHED", b"AEND", b"ANXT"
    FHED, PHSF, FDAT, FEND = b"FHED", b"PHSF", b"FDAT", b"FEND"
    SHED, SDAT, SEND = b"SHED", b"SDAT", b"SEND"
    fSIZ = b"fSIZ"


class Compression:
    NO, DEFLATE, ZSTANDARD, XZ = 0, 1, 2, 4


class Encryption:
    NO, AES, CAMELLIA = 0, 1, 2


class CipherMode:
    CBC, CTR, GCM = 0, 1, 2


class DataKind:
    FILE, DIRECTORY, SYMLINK, HARDLINK = 0, 1, 2, 3


# ------------------------------------------------------------------------------------ KDF (host work)
def _b64(s: str) -> bytes:
    return base64.b64decode(s + "=" * (-len(s) % 4))


def _b64e(b: bytes) -> str:
    return base64.b64encode(b).decode().rstrip("=")


def derive_key(phsf: str, password: bytes) -> bytes:
    """lib/src/hash.rs:45-85: key = PHC hash of the password under the parameters+salt recorded in PHSF."""
    parts = phsf.split("$")
    if len(parts) < 4:
        raise PnaError(_ffi.E_INVALID_DATA, "malformed PHSF")
    alg = parts[1]
    if alg.startswith("argon2"):
        from cryptography.hazmat.primitives.kdf.argon2 import Argon2id
        if alg != "argon2id":
            raise PnaError(_ffi.E_UNSUPPORTED, f"unsupported algorithm {alg}")
        idx = 3 if parts[2].startswith("v=") else 2
        params = dict(kv.split("=") for kv in parts[idx].split(","))
        return Argon2id(salt=_b64(parts[idx + 1]), length=32, iterations=int(params["t"]), lanes=int(params["p"]),
                        memory_cost=int(params["m"])).derive(password)
    if alg in ("pbkdf2-sha256", "pbkdf2-sha512"):
        params = dict(kv.split("=") for kv in parts[2].split(","))
        return hashlib.pbkdf2_hmac(alg.split("-")[1], password, _b64(parts[3]), int(params.get("i", 600000)),
                                   int(params.get("l", 32)))
    raise PnaError(_ffi.E_UNSUPPORTED, f"unsupported algorithm {alg}")


class ReadOptions:
    """options.rs:1344 -- password plus a per-PHSF key cache (KeyCache, options.rs:79)."""

    def __init__(self, password: bytes | str | None = None):
        self.password = password.encode() if isinstance(password, str) else password
        self._keys: dict[str, bytes] = {}

    @classmethod
    def with_password(cls, password):
        return cls(password)

    @classmethod
    def builder(cls):
        return cls()

    def key_for(self, phsf: str | None) -> bytes:
        if phsf is None:
            raise PnaError(_ffi.E_INVALID_DATA, "`PHSF` chunk not found")   # entry/read.rs:74
        if self.password is None:
            raise PnaError(_ffi.E_INVALID_INPUT, "password was not provided")  # entry/read.rs:50
        k = self._keys.get(phsf)
        if k is None:
            k = self._keys[phsf] = derive_key(phsf, self.password)
        return k


class WriteOptions:
    """options.rs:1035 -- compression, level, encryption, cipher mode, hash algorithm, password.
    The KDF runs once per options object (options.rs:1239-1274)."""

    def __init__(self, compression=Compression.NO, level=-1, encryption=Encryption.NO, cipher_mode=CipherMode.CTR,
                 password=None, hash_algorithm="pbkdf2-sha256", kdf_params=None, salt=None, segment_size=1 << 20):
        self.compression, self.level = compression, level
        self.segment_size = segment_size     # GCM datastream segment size (options.rs:1200); unused by CBC/CTR
        self.encryption, self.cipher_mode = encryption, cipher_mode
        self.password = password.encode() if isinstance(password, str) else password
        self.phsf, self.key = None, None
        if encryption != Encryption.NO:
            if self.password is None:
                raise PnaError(_ffi.E_INVALID_INPUT, "password is required for encryption")
            salt = salt or os.urandom(16)
            if hash_algorithm == "argon2id":
                p = {"m": 19456, "t": 2, "p": 1}
                p.update(kdf_params or {})
                self.phsf = f"$argon2id$v=19$m={p['m']},t={p['t']},p={p['p']}${_b64e(salt)}"
            else:
                p = {"i": 600000, "l": 32}
                p.update(kdf_params or {})
                self.phsf = f"$pbkdf2-sha256$i={p['i']},l={p['l']}${_b64e(salt)}"
            self.key = derive_key(self.phsf, self.password)

    @classmethod
    def store(cls):
        return cls()

    @classmethod
    def builder(cls, **kw):
        return cls(**kw)


# ------------------------------------------------------------------------------------ index pass
class RawChunk:
    __slots__ = ("ty", "off", "length", "crc", "buf")

    def __init__(self, ty, off, length, crc, buf=None):
        self.ty, self.off, self.length, self.crc = ty, off, length, crc   # off = offset of the data field
        self.buf = buf      # the part this chunk lives in when an archive is read from several parts (else the entry's buffer)


def index_archive(buf: np.ndarray, pos: int = 0, end: int | None = None):
    """Walk chunk headers WITHOUT touching chunk data (bytes::skip_chunk semantics, bytes.rs:90): yields RawChunk
    records.  CRC validation of type||data is a separate batched GPU call (Archive.verify_chunks)."""
    mv = memoryview(buf)
    n = len(mv) if end is None else end
    out = []
    while pos < n:
        if n - pos < MIN_CHUNK_BYTES_SIZE:
            raise PnaError(_ffi.E_UNEXPECTED_EOF, "truncated chunk")
        length = int.from_bytes(mv[pos:pos + 4], "big")
        ty = bytes(mv[pos + 4:pos + 8])
        if not ty.isalpha():                                   # chunk/types.rs:204
            raise PnaError(_ffi.E_INVALID_DATA, f"invalid chunk type {ty!r}")
        if n - pos - 12 < length:
            raise PnaError(_ffi.E_UNEXPECTED_EOF, "truncated chunk body")
        crc = int.from_bytes(mv[pos + 8 + length:pos + 12 + length], "big")
        out.append(RawChunk(ty, pos + 8, length, crc))
        pos += 12 + length
    return out


GCM_STREAM_HEADER_LEN = 75      # aead.rs:15
GCM_DEFAULT_SEGMENT_SIZE = 1 << 20  # aead.rs:18


def gcm_stream_key(k_master: bytes, stream_header: bytes, header_type: bytes, header_data: bytes, phsf: bytes) -> bytes:
    """Per-stream key of a GCM entry (aead.rs:166-208) through the library's host-side key schedule."""
    out = C.create_string_buffer(32)
    rc = _ffi.lib().pna_cuda_gcm_stream_key(k_master, stream_header, len(stream_header), header_type, header_data,
                                            len(header_data), phsf, len(phsf), out)
    if rc != _ffi.OK:
        raise PnaError(rc, "malformed AEAD datastream header or key mismatch")
    return out.raw


def gcm_stream_header(k_master: bytes, salt: bytes, nonce_prefix: bytes, segment_size: int = GCM_DEFAULT_SEGMENT_SIZE) -> bytes:
    """salt || nonce_prefix || segment_size || key confirmation (aead.rs:124-132, entry/write.rs:81-99)."""
    out = C.create_string_buffer(GCM_STREAM_HEADER_LEN)
    rc = _ffi.lib().pna_cuda_gcm_stream_header(k_master, salt, nonce_prefix, segment_size, out)
    if rc != _ffi.OK:
        raise PnaError(rc, "bad GCM stream parameters")
    return out.raw


class _EntryBase:
    def __init__(self, archive_buf, chunks):
        self._buf = archive_buf
        self.chunks = chunks
        self.phsf = None
        self.bodies = []          # numpy views of the FDAT/SDAT bodies, in order

    def _body(self, ch: RawChunk) -> np.ndarray:
        buf = ch.buf if ch.buf is not None else self._buf
        return buf[ch.off:ch.off + ch.length]

    def _stream_prefix(self, n: int) -> bytes:
        out = bytearray()
        for b in self.bodies:
            if len(out) >= n:
                break
            out += bytes(b[:n - len(out)])
        return bytes(out)

    def _desc(self, options: ReadOptions | None):
        key = None
        if self.encryption not in (Encryption.NO,) and self.encryption in (Encryption.AES, Encryption.CAMELLIA) \
                and self.cipher_mode in (CipherMode.CBC, CipherMode.CTR, CipherMode.GCM):
            key = (options or ReadOptions()).key_for(self.phsf)
            if self.cipher_mode == CipherMode.GCM:
                # decrypt_reader's GCM branch (entry/read.rs:105-139): header checks, key confirmation, stream key
                key = gcm_stream_key(key, self._stream_prefix(GCM_STREAM_HEADER_LEN), self.chunks[0].ty, self.header_bytes,
                                     self.phsf.encode("utf-8"))
        return {"bodies": self.bodies, "compression": self.compression, "encryption": self.encryption,
                "cipher_mode": self.cipher_mode, "key": key, "raw_size_hint": getattr(self, "raw_file_size", None)}


_KIND_ERR = {
    _ffi.E_INVALID_DATA: "invalid data", _ffi.E_UNEXPECTED_EOF: "unexpected end of file",
    _ffi.E_INVALID_INPUT: "corrupt deflate stream", _ffi.E_UNSUPPORTED: "unsupported method",
    _ffi.E_NOSPACE: "output buffer too small", _ffi.E_OOM: "out of memory", _ffi.E_INTERNAL: "internal error",
}


class NormalEntry(_EntryBase):
    """entry.rs:741 -- FHED .. FEND."""

    def __init__(self, archive_buf, chunks, ctx=None):
        super().__init__(archive_buf, chunks)
        self._ctx = ctx
        if not chunks or chunks[0].ty != ChunkType.FHED:
            raise PnaError(_ffi.E_INVALID_DATA, "expected `FHED` chunk")
        h = bytes(self._body(chunks[0]))
        if len(h) < 6:
            raise PnaError(_ffi.E_INVALID_DATA, "entry header too short")
        self.header_bytes = h
        self.major, self.minor, self.data_kind, self.compression, self.encryption, self.cipher_mode = h[:6]
        if self.major != 0 or self.minor != 0:
            raise PnaError(_ffi.E_UNSUPPORTED, f"entry version {self.major}.{self.minor} is not supported")
        self.name = h[6:].decode("utf-8")
        self.raw_file_size = None
        self.extra = []
        for ch in chunks[1:]:
            if ch.ty == ChunkType.FEND:
                break
            if ch.ty == ChunkType.FDAT:
                self.bodies.append(self._body(ch))
            elif ch.ty == ChunkType.PHSF:
                self.phsf = bytes(self._body(ch)).decode("utf-8")
            elif ch.ty == ChunkType.fSIZ:
                self.raw_file_size = int.from_bytes(bytes(self._body(ch)), "big")
            else:
                if ch.ty[0:1].isupper() and ch.ty not in (b"FHED",):   # unknown critical chunk, entry.rs:851
                    raise PnaError(_ffi.E_INVALID_DATA, f"unknown critical chunk type: {ch.ty!r}")
                self.extra.append(ch)

    @property
    def compressed_size(self) -> int:
        return sum(int(b.size) for b in self.bodies)

    def reader(self, options: ReadOptions | None = None, ctx=None) -> bytes:
        """NormalEntry::reader (entry.rs:1150): a batch of one on the GPU.  Prefer Archive.read_all for throughput."""
        from . import default_context
        ctx = ctx or self._ctx or default_context()
        outs, st, _ = ctx.decode_batch([self._desc(options)])
        if st[0] != _ffi.OK:
            raise PnaError(st[0], _KIND_ERR.get(st[0], "error"))
        return outs[0].tobytes()


class SolidEntry(_EntryBase):
    """entry.rs:458 -- SHED .. SEND; the decoded SDAT stream is a sequence of normal-entry chunks."""

    def __init__(self, archive_buf, chunks, ctx=None):
        super().__init__(archive_buf, chunks)
        self._ctx = ctx
        h = bytes(self._body(chunks[0]))
        if len(h) != 5:
            raise PnaError(_ffi.E_INVALID_DATA, "solid header must be 5 bytes")
        self.header_bytes = h
        self.major, self.minor, self.compression, self.encryption, self.cipher_mode = h
        if self.major != 0 or self.minor != 0:
            raise PnaError(_ffi.E_UNSUPPORTED, f"entry version {self.major}.{self.minor} is not supported")
        for ch in chunks[1:]:
            if ch.ty == ChunkType.SEND:
                break
            if ch.ty == ChunkType.SDAT:
                self.bodies.append(self._body(ch))
            elif ch.ty == ChunkType.PHSF:
                self.phsf = bytes(self._body(ch)).decode("utf-8")
            elif ch.ty[0:1].isupper():   # unknown critical chunk, entry.rs:716
                raise PnaError(_ffi.E_INVALID_DATA, f"unknown critical chunk type: {ch.ty!r}")

    def entries(self, options: ReadOptions | None = None, ctx=None):
        """SolidEntry::entries (entry.rs:567): decode the solid stream on the GPU, then re-parse the inner
        chunks -- with their CRCs checked on the GPU as read_chunk does (entry.rs:401-423)."""
        from . import default_context
        ctx = ctx or self._ctx or default_context()
        outs, st, _ = ctx.decode_batch([self._desc(options)])
        if st[0] != _ffi.OK:
            raise PnaError(st[0], _KIND_ERR.get(st[0], "error"))
        inner = outs[0]
        chunks = index_archive(inner, 0)
        _verify(ctx, inner, chunks)
        return [e for e in _group(inner, chunks, ctx) if isinstance(e, NormalEntry)]


def _verify(ctx, buf, chunks):
    """validate_chunk_crc (format/chunk.rs:16) for every chunk, one GPU batch: CRC over type||data."""
    if not chunks:
        return
    offs = np.fromiter((c.off - 4 for c in chunks), dtype=np.uint64, count=len(chunks))
    lens = np.fromiter((c.length + 4 for c in chunks), dtype=np.uint64, count=len(chunks))
    want = np.fromiter((c.crc for c in chunks), dtype=np.uint32, count=len(chunks))
    got = ctx.crc32_image(buf, offs, lens)
    bad = np.nonzero(got != want)[0]
    if bad.size:
        raise PnaError(_ffi.E_INVALID_DATA, f"broken chunk (#{int(bad[0])} `{chunks[int(bad[0])].ty.decode()}`)")


def _group(buf, chunks, ctx):
    """next_raw_item (archive/read.rs:46-73): gather chunks up to FEND / SEND."""
    cur, kind = None, None
    for ch in chunks:
        if ch.ty == ChunkType.AEND:   # archive/read.rs:58: end of this archive -- nothing behind it is an entry
            return
        if ch.ty == ChunkType.ANXT:   # archive/read.rs:57: only flags that another part follows
            continue
        if cur is None:
            if ch.ty == ChunkType.FHED:
                cur, kind = [ch], "F"
            elif ch.ty == ChunkType.SHED:
                cur, kind = [ch], "S"
            continue
        cur.append(ch)
        if kind == "F" and ch.ty == ChunkType.FEND:
            yield NormalEntry(buf, cur, ctx)
            cur = None
        elif kind == "S" and ch.ty == ChunkType.SEND:
            yield SolidEntry(buf, cur, ctx)
            cur = None
    if cur is not None:
        raise PnaError(_ffi.E_UNEXPECTED_EOF, "entry without end chunk")


class Archive:
    """archive.rs:79 -- reader over a byte buffer (the reference's mmap/slice path, read/slice.rs:17) and
    writer into a bytearray."""

    # ---- read side
    def __init__(self):
        self._buf = None
        self._chunks = None
        self._ctx = None
        self._out = None
        self.max_chunk_size = 0xFFFFFFFF

    @classmethod
    def read_header(cls, data, ctx=None, verify: bool = True) -> "Archive":
        """Archive::read_header_from_slice: signature + AHED, then index every chunk and (verify=True) check
        all chunk CRCs in one GPU batch, as read_chunk does per chunk (bytes.rs:64)."""
        from . import default_context
        a = cls()
        a._ctx = ctx or default_context()
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        a._buf = buf
        if buf.size < 8 or bytes(buf[:8]) != SIGNATURE:
            raise PnaError(_ffi.E_INVALID_DATA, "it is not PNA")
        a._chunks = index_archive(buf, 8)
        if not a._chunks or a._chunks[0].ty != ChunkType.AHED:
            raise PnaError(_ffi.E_INVALID_DATA, "expected `AHED` chunk")
        ah = bytes(buf[a._chunks[0].off:a._chunks[0].off + a._chunks[0].length])
        if len(ah) != 8:
            raise PnaError(_ffi.E_INVALID_DATA, "bad archive header")
        a.major, a.minor, a.archive_number = ah[0], ah[1], int.from_bytes(ah[4:8], "big")
        if verify:
            _verify(a._ctx, buf, a._chunks)
        return a

    @classmethod
    def read_multipart(cls, parts, ctx=None, verify: bool = True) -> "Archive":
        """A split archive (archive/read.rs:105-165 read_next_archive; writer archive/split_parts.rs): every part is a complete
        chunk stream signature + AHED(archive_number = part index) ... [ANXT] AEND, and an entry's chunks -- its FDAT stream
        included -- simply continue in the next part.  Parts are indexed and CRC-checked one by one; the entry iterator then
        runs over the concatenated chunk list without the archive-level chunks."""
        from . import default_context
        a = cls()
        a._ctx = ctx or default_context()
        a._chunks = []
        a._multipart = True
        for k, data in enumerate(parts):
            buf = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
            if buf.size < 8 or bytes(buf[:8]) != SIGNATURE:
                raise PnaError(_ffi.E_INVALID_DATA, "it is not PNA")
            chunks = index_archive(buf, 8)
            if not chunks or chunks[0].ty != ChunkType.AHED or chunks[0].length != 8:
                raise PnaError(_ffi.E_INVALID_DATA, "expected `AHED` chunk")
            ah = bytes(buf[chunks[0].off:chunks[0].off + 8])
            if int.from_bytes(ah[4:8], "big") != k:
                raise PnaError(_ffi.E_INVALID_DATA, f"part {k} carries archive number {int.from_bytes(ah[4:8], 'big')}")
            if k == 0:
                a._buf, a.major, a.minor, a.archive_number = buf, ah[0], ah[1], 0
            if verify:
                _verify(a._ctx, buf, chunks)
            if chunks[-1].ty != ChunkType.AEND:
                raise PnaError(_ffi.E_UNEXPECTED_EOF, "part without `AEND`")
            has_next = any(c.ty == ChunkType.ANXT for c in chunks)
            if has_next != (k + 1 < len(parts)):
                raise PnaError(_ffi.E_UNEXPECTED_EOF if has_next else _ffi.E_INVALID_DATA,
                               "next part missing" if has_next else "part does not announce a next archive (`ANXT`)")
            for c in chunks:
                if c.ty in (ChunkType.AHED, ChunkType.ANXT, ChunkType.AEND):
                    continue
                c.buf = buf
                a._chunks.append(c)
        return a

    def entries(self):
        """Entries iterator: NormalEntry | SolidEntry in archive order."""
        return _group(self._buf, self._chunks, self._ctx)

    def read_all(self, options: ReadOptions | None = None, batch_bytes: int = 1 << 30):
        """Batched extract of every FILE entry (the CLI's extract loop, extract.rs:868-1019, with the per-entry
        tasks folded into GPU batches).  Yields (entry, bytes)."""
        pend, size = [], 0
        for e in self.entries():
            if isinstance(e, SolidEntry):
                yield from self._flush(pend, options)
                pend, size = [], 0
                for ie in e.entries(options, self._ctx):
                    if ie.data_kind == DataKind.FILE:
                        pend.append(ie)
                yield from self._flush(pend, None)
                pend = []
                continue
            if e.data_kind != DataKind.FILE:
                continue
            pend.append(e)
            size += e.compressed_size
            if size >= batch_bytes:
                yield from self._flush(pend, options)
                pend, size = [], 0
        yield from self._flush(pend, options)

    def extract_plan(self, options: ReadOptions | None = None):
        """The whole extract hot path as ONE device plan: CRC check of every chunk of the archive (seam 1) fused
        with decrypt+decompress of every FILE entry (seam 2) over a single upload of the archive bytes.
        Returns (plan, entries); plan.run(); plan.fetch(sizes)."""
        if getattr(self, "_multipart", False):
            raise PnaError(_ffi.E_UNSUPPORTED, "extract_plan works on one contiguous archive buffer; use read_all for split archives")
        ents = [e for e in self.entries() if isinstance(e, NormalEntry) and e.data_kind == DataKind.FILE]
        owner = {}
        for i, e in enumerate(ents):
            for ch in e.chunks:
                owner[ch.off] = i
        base = self._buf.ctypes.data
        n = len(self._chunks)
        ptrs = np.fromiter((base + c.off - 4 for c in self._chunks), dtype=np.uint64, count=n)
        lens = np.fromiter((c.length + 4 for c in self._chunks), dtype=np.uint64, count=n)
        expect = np.fromiter((c.crc for c in self._chunks), dtype=np.uint32, count=n)
        entry_of = np.fromiter((owner.get(c.off, -1) for c in self._chunks), dtype=np.int32, count=n)
        plan = self._ctx.decode_plan([e._desc(options) for e in ents], crc=(ptrs, lens, expect, entry_of))
        return plan, ents

    def _flush(self, pend, options):
        if not pend:
            return
        outs, st, _ = self._ctx.decode_batch([e._desc(options) for e in pend])
        for e, o, s in zip(pend, outs, st):
            if s != _ffi.OK:
                raise PnaError(s, f"{e.name}: {_KIND_ERR.get(s, 'error')}")
            yield e, o.tobytes()

    # ---- write side
    @classmethod
    def write_header(cls, ctx=None, archive_number: int = 0) -> "Archive":
        """Archive::write_header (archive/write.rs:92): signature + AHED."""
        from . import default_context
        a = cls()
        a._ctx = ctx or default_context()
        a._out = bytearray(SIGNATURE)
        a._pending = []   # (type, data) waiting for their CRC batch
        a._emit(ChunkType.AHED, bytes([0, 0, 0, 0]) + struct.pack(">I", archive_number))
        return a

    def set_max_chunk_size(self, n: int):
        self.max_chunk_size = max(1, min(int(n), 0xFFFFFFFF))

    def _emit(self, ty: bytes, data, crc: int | None = None):
        self._pending.append((ty, data, crc))

    def add_entry(self, entry):
        """Archive::add_entry (archive/write.rs:368) -> write_chunks_to (entry.rs:895-912).  A freshly built entry brings
        the CRCs of its data chunks from the encode kernels; an entry READ from another archive (NormalEntry / SolidEntry) is
        re-serialised chunk by chunk in its wire order, every CRC recomputed by the finalize batch -- what makes
        lib/tests/copy_entries.rs:15-21 byte exact."""
        if isinstance(entry, _EntryBase):
            for ch in entry.chunks:
                self._emit(ch.ty, bytes(entry._body(ch)), None)
            return
        for ty, data, crc in entry.chunks(self.max_chunk_size):
            self._emit(ty, data, crc)

    def finalize(self) -> bytes:
        """Archive::finalize (archive/write.rs:545): AEND.  All chunk CRCs that were not produced by the encode
        kernels are computed here in one GPU batch (io::write_chunk, io.rs:183-197)."""
        self._emit(ChunkType.AEND, b"")
        need = [i for i, (_, _, crc) in enumerate(self._pending) if crc is None]
        if need:
            spans = [self._pending[i][0] + bytes(self._pending[i][1]) for i in need]
            crcs = self._ctx.crc32(spans)
            for i, c in zip(need, crcs):
                ty, data, _ = self._pending[i]
                self._pending[i] = (ty, data, int(c))
        out = self._out
        for ty, data, crc in self._pending:
            out += struct.pack(">I", len(data)) + ty
            out += bytes(data) if not isinstance(data, (bytes, bytearray)) else data
            out += struct.pack(">I", crc)
        self._pending = []
        return bytes(out)


def _cipher_fields(o: "WriteOptions", head_ty: bytes, header: bytes, iv) -> tuple[dict, int]:
    """to_hashed (entry/write.rs:75-125): what the encode seam needs for this entry's cipher and the length of the stream
    prefix that becomes its own chunk (builder.rs:62-69) -- the IV for CBC/CTR, the stream header for GCM, whose per-stream
    key is bound to the header chunk (aead.rs:166-208)."""
    if o.encryption == Encryption.NO:
        return {"key": None, "iv": None}, 0
    if o.cipher_mode == CipherMode.GCM:
        sh = gcm_stream_header(o.key, os.urandom(32), os.urandom(7), o.segment_size)
        return {"key": gcm_stream_key(o.key, sh, head_ty, header, o.phsf.encode()), "iv": None, "stream_header": sh}, GCM_STREAM_HEADER_LEN
    return {"key": o.key, "iv": iv}, 16


class BuiltEntry:
    """A NormalEntry / SolidEntry ready to be written: header, metadata chunks, PHSF and data bodies."""

    def __init__(self, head_ty, header, extra, phsf, data_ty, stream, end_ty, data_crcs=None, iv_len=0):
        self.head_ty, self.header, self.extra, self.phsf = head_ty, header, extra, phsf
        self.data_ty, self.stream, self.end_ty = data_ty, stream, end_ty
        self.data_crcs, self.iv_len = data_crcs, iv_len

    def chunks(self, max_chunk_size):
        yield self.head_ty, self.header, None
        for ty, d in self.extra:
            yield ty, d, None
        if self.phsf:
            yield ChunkType.PHSF, self.phsf.encode(), None
        s = self.stream
        pos = 0
        # the IV is its own FDAT chunk (builder.rs:62-69), then bodies of <= max_chunk_size
        if self.iv_len:
            yield self.data_ty, bytes(s[:self.iv_len]), None
            pos = self.iv_len
        k = 0
        while pos < len(s):
            n = min(max_chunk_size, len(s) - pos)
            crc = None
            if self.data_crcs is not None and k < len(self.data_crcs):
                crc = int(self.data_crcs[k])
            yield self.data_ty, s[pos:pos + n], crc
            pos += n
            k += 1
        yield self.end_ty, b"", None


class FileEntryBuilder:
    """builder/file.rs:41 -- io::Write-like: write() buffers plaintext, build() joins a GPU encode batch."""

    def __init__(self, name: str, options: WriteOptions, data_kind=DataKind.FILE):
        self.name, self.options, self.data_kind = name, options, data_kind
        self._parts = []
        self.iv = os.urandom(16) if options.encryption != Encryption.NO else None   # entry/write.rs:108-111

    @classmethod
    def new_with_options(cls, name: str, options: WriteOptions):
        return cls(name, options)

    def write(self, data) -> int:
        self._parts.append(bytes(data))
        return len(data)

    def plain(self) -> bytes:
        return b"".join(self._parts)

    def header_bytes(self) -> bytes:
        o = self.options
        return bytes([0, 0, self.data_kind, o.compression, o.encryption, o.cipher_mode]) + self.name.encode()

    def _encode_desc(self, max_chunk_size=0):
        o = self.options
        fields, self._prefix_len = _cipher_fields(o, ChunkType.FHED, self.header_bytes(), self.iv)
        return dict({"plain": self.plain(), "compression": o.compression, "level": o.level, "encryption": o.encryption,
                     "cipher_mode": o.cipher_mode, "max_chunk_size": max_chunk_size}, **fields)

    def build(self, ctx=None, max_chunk_size: int = 0) -> BuiltEntry:
        return EntryBuilder.build_many([self], ctx, max_chunk_size)[0]


class EntryBuilder:
    """Batching front of the encode seam: many FileEntryBuilders -> one pna_cuda_encode_batch."""

    @staticmethod
    def build_many(builders, ctx=None, max_chunk_size: int = 0):
        from . import default_context
        ctx = ctx or default_context()
        streams, crcs, st = ctx.encode_batch([b._encode_desc(max_chunk_size) for b in builders])
        out = []
        for b, s, c, code in zip(builders, streams, crcs, st):
            if code != _ffi.OK:
                raise PnaError(code, f"{b.name}: encode failed")
            o = b.options
            header = b.header_bytes()
            size = len(b.plain())
            extra = [(ChunkType.fSIZ, size.to_bytes(16, "big").lstrip(b"\0"))]   # entry.rs:901-903 minimal BE
            iv_len = b._prefix_len
            # encode kernels return CRCs for the bodies after the IV when max_chunk_size matches the writer's
            out.append(BuiltEntry(ChunkType.FHED, header, extra, o.phsf, ChunkType.FDAT, bytes(s), ChunkType.FEND,
                                  c if len(c) else None, iv_len))
        return out


class SolidEntryBuilder:
    """builder/solid.rs:68 -- inner entries are written STORE into one stream that is compressed+encrypted as a
    whole (archive.rs:206-210)."""

    def __init__(self, options: WriteOptions, ctx=None):
        from . import default_context
        self.options = options
        self._ctx = ctx or default_context()
        self._inner = Archive()
        self._inner._ctx = self._ctx
        self._inner._out = bytearray()
        self._inner._pending = []
        self.iv = os.urandom(16) if options.encryption != Encryption.NO else None

    def add_entry(self, entry: BuiltEntry):
        self._inner.add_entry(entry)

    def build(self, sdat_size: int = 32 * 1024) -> BuiltEntry:
        a = self._inner
        need = [i for i, (_, _, crc) in enumerate(a._pending) if crc is None]
        if need:
            crcs = self._ctx.crc32([a._pending[i][0] + bytes(a._pending[i][1]) for i in need])
            for i, c in zip(need, crcs):
                ty, data, _ = a._pending[i]
                a._pending[i] = (ty, data, int(c))
        inner = bytearray()
        for ty, data, crc in a._pending:
            inner += struct.pack(">I", len(data)) + ty + bytes(data) + struct.pack(">I", crc)
        o = self.options
        header = bytes([0, 0, o.compression, o.encryption, o.cipher_mode])
        fields, iv_len = _cipher_fields(o, ChunkType.SHED, header, self.iv)
        streams, _, st = self._ctx.encode_batch([dict({"plain": bytes(inner), "compression": o.compression, "level": o.level,
                                                       "encryption": o.encryption, "cipher_mode": o.cipher_mode}, **fields)])
        if st[0] != _ffi.OK:
            raise PnaError(st[0], "solid encode failed")
        be = BuiltEntry(ChunkType.SHED, header, [], o.phsf, ChunkType.SDAT, bytes(streams[0]), ChunkType.SEND, None, iv_len)
        be._sdat = sdat_size
        return be
