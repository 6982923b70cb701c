"""CPU ORACLE (test infrastructure only -- never imported by the product package).

ctypes front-end of ``oracle/libpna_oracle.so`` (see ``pna_oracle.c`` for what is restated and from
where) plus a small, independent restatement of the PNA container walk so that the oracle can
decode the reference's golden archives end to end:

* signature / chunk framing   /root/reference/lib/src/format/signature.rs:6, lib/src/io.rs:117-197
* archive header, entry gather lib/src/archive/header.rs:27-56, lib/src/archive/read.rs:22-73
* NormalEntry / SolidEntry     lib/src/entry.rs:757-886 (FHED..FEND), :665-737 (SHED..SEND), :401-423
* header bytes                 lib/src/entry/header.rs:123-162, :274-296
* key derivation (PHSF)        lib/src/hash.rs:45-85  (argon2id / pbkdf2-sha256 via `cryptography`)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import base64
import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass, field

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpna_oracle.so")

OK, INVALID_DATA, UNEXPECTED_EOF, INVALID_INPUT, UNSUPPORTED, NOSPACE, OOM, INTERNAL = range(8)
SIGNATURE = b"\x89PNA\r\n\x1a\n"

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "pna_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return SO


class Job(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("len", C.c_uint64), ("compression", C.c_uint8), ("encryption", C.c_uint8),
                ("cipher_mode", C.c_uint8), ("_pad", C.c_uint8), ("key", C.c_uint8 * 32), ("out", C.c_void_p),
                ("cap", C.c_uint64), ("out_len", C.c_uint64), ("status", C.c_int32)]


class EncJob(C.Structure):
    _fields_ = [("plain", C.c_void_p), ("len", C.c_uint64), ("compression", C.c_uint8), ("encryption", C.c_uint8),
                ("cipher_mode", C.c_uint8), ("_pad", C.c_uint8), ("level", C.c_int32), ("key", C.c_uint8 * 32),
                ("iv", C.c_uint8 * 16), ("out", C.c_void_p), ("cap", C.c_uint64), ("out_len", C.c_uint64),
                ("status", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.pna_oracle_crc32_update.restype = C.c_uint32
        L.pna_oracle_crc32_update.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.pna_oracle_crc32_zlib.restype = C.c_uint32
        L.pna_oracle_crc32_zlib.argtypes = [C.c_void_p, C.c_size_t]
        L.pna_oracle_chunk_crc.restype = C.c_uint32
        L.pna_oracle_chunk_crc.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.pna_oracle_ctr.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p]
        L.pna_oracle_ctr_restated.argtypes = L.pna_oracle_ctr.argtypes
        L.pna_oracle_cbc_decrypt.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                             C.POINTER(C.c_size_t)]
        L.pna_oracle_cbc_encrypt.argtypes = L.pna_oracle_cbc_decrypt.argtypes
        L.pna_oracle_ecb.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p]
        L.pna_oracle_decompress.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                            C.POINTER(C.c_size_t)]
        L.pna_oracle_compress.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                          C.POINTER(C.c_size_t)]
        L.pna_oracle_compress_bound.restype = C.c_size_t
        L.pna_oracle_compress_bound.argtypes = [C.c_int, C.c_size_t]
        L.pna_oracle_encode_bound.restype = C.c_size_t
        L.pna_oracle_encode_bound.argtypes = [C.c_int, C.c_size_t]
        L.pna_oracle_decode_stream.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                               C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pna_oracle_encode_stream.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                               C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.pna_oracle_decode_batch_mt.argtypes = [C.POINTER(Job), C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.pna_oracle_encode_batch_mt.argtypes = [C.POINTER(EncJob), C.c_uint32, C.c_int, C.c_void_p]
        L.pna_oracle_zstd_version.restype = C.c_uint
        L.pna_oracle_gcm_decrypt_stream.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                                    C.POINTER(C.c_size_t)]
        L.pna_oracle_gcm_encrypt_bound.restype = C.c_size_t
        L.pna_oracle_gcm_encrypt_bound.argtypes = [C.c_size_t, C.c_uint32]
        L.pna_oracle_gcm_encrypt_stream.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                                    C.POINTER(C.c_size_t)]
        L.pna_oracle_gcm_openssl.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p]
        L.pna_oracle_gcm_segment.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p,
                                             C.c_char_p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status: int, what: str = ""):
        super().__init__(f"oracle status {status} {what}")
        self.status = status


# ----------------------------------------------------------------------------- primitives
def crc32(data: bytes, crc: int = 0) -> int:
    return lib().pna_oracle_crc32_update(crc, data, len(data))


def chunk_crc(ty: bytes, data: bytes) -> int:
    """format/chunk.rs:7-12"""
    return lib().pna_oracle_chunk_crc(ty, data, len(data))


def ctr(encryption: int, key: bytes, iv: bytes, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) or 1)
    rc = lib().pna_oracle_ctr(encryption, key, iv, data, len(data), out)
    if rc:
        raise OracleError(rc, "ctr")
    return out.raw[:len(data)]


def ctr_restated(encryption: int, key: bytes, iv: bytes, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) or 1)
    rc = lib().pna_oracle_ctr_restated(encryption, key, iv, data, len(data), out)
    if rc:
        raise OracleError(rc, "ctr")
    return out.raw[:len(data)]


def cbc_decrypt(encryption: int, key: bytes, iv: bytes, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) or 1)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_cbc_decrypt(encryption, key, iv, data, len(data), out, C.byref(n))
    if rc:
        raise OracleError(rc, "cbc_decrypt")
    return out.raw[:n.value]


def cbc_encrypt(encryption: int, key: bytes, iv: bytes, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) + 32)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_cbc_encrypt(encryption, key, iv, data, len(data), out, C.byref(n))
    if rc:
        raise OracleError(rc, "cbc_encrypt")
    return out.raw[:n.value]


def ecb(encryption: int, encrypt: bool, key: bytes, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) or 1)
    rc = lib().pna_oracle_ecb(encryption, int(encrypt), key, data, len(data), out)
    if rc:
        raise OracleError(rc, "ecb")
    return out.raw[:len(data) & ~15]


def _xz_decompress(data: bytes, cap: int | None = None) -> bytes:
    """Compression::XZ (entry/read.rs:182: liblzma::bufread::XzDecoder::new -- ONE stream, whatever follows it is not read).
    Python's lzma module is liblzma, the C library the reference links through liblzma-sys, so this arm is the reference's own
    decoder rather than a restatement.  Error classes as liblzma-rs maps them: a stream that ends early is "premature eof"
    (UnexpectedEof), LZMA_OPTIONS_ERROR is Unsupported, every other liblzma error InvalidData."""
    import lzma
    d = lzma.LZMADecompressor(format=lzma.FORMAT_XZ)
    try:
        out = d.decompress(data)
    except lzma.LZMAError as e:
        msg = str(e).lower()
        raise OracleError(UNSUPPORTED if ("unsupported" in msg or "invalid or unsupported options" in msg) else INVALID_DATA, f"xz: {e}")
    if not d.eof:
        raise OracleError(UNEXPECTED_EOF, "xz: premature eof")
    if cap is not None and len(out) > cap:
        raise OracleError(NOSPACE, "xz")
    return out


def decompress(compression: int, data: bytes, cap: int | None = None) -> bytes:
    if compression == 4:
        return _xz_decompress(data, cap)
    n = C.c_size_t(0)
    if cap is None:
        rc = lib().pna_oracle_decompress(compression, data, len(data), None, 0, C.byref(n))
        if rc not in (OK, NOSPACE):
            raise OracleError(rc, "decompress(size)")
        cap = n.value
    out = C.create_string_buffer(cap or 1)
    rc = lib().pna_oracle_decompress(compression, data, len(data), out, cap, C.byref(n))
    if rc:
        raise OracleError(rc, "decompress")
    return out.raw[:n.value]


def compress(compression: int, data: bytes, level: int = -1) -> bytes:
    cap = lib().pna_oracle_compress_bound(compression, len(data))
    out = C.create_string_buffer(cap or 1)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_compress(compression, level, data, len(data), out, cap, C.byref(n))
    if rc:
        raise OracleError(rc, "compress")
    return out.raw[:n.value]


def decode_stream(stream: bytes, compression: int, encryption: int, cipher_mode: int, key: bytes | None,
                  cap: int | None = None) -> bytes:
    """NormalEntry::reader (entry.rs:1150) over the concatenated FDAT bodies."""
    key = key or bytes(32)
    if compression == 4:   # decrypt through the C oracle (as a stored stream), then liblzma
        return _xz_decompress(decode_stream(stream, 0, encryption, cipher_mode, key), cap)
    n = C.c_size_t(0)
    if cap is None:
        rc = lib().pna_oracle_decode_stream(stream, len(stream), compression, encryption, cipher_mode, key, None, 0,
                                            C.byref(n))
        if rc not in (OK, NOSPACE):
            raise OracleError(rc, "decode_stream(size)")
        cap = n.value
    out = C.create_string_buffer(cap or 1)
    rc = lib().pna_oracle_decode_stream(stream, len(stream), compression, encryption, cipher_mode, key, out, cap,
                                        C.byref(n))
    if rc:
        raise OracleError(rc, "decode_stream")
    return out.raw[:n.value]


def encode_stream(plain: bytes, compression: int, level: int, encryption: int, cipher_mode: int, key: bytes | None,
                  iv: bytes | None) -> bytes:
    """FileEntryBuilder data_writer (builder.rs:45) -> IV || cipher(compress(plain))."""
    key = key or bytes(32)
    iv = iv or bytes(16)
    if compression == 4:   # XzEncoder::new(writer, level) (entry/write.rs:263): liblzma's easy encoder, CRC64 check; default preset 6
        import lzma
        return encode_stream(lzma.compress(plain, format=lzma.FORMAT_XZ, preset=6 if level < 0 else level), 0, -1, encryption,
                             cipher_mode, key, iv)
    cap = lib().pna_oracle_encode_bound(compression, len(plain))
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_encode_stream(plain, len(plain), compression, level, encryption, cipher_mode, key, iv, out,
                                        cap, C.byref(n))
    if rc:
        raise OracleError(rc, "encode_stream")
    return out.raw[:n.value]


# ----------------------------------------------------------------------------- GCM STREAM (cipher mode 2)
GCM_HEADER_LEN = 75
GCM_TAG_LEN = 16
GCM_DEFAULT_SEGMENT = 1 << 20


def hkdf_sha256(ikm: bytes, salt: bytes, info: bytes) -> bytes:
    """aead.rs:152-158: RFC 5869 extract + one expand block (32 bytes out).  hkdf 0.12: an empty salt is HashLen zeros."""
    import hashlib, hmac
    prk = hmac.new(salt if salt else bytes(32), ikm, hashlib.sha256).digest()
    return hmac.new(prk, info + b"\x01", hashlib.sha256).digest()


def gcm_key_confirmation(k_master: bytes) -> bytes:
    """aead.rs:162-164"""
    return hkdf_sha256(k_master, b"", b"PNA-KC-v1")


def gcm_stream_header(salt: bytes, nonce_prefix: bytes, segment_size: int, k_master: bytes) -> bytes:
    """aead.rs:124-132 to_bytes: salt(32) || nonce_prefix(7) || segment_size(u32 BE) || key_confirmation(32)"""
    assert len(salt) == 32 and len(nonce_prefix) == 7
    return salt + nonce_prefix + struct.pack(">I", segment_size) + gcm_key_confirmation(k_master)


def gcm_derive_stream_key(k_master: bytes, header: bytes, header_type: bytes, header_data: bytes, phsf: bytes) -> bytes:
    """aead.rs:166-208 entry_context + derive_stream_key; entry/read.rs:105-139 order of checks.
    Raises INVALID_DATA for a malformed header or a key-confirmation mismatch (AeadError -> InvalidData, error.rs:67)."""
    import hashlib
    if len(header) < GCM_HEADER_LEN:
        raise OracleError(INVALID_DATA, "datastream shorter than the stream header")
    header = header[:GCM_HEADER_LEN]
    (seg,) = struct.unpack_from(">I", header, 39)
    if seg == 0 or seg > 67108864:
        raise OracleError(INVALID_DATA, "segment size out of range")
    if len(k_master) != 32:
        raise OracleError(INVALID_DATA, "K_master is not 32 bytes")
    if gcm_key_confirmation(k_master) != header[43:75]:
        raise OracleError(INVALID_DATA, "key mismatch")
    ctx = (b"PNA-STREAM-v1" + hashlib.sha256(header_type + header_data).digest() + hashlib.sha256(phsf).digest()
           + header[32:39] + header[39:43])
    assert len(ctx) == 88
    return hkdf_sha256(k_master, header[:32], ctx)


def gcm_decrypt_stream(encryption: int, k_stream: bytes, stream: bytes) -> bytes:
    out = C.create_string_buffer(len(stream) or 1)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_gcm_decrypt_stream(encryption, k_stream, stream, len(stream), out, C.byref(n))
    if rc:
        raise OracleError(rc, "gcm_decrypt_stream")
    return out.raw[:n.value]


def gcm_encrypt_stream(encryption: int, k_stream: bytes, header: bytes, plain: bytes) -> bytes:
    """GcmEncryptWriter (gcm.rs:44-90): header || { ciphertext || tag } per segment, the last one flagged final."""
    (seg,) = struct.unpack_from(">I", header, 39)
    cap = lib().pna_oracle_gcm_encrypt_bound(len(plain), seg)
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    rc = lib().pna_oracle_gcm_encrypt_stream(encryption, k_stream, header, plain, len(plain), out, C.byref(n))
    if rc:
        raise OracleError(rc, "gcm_encrypt_stream")
    return out.raw[:n.value]


# ----------------------------------------------------------------------------- KDF (lib/src/hash.rs:45-85)
def _b64(s: str) -> bytes:
    return base64.b64decode(s + "=" * (-len(s) % 4))


def derive_key(phsf: str, password: bytes) -> bytes:
    parts = phsf.split("$")
    alg = parts[1]
    if alg.startswith("argon2"):
        from cryptography.hazmat.primitives.kdf.argon2 import Argon2id
        idx = 2
        if parts[idx].startswith("v="):
            idx += 1
        params = dict(kv.split("=") for kv in parts[idx].split(","))
        salt = _b64(parts[idx + 1])
        return Argon2id(salt=salt, length=32, iterations=int(params["t"]), lanes=int(params["p"]),
                        memory_cost=int(params["m"])).derive(password)
    if alg.startswith("pbkdf2"):
        import hashlib
        params = dict(kv.split("=") for kv in parts[2].split(","))
        salt = _b64(parts[3])
        return hashlib.pbkdf2_hmac("sha256", password, salt, int(params.get("i", 600000)), 32)
    raise OracleError(UNSUPPORTED, f"kdf {alg}")


# ----------------------------------------------------------------------------- container walk
@dataclass
class Chunk:
    ty: bytes
    data: bytes
    crc: int
    offset: int  # of the length field inside the archive


def read_chunks(buf: bytes, pos: int = 0, verify: bool = True):
    """io.rs:117-149 read_chunk: [len BE][type][data][crc BE]; CRC over type||data (io.rs:141)."""
    n = len(buf)
    while pos < n:
        if pos + 8 > n:
            raise OracleError(UNEXPECTED_EOF, "chunk header")
        (length,) = struct.unpack_from(">I", buf, pos)
        ty = bytes(buf[pos + 4:pos + 8])
        if pos + 12 + length > n:
            raise OracleError(UNEXPECTED_EOF, "chunk body")
        data = bytes(buf[pos + 8:pos + 8 + length])
        (crc,) = struct.unpack_from(">I", buf, pos + 8 + length)
        if verify and chunk_crc(ty, data) != crc:
            raise OracleError(INVALID_DATA, "broken chunk")
        yield Chunk(ty, data, crc, pos)
        pos += 12 + length


@dataclass
class Entry:
    solid: bool
    header: bytes
    name: str = ""
    data_kind: int = 0
    compression: int = 0
    encryption: int = 0
    cipher_mode: int = 0
    phsf: str | None = None
    raw_file_size: int | None = None
    bodies: list = field(default_factory=list)
    chunks: list = field(default_factory=list)

    @property
    def stream(self) -> bytes:
        return b"".join(self.bodies)


def parse_entries(chunks):
    """archive/read.rs:46-73 next_raw_item + entry.rs:757-886 / :665-737."""
    cur = None
    for ch in chunks:
        if ch.ty == b"FHED":
            h = ch.data
            cur = Entry(False, h, name=h[6:].decode("utf-8", "replace"), data_kind=h[2], compression=h[3],
                        encryption=h[4], cipher_mode=h[5])
        elif ch.ty == b"SHED":
            h = ch.data
            cur = Entry(True, h, compression=h[2], encryption=h[3], cipher_mode=h[4])
        elif cur is None:
            if ch.ty in (b"AHED", b"AEND", b"ANXT"):
                continue
            continue
        elif ch.ty in (b"FEND", b"SEND"):
            cur.chunks.append(ch)
            yield cur
            cur = None
            continue
        elif ch.ty in (b"FDAT", b"SDAT"):
            cur.bodies.append(ch.data)
        elif ch.ty == b"PHSF":
            cur.phsf = ch.data.decode()
        elif ch.ty == b"fSIZ":
            cur.raw_file_size = int.from_bytes(ch.data, "big")
        if cur is not None:
            cur.chunks.append(ch)


def read_archive(buf: bytes, verify: bool = True):
    if buf[:8] != SIGNATURE:
        raise OracleError(INVALID_DATA, "signature")
    return list(parse_entries(read_chunks(buf, 8, verify)))


def extract_all(buf: bytes, password: bytes | None = None, _keys=None):
    """Yield (name, bytes) for every FILE entry, descending into solid entries (entry.rs:567-583)."""
    keys = {} if _keys is None else _keys
    for e in read_archive(buf):
        key = None
        if e.encryption:
            if e.phsf is None:
                raise OracleError(INVALID_DATA, "`PHSF` chunk not found")
            if password is None:
                raise OracleError(INVALID_INPUT, "password")
            if e.phsf not in keys:
                keys[e.phsf] = derive_key(e.phsf, password)
            key = keys[e.phsf]
        if e.encryption and e.cipher_mode == 2:
            key = gcm_derive_stream_key(key, e.stream[:GCM_HEADER_LEN], b"SHED" if e.solid else b"FHED", e.header,
                                        e.phsf.encode())
        if e.solid:
            inner = decode_stream(e.stream, e.compression, e.encryption, e.cipher_mode, key)
            for ie in parse_entries(read_chunks(inner, 0, True)):
                if ie.data_kind == 0:
                    yield ie.name, decode_stream(ie.stream, ie.compression, ie.encryption, ie.cipher_mode, None)
        elif e.data_kind == 0:
            yield e.name, decode_stream(e.stream, e.compression, e.encryption, e.cipher_mode, key)
