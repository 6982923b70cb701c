"""CPU tier: the oracle against the reference's own KATs and golden archives (parity pinning)."""
import hashlib
import os
import zlib

import pytest


def test_crc_kats(oracle):
    # lib/src/format/chunk.rs:31, lib/src/chunk.rs:68, lib/src/chunk/traits.rs:24, lib/src/io.rs:179
    assert oracle.chunk_crc(b"FDAT", bytes([0xAA, 0xBB, 0xCC, 0xDD])) == 0x47F32B10 == 1207118608
    assert oracle.chunk_crc(b"FDAT", bytes([1, 2, 3])) == 2776590148
    assert oracle.chunk_crc(b"AEND", b"") == 0x6BF6486D
    for n in (0, 1, 7, 8, 9, 1000, 70001):
        d = os.urandom(n)
        assert oracle.crc32(d) == zlib.crc32(d)


KAT_CBC = {   # lib/src/cipher.rs:261-266 (AES) and :278-283 (Camellia)
    1: "b4ea96c2fc15825ce85690385d8b6c5f92bf896b07e1ebeee0f68438aed6b63e",
    2: "47d8900ace4556eff9ff32a5b9605329feabcb5593910cb9acfc2fcb86c8a78b",
}


def test_cbc_kats(oracle):
    # lib/src/cipher.rs:256-292: key 0x11*32, iv 0x22*16, pt "PNA test vector!" (one block + PKCS#7 pad block)
    key, iv, pt = bytes([0x11]) * 32, bytes([0x22]) * 16, b"PNA test vector!"
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    for enc, alg in ((1, algorithms.AES), (2, algorithms.Camellia)):
        ct = oracle.cbc_encrypt(enc, key, iv, pt)
        assert ct.hex() == KAT_CBC[enc]          # expected bytes as written in the reference test
        e = Cipher(alg(key), modes.CBC(iv)).encryptor()
        want = e.update(pt + bytes([16]) * 16) + e.finalize()
        assert ct == want
        assert oracle.cbc_decrypt(enc, key, iv, ct) == pt


def test_ctr_matches_full_width_be_counter(oracle):
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    key = os.urandom(32)
    for iv in (os.urandom(16), bytes([0xFF]) * 16, bytes(8) + bytes([0xFF]) * 8):
        d = os.urandom(1000)
        e = Cipher(algorithms.AES(key), modes.CTR(iv)).encryptor()
        assert oracle.ctr(1, key, iv, d) == e.update(d) + e.finalize()
        for enc in (1, 2):   # OpenSSL's CTR mode == the restated Ctr128BE counter, both ciphers
            assert oracle.ctr(enc, key, iv, d) == oracle.ctr_restated(enc, key, iv, d)


def test_golden_archives(oracle, golden):
    """Every hot-path fixture decodes to the bytes the reference's tests assert (extract_compatibility.rs:104-213)."""
    raw = golden["raw_sha256"]
    for name, info in golden["archives"].items():
        buf = open(os.path.join(golden["dir"], info["file"]), "rb").read()
        if info["expect"] != "ok":
            with pytest.raises(oracle.OracleError):
                list(oracle.extract_all(buf, golden["password"].encode()))
            continue
        keys = {k: bytes.fromhex(v) for k, v in info["keys"].items()}
        got = list(oracle.extract_all(buf, golden["password"].encode(), keys))
        assert len(got) == len(info["entries"])
        for (n, d), e in zip(got, info["entries"]):
            assert n == e["name"] and hashlib.sha256(d).hexdigest() == e["sha256"]
            if n in raw:
                assert e["sha256"] == raw[n]
            if n.endswith("icon.bmp"):   # missing blob: three fixtures agree on this hash
                assert len(d) == 4194442 and e["sha256"].startswith("2948a8f585e46a53")


def test_multipart(oracle, golden):
    p1 = open(os.path.join(golden["dir"], "ref", "multipart.part1.pna"), "rb").read()
    p2 = open(os.path.join(golden["dir"], "ref", "multipart.part2.pna"), "rb").read()
    want = open(os.path.join(golden["dir"], "ref", "multipart_test.txt"), "rb").read()
    bodies, hdr = [], None
    for part in (p1, p2):
        for ch in oracle.read_chunks(part, 8):
            if ch.ty == b"FHED":
                hdr = ch.data
            if ch.ty == b"FDAT":
                bodies.append(ch.data)
    assert oracle.decode_stream(b"".join(bodies), hdr[3], hdr[4], hdr[5], None) == want


def test_roundtrip_cross_product(oracle):
    """archive.rs:221-362 pattern: every compression x cipher x mode round-trips."""
    key = os.urandom(32)
    for n in (0, 1, 15, 16, 17, 5000):
        plain = os.urandom(n // 2) + bytes(n - n // 2)
        for comp in (0, 1, 2):
            for enc, mode in ((0, 0), (1, 0), (1, 1), (2, 0), (2, 1)):
                s = oracle.encode_stream(plain, comp, -1, enc, mode, key, os.urandom(16))
                assert oracle.decode_stream(s, comp, enc, mode, key) == plain


def test_folding_crc_matches_zlib_and_reference_kats(oracle):
    """The CPU baseline's CRC is the PCLMULQDQ folding method crc32fast uses (lib/src/format/chunk.rs:8-11): same values as
    zlib's table CRC on every length / alignment class, and the reference's known answers (chunk.rs:31, traits.rs:24, io.rs:179)."""
    import ctypes as C
    import os
    import zlib
    L = oracle.lib()
    L.pna_oracle_crc32_fold.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
    L.pna_oracle_crc32_fold.restype = C.c_uint32
    for n in [0, 1, 15, 16, 17, 63, 64, 65, 79, 80, 81, 127, 128, 129, 255, 1000, 4095, 4096, 65537, 1 << 20]:
        for off in (0, 1, 3, 8, 15):
            b = os.urandom(n + off)[off:]
            for init in (0, 0xFFFFFFFF, 0x12345678):
                assert L.pna_oracle_crc32_fold(init, b, len(b)) == zlib.crc32(b, init), (n, off, init)
    fdat = zlib.crc32(b"FDAT")
    assert L.pna_oracle_crc32_fold(fdat, bytes([0xAA, 0xBB, 0xCC, 0xDD]), 4) == 0x47F32B10 == 1207118608
    assert L.pna_oracle_crc32_fold(fdat, bytes([1, 2, 3]), 3) == 2776590148
    assert L.pna_oracle_crc32_fold(zlib.crc32(b"AEND"), b"", 0) == 0x6BF6486D
    big = os.urandom(3 << 20) * 3
    assert L.pna_oracle_crc32_fold(fdat, big, len(big)) == zlib.crc32(big, fdat)
