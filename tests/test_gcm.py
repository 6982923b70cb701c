"""GCM STREAM cipher mode (cipher mode 2; lib/src/cipher/gcm.rs, lib/src/cipher/aead.rs).

CPU tier: the oracle's restated GCM against OpenSSL's AES-256-GCM, the reference's HKDF known answer (aead.rs tests,
K_STREAM_FHED), and the library's host-side key schedule (no GPU work).  GPU tier: the decode seam against the oracle.
"""
import ctypes as C
import hashlib
import os
import struct

import numpy as np
import pytest

import corpus

# aead.rs tests: SALT, PREFIX, SEGMENT_SIZE, K_MASTER, HEADER_DATA, PHSF_DATA -> K_STREAM_FHED
KAT_SALT, KAT_PREFIX, KAT_SEG = bytes([0x42]) * 32, bytes([0x5A]) * 7, 0x01020304
KAT_K_STREAM_FHED = bytes.fromhex("b88e2edc07538bdd2b9afff57fb0d3433a1f4498d22a5911507e6827590fadb5")


def test_oracle_gcm_segment_equals_openssl(oracle):
    key, nonce = os.urandom(32), os.urandom(12)
    for n in (0, 1, 15, 16, 17, 255, 4096, 70001):
        data = os.urandom(n)
        o1, t1 = C.create_string_buffer(n or 1), C.create_string_buffer(16)
        o2, t2 = C.create_string_buffer(n or 1), C.create_string_buffer(16)
        assert oracle.lib().pna_oracle_gcm_openssl(key, nonce, data, n, o1, t1) == 0
        assert oracle.lib().pna_oracle_gcm_segment(1, key, nonce, data, n, o2, t2) == 0
        assert o1.raw[:n] == o2.raw[:n] and t1.raw == t2.raw


def test_oracle_hkdf_known_answer(oracle):
    """aead.rs tests: derive_stream_key(K_MASTER, header, FHED, "header", "phsf") == K_STREAM_FHED (external RFC 5869 vector)."""
    ctx = (b"PNA-STREAM-v1" + hashlib.sha256(b"FHED" + b"header").digest() + hashlib.sha256(b"phsf").digest() + KAT_PREFIX
           + struct.pack(">I", KAT_SEG))
    assert oracle.hkdf_sha256(b"master_key", KAT_SALT, ctx) == KAT_K_STREAM_FHED


def test_library_key_schedule_matches_oracle(oracle, pna):
    """pna_cuda_gcm_stream_key / _header are host code: same answers as the oracle's hashlib/hmac restatement."""
    mod = __import__("importlib").import_module("portable-network-archive_b200.archive")
    for _ in range(8):
        km, salt, prefix = os.urandom(32), os.urandom(32), os.urandom(7)
        seg = int.from_bytes(os.urandom(3), "big") + 1
        hdr = mod.gcm_stream_header(km, salt, prefix, seg)
        assert hdr == oracle.gcm_stream_header(salt, prefix, seg, km)
        hd, ph = os.urandom(int.from_bytes(os.urandom(1), "big") + 6), os.urandom(70)
        for ty in (b"FHED", b"SHED"):
            assert mod.gcm_stream_key(km, hdr, ty, hd, ph) == oracle.gcm_derive_stream_key(km, hdr, ty, hd, ph)
        with pytest.raises(pna.PnaError) as ei:                       # wrong password -> KeyMismatch (InvalidData)
            mod.gcm_stream_key(os.urandom(32), hdr, b"FHED", hd, ph)
        assert ei.value.kind == pna.E_INVALID_DATA
        bad = hdr[:39] + struct.pack(">I", 0) + hdr[43:]
        with pytest.raises(pna.PnaError):                             # segment size out of range (aead.rs:141)
            mod.gcm_stream_key(km, bad, b"FHED", hd, ph)
        with pytest.raises(pna.PnaError):
            mod.gcm_stream_key(km, hdr[:74], b"FHED", hd, ph)         # shorter than the stream header


def test_oracle_gcm_stream_round_trip_and_layout(oracle):
    """gcm.rs tests: empty plaintext = one tag-only segment; an exact multiple of the segment size ends on a FULL final segment."""
    key = os.urandom(32)
    for enc in (1, 2):
        hdr = os.urandom(39) + struct.pack(">I", 4) + os.urandom(32)
        assert len(oracle.gcm_encrypt_stream(enc, key, hdr, b"")) == 75 + 16
        assert len(oracle.gcm_encrypt_stream(enc, key, hdr, b"abcd")) == 75 + 4 + 16
        assert len(oracle.gcm_encrypt_stream(enc, key, hdr, b"abcde")) == 75 + 2 * 16 + 5
        for n in (0, 1, 4, 5, 8, 9, 1000):
            p = os.urandom(n)
            s = oracle.gcm_encrypt_stream(enc, key, hdr, p)
            assert oracle.gcm_decrypt_stream(enc, key, s) == p
            if n > 4:   # dropping the final segment leaves a non-final one at the end: its nonce flag no longer matches
                cut = s[:75 + 4 + 16]
                with pytest.raises(oracle.OracleError):
                    oracle.gcm_decrypt_stream(enc, key, cut)


# ------------------------------------------------------------------------------------------------------------ GPU tier
def _gcm_entry(oracle, plain, comp, enc, seg, key, split=None, hint=True):
    hdr = os.urandom(39) + struct.pack(">I", seg) + os.urandom(32)   # the key confirmation belongs to the key schedule, not to this seam
    body = oracle.compress(comp, plain) if comp else plain
    stream = oracle.gcm_encrypt_stream(enc, key, hdr, body)
    if split is None:
        bodies = [stream]
    else:
        cuts = sorted(set(min(c, len(stream)) for c in split))
        bodies = [stream[a:b] for a, b in zip([0] + cuts, cuts + [len(stream)])]
    return {"bodies": bodies, "compression": comp, "encryption": enc, "cipher_mode": 2, "key": key,
            "raw_size_hint": len(plain) if hint else None}


@pytest.mark.gpu
@pytest.mark.parametrize("enc", [1, 2])
def test_gcm_decode_cross_product(ctx, oracle, enc):
    """Segment sizes around the tile (16 KiB) and block (16 B) boundaries x payload sizes incl. empty and exact multiples."""
    entries, want = [], []
    k = 0
    for seg in (1, 5, 16, 100, 4096, 16384, 16400, 50000, 1 << 20):
        for n in (0, 1, 15, 16, 17, 4096, 16384, 16385, 50000, 100000, 300000):
            if seg < 16 and n > 300:
                continue
            for comp in (0, 2):
                key = os.urandom(32)    # stream keys are per entry (aead.rs:188)
                plain = corpus.make_file(7000 + k, n)
                k += 1
                split = None if k % 3 else [40, 75, 76, 75 + seg, 75 + seg + 16, 75 + seg + 17, 99999]
                entries.append(_gcm_entry(oracle, plain, comp, enc, seg, key, split, hint=(k % 2 == 0)))
                want.append(plain)
    outs, st, lens = ctx.decode_batch(entries)
    assert st == [0] * len(entries)
    for o, w in zip(outs, want):
        assert o.tobytes() == w


@pytest.mark.gpu
def test_gcm_multi_tile_segments_and_mixed_batch(ctx, oracle):
    """1 MiB segments (64 tiles chained with H^1024), 4 MiB entries, AES and Camellia and CTR entries in one batch."""
    entries, want = [], []
    for i, (enc, mode) in enumerate([(1, 2), (2, 2), (1, 1), (1, 2), (2, 0), (2, 2)]):
        key = os.urandom(32)
        plain = corpus.make_file(9000 + i, (4 << 20) + 123 * i)
        if mode == 2:
            entries.append(_gcm_entry(oracle, plain, 2 if i % 2 else 0, enc, 1 << 20, key))
        else:
            s = oracle.encode_stream(plain, 2, -1, enc, mode, key, os.urandom(16))
            entries.append({"bodies": [s], "compression": 2, "encryption": enc, "cipher_mode": mode, "key": key,
                            "raw_size_hint": len(plain)})
        want.append(plain)
    outs, st, lens = ctx.decode_batch(entries)
    assert st == [0] * len(entries)
    for o, w in zip(outs, want):
        assert hashlib.sha256(o.tobytes()).digest() == hashlib.sha256(w).digest()


@pytest.mark.gpu
def test_gcm_error_classes(ctx, oracle, pna):
    """Every AEAD failure is InvalidData (error.rs:67-74): tampering, truncation, layout violations."""
    key = os.urandom(32)
    plain = corpus.make_file(5, 40000)
    good = _gcm_entry(oracle, plain, 0, 1, 16384, key)
    s = bytes(good["bodies"][0])
    seg_len = 16384 + 16

    def flip(pos):
        b = bytearray(s)
        b[pos] ^= 1
        return bytes(b)

    cases = [
        ([s], 0),
        ([flip(75 + 100)], pna.E_INVALID_DATA),                   # ciphertext bit           gcm.rs:283
        ([flip(75 + seg_len - 1)], pna.E_INVALID_DATA),           # tag bit of segment 0
        ([flip(len(s) - 1)], pna.E_INVALID_DATA),                 # tag bit of the final segment
        ([flip(33)], pna.E_INVALID_DATA),                         # nonce prefix in the header
        ([s[:75 + 2 * seg_len]], pna.E_INVALID_DATA),             # final segment dropped: flag mismatch on the new last one
        ([s[:75 + 2 * seg_len + 5]], pna.E_INVALID_DATA),         # 5-byte tail after verified segments: truncation  gcm.rs:254
        ([s[:len(s) - 3]], pna.E_INVALID_DATA),                   # final segment cut short
        ([s[:60]], pna.E_INVALID_DATA),                           # shorter than the stream header  entry/read.rs:108
        ([s[:75]], pna.E_INVALID_DATA),                           # no segment at all          gcm.rs:250
        ([s[:75 + 9]], pna.E_INVALID_DATA),                       # shorter than one empty final segment
        ([s[:39] + bytes(4) + s[43:]], pna.E_INVALID_DATA),       # segment size 0             aead.rs:141
        ([s[:39] + struct.pack(">I", (64 << 20) + 1) + s[43:]], pna.E_INVALID_DATA),
        ([s + s[75:]], pna.E_INVALID_DATA),                       # bytes after the final segment: it is no longer final
    ]
    entries = [dict(good, bodies=b) for b, _ in cases]
    outs, st, lens = ctx.decode_batch(entries)
    for (b, want), got, o in zip(cases, st, outs):
        assert got == want, (len(b[0]), got, want)
        try:
            ref = oracle.gcm_decrypt_stream(1, key, b[0])
            assert want == 0 and o.tobytes() == ref
        except oracle.OracleError as e:
            assert want == e.status
    wrong = dict(good, key=os.urandom(32))
    _, st, _ = ctx.decode_batch([wrong, good])
    assert st[0] == pna.E_INVALID_DATA and st[1] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("enc", [1, 2])
def test_gcm_encode_reference_readable(ctx, oracle, enc):
    """GcmEncryptWriter (gcm.rs:44-90) on the GPU: the reference pipeline reads it back, tags verify, store entries are
    bit-exact with the reference dataflow, chunk CRCs cover the bodies after the 75-byte prefix (builder.rs:62-69)."""
    mod = __import__("importlib").import_module("portable-network-archive_b200.archive")
    entries, plains = [], []
    k = 0
    for seg in (1, 16, 100, 4096, 16384, 16400, 1 << 20):
        for n in (0, 1, 15, 16, 17, 4096, 16384, 16385, 100000, 300000, (1 << 20) + 5):
            if seg < 16 and n > 300:
                continue
            for comp in (0, 1, 2):
                hdr = mod.gcm_stream_header(os.urandom(32), os.urandom(32), os.urandom(7), seg)
                p = corpus.make_file(8000 + k, n)
                entries.append({"plain": p, "compression": comp, "level": -1, "encryption": enc, "cipher_mode": 2, "key": os.urandom(32),
                                "stream_header": hdr, "max_chunk_size": [0, 16, 1000, 65536][k % 4]})
                plains.append(p)
                k += 1
    streams, crcs, st = ctx.encode_batch(entries)
    assert st == [0] * len(entries)
    for e, s, c, p in zip(entries, streams, crcs, plains):
        s = s.tobytes()
        assert s[:75] == e["stream_header"]
        body = oracle.gcm_decrypt_stream(enc, e["key"], s)      # every tag verifies, final flag on the last segment only
        assert (oracle.decompress(e["compression"], body, len(p)) if e["compression"] else body) == p
        mcs = e["max_chunk_size"] or 0xFFFFFFFF
        bodies = [s[o:o + mcs] for o in range(75, len(s), mcs)]
        assert [int(x) for x in c] == [oracle.chunk_crc(b"FDAT", b) for b in bodies]
        if e["compression"] == 0:
            assert s == oracle.gcm_encrypt_stream(enc, e["key"], e["stream_header"], p)
    back, st2, _ = ctx.decode_batch([{"bodies": [s], "compression": e["compression"], "encryption": enc, "cipher_mode": 2,
                                      "key": e["key"], "raw_size_hint": None} for e, s in zip(entries, streams)])
    assert st2 == [0] * len(entries)
    assert all(b.tobytes() == p for b, p in zip(back, plains))


@pytest.mark.gpu
def test_gcm_encode_rejects_bad_header(ctx, pna):
    good = {"plain": b"abc", "compression": 0, "encryption": 1, "cipher_mode": 2, "key": os.urandom(32)}
    _, _, st = ctx.encode_batch([dict(good), dict(good, stream_header=bytes(39) + struct.pack(">I", 0) + bytes(32)),
                                 dict(good, stream_header=bytes(39) + struct.pack(">I", (64 << 20) + 1) + bytes(32))])
    assert st == [pna.E_INVALID_INPUT] * 3


@pytest.mark.gpu
@pytest.mark.parametrize("enc", [1, 2])
def test_gcm_archive_round_trip_through_builder_api(ctx, pna, oracle, enc):
    """cli/tests/cli/encrypt.rs-style round trip for cipher mode GCM: created through the host mirror (stream header, key
    confirmation and per-entry stream key from the library's key schedule, data on the GPU), read back by the oracle's
    restatement of the reference reader (which re-derives every key from the password and verifies every tag) and by our own."""
    opts = pna.WriteOptions(compression=2, encryption=enc, cipher_mode=2, password=b"pw", kdf_params={"i": 1000}, segment_size=65536)
    files = {f"d/f{i}.bin": corpus.make_file(1900 + i, n) for i, n in enumerate([0, 10, 5000, 70_000, 200_000, 65536, 131072] * 3)}
    builders = []
    for name, data in files.items():
        b = pna.FileEntryBuilder.new_with_options(name, opts)
        b.write(data)
        builders.append(b)
    a = pna.Archive.write_header(ctx)
    a.set_max_chunk_size(50_000)
    for be in pna.EntryBuilder.build_many(builders, ctx, max_chunk_size=50_000):
        a.add_entry(be)
    blob = a.finalize()
    assert dict(oracle.extract_all(blob, b"pw")) == files
    ar = pna.Archive.read_header(np.frombuffer(blob, dtype=np.uint8), ctx)
    assert {e.name: d for e, d in ar.read_all(pna.ReadOptions.with_password(b"pw"))} == files
    with pytest.raises(pna.PnaError) as ei:      # wrong password: KeyMismatch before any segment is touched (entry/read.rs:127)
        list(ar.read_all(pna.ReadOptions.with_password(b"not pw")))
    assert ei.value.kind == pna.E_INVALID_DATA
    # an entry moved under another name no longer decrypts: the header chunk is bound into the stream key (aead.rs:1-6)
    moved = bytearray(blob)
    at = moved.find(b"d/f2.bin")
    moved[at:at + 8] = b"d/fX.bin"
    crc_at = at - 6 - 4   # FHED chunk: [len][type][data][crc]; recompute its CRC so only the key derivation notices
    ln = int.from_bytes(moved[crc_at - 4:crc_at], "big")
    import zlib
    moved[crc_at + 4 + ln:crc_at + 8 + ln] = zlib.crc32(bytes(moved[crc_at:crc_at + 4 + ln])).to_bytes(4, "big")
    ar2 = pna.Archive.read_header(np.frombuffer(bytes(moved), dtype=np.uint8), ctx)
    with pytest.raises(pna.PnaError):
        list(ar2.read_all(pna.ReadOptions.with_password(b"pw")))
    with pytest.raises(oracle.OracleError):
        dict(oracle.extract_all(bytes(moved), b"pw"))


@pytest.mark.gpu
def test_gcm_solid_archive_round_trip(ctx, pna, oracle):
    """Solid entry under GCM: the stream key is bound to the SHED chunk (entry.rs:567, aead.rs:166)."""
    store = pna.WriteOptions.store()
    opts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=2, password=b"pw", kdf_params={"i": 1000}, segment_size=100_000)
    files = {f"s/f{i}.txt": corpus.make_file(2900 + i, n) for i, n in enumerate([0, 1, 4000, 150_000, 333_333])}
    sb = pna.SolidEntryBuilder(opts, ctx)
    builders = []
    for name, data in files.items():
        b = pna.FileEntryBuilder.new_with_options(name, store)
        b.write(data)
        builders.append(b)
    for be in pna.EntryBuilder.build_many(builders, ctx):
        sb.add_entry(be)
    a = pna.Archive.write_header(ctx)
    a.set_max_chunk_size(32 * 1024)
    a.add_entry(sb.build())
    blob = a.finalize()
    assert dict(oracle.extract_all(blob, b"pw")) == files
    ar = pna.Archive.read_header(np.frombuffer(blob, dtype=np.uint8), ctx)
    assert {e.name: d for e, d in ar.read_all(pna.ReadOptions.with_password(b"pw"))} == files


def test_encode_bound_counts_header_and_tags(pna):
    """pna_cuda_encode_bound / _crc_count are host arithmetic: a GCM stream is header(75) + payload + one tag per segment
    (at least one: the empty final segment, gcm.rs tests `empty_plaintext_emits_single_tag_only_segment`)."""
    import importlib
    ffi = importlib.import_module("portable-network-archive_b200._ffi")
    L = ffi.lib()
    for n, seg in [(0, 1 << 20), (1, 1 << 20), (1 << 20, 1 << 20), ((1 << 20) + 1, 1 << 20), (5 << 20, 65536), (1000, 16)]:
        d = ffi.EncodeDesc()
        hdr = C.create_string_buffer(bytes(39) + struct.pack(">I", seg) + bytes(32), 75)
        d.plain.len, d.compression, d.encryption, d.cipher_mode = n, 0, 1, 2
        d.stream_header = C.cast(hdr, C.c_void_p)
        bound = L.pna_cuda_encode_bound(C.byref(d))
        exact = 75 + n + 16 * max(1, -(-n // seg))          # store: the payload is the plaintext itself
        assert exact <= bound <= exact + 16, (n, seg, bound, exact)
        d.max_chunk_size = 1000
        assert L.pna_cuda_encode_crc_count(C.byref(d)) >= -(-(exact - 75) // 1000)


def test_multipart_index_without_gpu(golden):
    """Archive.read_multipart's part walk is host logic: the reference's two-part fixture yields ONE entry whose FDAT stream
    spans both parts; part order and completeness are checked (archive/read.rs:105-165)."""
    import importlib
    mod = importlib.import_module("portable-network-archive_b200.archive")
    p1 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part1.pna"), dtype=np.uint8)
    p2 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part2.pna"), dtype=np.uint8)
    a = mod.Archive.read_multipart([p1, p2], ctx=object(), verify=False)
    es = list(a.entries())
    assert len(es) == 1 and es[0].name == "multipart_test.txt" and len(es[0].bodies) == 2
    assert {id(ch.buf) for ch in es[0].chunks} == {id(p1), id(p2)}
    for bad in ([p2, p1], [p1], [p1, p1]):
        with pytest.raises(mod.PnaError):
            mod.Archive.read_multipart(bad, ctx=object(), verify=False)
