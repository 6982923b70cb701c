"""Round-2 hardening: untrusted fSIZ (size hints), fresh IVs from the writer, end-marker handling of the API mirror,
contexts over several GPUs.  CPU tier: the rules that need no device; GPU tier: the behaviour through the C ABI and both host layers."""
import ctypes as C
import importlib
import os
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = bytes(range(32))
PHSF = "$pbkdf2-sha256$i=1000$c2FsdHNhbHRzYWx0$"


def _chunk(ty, data):
    return struct.pack(">I", len(data)) + ty + data + struct.pack(">I", zlib.crc32(data, zlib.crc32(ty)))


def _archive(entries):
    """entries: (name, stream bytes, compression, encryption, mode, fSIZ or None).  Layout FHED,[fSIZ],[PHSF],FDAT..,FEND."""
    out = [b"\x89PNA\r\n\x1a\n", _chunk(b"AHED", bytes(8))]
    for name, stream, comp, enc, mode, fsiz in entries:
        out.append(_chunk(b"FHED", bytes([0, 0, 0, comp, enc, mode]) + name.encode()))
        if fsiz is not None:
            out.append(_chunk(b"fSIZ", fsiz.to_bytes(8, "big").lstrip(b"\0") or b"\0"))
        if enc:
            out.append(_chunk(b"PHSF", PHSF.encode()))
            out.append(_chunk(b"FDAT", stream[:16]))
            stream = stream[16:]
        half = len(stream) // 2
        out.append(_chunk(b"FDAT", stream[:half]))
        out.append(_chunk(b"FDAT", stream[half:]))
        out.append(_chunk(b"FEND", b""))
    out.append(_chunk(b"AEND", b""))
    return b"".join(out)


@pytest.fixture(scope="module")
def host(pna):
    return importlib.import_module("portable-network-archive_b200._host")


# ---------------------------------------------------------------------------------------------- CPU tier
def test_size_hint_trust_rule(pna):
    """fSIZ is untrusted: beyond what the stream can decode to, or huge and out of proportion, it is ignored (host code, no GPU)."""
    L = importlib.import_module("portable-network-archive_b200._ffi").lib()
    U64 = (1 << 64) - 1
    assert L.pna_cuda_decode_size_bound(0, 1000) == 1000
    assert L.pna_cuda_decode_size_bound(1, 1000) == 1000 * 1032 + 1024
    assert L.pna_cuda_decode_size_bound(2, 12) == (12 // 3 + 2) * 131072
    for comp in (0, 1, 2):
        assert L.pna_cuda_size_hint_trusted(comp, 16, U64 - 16) == 0            # the ADVICE case: 2^64 - 16 on a 16-byte stream
        assert L.pna_cuda_size_hint_trusted(comp, 16, U64) == 0                 # absent
        assert L.pna_cuda_size_hint_trusted(comp, 1 << 20, 1 << 20) == 1
    assert L.pna_cuda_size_hint_trusted(0, 100, 101) == 0                        # store cannot grow
    assert L.pna_cuda_size_hint_trusted(2, 4 << 20, 4 << 30) == 0                # 4 GiB from 4 MiB: plausible for zstd, sized exactly anyway
    assert L.pna_cuda_size_hint_trusted(2, 2 << 20, 4 << 20) == 1


def test_api_mirror_stops_at_end_marker(pna, oracle):
    """next_raw_item (archive/read.rs:46-73) returns None at AEND: chunks behind it are not entries; both host layers agree."""
    mod = importlib.import_module("portable-network-archive_b200.archive")
    host = importlib.import_module("portable-network-archive_b200._host")
    good = _archive([("a.txt", b"hello", 0, 0, 0, 5)])
    trailing = good + _chunk(b"FHED", bytes([0, 0, 0, 0, 0, 0]) + b"ghost") + _chunk(b"FDAT", b"boo") + _chunk(b"FEND", b"")
    buf = np.frombuffer(trailing, dtype=np.uint8)
    chunks = mod.index_archive(buf, 8)
    names = [e.name for e in mod._group(buf, chunks, None)]
    assert names == ["a.txt"]
    assert [e["name"] for e in host.HostArchive(buf).entries()] == ["a.txt"]
    # an unknown critical chunk inside a solid entry is rejected like in a normal one (entry.rs:716)
    solid = b"\x89PNA\r\n\x1a\n" + _chunk(b"AHED", bytes(8)) + _chunk(b"SHED", bytes(5)) + _chunk(b"XBAD", b"?") + _chunk(b"SDAT", b"") + \
        _chunk(b"SEND", b"") + _chunk(b"AEND", b"")
    sbuf = np.frombuffer(solid, dtype=np.uint8)
    with pytest.raises(mod.PnaError) as ei:
        list(mod._group(sbuf, mod.index_archive(sbuf, 8), None))
    assert ei.value.kind == pna.E_INVALID_DATA
    with pytest.raises(host.HostError) as ei2:
        host.HostArchive(sbuf)
    assert ei2.value.kind == pna.E_INVALID_DATA


# ---------------------------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
def test_lying_fsiz_cannot_overflow_and_real_lengths_come_back(ctx, pna, host, oracle):
    """The reference never uses fSIZ to extract (stream-driven readers): whatever the chunk claims, every file comes out with
    the bytes and the length its stream decodes to, and no size sum wraps."""
    rng = np.random.default_rng(11)
    plain = [(b"fsiz test %d " % i) * (200 + 37 * i) + rng.integers(0, 256, 300, dtype=np.uint8).tobytes() for i in range(6)]
    iv = bytes(range(16))
    U64 = (1 << 64) - 1
    cases = [   # (codec, cipher, mode, fSIZ claim)
        (2, 0, 0, len(plain[0])),            # honest
        (2, 1, 1, U64 - 16),                 # ADVICE: wraps every offset sum when trusted
        (1, 0, 0, len(plain[2]) + 4096),     # claims more than there is: no stale tail may come back
        (2, 2, 0, len(plain[3]) - 100),      # claims less: the reference extracts it all the same
        (0, 0, 0, len(plain[4]) + 1),        # store cannot grow
        (1, 1, 1, 1 << 62),
    ]
    ents = []
    for i, (comp, enc, mode, claim) in enumerate(cases):
        ents.append((f"f{i}.bin", oracle.encode_stream(plain[i], comp, -1, enc, mode, KEY, iv), comp, enc, mode, claim))
    buf = np.frombuffer(_archive(ents), dtype=np.uint8)
    a = host.HostArchive(buf)
    a.set_key(PHSF, KEY)
    got = a.read_all(device=0, workers=2, verify=True)
    assert [g[0] for g in got] == [e[0] for e in ents]
    for (name, st, data), want in zip(got, plain):
        assert st == 0, (name, st)
        assert data == want, name
    assert [f[1] for f in a.files()] == [len(p) for p in plain]     # files() now reports the decoded lengths
    a.close()
    # straight through the C ABI: a hint that cannot be true is ignored (sizing pass), a too small one is PNA_E_NOSPACE + length
    descs = [{"bodies": [e[1]], "compression": e[2], "encryption": e[3], "cipher_mode": e[4], "key": KEY, "raw_size_hint": e[5]} for e in ents]
    plan = ctx.decode_plan(descs)
    plan.run()
    lens, st = plan.lengths()
    assert st[0] == 0 and st[1] == 0 and st[2] == 0 and st[4] == 0 and st[5] == 0, st
    assert st[3] == pna.E_NOSPACE and lens[3] == len(plain[3])
    assert [lens[i] for i in (0, 1, 2, 4, 5)] == [len(plain[i]) for i in (0, 1, 2, 4, 5)]
    plan.close()
    # python API mirror
    mod = importlib.import_module("portable-network-archive_b200.archive")
    arch = mod.Archive.read_header(buf, ctx)
    ro = mod.ReadOptions.with_password(b"x")
    ro._keys[PHSF] = KEY
    for (name, data), want in zip(arch.read_all(ro), plain):
        assert data == want, name


@pytest.mark.gpu
def test_extract_to_dir_with_lying_fsiz(ctx, pna, host, oracle, tmp_path):
    plain = [b"alpha " * 500, b"beta " * 900, b"gamma " * 100]
    ents = [("a.txt", oracle.encode_stream(plain[0], 2, -1, 0, 0, None, None), 2, 0, 0, len(plain[0]) + 999),
            ("b.txt", oracle.encode_stream(plain[1], 1, -1, 0, 0, None, None), 1, 0, 0, 17),
            ("c.txt", oracle.encode_stream(plain[2], 2, -1, 0, 0, None, None), 2, 0, 0, (1 << 64) - 16)]
    p = tmp_path / "lying.pna"
    p.write_bytes(_archive(ents))
    a = host.HostArchive.open_file(str(p))
    stats, st = a.extract_to_dir(str(tmp_path / "out"), device=0, workers=2, io_threads=2)
    assert st == [0, 0, 0], st
    for (name, *_), want in zip(ents, plain):
        assert (tmp_path / "out" / name).read_bytes() == want, name
    a.close()


@pytest.mark.gpu
def test_writer_draws_a_fresh_iv_per_entry(ctx, pna, host, oracle):
    """entry/write.rs:108-111: every encrypted entry gets its own random IV; a caller that passes none must not get zeros."""
    files = [(f"f{i}", b"same plaintext in every file " * 64) for i in range(8)]
    for mode in (0, 1):
        blobs = [host.create_archive(files, compression=2, encryption=1, cipher_mode=mode, key=KEY, phsf=PHSF, ivs=None, device=0, workers=2)
                 for _ in range(2)]
        ivs = []
        for blob in blobs:
            chunks = list(oracle.read_chunks(blob.tobytes(), 8))
            fdat = [c for c in chunks if c.ty == b"FDAT"]
            ivs += [bytes(c.data) for c in fdat if len(c.data) == 16]
            got = oracle.extract_all(blob.tobytes(), b"x", _keys={PHSF: KEY})
            assert [d for _, d in got] == [f[1] for f in files]
        assert len(ivs) == 16 and len(set(ivs)) == 16, "IVs repeat"
        assert bytes(16) not in ivs
    # caller-supplied IVs are still honoured (tests that pin ciphertext)
    fixed = bytes(range(16)) * len(files)
    blob = host.create_archive(files, compression=2, encryption=1, cipher_mode=1, key=KEY, phsf=PHSF, ivs=fixed, device=0)
    chunks = list(oracle.read_chunks(blob.tobytes(), 8))
    assert [bytes(c.data) for c in chunks if c.ty == b"FDAT" and len(c.data) == 16] == [bytes(range(16))] * len(files)


def _n_devices():
    if os.environ.get("PNA_EMU_DEVICES"):
        return int(os.environ["PNA_EMU_DEVICES"])
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_context_over_several_devices(pna, host, oracle, golden):
    """pna_cuda_init(device list): batch calls shard by entry inside the library, results come back in caller order; the
    C++ host layer partitions entry groups over the same list (cli/src/command/extract.rs:868-1019 in one process)."""
    nd = _n_devices()
    if nd < 2:
        pytest.skip("needs two GPUs (driver's multi-GPU run; `PNA_EMU_DEVICES=2 pytest --emu` on the CPU box)")
    devs = list(range(min(nd, 4)))
    mctx = pna.Context(devices=devs)
    assert mctx.L.pna_cuda_device_count(mctx.h) == len(devs)
    rng = np.random.default_rng(5)
    plain, descs = [], []
    for i in range(23):
        comp, enc, mode = [(2, 1, 1), (1, 2, 0), (0, 0, 0), (2, 0, 0), (2, 1, 2)][i % 5] if i % 5 != 4 else (2, 2, 1)
        p = (b"multi %d " % i) * (50 + 211 * (i % 7)) + rng.integers(0, 256, 100 * (i % 3), dtype=np.uint8).tobytes()
        s = oracle.encode_stream(p, comp, -1, enc, mode, KEY, rng.bytes(16))
        plain.append(p)
        descs.append({"bodies": [s[:len(s) // 3], s[len(s) // 3:]], "compression": comp, "encryption": enc, "cipher_mode": mode, "key": KEY,
                      "raw_size_hint": len(p) if i % 2 else None})
    outs, st, _ = mctx.decode_batch(descs)
    assert st == [0] * len(descs)
    assert [o.tobytes() for o in outs] == plain
    spans = [b"FDAT" + p[:777] for p in plain]
    assert [int(c) for c in mctx.crc32(spans)] == [oracle.crc32(s) for s in spans]
    # plans: create, run, lengths, fetch, counts
    plan = mctx.decode_plan(descs)
    plan.run()
    lens, st = plan.lengths()
    assert st == [0] * len(descs) and lens == [len(p) for p in plain]
    outs, st, _ = plan.fetch(lens)
    assert [o.tobytes() for o in outs] == plain
    plan.close()
    # encode: sharded too, streams + FDAT CRCs in caller order, readable by the reference pipeline
    ents = [{"plain": p, "compression": 2, "encryption": 1, "cipher_mode": 1, "key": KEY, "iv": rng.bytes(16), "max_chunk_size": 1000} for p in plain]
    streams, crcs, st = mctx.encode_batch(ents)
    assert st == [0] * len(ents)
    for s, c, p in zip(streams, crcs, plain):
        s = s.tobytes()
        assert oracle.decode_stream(s, 2, 1, 1, KEY, None) == p
        assert [int(x) for x in c] == [oracle.chunk_crc(b"FDAT", s[o:o + 1000]) for o in range(16, len(s), 1000)]
    mctx.close()
    # host layer over the device list: golden archive + a created one
    info = golden["archives"]["zstd_aes_ctr.pna"]
    buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
    a = host.HostArchive(buf)
    a.set_password(b"password")
    ref = oracle.extract_all(buf.tobytes(), b"password")
    got = a.read_all(devices=devs, workers=2, group_bytes=64 << 10, verify=True)
    assert [(n, d) for n, s, d in got] == [(n, d) for n, d in ref]
    a.close()
    files = [(f"dir/f{i}", p) for i, p in enumerate(plain)]
    blob = host.create_archive(files, compression=2, encryption=2, cipher_mode=1, key=KEY, phsf=PHSF, devices=devs, workers=2, group_bytes=4096)
    assert [d for _, d in oracle.extract_all(blob.tobytes(), b"x", _keys={PHSF: KEY})] == plain
