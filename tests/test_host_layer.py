"""C++ host layer (include/pna_host.hpp, libpna_host.so): index pass + entry grouping on the CPU tier, pipelined extract and
archive creation on the GPU tier.  Reads like lib/tests/extract_compatibility.rs / extract_solid_compatibility.rs."""
import ctypes as C
import hashlib
import importlib
import os
import re

import numpy as np
import pytest

import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(pna):
    return importlib.import_module("portable-network-archive_b200._host")


def test_host_exports_match_header(host):
    hdr = open(os.path.join(ROOT, "include", "pna_host.hpp")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pnah_\w+)\s*\(", hdr))
    lib = C.CDLL(host.LIB_PATH)   # the in-tree build (under --emu: the emulator build -- two libraries of one soname cannot share a process)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(host.EXPORTS)


def test_index_pass_matches_python_mirror_on_every_fixture(host, pna, golden):
    """The index pass touches only the 12-byte chunk frames (bytes.rs:90); entry grouping = archive/read.rs:46-73."""
    mod = importlib.import_module("portable-network-archive_b200.archive")
    for name, info in golden["archives"].items():
        buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
        if name.startswith("multipart.part1"):   # the entry continues in part 2: a single slice ends inside it
            with pytest.raises(host.HostError) as ei:
                host.HostArchive(buf)
            assert ei.value.kind == pna.E_UNEXPECTED_EOF
            continue
        a = host.HostArchive(buf)
        chunks = mod.index_archive(buf, 8)
        assert a.n_chunks == len(chunks), name
        ents = a.entries()
        if name.startswith("multipart"):
            continue
        n_fhed = sum(1 for c in chunks if c.ty == b"FHED")
        n_shed = sum(1 for c in chunks if c.ty == b"SHED")
        assert sum(1 for e in ents if e["kind"] == 0) == n_fhed and sum(1 for e in ents if e["kind"] == 1) == n_shed, name
        assert sum(e["n_bodies"] for e in ents) == sum(1 for c in chunks if c.ty in (b"FDAT", b"SDAT")), name


def test_index_pass_errors(host):
    with pytest.raises(host.HostError) as ei:
        host.HostArchive(b"not a pna archive at all")
    assert ei.value.kind == 1   # InvalidData "it is not PNA"
    good = np.fromfile(os.path.join(ROOT, "tests", "golden", "ref", "zstd.pna"), dtype=np.uint8)
    with pytest.raises(host.HostError) as ei:
        host.HostArchive(good[:1000].copy())
    assert ei.value.kind == 2   # UnexpectedEof: truncated chunk
    bad = good.copy()
    bad[12] = ord("1")          # chunk type must be ASCII letters (chunk/types.rs:204)
    with pytest.raises(host.HostError) as ei:
        host.HostArchive(bad)
    assert ei.value.kind == 1


def _synthetic_archive(n_entries, rnd, decoys=True):
    """Entries with 1..3 FDAT bodies of ragged sizes; bodies carry well-formed chunk headers as DATA (decoys for the
    speculative chunk walk) and a few are megabytes long (a thread's region can start deep inside one)."""
    import struct
    import zlib
    parts = [b"\x89PNA\r\n\x1a\n"]

    def chunk(ty, data):
        parts.append(struct.pack(">I", len(data)) + ty + data + struct.pack(">I", zlib.crc32(ty + data)))
    decoy = b"".join(struct.pack(">I", 8) + b"FDAT" + bytes(8) + bytes(4) for _ in range(12))   # 12 plausible chunks
    chunk(b"AHED", bytes(8))
    want = []
    for i in range(n_entries):
        name = f"dir/{i:07d}.bin"
        chunk(b"FHED", bytes([0, 0, 0, 0, 0, 0]) + name.encode())
        nb = rnd.randrange(1, 4)
        total = 0
        sizes = []
        for b in range(nb):
            n = rnd.choice([0, 1, 17, 300, 4000, 70_000]) if i % 997 else 3_000_000
            body = (decoy * (n // len(decoy) + 1))[:n] if decoys else bytes(n)
            chunk(b"FDAT", body)
            sizes.append(n)
            total += n
        chunk(b"fSIZ", total.to_bytes(8, "big").lstrip(b"\0") or b"")
        chunk(b"FEND", b"")
        want.append((name, nb, total))
    chunk(b"AEND", b"")
    return b"".join(parts), want


def test_parallel_index_pass_equals_serial_walk(host, monkeypatch):
    """Archives >= 32 MiB are indexed by several threads that guess a chunk boundary inside their region and are only
    believed when the previous thread's walk lands exactly there (host_api.cpp index_chunks / group_entries): the result
    must be the serial walk's, decoy chunk headers inside bodies and multi-megabyte chunks notwithstanding."""
    import random
    blob, want = _synthetic_archive(30_000, random.Random(12))
    assert len(blob) > (64 << 20)
    buf = np.frombuffer(blob, dtype=np.uint8)
    views = []
    for serial in (False, True):
        if serial:
            monkeypatch.setenv("PNA_INDEX_SERIAL", "1")
        a = host.HostArchive(buf)
        ents = a.entries()
        views.append((a.n_chunks, [(e["name"], e["n_bodies"], e["compressed_size"], e["raw_file_size"]) for e in ents]))
        a.close()
    assert views[0] == views[1]
    assert [(n, b, t, t) for n, b, t in want] == views[0][1]
    # a truncated archive reports the same error either way
    for serial in (False, True):
        if serial:
            monkeypatch.setenv("PNA_INDEX_SERIAL", "1")
        else:
            monkeypatch.delenv("PNA_INDEX_SERIAL", raising=False)
        with pytest.raises(host.HostError) as ei:
            host.HostArchive(buf[: len(blob) - 40_000_001].copy())
        assert ei.value.kind == 2


@pytest.mark.gpu
def test_golden_archives_through_cpp_host(host, pna, ctx, golden):
    for name, info in golden["archives"].items():
        if name.startswith("multipart"):
            continue
        buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
        a = host.HostArchive(buf)
        for phsf, key in info["keys"].items():
            a.set_key(phsf, bytes.fromhex(key))
        if info["expect"] != "ok":
            with pytest.raises(host.HostError) as ei:
                r = a.read_all()
                bad = [s for _, s, _ in r if s]
                if bad:
                    raise host.HostError(bad[0], "entry status")
            assert ei.value.kind == pna.E_UNSUPPORTED, name
            continue
        got = a.read_all(workers=2, group_bytes=20_000)
        assert [n for n, _, _ in got] == [e["name"] for e in info["entries"]], name
        for (n, st, d), e in zip(got, info["entries"]):
            assert st == 0 and len(d) == e["size"] and hashlib.sha256(d).hexdigest() == e["sha256"], (name, n)


@pytest.mark.gpu
def test_broken_chunk_detected_by_cpp_host(host, pna, ctx, golden):
    info = golden["archives"]["zstd.pna"]
    buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8).copy()
    buf[5000] ^= 0x40
    a = host.HostArchive(buf)
    out, offs, st = a.extract_files()
    assert st.count(pna.E_INVALID_DATA) == 1 and st.count(0) == len(st) - 1     # the entry owning the broken chunk
    # a broken archive-level chunk (AHED data) is an archive error
    buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8).copy()
    buf[8 + 8 + 5] ^= 1
    with pytest.raises(host.HostError) as ei:
        host.HostArchive(buf).extract_files()
    assert ei.value.kind == pna.E_INVALID_DATA


@pytest.mark.gpu
@pytest.mark.parametrize("comp,enc,mode", [(2, 1, 1), (1, 2, 0), (0, 0, 0), (2, 0, 0), (2, 1, 2), (1, 2, 2), (4, 1, 1), (4, 0, 0)])
def test_create_with_cpp_host_is_reference_readable(host, pna, ctx, oracle, comp, enc, mode):
    """create path end to end (FileEntryBuilder -> add_entry -> finalize in C++): the oracle's restatement of the reference
    reader must list and extract the same files; then our own C++ reader too (cli/tests/cli/encrypt.rs round trip)."""
    opts = pna.WriteOptions(compression=comp, encryption=enc, cipher_mode=mode, password=b"pw", kdf_params={"i": 1000})
    files = [(f"dir/f{i:03d}.bin", corpus.make_file(300 + i, n)) for i, n in enumerate([0, 1, 100, 70_000, 400_000, 33_000] * 5)]
    blob = host.create_archive(files, compression=comp, encryption=enc, cipher_mode=mode, key=opts.key, phsf=opts.phsf,
                               max_chunk_size=60_000, workers=2, group_bytes=300_000)
    got = list(oracle.extract_all(blob.tobytes(), b"pw"))
    assert got == files
    a = host.HostArchive(blob)
    if enc:
        a.set_key(opts.phsf, opts.key)
    back = a.read_all(workers=2, group_bytes=300_000)
    assert [(n, d) for n, _, d in back] == files and all(s == 0 for _, s, _ in back)


@pytest.mark.gpu
@pytest.mark.parametrize("comp,enc,mode,mcs", [(2, 0, 0, 0), (2, 1, 1, 0), (1, 2, 0, 0), (0, 0, 0, 0), (2, 0, 0, 50_000)])
def test_create_from_packed_small_files_uses_coalesced_transfers(host, pna, ctx, oracle, comp, enc, mode, mcs):
    """BASELINE config 1's shape in small: many small files that are slices of ONE buffer.  Their plaintext goes up as one copy
    per run of adjacent files (scattered on the device), their streams come down as one copy per group into the archive region
    (pna_cuda_encode_plan_fetch_region) -- the archive must be what the reference reader expects, byte for byte of content."""
    sizes = [3000, 1, 70_000, 0, 0, 12_345, 999_999, 16, 15, 17, 2_000_000, 400, 500, 600, 700, 131_072, 32_768, 32_769] + [1000 + 37 * i for i in range(60)]
    total = sum(sizes)
    plain = ctx.pinned(total)
    plain[:] = np.frombuffer(corpus.make_file(77, total), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    views = [plain[int(offs[i]):int(offs[i + 1])] for i in range(len(sizes))]
    files = [(f"p/{i:03d}", v) for i, v in enumerate(views)]
    opts = pna.WriteOptions(compression=comp, encryption=enc, cipher_mode=mode, password=b"pw", kdf_params={"i": 1000})
    out = ctx.pinned(int(total * 1.1) + (1 << 20))
    for form in (files, ([n for n, _ in files], plain, offs)):     # list of (name, view) and the packed (names, buffer, offsets) form
        blob = host.create_archive(form, compression=comp, encryption=enc, cipher_mode=mode, key=opts.key, phsf=opts.phsf,
                                   max_chunk_size=mcs, workers=2, group_bytes=1 << 20, out=out)
        got = list(oracle.extract_all(blob.tobytes(), b"pw"))
        assert [n for n, _ in got] == [n for n, _ in files]
        assert all(d == v.tobytes() for (_, d), v in zip(got, views))
    ctx.pinned_free(out)
    ctx.pinned_free(plain)


@pytest.mark.gpu
def test_extract_to_dir_and_create_from_files(host, pna, ctx, oracle, golden, tmp_path):
    """The file-system side of the path (cli/src/command/extract.rs:868-1019, core.rs:889-913): golden archives extracted to
    a directory through an mmap equal the reference's raw files; a directory packed by create_from_files is read back by the
    oracle's restatement of the reference reader and by extract_to_dir again (cli/tests/cli/encrypt.rs round trip shape).
    Small windows force several decode / write rounds."""
    for name in ("zstd.pna", "deflate.pna", "zstd_aes_ctr.pna", "solid_zstd.pna", "zstd_keep_all.pna", "zstd_camellia_gcm.pna",
                 "solid_zstd_aes_gcm.pna"):
        info = golden["archives"][name]
        a = host.HostArchive.open_file(os.path.join(golden["dir"], info["file"]))
        for phsf, key in info["keys"].items():
            a.set_key(phsf, bytes.fromhex(key))
        out = tmp_path / name
        stats, st = a.extract_to_dir(str(out), window_bytes=1 << 20, io_threads=4)
        a.close()
        files = [e for e in info["entries"]]
        assert st == [0] * len(files) and stats["files"] == len(files), name
        for e in files:
            p = out / host_sanitize(e["name"])
            d = p.read_bytes()
            assert len(d) == e["size"] and hashlib.sha256(d).hexdigest() == e["sha256"], (name, e["name"])
    # create from a directory tree, encrypted, then read it back both ways
    src = tmp_path / "src"
    want = {}
    for i, n in enumerate([0, 1, 5000, 70_000, 300_000, 1_200_000] * 3):
        rel = f"tree/d{i % 3}/f{i:02d}.bin"
        (src / rel).parent.mkdir(parents=True, exist_ok=True)
        (src / rel).write_bytes(corpus.make_file(900 + i, n))
        want[rel] = corpus.make_file(900 + i, n)
    opts = pna.WriteOptions(compression=2, encryption=1, cipher_mode=1, password=b"pw", kdf_params={"i": 1000})
    arch = tmp_path / "made.pna"
    stats = host.create_from_files([(rel, str(src / rel)) for rel in want], str(arch), compression=2, level=3, encryption=1, cipher_mode=1,
                                   key=opts.key, phsf=opts.phsf, group_bytes=1 << 20, io_threads=4)
    assert stats["files"] == len(want) and stats["bytes"] == sum(len(v) for v in want.values())
    got = dict(oracle.extract_all(arch.read_bytes(), b"pw", _keys={opts.phsf: opts.key}))
    assert got == want
    b = host.HostArchive.open_file(str(arch))
    b.set_key(opts.phsf, opts.key)
    out2 = tmp_path / "back"
    stats2, st2 = b.extract_to_dir(str(out2), window_bytes=1 << 20)
    b.close()
    assert st2 == [0] * len(want)
    for rel, data in want.items():
        assert (out2 / rel).read_bytes() == data
    # entry names are sanitised like lib/src/entry/name.rs:148 (only normal components survive)
    assert host_sanitize("../../etc//./passwd") == "etc/passwd" and host_sanitize("/abs/x") == "abs/x"


def host_sanitize(name):
    return "/".join(c for c in name.split("/") if c not in ("", ".", ".."))


@pytest.mark.gpu
@pytest.mark.parametrize("comp,enc,mode", [(2, 0, 0), (2, 1, 1), (2, 2, 2), (1, 1, 0), (0, 0, 0), (4, 0, 0), (4, 2, 1)])
def test_create_solid_with_cpp_host_is_reference_readable(host, pna, ctx, oracle, comp, enc, mode):
    """Solid create in C++ (write_solid_header -> add_entry -> finalize, archive/write.rs:438-471): the oracle's restatement of the
    reference reader (SolidEntry::entries, inner chunk CRCs included) and our own C++ reader extract the same files
    (cli/tests/cli/solid_mode.rs round trip shape)."""
    opts = pna.WriteOptions(compression=comp, encryption=enc, cipher_mode=mode, password=b"pw", kdf_params={"i": 1000})
    files = [(f"solid/f{i:03d}.bin", corpus.make_file(700 + i, n)) for i, n in enumerate([0, 1, 100, 70_000, 1_400_000, 33_000, 2_500_000])]
    blob = host.create_solid_archive(files, compression=comp, encryption=enc, cipher_mode=mode, key=opts.key, phsf=opts.phsf,
                                     max_chunk_size=32 * 1024)
    assert list(oracle.extract_all(blob.tobytes(), b"pw")) == files
    a = host.HostArchive(blob)
    if enc:
        a.set_key(opts.phsf, opts.key)
    back = a.read_all(workers=2, group_bytes=300_000)
    assert [(n, d) for n, _, d in back] == files and all(s == 0 for _, s, _ in back)


# ---------------------------------------------------------------------------------------------- split archives (C++ reader)
def _frame(ty: bytes, data: bytes = b"") -> bytes:
    import struct
    import zlib
    return struct.pack(">I", len(data)) + ty + data + struct.pack(">I", zlib.crc32(ty + data))


def _split_at_chunks(buf: np.ndarray, n_parts: int):
    """What archive/split_parts.rs produces when no chunk has to be cut: part k = signature, AHED(archive_number k), a run of
    the archive's chunks, [ANXT], AEND.  The cuts fall between ANY two chunks, so entries and their FDAT streams straddle parts."""
    raw = buf.tobytes()
    pos, frames = 8, []
    while pos < len(raw):
        ln = int.from_bytes(raw[pos:pos + 4], "big")
        frames.append(raw[pos:pos + 12 + ln])
        pos += 12 + ln
    ahed, body = frames[0], frames[1:-1]
    assert ahed[4:8] == b"AHED" and frames[-1][4:8] == b"AEND"
    per = -(-len(body) // n_parts)
    parts = []
    for k in range(n_parts):
        head = _frame(b"AHED", ahed[8:12] + k.to_bytes(4, "big"))
        tail = (_frame(b"ANXT") if k + 1 < n_parts else b"") + _frame(b"AEND")
        parts.append(np.frombuffer(raw[:8] + head + b"".join(body[k * per:(k + 1) * per]) + tail, dtype=np.uint8))
    return parts


def test_cpp_multipart_index(host, pna, golden):
    """archive/read.rs:105-165: the reference's two-part fixture is ONE entry with two FDAT bodies; part order, a missing
    part and a part behind the last one are errors.  Every chunk of every part stays in the index (CRC-checked later)."""
    p1 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part1.pna"), dtype=np.uint8)
    p2 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part2.pna"), dtype=np.uint8)
    a = host.HostArchive.open_multipart([p1, p2])
    es = a.entries()
    assert len(es) == 1 and es[0]["name"] == "multipart_test.txt" and es[0]["n_bodies"] == 2
    mod = importlib.import_module("portable-network-archive_b200.archive")
    assert a.n_chunks == len(mod.index_archive(p1, 8)) + len(mod.index_archive(p2, 8))
    for bad, kind in (([p2, p1], pna.E_INVALID_DATA), ([p1], pna.E_UNEXPECTED_EOF), ([p1, p2, p2], pna.E_INVALID_DATA),
                      ([p1, p2[:-5]], pna.E_UNEXPECTED_EOF), ([], pna.E_INVALID_INPUT)):
        with pytest.raises(host.HostError) as ei:
            host.HostArchive.open_multipart(bad)
        assert ei.value.kind == kind
    # a fixture cut into three parts between arbitrary chunks groups into the same entries as the single archive
    info = golden["archives"]["zstd_aes_ctr.pna"]
    buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
    one = host.HostArchive(buf).entries()
    three = host.HostArchive.open_multipart(_split_at_chunks(buf, 3)).entries()
    assert [(e["name"], e["compressed_size"], e["n_bodies"]) for e in one] == [(e["name"], e["compressed_size"], e["n_bodies"]) for e in three]


@pytest.mark.gpu
def test_cpp_multipart_extract(host, pna, ctx, golden):
    """extract_multipart_compatibility.rs:36 through the C++ host layer, plus golden archives re-split into 2..4 parts."""
    p1 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part1.pna"), dtype=np.uint8)
    p2 = np.fromfile(os.path.join(golden["dir"], "ref", "multipart.part2.pna"), dtype=np.uint8)
    want = open(os.path.join(golden["dir"], "ref", "multipart_test.txt"), "rb").read()
    got = host.HostArchive.open_multipart([p1, p2]).read_all()
    assert [(n, s) for n, s, _ in got] == [("multipart_test.txt", 0)] and got[0][2] == want
    for name, n_parts in (("zstd.pna", 2), ("zstd_aes_ctr.pna", 3), ("deflate.pna", 4), ("solid_zstd.pna", 2), ("zstd_camellia_cbc.pna", 3)):
        info = golden["archives"][name]
        buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
        a = host.HostArchive.open_multipart(_split_at_chunks(buf, n_parts), pinned_device=0 if n_parts != 3 else -1)
        for phsf, key in info["keys"].items():
            a.set_key(phsf, bytes.fromhex(key))
        got = a.read_all(workers=2, group_bytes=20_000)
        assert [n for n, _, _ in got] == [e["name"] for e in info["entries"]], name
        for (n, st, d), e in zip(got, info["entries"]):
            assert st == 0 and len(d) == e["size"] and hashlib.sha256(d).hexdigest() == e["sha256"], (name, n)
    # a flipped bit in a later part's archive-level chunk (its AHED) is an archive error, in an entry chunk an entry error
    buf = np.fromfile(os.path.join(golden["dir"], golden["archives"]["zstd.pna"]["file"]), dtype=np.uint8)
    parts = [p.copy() for p in _split_at_chunks(buf, 2)]
    parts[1][8 + 8 + 1] ^= 1      # minor version byte of part 2's AHED: framing intact, CRC wrong
    with pytest.raises(host.HostError) as ei:
        host.HostArchive.open_multipart(parts).extract_files()
    assert ei.value.kind == pna.E_INVALID_DATA
    parts = [p.copy() for p in _split_at_chunks(buf, 2)]
    parts[1][len(parts[1]) // 2] ^= 0x40
    out, offs, st = host.HostArchive.open_multipart(parts).extract_files()
    assert st.count(pna.E_INVALID_DATA) >= 1 and st.count(0) >= 1


# ---------------------------------------------------------------------------------------------- split writer
def _split_parts_restated(raw: bytes, max_part_bytes: int):
    """lib/src/archive/split_parts.rs:90-188 restated (SplitParts::new / put_chunk / put_stream / roll_over / finalize), CRCs by
    zlib: the checker for pna::split_archive."""
    if max_part_bytes < 64:
        raise ValueError("max_part_bytes")
    budget = max_part_bytes - 52
    parts, st = [], {"remaining": 0}

    def open_part():
        parts.append(bytearray(raw[:8] + _frame(b"AHED", bytes(4) + len(parts).to_bytes(4, "big"))))
        st["remaining"] = budget

    def roll_over():
        parts[-1] += _frame(b"ANXT") + _frame(b"AEND")
        open_part()

    def write(ty, data):
        parts[-1] += _frame(ty, data)
        st["remaining"] -= 12 + len(data)

    open_part()
    pos = 8
    first = True
    while pos < len(raw):
        ln = int.from_bytes(raw[pos:pos + 4], "big")
        ty, data = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if first:
            first = False
            continue                  # the source AHED: every part gets its own
        if ty == b"AEND":
            break
        clen = 12 + ln
        if clen <= st["remaining"]:
            write(ty, data)
        elif ty not in (b"FDAT", b"SDAT"):
            if clen > budget:
                raise ValueError("does not fit")
            roll_over()
            write(ty, data)
        elif clen <= budget and st["remaining"] <= 12:
            roll_over()
            write(ty, data)
        else:
            while True:
                if 12 + len(data) <= st["remaining"]:
                    write(ty, data)
                    break
                if st["remaining"] > 12:
                    take = st["remaining"] - 12
                    write(ty, data[:take])
                    data = data[take:]
                elif budget <= 12:
                    raise ValueError("does not fit")
                roll_over()
    parts[-1] += _frame(b"AEND")
    return [bytes(p) for p in parts]


def test_split_writer_rejects_before_any_gpu_work(host, pna, golden):
    """split_parts.rs:91-96 (below MIN_SPLIT_PART_BYTES) and :148-150 (a non-stream chunk larger than a part): InvalidInput."""
    buf = np.fromfile(os.path.join(golden["dir"], golden["archives"]["zstd.pna"]["file"]), dtype=np.uint8)
    for size in (0, 63, 64, 65):         # at 64/65 the budget (12/13 bytes) cannot hold the first FHED chunk
        with pytest.raises(host.HostError) as ei:
            host.split_archive(buf, size)
        assert ei.value.kind == pna.E_INVALID_INPUT, size
    with pytest.raises(host.HostError) as ei:
        host.split_archive(buf[:7], 1000)
    assert ei.value.kind == pna.E_INVALID_DATA


@pytest.mark.gpu
def test_split_writer_matches_restated_reference_and_reads_back(host, pna, ctx, golden):
    """Byte-exact parts against the restated SplitParts at part sizes from the reference's own tests (172 = ROLLOVER_PART_MAX,
    split_parts.rs:374) up to one larger than the archive; every split reads back through the multi-part reader."""
    for name in ("zstd.pna", "zstd_aes_ctr.pna", "solid_zstd.pna", "deflate.pna"):
        info = golden["archives"][name]
        buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
        raw = buf.tobytes()
        for size in (172, 173, 257, 1000, 4096, 50_000, len(raw) + 100):
            try:
                want = _split_parts_restated(raw, size)
            except ValueError:
                with pytest.raises(host.HostError) as ei:
                    host.split_archive(buf, size)
                assert ei.value.kind == pna.E_INVALID_INPUT
                continue
            got = host.split_archive(buf, size)
            assert [p.tobytes() for p in got] == want, (name, size)
            assert all(p.size <= size for p in got)
            if size in (172, 4096, len(raw) + 100) or name == "solid_zstd.pna":
                a = host.HostArchive.open_multipart(got)
                for phsf, key in info["keys"].items():
                    a.set_key(phsf, bytes.fromhex(key))
                res = a.read_all(workers=2, group_bytes=20_000)
                assert [n for n, _, _ in res] == [e["name"] for e in info["entries"]], (name, size)
                for (n, st, d), e in zip(res, info["entries"]):
                    assert st == 0 and hashlib.sha256(d).hexdigest() == e["sha256"], (name, size, n)


def test_split_layout_matches_restated_reference_without_gpu(host, pna, golden):
    """The budget arithmetic of SplitParts (split_parts.rs:140-188) is host logic: part count and every part length equal the
    restated writer's, on every single-archive fixture and a sweep of part sizes (sizing call: no copy, no GPU work)."""
    for name, info in golden["archives"].items():
        if name.startswith("multipart"):
            continue
        buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
        raw = buf.tobytes()
        for size in (100, 172, 173, 200, 511, 4096, 65_536, len(raw) + 52, len(raw) + 1000):
            try:
                want = [len(p) for p in _split_parts_restated(raw, size)]
            except ValueError:
                with pytest.raises(host.HostError) as ei:
                    host.split_layout(buf, size)
                assert ei.value.kind == pna.E_INVALID_INPUT, (name, size)
                continue
            lens, total = host.split_layout(buf, size)
            assert lens == want and total == sum(want), (name, size)
