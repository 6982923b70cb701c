"""Property round trips mirroring the reference's fuzz targets (fuzz/fuzz_targets/{aes,camellia}_{ctr,cbc,gcm}.rs: arbitrary bytes
-> FileEntryBuilder -> NormalEntry::reader == the bytes) and lib/tests/copy_entries.rs (byte-exact re-serialisation).  Every case
runs through the GPU: builder -> pna_cuda_encode_batch, reader -> pna_cuda_decode_batch; the reference pipeline (oracle) must read
what the builder wrote."""
import importlib
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

# byte strings with the shapes codecs care about: empty, tiny, block-size neighbours, runs, repeats of a short period
payloads = st.one_of(
    st.binary(min_size=0, max_size=300),
    st.builds(lambda b, n: b * n, st.binary(min_size=1, max_size=40), st.integers(1, 3000)),
    st.builds(lambda n, seed: np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes(),
              st.sampled_from([15, 16, 17, 31, 32, 33, 4095, 4096, 4097, 65535, 65536, 131071, 131072, 131073, 300001]), st.integers(0, 2**31)),
    st.builds(lambda a, b, n: (a + bytes(n) + b) * 3, st.binary(max_size=50), st.binary(max_size=50), st.integers(0, 70000)),
)


@pytest.fixture(scope="module")
def mod(pna):
    return importlib.import_module("portable-network-archive_b200.archive")


@pytest.mark.parametrize("encryption,mode", [(1, 1), (1, 0), (2, 1), (2, 0), (1, 2), (2, 2)])
@pytest.mark.parametrize("compression", [0, 2, 1, 4])
@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(data=payloads)
def test_builder_reader_round_trip(ctx, mod, oracle, encryption, mode, compression, data):
    """fuzz_targets/aes_ctr.rs:7-21 and its five siblings, with every codec in front of the cipher"""
    wo = mod.WriteOptions(compression=compression, encryption=encryption, cipher_mode=mode, password=b"password", kdf_params={"i": 1000})
    b = mod.FileEntryBuilder.new_with_options("fuzz", wo)
    b.write(data)
    built = b.build(ctx)
    a = mod.Archive.write_header(ctx)
    a.add_entry(built)
    blob = a.finalize()
    arch = mod.Archive.read_header(np.frombuffer(blob, dtype=np.uint8), ctx)
    ro = mod.ReadOptions.with_password(b"password")
    (entry,) = list(arch.entries())
    assert entry.reader(ro, ctx) == data
    # and the reference pipeline reads it too
    got = list(oracle.extract_all(blob, b"password"))
    assert [d for _, d in got] == [data]


def test_copy_entries_is_byte_exact(ctx, mod, golden):
    """lib/tests/copy_entries.rs:15-21: read every entry of deflate.pna, add it to a new archive, the bytes are the same
    (chunk framing, wire order and every recomputed CRC)."""
    for name in ("deflate.pna", "zstd_aes_ctr.pna", "solid_zstd.pna", "zstd_keep_all.pna" if "zstd_keep_all.pna" in golden["archives"] else "zstd.pna"):
        src = np.fromfile(os.path.join(golden["dir"], golden["archives"][name]["file"]), dtype=np.uint8)
        reader = mod.Archive.read_header(src, ctx)
        writer = mod.Archive.write_header(ctx)
        for e in reader.entries():
            writer.add_entry(e)
        assert bytes(writer.finalize()) == src.tobytes(), name
