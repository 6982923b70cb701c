"""CPU tier: libpna_cuda.so loads and exports exactly what include/pna_cuda.h declares; no GPU -> loud failure."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(pna):
    hdr = open(os.path.join(ROOT, "include", "pna_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pna_cuda_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(os.path.join(ROOT, "portable-network-archive_b200", "libpna_cuda.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pna_cuda.h but not exported"
    from importlib import import_module
    ffi = import_module("portable-network-archive_b200._ffi")
    assert set(ffi.EXPORTS) == declared


def test_no_cpu_fallback(pna):
    import torch
    if torch.cuda.is_available():
        return
    try:
        pna.Context(0)
    except pna.PnaCudaError as e:
        assert e.code == pna.E_CUDA
    else:
        raise AssertionError("Context() must fail without a GPU")


def test_struct_layout(pna):
    from importlib import import_module
    ffi = import_module("portable-network-archive_b200._ffi")
    assert C.sizeof(ffi.Span) == 16 and C.sizeof(ffi.Buf) == 24
    assert C.sizeof(ffi.DecodeDesc) == 8 + 4 + 4 + 32 + 8 and C.sizeof(ffi.EncodeDesc) == 16 + 4 + 4 + 32 + 16 + 4 + 4 + 8   # + stream_header pointer (GCM)


def test_index_pass_on_golden(pna, golden):
    """Host index pass restates chunk framing without touching data (bytes.rs:90)."""
    import numpy as np
    mod = __import__("importlib").import_module("portable-network-archive_b200.archive")
    info = golden["archives"]["zstd_aes_ctr.pna"]
    buf = np.fromfile(os.path.join(golden["dir"], info["file"]), dtype=np.uint8)
    chunks = mod.index_archive(buf, 8)
    assert chunks[0].ty == b"AHED" and chunks[-1].ty == b"AEND"
    assert sum(1 for c in chunks if c.ty == b"FHED") == 9 == sum(1 for c in chunks if c.ty == b"FDAT")
    assert sum(1 for c in chunks if c.ty == b"PHSF") == 9
