#!/usr/bin/env python
"""build_emu.py -- TEST INFRASTRUCTURE ONLY: rewrite the CUDA sources of portable-network-archive_b200/csrc into plain
C++ over tests/emu/cuda_emu.h (kernel launches -> emu::launch, dynamic shared memory -> emu::dyn_smem) and build
tests/emu/_gen/libpna_cuda.so + libpna_host.so with g++.  Used by `pytest --emu` to debug kernel logic on a box
without a GPU; never loaded by the package itself."""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "portable-network-archive_b200", "csrc")
GEN = os.path.join(HERE, "_gen")

IDENT = re.compile(r"[A-Za-z0-9_:]")


def _match_back_template(s: str, i: int) -> int:
    """s[i] == '>' ; return index of the matching '<'."""
    depth = 0
    while i >= 0:
        c = s[i]
        if c == '>':
            depth += 1
        elif c == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template brackets before <<<")


def _match_fwd(s: str, i: int, open_c: str, close_c: str) -> int:
    depth = 0
    while i < len(s):
        c = s[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + open_c)


def _split_top(s: str):
    parts, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == ',' and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur).strip())
    return parts


def rewrite_launches(src: str) -> str:
    out = []
    pos = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # callee: identifier (with :: and template arguments) right before <<<
        j = k - 1
        while j >= 0 and src[j].isspace():
            j -= 1
        if src[j] == '>':
            j = _match_back_template(src, j) - 1
        while j >= 0 and IDENT.match(src[j]):
            j -= 1
        callee = src[j + 1:k].strip()
        # launch configuration up to the first >>> at parenthesis depth 0
        depth, e = 0, k + 3
        while True:
            c = src[e]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif depth == 0 and src.startswith(">>>", e):
                break
            e += 1
        cfg = _split_top(src[k + 3:e])
        a0 = e + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == '(', (callee, src[a0:a0 + 20])
        a1 = _match_fwd(src, a0, '(', ')')
        args = src[a0 + 1:a1]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:j + 1])
        name = callee.replace('"', "")
        out.append(f"emu::launch(\"{name}\", dim3({grid}), dim3({block}), (size_t)({smem}), [&]() {{ {callee}({args}); }})")
        pos = a1 + 1
    return "".join(out)


DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];")


def convert(text: str) -> str:
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_emu.h"')
    text = text.replace('#include "../../include/', '#include "' + os.path.join(ROOT, "include") + '/')
    text = DYN.sub(r"\1* const \2 = reinterpret_cast<\1*>(emu::dyn_smem());", text)
    return rewrite_launches(text)


def build(force: bool = False) -> str:
    os.makedirs(GEN, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".cpp", ".h")))
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in srcs)
    newest = max(newest, *(os.path.getmtime(os.path.join(HERE, f)) for f in ("cuda_emu.h", "emu_switch.S", "build_emu.py")),
                 *(os.path.getmtime(os.path.join(ROOT, "include", f)) for f in os.listdir(os.path.join(ROOT, "include"))))
    lib = os.path.join(GEN, "libpna_cuda.so")
    host = os.path.join(GEN, "libpna_host.so")
    if not force and os.path.exists(lib) and os.path.exists(host) and min(os.path.getmtime(lib), os.path.getmtime(host)) > newest:
        return GEN
    for f in srcs:
        text = open(os.path.join(CSRC, f)).read()
        name = f[:-3] + ".cpp" if f.endswith(".cu") else f
        if f.endswith((".cu", ".cuh")):
            text = convert(text)
        else:
            text = text.replace('#include "../../include/', '#include "' + os.path.join(ROOT, "include") + '/')
        with open(os.path.join(GEN, name), "w") as o:
            o.write(text)
    opt = os.environ.get("PNA_EMU_OPT", "-O2")
    subprocess.check_call(["g++", *opt.split(), "-g", "-std=c++17", "-shared", "-fPIC", "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-Wno-attributes",
                           "-I" + HERE, "-o", lib, os.path.join(GEN, "abi.cpp"), os.path.join(HERE, "emu_switch.S"), "-lpthread"])
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-o", host, os.path.join(GEN, "host_api.cpp"), "-L" + GEN,
                           "-lpna_cuda", "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return GEN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
