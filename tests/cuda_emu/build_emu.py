"""Builds tests/cuda_emu/libmagcore_emu.so — TEST INFRASTRUCTURE ONLY.

The product's own sources (live-video-magnification_b200/csrc/*.cu, mc_tables.cpp) compiled with g++ against the
CUDA-on-CPU emulation in tests/cuda_emu/include, so the kernels' *logic* (indexing, borders, tile staging, the
arithmetic of the device code paths) can be exercised by the parity tests in a container without a GPU, under
AddressSanitizer if wanted.  The only source transformation is the launch syntax: `k<<<cfg>>>(args)` becomes
`cuda_emu::Launcher(cfg).run("k", (k), args)`.  The result exports the same C ABI as libmagcore_b200.so
but is never loaded by the product (tests/conftest.py points lvm_b200.capi at it only when MC_EMU=1).

    python tests/cuda_emu/build_emu.py [--asan]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "live-video-magnification_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
UNITS = ["mc_core.cu", "mc_laplace.cu", "mc_motion.cu", "mc_color.cu", "mc_riesz.cu", "mc_preprocess.cu", "mc_tables.cpp"]


def lib_path(asan: bool = False) -> str:
    if asan and os.environ.get("MC_EMU_UBSAN") == "1":
        return os.path.join(HERE, "libmagcore_emu_ubsan.so")
    return os.path.join(HERE, "libmagcore_emu_asan.so" if asan else "libmagcore_emu.so")


def _match_back(s: str, i: int) -> int:
    """s[i] == '>' closing a template argument list: index of the matching '<'."""
    depth = 0
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _match_paren(s: str, i: int) -> int:
    """s[i] == '(': index of the matching ')'."""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced launch arguments")


def rewrite_launches(src: str) -> str:
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression to the left: identifier [<template args>]
        e = k
        while src[e - 1].isspace():
            e -= 1
        b = e
        if src[b - 1] == ">":
            b = _match_back(src, b - 1)
        while b > 0 and (src[b - 1].isalnum() or src[b - 1] in "_:"):
            b -= 1
        kernel = src[b:e]
        close = src.index(">>>", k)
        cfg = src[k + 3:close]
        lp = close + 3
        while src[lp].isspace():
            lp += 1
        assert src[lp] == "(", f"launch of {kernel}: expected '(' after >>>"
        rp = _match_paren(src, lp)
        args = src[lp + 1:rp]
        name = " ".join(kernel.split())   # keeps the template arguments: visible with CUDA_EMU_TRACE=1
        out.append(src[pos:b])
        out.append(f'cuda_emu::Launcher({cfg}).run("{name}", ({kernel}){", " if args.strip() else ""}{args})')
        pos = rp + 1


def build(asan: bool = False, force: bool = False) -> str:
    out = lib_path(asan)
    inc = os.path.join(HERE, "include")
    srcs = [os.path.join(CSRC, u) for u in UNITS]
    deps = srcs + [os.path.join(CSRC, h) for h in ("mc_internal.h", "mc_math.cuh", "mc_modes.h")] + \
        [os.path.join(inc, h) for h in ("cuda_runtime.h", "cuda.h", "cufft.h")] + \
        [os.path.join(HERE, "emu_runtime.cpp"), __file__, os.path.join(ROOT, "include", "magcore_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    ubsan = asan and os.environ.get("MC_EMU_UBSAN") == "1"   # MC_EMU_ASAN=1 MC_EMU_UBSAN=1: UndefinedBehaviorSanitizer instead
    gen = GEN + ("_ubsan" if ubsan else "_asan" if asan else "")
    os.makedirs(gen, exist_ok=True)
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-D__CUDACC__", "-D__CUDA_ARCH__=1000",
             "-I", inc, "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-Wno-unknown-pragmas", "-Wno-attributes"]
    if ubsan:
        flags += ["-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
    elif asan:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    jobs = []
    for u, path in zip(UNITS, srcs):
        with open(path) as f:
            text = rewrite_launches(f.read())
        cpp = os.path.join(gen, u.replace(".cu", "_cu") .replace(".cpp", "_cpp") + ".cpp")
        with open(cpp, "w") as f:
            f.write(f'#line 1 "{path}"\n' + text)
        jobs.append(cpp)
    jobs.append(os.path.join(HERE, "emu_runtime.cpp"))

    def cc(src):
        obj = os.path.join(gen, os.path.basename(src)[:-4] + ".o")
        r = subprocess.run(["g++", *flags, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-6000:])
            raise RuntimeError(f"cuda_emu: compiling {src} failed")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(cc, jobs))
    lut_o = os.path.join(gen, "lab_lut_s16.o")
    subprocess.run(["ld", "-r", "-b", "binary", "-z", "noexecstack", "-o", lut_o, "lab_lut_s16.bin"], cwd=CSRC, check=True)
    link = ["g++", "-shared", "-o", out, *objs, lut_o] + (["-fsanitize=undefined"] if ubsan else ["-fsanitize=address"] if asan else [])
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        raise RuntimeError("cuda_emu: link failed")
    return out


if __name__ == "__main__":
    print(build(asan="--asan" in sys.argv, force="--force" in sys.argv))
