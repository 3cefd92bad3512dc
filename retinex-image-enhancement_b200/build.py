"""In-tree build of libupretinex_b200.so (hand-written CUDA for sm_100a, C ABI in include/).

    python retinex-image-enhancement_b200/build.py [--force] [--verbose]

Steps: (1) compile + run the host table generator (csrc/gen_tables.cpp -> csrc/upr_tables_gen.h),
(2) nvcc every csrc/*.cu into one shared library next to this file.  nvcc cross-compiles without a
GPU; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libupretinex_b200.so")
GEN_HEADER = os.path.join(CSRC, "upr_tables_gen.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2",
    "--expt-relaxed-constexpr", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libupretinex_b200.so")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    root = os.path.dirname(PKG_DIR)
    return (_sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.cpp")) +
            glob.glob(os.path.join(root, "include", "*.h")) + [os.path.abspath(__file__)])


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def gen_tables(verbose: bool = False) -> None:
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(CSRC, "_gen_tables.bin")
    cmd = [cxx, "-O2", "-ffp-contract=off", "-o", exe, os.path.join(CSRC, "gen_tables.cpp")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    try:
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    finally:
        os.remove(exe)
    with open(GEN_HEADER, "w") as f:
        f.write(out)


def build_native(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not is_stale():
        return LIB_PATH
    gen_tables(verbose)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", LIB_PATH, *_sources()]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libupretinex_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    print(build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv or bool(flags), extra_flags=flags))
