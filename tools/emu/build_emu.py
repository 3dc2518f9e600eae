"""DEVELOPMENT TOOL: builds the plain-CUDA kernels as host C++ against tools/emu/cuda_emu.h.

    python tools/emu/build_emu.py [--asan]

Output: tools/emu/_build/libcarca_emu.so (git-ignored, gpurun-ignored).  It is loaded only by
tests/test_emu_*.py, which swap it in by monkeypatching; the product loader never looks for it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "carca_replication_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["api.cu", "gemm.cu", "umma_selftest.cu", "rows.cu", "peer.cu"]


def build(asan: bool = False, force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libcarca_emu_asan.so" if asan else "libcarca_emu.so")
    newest = max(os.path.getmtime(os.path.join(r, f))
                 for r in (CSRC, os.path.join(ROOT, "include"), HERE)
                 for f in os.listdir(r) if f.endswith((".cu", ".cuh", ".h")))
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest:
        return lib
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-x", "c++", "-DCARCA_EMU=1",
           "-include", os.path.join(HERE, "cuda_emu.h"), "-Wno-unknown-pragmas", "-Wno-attributes"]
    if asan:
        cmd += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", lib]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(asan="--asan" in sys.argv, force="--force" in sys.argv))
