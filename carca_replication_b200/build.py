"""Ahead-of-time build of libcarca_b200.so (sm_100a only) with nvcc.

    python -m carca_replication_b200.build          # or __graft_entry__.build()

The library is built in-tree (carca_replication_b200/libcarca_b200.so) so that it travels with
the repo snapshot to the GPU box; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcarca_b200.so")
SOURCES = ["api.cu", "gemm.cu", "umma_selftest.cu", "rows.cu", "peer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
]


def _newest_source_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, f)))
    return newest


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds the library if it is missing or older than a source.  Safe under `torchrun` (several ranks calling
    it at once): one process builds under a file lock into temporary names and renames into place."""
    import fcntl
    import tempfile

    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
                return LIB                     # another rank built it while this one waited
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = tempfile.mkdtemp(prefix="carca_build_", dir=CSRC)
            try:
                objs, procs = [], []
                for src in SOURCES:                      # the translation units compile side by side
                    obj = os.path.join(tmp, src.replace(".cu", ".o"))
                    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
                    if verbose:
                        cmd.insert(1, "-Xptxas=-v")
                        print(" ".join(cmd), flush=True)
                        subprocess.run(cmd, check=True)  # (one at a time: readable ptxas output)
                    else:
                        procs.append((cmd, subprocess.Popen(cmd)))
                    objs.append(obj)
                for cmd, pr in procs:
                    if pr.wait() != 0:
                        raise subprocess.CalledProcessError(pr.returncode, cmd)
                out = os.path.join(tmp, "libcarca_b200.so")
                cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs, "-Xcompiler",
                       "-fPIC", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
                subprocess.run(cmd, check=True)
                os.replace(out, LIB)
            finally:
                import shutil

                shutil.rmtree(tmp, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
