"""The C-ABI library loads and exports every symbol include/carca_b200.h declares (no compute)."""
import ctypes
import os
import re

from carca_replication_b200 import _native as N
from carca_replication_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "carca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(carca_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = B.build()                      # nvcc cross-compiles for sm_100a without a GPU
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in carca_b200.h but not exported"
    assert lib.carca_abi_version() == 2


def test_python_binding_covers_the_header():
    declared = set(_declared_symbols()) - {"carca_last_error"}
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)


def test_no_cpu_fallback(monkeypatch):
    """CPU tensors are rejected loudly; the product never routes around the CUDA library."""
    import pytest
    import torch

    import carca_replication_b200 as cb

    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        cb.DotProduct().eval().forward(torch.zeros(1, 2, 4), torch.ones(1, 2), torch.zeros(1, 3, 4), torch.ones(1, 3))
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/libcarca_b200.so")
    monkeypatch.setattr(N, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        N.lib()


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """bind_host_to_gpu_numa_node never raises and never narrows the affinity when it cannot find the GPU's node."""
    from carca_replication_b200.parallel import bind_host_to_gpu_numa_node

    before = os.sched_getaffinity(0)
    info = bind_host_to_gpu_numa_node(0)
    assert info["gpu"] == 0 and info["bound"] is False
    assert os.sched_getaffinity(0) == before
