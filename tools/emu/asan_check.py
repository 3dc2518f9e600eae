"""DEVELOPMENT TOOL: runs the fused-training parity checks through the emulator built with -fsanitize=address
(out-of-bounds accesses of global / shared memory in the plain-CUDA kernels show up as ASan reports).

    ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tools/emu/asan_check.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_emu  # noqa: E402

from carca_replication_b200 import _native as N  # noqa: E402

N._LIB = N.bind(ctypes.CDLL(build_emu.build(asan=True)))
N.require_device = lambda *t: None
N.stream = lambda: 0
N.is_device_tensor = lambda t: True
import parity_suite as S  # noqa: E402

S.check_fused_train_vs_per_op("cpu", decoder="ca", p=0.3)
S.check_fused_train_vs_per_op("cpu", decoder="dot", p=0.25, odd_masks=True)
S.check_fused_train_vs_per_op("cpu", decoder="ca", p=0.3, B=9, all_valid=True)
S.check_fused_train_vs_per_op("cpu", decoder="ca", p=0.0, heads=4)
S.check_fused_train_edges("cpu")
S.check_fused_adam("cpu", weight_decay=0.01)
S.check_train("learnable_ca", "cpu", "csr")
print("ASAN RUN OK")
