"""Peer-memory all-reduce (csrc/peer.cu, parallel.PeerAllReduce) against NCCL on the same tensors — needs two GPUs of
one NVLink box, so it is skipped on a single-GPU machine (the 2 / 4 / 8-GPU runs are recorded in profiles/r02/)."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_equals_nccl_on_two_gpus():
    env = dict(os.environ, PEER_QUICK="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "peer_check.py")],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    text = out.stdout + out.stderr
    assert "available True" in text, text[-2000:]
    errs = [float(x) for x in re.findall(r"max \|peer - nccl\| ([0-9.e+-]+)", text)]
    assert len(errs) >= 10 and max(errs) == 0.0, text[-2000:]       # two ranks: a + b is the same sum in any order
    assert "ranks bit-equal False" not in text and "timed out True" not in text, text[-2000:]
