"""Parity gate on a real B200: CUDA path (through the C ABI) vs the committed golden fixtures of
the reference and vs the oracle on seeded inputs.  Run with `pytest -m gpu`."""
import numpy as np
import pytest
import torch

import parity_suite as S
from helpers import MODEL_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _modes(name):
    return ["dense", "table"] + (["csr"] if "men" not in name else [])


@pytest.mark.parametrize("name,mode", [(n, m) for n in MODEL_CASES for m in _modes(n)])
def test_eval_vs_reference_fixture(name, mode):
    S.check_eval(name, DEV, mode)


@pytest.mark.parametrize("name,mode", [(n, m) for n in MODEL_CASES for m in _modes(n)])
def test_train_step_vs_reference_fixture(name, mode):
    S.check_train(name, DEV, mode)


@pytest.mark.parametrize("name", ["beauty_ca", "beauty_dot", "men_ca", "learnable_ca"])
@pytest.mark.parametrize("p", [0.5, 0.2])
def test_train_step_with_dropout_vs_oracle(name, p):
    S.check_train_dropout(name, DEV, p=p, mode="dense" if "men" in name else "csr")


@pytest.mark.parametrize("tag", ["self", "cross_eval", "cross_train"])
def test_mha_vs_reference_fixture(tag):
    S.check_mha(tag, DEV)


def test_metrics_and_bce_vs_reference_fixture():
    S.check_metrics(DEV)


def test_pickle_roundtrip_and_b1_shape():
    S.check_pickle_and_shapes(DEV)


def test_library_loaded_is_the_in_tree_cuda_build():
    from carca_replication_b200 import _native as N

    N.lib()
    with open("/proc/self/maps") as f:
        assert "libcarca_b200.so" in f.read()
