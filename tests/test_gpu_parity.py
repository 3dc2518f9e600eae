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


def test_knn_baseline_vs_reference_fixture():
    S.check_knn(DEV)


def test_pickle_roundtrip_and_b1_shape():
    S.check_pickle_and_shapes(DEV)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_evaluate_vs_oracle(decoder):
    S.check_evaluate_vs_oracle(DEV, decoder)


@pytest.mark.parametrize("d", [64, 32])
def test_weight_derived_caches_follow_fused_adam(d):
    S.check_weights_epoch(DEV, d=d)


def test_train_loop_writes_state_dict_checkpoint(tmp_path):
    S.check_train_loop_checkpoint(DEV, tmp_path)


def test_library_loaded_is_the_in_tree_cuda_build():
    from carca_replication_b200 import _native as N

    N.lib()
    with open("/proc/self/maps") as f:
        assert "libcarca_b200.so" in f.read()


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("B,T,all_valid", [(64, 101, False), (33, 101, True), (7, 333, False)])
def test_fused_eval_vs_per_op_kernels_beauty_shape(decoder, B, T, all_valid):
    """Full Beauty sizes (57,290 items, A=6,507 CSR, L=50, d=64, g=256, 3 blocks)."""
    S.check_fused_vs_modular(DEV, "beauty", B=B, T=T, decoder=decoder, all_valid=all_valid)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_beauty_shape_eval_vs_oracle(decoder):
    """Seeded Beauty-shaped batch: CUDA (fused and per-op) vs the CPU oracle on identical inputs,
    scores 1e-4 rel, top-10 identical up to ties, HR@10 / NDCG@10 equal to 3 decimals."""
    from carca_replication_b200 import synth
    from helpers import FP32_RTOL, rel_err, topk_equal_up_to_ties
    from oracle import carca_oracle as O
    import carca_replication_b200 as cb

    shape, B = synth.BEAUTY, 48
    model = synth.build_model(shape, decoder, p=0.5, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=5)
    b = synth.make_eval_batch(shape, B, seed=5)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder)
    dense = (b["p_x"], table.gather_dense(b["p_x"]), b["p_c"], b["o_x"], table.gather_dense(b["o_x"]), b["o_c"],
             b["y_true"])
    with torch.no_grad():
        y_ref = O.carca_forward(sd, cfg, dense[:3], [dense[3:6]], training=False)
    hits, ndcg, _, _ = O.eval_batch(sd, cfg, dense)
    model = model.to(DEV).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=5).to(DEV))
    d = {k: v.to(DEV) for k, v in b.items()}
    for fused_on in (True, False):
        model.use_fused_eval = fused_on
        with torch.no_grad():
            y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])])
        assert rel_err(y.cpu().numpy(), y_ref.numpy()) < FP32_RTOL
        assert topk_equal_up_to_ties(y.cpu().numpy(), y_ref.numpy(), 10, tol=1e-6)
        assert cb.compute_HR(y, d["y_true"], 10) == hits
        assert round(cb.compute_NDCG(y, d["y_true"], 10) / B, 3) == round(ndcg / B, 3)
    # the dense-attribute tensor of the reference API gives the same scores as the device table
    model.use_fused_eval = False
    with torch.no_grad():
        y_dense = model.forward((d["p_x"], dense[1].to(DEV), d["p_c"]), [(d["o_x"], dense[4].to(DEV), d["o_c"])])
    assert rel_err(y_dense.cpu().numpy(), y_ref.numpy()) < FP32_RTOL
