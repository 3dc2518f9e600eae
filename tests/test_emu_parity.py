"""CPU-side check of the kernels' logic and the Python/ctypes glue through the development
emulator (tools/emu).  Small cases only; the real parity gate is tests/test_gpu_parity.py."""
import pytest

import parity_suite as S


@pytest.mark.parametrize("name,mode", [("beauty_ca", "dense"), ("beauty_dot", "csr"), ("men_ca", "table"),
                                       ("learnable_ca", "csr"), ("noresid_ca", "dense")])
def test_eval(emu_backend, name, mode):
    S.check_eval(name, emu_backend, mode)


@pytest.mark.parametrize("name,mode", [("beauty_ca", "csr"), ("beauty_dot", "dense"), ("men_ca", "table"),
                                       ("sinus_dot", "csr"), ("learnable_ca", "dense"), ("noresid_dot", "csr")])
def test_train(emu_backend, name, mode):
    S.check_train(name, emu_backend, mode)


@pytest.mark.parametrize("name", ["beauty_ca", "beauty_dot"])
def test_train_dropout(emu_backend, name):
    S.check_train_dropout(name, emu_backend, mode="csr")


@pytest.mark.parametrize("tag", ["self", "cross_eval", "cross_train"])
def test_mha(emu_backend, tag):
    S.check_mha(tag, emu_backend)


def test_metrics(emu_backend):
    S.check_metrics(emu_backend)


def test_pickle_and_shapes(emu_backend):
    S.check_pickle_and_shapes(emu_backend)


@pytest.mark.parametrize("decoder,B,T", [("ca", 3, 230), ("dot", 2, 21), ("ca", 1, 105)])
def test_fused_eval_vs_per_op_kernels(emu_backend, decoder, B, T):
    S.check_fused_vs_modular(emu_backend, "tiny", B=B, T=T, decoder=decoder)


def test_fused_eval_single_user_fixture(emu_backend):
    S.check_eval("single_user_ca", emu_backend, "csr")


VARIANTS = [("idemb_dot", "dense"), ("mlpid_ca", "dense"), ("attr_dot", "csr"), ("attrctx_ca", "dense"),
            ("attrctx_dense_ca", "table"), ("wdot_all", "csr"), ("wdotnorm_all", "dense")]


@pytest.mark.parametrize("name,mode", VARIANTS)
def test_module_variants_eval(emu_backend, name, mode):
    S.check_eval(name, emu_backend, mode)


@pytest.mark.parametrize("name,mode", VARIANTS)
def test_module_variants_train(emu_backend, name, mode):
    S.check_train(name, emu_backend, mode)


def test_knn(emu_backend):
    S.check_knn(emu_backend)


@pytest.mark.parametrize("kw", [dict(decoder="ca", p=0.3), dict(decoder="dot", p=0.3), dict(decoder="ca", p=0.0, heads=4),
                                dict(decoder="ca", p=0.25, heads=1, odd_masks=True),
                                dict(decoder="dot", p=0.25, odd_masks=True), dict(decoder="ca", p=0.5, n_tuples=1),
                                dict(decoder="ca", p=0.3, B=9, all_valid=True)])
def test_fused_train_vs_per_op_kernels(emu_backend, kw):
    S.check_fused_train_vs_per_op(emu_backend, **kw)


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_vs_torch(emu_backend, wd):
    S.check_fused_adam(emu_backend, weight_decay=wd)


def test_fused_train_edge_shapes(emu_backend):
    S.check_fused_train_edges(emu_backend)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_evaluate_vs_oracle(emu_backend, decoder):
    S.check_evaluate_vs_oracle(emu_backend, decoder)


@pytest.mark.parametrize("d", [64, 32])
def test_weight_derived_caches_follow_fused_adam(emu_backend, d):
    S.check_weights_epoch(emu_backend, d=d)


def test_train_loop_writes_state_dict_checkpoint(emu_backend, tmp_path):
    S.check_train_loop_checkpoint(emu_backend, tmp_path)
