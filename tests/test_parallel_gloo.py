"""Data-parallel path over users on CPU: world_size 2, gloo, kernels' logic through tools/emu.

Checks SURVEY.md §8e: gradients after the bucket all-reduce equal the single-process gradients on
the concatenated batch (including the GLOBAL sum(mask) normaliser of the loss), and the metric
accumulators reduce to the single-process values.  The reference fixture (full batch, reference
gradients) is the ground truth on both ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _use_emulator():
    import ctypes
    import importlib.util

    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tools", "emu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from carca_replication_b200 import _native as N

    N._LIB = N.bind(ctypes.CDLL(mod.build()))
    N.require_device = lambda *t: None
    N.stream = lambda: 0
    N.is_device_tensor = lambda t: True
    N.is_emulated = lambda: True


def _worker(rank, world, port, name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _use_emulator()
        import carca_replication_b200 as cb
        from carca_replication_b200.parallel import UserDataParallel
        from helpers import batch_of, grad_err, grad_floor, load_case
        from parity_suite import build_model

        cfg, sd, z = load_case(name)
        model = build_model(cfg, sd, "cpu")
        if rank == 1:                                   # rank 0's weights must win the broadcast
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        dp = UserDataParallel(model)
        assert dp.world == world
        # ---- train step on this rank's users
        p_x, p_a, p_c, o_x, o_a, o_c, y_true = dp.shard(batch_of(z, "train"))
        L = p_x.shape[1]
        model.train()
        y = model.forward((p_x, p_a, p_c), [(o_x[:, :L], o_a[:, :L], o_c[:, :L]), (o_x[:, L:], o_a[:, L:], o_c[:, L:])])
        loss = dp.loss_fn.forward(y, y_true, cb.get_mask(o_x))
        loss.backward()
        assert abs(loss.item() - float(z["train/loss"])) < 1e-5       # the GLOBAL masked mean on every rank
        for k, prm in model.named_parameters():
            e = grad_err(prm.grad.numpy(), z["train/grad/" + k], grad_floor(k))
            assert e < 2e-4, (rank, k, e)
        # ---- evaluation: accumulators all-reduced once
        eb = batch_of(z, "eval")
        hr, ndcg, _ = dp.evaluate([eb], "cpu", cfg["k"])
        B = eb[0].shape[0]
        assert abs(hr - float(z["eval/HR"]) / B) < 1e-9
        assert abs(ndcg - float(z["eval/NDCG"]) / B) < 1e-5
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["beauty_ca", "beauty_dot"])
def test_two_rank_gradients_and_metrics_match_single_process(name):
    mp.spawn(_worker, args=(2, _free_port(), name), nprocs=2, join=True)


def _fused_worker(rank, world, port, name):
    """Device-resident CSR attributes: the fused training step produces every gradient inside one flat buffer,
    which the wrapper all-reduces in place; then one FusedAdam step must leave both ranks with the weights a
    single process gets from the reference gradients of the whole batch."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _use_emulator()
        import carca_replication_b200 as cb
        from carca_replication_b200.parallel import UserDataParallel
        from helpers import batch_of, grad_err, grad_floor, load_case
        from parity_suite import build_model

        cfg, sd, z = load_case(name)
        model = build_model(cfg, sd, "cpu")
        model.embeds.set_attr_table(cb.ItemAttrTable.from_dense(z["attr_table"], sparse=True))
        dp = UserDataParallel(model)
        optim = cb.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
        p_x, p_a, p_c, o_x, o_a, o_c, y_true = dp.shard(batch_of(z, "train"))
        L = p_x.shape[1]
        model.train()
        reduced = []
        orig = dp._shared_grad_buffer
        dp._shared_grad_buffer = lambda: reduced.append(orig()) or reduced[-1]
        y = model.forward((p_x, None, p_c), [(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = dp.loss_fn.forward(y, y_true, cb.get_mask(o_x))
        loss.backward()
        assert len(reduced) == 1 and reduced[0] is not None            # one in-place all-reduce, no bucket copies
        assert abs(loss.item() - float(z["train/loss"])) < 1e-5
        for k, prm in model.named_parameters():
            e = grad_err(prm.grad.numpy(), z["train/grad/" + k], grad_floor(k))
            assert e < 2e-4, (rank, k, e)
        before = {k: v.detach().clone() for k, v in model.named_parameters()}
        optim.step()
        ref = {k: v.detach().clone().requires_grad_(True) for k, v in before.items()}
        for k, v in ref.items():
            v.grad = torch.from_numpy(z["train/grad/" + k].copy())
        torch.optim.Adam(list(ref.values()), lr=1e-3, betas=(0.9, 0.98)).step()
        for k, prm in model.named_parameters():
            # first Adam step moves every weight by ~lr * sign(grad): compare the moves
            mv, mr = (prm.detach() - before[k]).numpy(), (ref[k].detach() - before[k]).numpy()
            big = np.abs(z["train/grad/" + k]) > 1e-6
            assert np.allclose(mv[big], mr[big], rtol=0, atol=2e-4), (rank, k)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["beauty_ca", "sinus_dot"])
def test_two_rank_fused_step_flat_allreduce_and_fused_adam(name):
    mp.spawn(_fused_worker, args=(2, _free_port(), name), nprocs=2, join=True)


def _catalog_worker(rank, world, port, decoder):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _use_emulator()
        import catalog_suite as S
        from carca_replication_b200 import catalog

        assert catalog.shard_bounds(300, 0, 2) == (1, 150) and catalog.shard_bounds(300, 1, 2) == (150, 300)
        got, ref = S.check_catalog("cpu", "tiny", decoder, B=6)      # item table sharded over the 2 ranks
        assert got.shape == ref.shape
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_two_rank_item_sharded_full_catalog_ranks_match_oracle(decoder):
    mp.spawn(_catalog_worker, args=(2, _free_port(), decoder), nprocs=2, join=True)
