"""Checks of the device-side batch construction (csrc/batch.cuh) against the fixtures the REAL reference
loader produced (tests/golden/data_sequences.npz).  Shared by the emulator (CPU) and the GPU test."""
import os

import numpy as np
import torch

from helpers import GOLDEN


def load():
    z = np.load(os.path.join(GOLDEN, "data_sequences.npz"))
    n_items, C, L, T = [int(v) for v in z["cfg"]]
    users = sorted({int(k.split("/")[0][1:]) for k in z.files if k.startswith("u")})
    profiles = {u: z[f"u{u}/profile"] for u in users}
    ctx_rows = {u: z[f"u{u}/ctx_rows"] for u in users}
    return z, n_items, C, L, T, users, profiles, ctx_rows


def build_log(device):
    from carca_replication_b200.device_data import DeviceInteractions

    z, n_items, C, L, T, users, profiles, ctx_rows = load()
    lens = [len(profiles[u]) for u in users]
    rowptr = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32)
    items = torch.from_numpy(np.concatenate([profiles[u] for u in users]).astype(np.int32))
    ctx = torch.from_numpy(np.concatenate([ctx_rows[u] for u in users]).astype(np.float32))
    return DeviceInteractions(rowptr, items, ctx).to(device), (z, n_items, C, L, T, users, profiles)


def _check_negatives(negs, profile, n_items, need):
    negs = [int(v) for v in negs]
    assert len(negs) == need and all(1 <= v <= n_items - 1 for v in negs)
    assert len(set(negs)) == need                                  # distinct (src/data.py:84)
    assert not (set(negs) & set(int(v) for v in profile))          # never an item of the WHOLE profile (:84)


def check_batches(device):
    log, (z, n_items, C, L, T, users, profiles) = build_log(device)
    ids = torch.tensor(users, dtype=torch.int32, device=device)
    for test in (True, False):
        for mode in ("val", "test"):
            p_x, p_a, p_c, o_x, o_a, o_c, y = log.eval_batch(ids, n_items, L, T, mode, test, seed=7)
            assert p_a is None and o_a is None and tuple(o_c.shape) == (len(users), T + 1, C)
            for r, u in enumerate(users):
                key = f"u{u}/{mode}/{int(test)}"
                if key + "/p_x" not in z.files:                    # too short for this mode: zero rows
                    assert int(p_x[r].abs().sum()) == 0 and int(o_x[r].abs().sum()) == 0
                    continue
                np.testing.assert_array_equal(p_x[r].cpu().numpy(), z[key + "/p_x"])
                np.testing.assert_array_equal(p_c[r].cpu().numpy(), z[key + "/p_c"])
                assert int(o_x[r, 0]) == int(z[key + "/o_x"][0])
                np.testing.assert_array_equal(o_c[r].cpu().numpy(), z[key + "/o_c"])      # positive's context everywhere
                np.testing.assert_array_equal(y[r].cpu().numpy(), z[key + "/y_true"])
                _check_negatives(o_x[r, 1:].cpu().numpy(), profiles[u], n_items, T)
        p_x, _, p_c, o_x, _, o_c, y = log.train_batch(ids, n_items, L, test, seed=9)
        for r, u in enumerate(users):
            key = f"u{u}/train/{int(test)}"
            g = {k: z[f"{key}/{k}"] for k in ("p_x", "p_c", "o_x", "o_c", "y_true")}
            np.testing.assert_array_equal(p_x[r].cpu().numpy(), g["p_x"])
            np.testing.assert_array_equal(p_c[r].cpu().numpy(), g["p_c"])
            np.testing.assert_array_equal(o_x[r, :L].cpu().numpy(), g["o_x"][:L])          # next items
            np.testing.assert_array_equal(o_c[r].cpu().numpy(), g["o_c"])                  # negatives share it
            np.testing.assert_array_equal(y[r].cpu().numpy(), g["y_true"])
            on = g["p_x"] > 0
            neg = o_x[r, L:].cpu().numpy()
            assert np.all(neg[~on] == 0)
            if on.any():
                _check_negatives(neg[on], profiles[u], n_items, int(on.sum()))
    # same seed -> same negatives; another seed -> different ones
    a = log.eval_batch(ids, n_items, L, T, "test", True, seed=3)[3]
    b = log.eval_batch(ids, n_items, L, T, "test", True, seed=3)[3]
    c = log.eval_batch(ids, n_items, L, T, "test", True, seed=4)[3]
    assert torch.equal(a, b) and not torch.equal(a, c)


def check_negatives_are_uniform(device, draws=400):
    """Chi-square of the sampled negatives over the free items of one user."""
    log, (z, n_items, C, L, T, users, profiles) = build_log(device)
    u = users[-1]
    ids = torch.full((draws,), u, dtype=torch.int32, device=device)
    counts = np.zeros(n_items, np.int64)
    for s in range(5):     # the stream is keyed by (seed, user): vary the seed
        o_x = log.eval_batch(ids[:1].repeat(1), n_items, L, T, "test", True, seed=100 + s)[3]
        for v in o_x[0, 1:].cpu().numpy():
            counts[int(v)] += 1
    for s in range(draws):
        o_x = log.eval_batch(ids[:1], n_items, L, T, "test", True, seed=1000 + s)[3]
        np.add.at(counts, o_x[0, 1:].cpu().numpy(), 1)
    free = np.setdiff1d(np.arange(1, n_items), profiles[u])
    assert counts[profiles[u]].sum() == 0 and counts[0] == 0
    exp = counts[free].sum() / free.size
    chi2 = float(((counts[free] - exp) ** 2 / exp).sum())
    assert chi2 < free.size + 6 * np.sqrt(2 * free.size), (chi2, free.size)


def check_loader_in_evaluate(device):
    """DeviceLoader drops into evaluate() (src/train.py:35-53 signature)."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import synth
    from carca_replication_b200.device_data import DeviceLoader

    log, (z, n_items, C, L, T, users, profiles) = build_log(device)
    shape = synth.Shape("golden", len(users), n_items, 7, "multihot", C, 64, 32, 2, 2, L, T + 1)
    model = synth.build_model(shape, "ca", seed=1).to(device)
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=1).to(device))
    loader = DeviceLoader(log, n_items, L, T, "val", batch_size=4, test=True, seed=5)
    assert len(loader) == 3 and loader.users.numel() == 9          # profiles of length 1 and 2 are filtered out
    hr, ndcg, loss = cb.evaluate(model, loader, device, 10)
    assert 0.0 <= hr <= 1.0 and 0.0 <= ndcg <= 1.0 and np.isfinite(loss)
    tl = DeviceLoader(log, n_items, L, T, "train", batch_size=4, shuffle=True, seed=5)
    n = sum(b[0].shape[0] for b in tl)
    assert n == tl.users.numel() == 10


def check_packed_eval_batch(device):
    """Host -> device transfer diet: pack_eval_batch / unpack_eval_batch reproduce the dense tensors of the batch
    (left-padded windows, candidates, one context row per user, the constant label row) bit for bit."""
    import torch

    from carca_replication_b200 import synth
    from carca_replication_b200.device_data import PackedEvalLayout, pack_eval_batch, unpack_eval_batch

    shape = synth.TINY
    b = synth.make_eval_batch(shape, 9, seed=12)
    b["p_x"][3] = 0                                   # an empty window
    b["p_c"][3] = 0.0
    lay = PackedEvalLayout(9, shape.seq_len, shape.n_targets, shape.n_ctx)
    arena = torch.zeros(lay.capacity, dtype=torch.uint8)
    used = pack_eval_batch(lay, arena, b["p_x"], b["p_c"], b["o_x"], b["o_c"])
    assert used <= lay.capacity and used < lay.dense_bytes()
    dev_arena = torch.zeros(lay.capacity, dtype=torch.uint8, device=device)
    dev_arena[:used].copy_(arena[:used])
    out = unpack_eval_batch(lay, dev_arena)
    for k in ("p_x", "p_c", "o_x", "o_c", "y_true"):
        assert torch.equal(out[k].cpu(), b[k]), k
    again = unpack_eval_batch(lay, dev_arena, out)     # reuses the buffers (what a captured graph replays)
    assert again["p_x"].data_ptr() == out["p_x"].data_ptr()
    holes = b["p_x"].clone()
    holes[0, -2] = 0
    try:
        pack_eval_batch(lay, arena, holes, b["p_c"], b["o_x"], b["o_c"])
        raise AssertionError("padding inside a window must be rejected")
    except ValueError:
        pass
