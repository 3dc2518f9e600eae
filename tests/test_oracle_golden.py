"""Pins oracle/carca_oracle.py to the fixtures the real reference produced (CPU, no GPU)."""
import json

import numpy as np
import pytest
import torch

from helpers import GOLDEN, MODEL_CASES, batch_of, grad_err, load_case, oracle_cfg, rel_err
from oracle import carca_oracle as O


@pytest.mark.parametrize("name", MODEL_CASES)
def test_eval_matches_reference(name):
    cfg, sd, z = load_case(name)
    batch = batch_of(z, "eval")
    oc = oracle_cfg(cfg)
    with torch.no_grad():
        y = O.carca_forward(sd, oc, batch[:3], [batch[3:6]], training=False)
    assert y.shape == tuple(z["eval/y_pred"].shape)
    assert rel_err(y.numpy(), z["eval/y_pred"]) < 2e-6
    hits, ndcg, loss, n = O.eval_batch(sd, oc, batch, k=cfg["k"])
    assert hits == float(z["eval/HR"])
    assert abs(ndcg - float(z["eval/NDCG"])) < 1e-5
    assert abs(loss - float(z["eval/loss"])) < 1e-5 * max(1.0, abs(loss))
    assert n == batch[0].shape[0]


@pytest.mark.parametrize("name", MODEL_CASES)
def test_train_forward_backward_matches_reference(name):
    cfg, sd, z = load_case(name)
    batch = batch_of(z, "train")
    oc = oracle_cfg(cfg)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "embeds.enc.pe") for k, v in sd.items()}
    y = O.carca_forward(sd, oc, batch[:3], O.train_step_targets(*batch[3:6]), training=True)
    assert rel_err(y.detach().numpy(), z["train/y_pred"]) < 2e-6
    loss = O.masked_bce(y, batch[6], O.padding_mask(batch[3]))
    assert abs(loss.item() - float(z["train/loss"])) < 1e-5
    loss.backward()
    for k in [f[len("train/grad/"):] for f in z.files if f.startswith("train/grad/")]:
        g = sd[k].grad
        g = torch.zeros_like(sd[k]) if g is None else g
        assert grad_err(g.numpy(), z["train/grad/" + k]) < 1e-5, k


@pytest.mark.parametrize("tag", ["self", "cross_eval", "cross_train"])
def test_mha_matches_reference(tag):
    z = np.load(f"{GOLDEN}/mha_ops.npz")
    c = json.loads(str(z[f"{tag}/cfg"]))
    sd = {"m." + k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}/sd/")}
    q, kv = torch.from_numpy(z[f"{tag}/q"]), torch.from_numpy(z[f"{tag}/kv"])
    qm, km = torch.from_numpy(z[f"{tag}/q_mask"]), torch.from_numpy(z[f"{tag}/k_mask"])
    w, o = O.multi_head_attention(sd, "m.", q, kv, kv, qm, km, c["H"], c["causal"], O.Dropper(0.0, 0, False), 0)
    np.testing.assert_allclose(w.numpy(), z[f"{tag}/w"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(o.numpy(), z[f"{tag}/out"], rtol=1e-5, atol=1e-6)
    # dead query rows and dead key rows give exactly zero
    assert np.all(o.numpy()[0] == 0.0) and np.all(o.numpy()[1] == 0.0)


def test_metrics_and_bce_match_reference():
    z = np.load(f"{GOLDEN}/metrics_ops.npz")
    y, yt, m, k = (torch.from_numpy(z["y_pred"]), torch.from_numpy(z["y_true"]), torch.from_numpy(z["mask"]),
                   int(z["k"]))
    assert O.hit_count(y, yt, k) == float(z["HR"])
    assert abs(O.ndcg_sum(y, yt, k) - float(z["NDCG"])) < 1e-5
    assert "HR_stable" in z.files and float(z["HR_stable"]) >= float(z["HR"])
    yv = y.clone().requires_grad_(True)
    loss = O.masked_bce(yv, yt, m)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-6
    np.testing.assert_allclose(yv.grad.numpy(), z["dy"], rtol=1e-6, atol=1e-9)


def test_philox_known_answer_and_rate():
    """Random123 known-answer vector for Philox4x32-10, plus keep-rate / determinism."""
    from oracle import philox

    out = philox.philox4x32_10(np.array([0xFFFFFFFF], np.uint32), 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF,
                               0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(o[0]) for o in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = philox.philox4x32_10(np.array([0], np.uint32), 0, 0, 0, 0, 0)
    assert [int(o[0]) for o in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    k1 = philox.keep_mask(200_000, 0.5, seed=42, site=3)
    k2 = philox.keep_mask(200_000, 0.5, seed=42, site=3)
    assert np.array_equal(k1, k2)
    assert abs(k1.mean() - 0.5) < 0.01
    assert abs(philox.keep_mask(200_000, 0.2, seed=7, site=0).mean() - 0.8) < 0.01
    assert not np.array_equal(k1, philox.keep_mask(200_000, 0.5, seed=42, site=4))
