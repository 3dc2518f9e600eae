"""Dense attribute rows of a large multi-hot vocabulary (the reference API's [B, N, A] tensors, src/carca.py:86):
the scan + gather-sum feature projection (csrc/embed.cuh: feat_dense_scan_fwd_kernel, A >= 1024) against the attribute
TABLE path (CSR gather-sum, the oracle-pinned one) on the same model and batch — forward scores and, because the kernel
is also the forward of the training path, one train step's loss and gradients."""
import dataclasses

import pytest
import torch


def _run(dev, B, n_items):
    import carca_replication_b200 as cb
    from carca_replication_b200 import synth

    shape = dataclasses.replace(synth.TINY, n_items=n_items, n_attrs=1100)
    table_cpu = synth.make_attr_table(shape, seed=4)
    model = synth.build_model(shape, "ca", p=0.0, seed=4).to(dev)
    model.use_fused_eval = False                  # both sides on the per-op kernels: only the projection differs
    model.use_fused_train = False
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=4).to(dev))
    b = synth.make_eval_batch(shape, B, seed=4)
    p_a, o_a = table_cpu.gather_dense(b["p_x"]).to(dev), table_cpu.gather_dense(b["o_x"]).to(dev)
    b = {k: v.to(dev) for k, v in b.items()}
    model.eval()
    with torch.no_grad():
        y_tab = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
        y_dense = model.forward((b["p_x"], p_a, b["p_c"]), [(b["o_x"], o_a, b["o_c"])])
    assert float((y_dense - y_tab).abs().max()) < 2e-6
    # training forward / backward with dense rows (forward through the scan kernel, weight gradient through the GEMM)
    model.train()
    grads = []
    for attrs in ((None, None), (p_a, o_a)):
        model.zero_grad()
        y = model.forward((b["p_x"], attrs[0], b["p_c"]), [(b["o_x"], attrs[1], b["o_c"])])
        loss = cb.BinaryCrossEntropy().forward(y, b["y_true"], cb.get_mask(b["o_x"]))
        loss.backward()
        grads.append((float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert grads[0][0] == pytest.approx(grads[1][0], rel=1e-5)
    for k in grads[0][1]:
        a, c = grads[0][1][k], grads[1][1][k]
        assert float((a - c).abs().max()) <= 2e-4 * max(1e-3, float(a.abs().max())), k


def test_dense_scan_projection_matches_the_table_path_on_the_emulator(emu_backend):
    _run(emu_backend, B=3, n_items=60)


@pytest.mark.gpu
def test_dense_scan_projection_matches_the_table_path():
    _run("cuda", B=40, n_items=900)
