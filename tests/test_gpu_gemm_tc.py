"""tcgen05 GEMM of the training path (csrc/gemm_tc.cuh) on a B200: every storage case against float64
matmuls through the C ABI, and a whole train step (gathers, dropout epilogues, split-K weight gradients)
against the CPU oracle at a size where every projection takes the tensor-core kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _err(got, ref, scale):
    return float((got.double().cpu() - ref).abs().max() / scale)


@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (777, 200, 96), (4096, 64, 70), (130, 256, 512), (5000, 33, 129)])
def test_linear_forward_backward_vs_float64(M, N, K):
    from carca_replication_b200 import _native as N_

    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xd, wd, bd, dyd = (t.cuda() for t in (x, w, b, dy))
    st = N_.stream()
    y = torch.empty(M, N, device="cuda")
    N_.call("carca_linear_fwd", N_.f32p(y), N_.f32p(xd), N_.f32p(wd), N_.f32p(bd), M, N, K, 1, st)
    ref = torch.nn.functional.leaky_relu(x.double() @ w.double().T + b.double(), 0.01)
    scale = float((x.double().abs() @ w.double().abs().T).max())
    assert _err(y, ref, scale) < 2e-6
    dx = torch.empty(M, K, device="cuda")
    N_.call("carca_linear_bwd_input", N_.f32p(dx), N_.f32p(dyd), N_.f32p(wd), M, N, K, 0, st)
    assert _err(dx, dy.double() @ w.double(), float((dy.double().abs() @ w.double().abs()).max())) < 2e-6
    dw = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    N_.call("carca_linear_bwd_weight", N_.f32p(dw), N_.f32p(db), N_.f32p(dyd), N_.f32p(xd), M, N, K, st)
    assert _err(dw, dy.double().T @ x.double(), float((dy.double().abs().T @ x.double().abs()).max())) < 2e-6
    assert _err(db, dy.double().sum(0), float(dy.double().abs().sum(0).max())) < 2e-6


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_train_step_on_tensor_core_gemms_vs_oracle(decoder):
    import carca_replication_b200 as cb
    from carca_replication_b200 import _native as N_
    from carca_replication_b200 import ops, synth
    from oracle import carca_oracle as O

    dev = "cuda"
    shape, B, p = synth.TINY, 160, 0.25          # 1920 positions: every [P,64]x[64,64] product is above the TC threshold
    model = synth.build_model(shape, decoder, p=p, seed=13)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=13)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder, p_drop=p, seed=77)
    bt = synth.make_train_batch(shape, B, seed=14)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref = O.train_batch_loss(sdo, cfg, (bt["p_x"], table.gather_dense(bt["p_x"]), bt["p_c"], bt["o_x"],
                                             table.gather_dense(bt["o_x"]), bt["o_c"], bt["y_true"]))
    loss_ref.backward()
    model = model.to(dev).train()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=13).to(dev))
    ops.set_dropout_seed(77)
    try:
        L = shape.seq_len
        o_x, o_c = bt["o_x"].to(dev), bt["o_c"].to(dev)
        n0 = N_.lib().carca_launch_count()
        y = model.forward((bt["p_x"].to(dev), None, bt["p_c"].to(dev)),
                          [(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = cb.BinaryCrossEntropy().forward(y, bt["y_true"].to(dev), cb.get_mask(o_x))
        loss.backward()
    finally:
        ops.set_dropout_seed(None)
    assert N_.lib().carca_launch_count() > n0
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    for k, prm in model.named_parameters():
        g_ref = sdo[k].grad.numpy()
        floor = 1e-3 if k.endswith("WK.bias") else 1e-5
        err = np.abs(prm.grad.cpu().numpy() - g_ref).max() / max(np.abs(g_ref).max(), floor)
        assert err < 3e-4, (k, err)
