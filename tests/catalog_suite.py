"""Full-catalog scoring checks shared by the GPU test and the CPU (emulator, gloo) test."""
import numpy as np
import torch

from helpers import FP32_RTOL, rel_err


def make_case(shape_name="tiny", decoder="ca", B=9, seed=21):
    from carca_replication_b200 import synth

    shape = synth.SHAPES[shape_name]
    model = synth.build_model(shape, decoder, p=0.5, seed=seed)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=seed)
    b = synth.make_eval_batch(shape, B, seed=seed)
    return shape, model, sd, table, b


def oracle_catalog(shape, sd, table, b, decoder, chunk=97):
    """The definition: the reference API over candidate chunks, then the position of
    the positive in a stable descending sort (src/train.py:16)."""
    from oracle import carca_oracle as O

    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder)
    B = b["p_x"].shape[0]
    n = shape.n_items
    pos, ctx = b["o_x"][:, 0], b["o_c"][:, 0]
    targets = []
    for lo in range(1, n, chunk):
        hi = min(n, lo + chunk)
        ids = torch.arange(lo, hi, dtype=torch.int32).unsqueeze(0).expand(B, hi - lo)
        targets.append((ids, table.gather_dense(ids), ctx.unsqueeze(1).expand(B, hi - lo, ctx.shape[-1])))
    with torch.no_grad():
        y = O.carca_forward(sd, cfg, (b["p_x"], table.gather_dense(b["p_x"]), b["p_c"]), targets, training=False)
    order = torch.sort(y, dim=1, descending=True, stable=True).indices
    ranks = (order == (pos.long() - 1).unsqueeze(1)).nonzero()[:, 1]
    return y.numpy(), ranks.numpy()


def check_catalog(device, shape_name="tiny", decoder="ca", B=9, fused_on=True, group=None):
    import carca_replication_b200 as cb
    from carca_replication_b200 import catalog, synth

    shape, model, sd, table, b = make_case(shape_name, decoder, B)
    y_ref, ranks_ref = oracle_catalog(shape, sd, table, b, decoder)
    model = model.to(device).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=21).to(device))
    model.use_fused_eval = fused_on
    d = {k: v.to(device) for k, v in b.items()}
    prof = (d["p_x"], None, d["p_c"])
    pos, ctx = d["o_x"][:, 0].contiguous(), d["o_c"][:, 0].contiguous()
    y = catalog.score_items(model, prof, ctx, 1, shape.n_items)
    assert tuple(y.shape) == (B, shape.n_items - 1)
    assert rel_err(y.cpu().numpy(), y_ref) < FP32_RTOL
    ranks = catalog.catalog_ranks(model, prof, pos, ctx, group=group)
    got = ranks.cpu().numpy()
    # exact unless two different items tie within fp32 noise at the positive's score
    for bi in np.nonzero(got != ranks_ref)[0]:
        yp = y_ref[bi, int(b["o_x"][bi, 0]) - 1]
        near = np.sum(np.abs(y_ref[bi] - yp) <= 2e-6 * abs(yp)) - 1
        assert abs(int(got[bi]) - int(ranks_ref[bi])) <= near, (bi, got[bi], ranks_ref[bi])
    hr, ndcg, _ = catalog.catalog_metrics(model, prof, pos, ctx, 10, group=group)
    hit = ranks_ref < 10
    assert abs(hr - hit.mean()) < 1.5 / B
    assert round(ndcg, 3) == round(float((hit / np.log2(ranks_ref + 2.0)).mean()), 3) or np.any(got != ranks_ref)
    return got, ranks_ref
