"""Stage-by-stage check of the tcgen05 fused inference kernel (csrc/fused_eval_tc.cuh) against a
float64 torch restatement of the same intermediates (src/carca.py:297-318, :411-421).  Used by
tests/test_gpu_fused_tc.py and runnable as a script on the GPU box:

    python tests/tc_stages.py [shape] [decoder]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STAGE_NAMES = {1: "LN1", 2: "Q", 3: "K", 4: "V", 5: "attn+res", 9: "block out"}


def reference_stages(model, table_dense_rows, p_x, p_c, users=(0, 1)):
    """float64 intermediates for the first two users: {stage: [n_users, L, d]}."""
    sd = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
    x = p_x[list(users)].cpu().long()
    c = p_c[list(users)].cpu().double()
    a = table_dense_rows(x).double()
    d = sd["embeds.items_embed.weight"].shape[1]
    H = model.encoder[0].attn.H
    dh = d // H
    mask = (x != 0).double()
    q = F.linear(torch.cat((a, c), -1), sd["embeds.feats_embed.weight"], sd["embeds.feats_embed.bias"])
    z = sd["embeds.items_embed.weight"][x] * d ** 0.5
    e = F.linear(torch.cat((z, q), -1), sd["embeds.joint_embed.weight"], sd["embeds.joint_embed.bias"])
    pos = model.embeds.enc.table(x.shape[1]) if hasattr(model.embeds.enc, "table") else None
    if pos is not None:
        e = e + pos.detach().double().cpu()[: x.shape[1]]
    cur = e * mask.unsqueeze(2)
    out = {0: cur}
    L = x.shape[1]
    for b, blk in enumerate(model.encoder):
        pre = f"encoder.{b}."
        qn = F.layer_norm(cur, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
        Q = F.linear(qn, sd[pre + "attn.WQ.weight"], sd[pre + "attn.WQ.bias"])
        K = F.linear(cur, sd[pre + "attn.WK.weight"], sd[pre + "attn.WK.bias"])
        V = F.linear(cur, sd[pre + "attn.WV.weight"], sd[pre + "attn.WV.bias"])
        allow = (mask.unsqueeze(2) * mask.unsqueeze(1)).bool() & torch.tril(torch.ones(L, L)).bool()
        s = torch.zeros_like(Q)
        for h in range(H):
            sl = slice(h * dh, (h + 1) * dh)
            w = Q[..., sl] @ K[..., sl].transpose(1, 2) / dh ** 0.5
            w = torch.softmax(w.masked_fill(~allow, -1e30), -1) * allow
            s[..., sl] = w @ V[..., sl]
        if blk.residual:
            s = s + qn
        s2 = F.layer_norm(s, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
        f = F.leaky_relu(F.linear(s2, sd[pre + "ffn_1.weight"][:, :, 0], sd[pre + "ffn_1.bias"]), 0.01)
        f = F.linear(f, sd[pre + "ffn_2.weight"][:, :, 0], sd[pre + "ffn_2.bias"])
        if blk.residual:
            f = f + s2
        out.update({10 * b + 1: qn, 10 * b + 2: Q, 10 * b + 3: K, 10 * b + 4: V, 10 * b + 5: s, 10 * b + 9: f})
        cur = f
    out[100] = F.layer_norm(cur, (d,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    return out


def stage_errors(shape_name="tiny", decoder="ca", B=5, seed=11, all_valid=False, variant=2):
    """-> ({stage: max abs err / max abs ref}, score error vs the per-op kernels, timed_out)."""
    from carca_replication_b200 import fused, synth

    dev = "cuda"
    shape = synth.SHAPES[shape_name]
    model = synth.build_model(shape, decoder, p=0.5, seed=seed).to(dev).eval()
    table_cpu = synth.make_attr_table(shape, seed=seed)
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=seed).to(dev))
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=seed, all_valid=all_valid).items()}
    ref = reference_stages(model, table_cpu.gather_dense, b["p_x"], b["p_c"], users=tuple(range(min(2, B))))
    L = shape.seq_len
    errs = {}
    prof, tgt = (b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])]
    with torch.no_grad():
        for stage in sorted(k for k in ref if k != 0):
            dbg = torch.full((128, 64), float("nan"), device=dev)
            fused.forward(model, prof, tgt, variant=variant, dbg=dbg, dbg_stage=stage)
            got = dbg.cpu().double().reshape(2, 64, 64)[: ref[stage].shape[0], :L]
            r = ref[stage]
            # the kernel packs and computes VALID positions only (padded rows feed nothing downstream)
            valid = (b["p_x"][: r.shape[0]].cpu() != 0)
            g, rr = got[valid], r[valid]
            errs[stage] = float((g - rr).abs().max() / rr.abs().max()) if g.numel() else 0.0
            if not torch.isfinite(g).all():
                errs[stage] = float("nan")
        y_tc = fused.forward(model, prof, tgt, variant=variant)
        model.use_fused_eval = False
        y_mod = model.forward(prof, tgt)
        model.use_fused_eval = True
    torch.cuda.synchronize()
    ym, yt = y_mod.cpu().double().numpy(), y_tc.cpu().double().numpy()
    score_err = float(np.max(np.abs(yt - ym) / np.maximum(np.abs(ym), 1e-12)))
    return errs, score_err, fused.mma_timed_out(model)


if __name__ == "__main__":
    shape = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    decoder = sys.argv[2] if len(sys.argv) > 2 else "ca"
    errs, score_err, timed_out = stage_errors(shape, decoder)
    for st, e in errs.items():
        name = "final LN" if st == 100 else f"block {st // 10} {STAGE_NAMES[st % 10]}"
        print(f"stage {st:3d} {name:20s} rel err {e:.3e}")
    print(f"scores vs per-op kernels: rel err {score_err:.3e}   mma timed out: {timed_out}")
