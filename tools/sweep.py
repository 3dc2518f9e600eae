"""Scaling sweep of BASELINE.json configs[4]: maxlen 50/100/200 x batch 128..8192 users, eval users/s of the
public API on one GPU, device-timed, inputs resident: evaluate()'s per-batch body (CARCA.forward + BCE + rank
metrics) issued eagerly and replayed through GraphedEvalStep (small batches are bound by the host-side dispatch).
    python tools/sweep.py [decoder]          (per-GPU numbers; multi-GPU runs shard users: see bench.py --gpus)"""
import dataclasses, json, sys; sys.path.insert(0, '.')
import torch
import carca_replication_b200 as cb
from carca_replication_b200 import ops, synth
decoder = sys.argv[1] if len(sys.argv) > 1 else "ca"
dev = torch.device("cuda")
rows = []
for L in (50, 100, 200):
    shape = dataclasses.replace(synth.BEAUTY, seq_len=L)
    model = synth.build_model(shape, decoder, p=0.5).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
    loss_fn = cb.BinaryCrossEntropy()
    for B in (128, 512, 2048, 8192):
        bs = []
        for i in range(4):
            b = synth.make_eval_batch(shape, B, seed=10 * L + i)       # UNFILTERED users (any number of valid positions)
            bs.append({k: v.contiguous().to(dev) for k, v in b.items()})
        Bn = bs[0]["p_x"].shape[0]
        acc = torch.zeros(3, dtype=torch.float64, device=dev)
        def step(b):
            y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
            loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
            ops.rank_metrics_(acc, y, b["y_true"], 10)
        from carca_replication_b200.graph import GraphedEvalStep
        with torch.no_grad():
            fused_path = model._fused_eval_applies((bs[0]["p_x"], None, bs[0]["p_c"]), [(bs[0]["o_x"], None, bs[0]["o_c"])])
            for i in range(3): step(bs[i % 4])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20 if B <= 2048 else 8
            torch.cuda.synchronize(); e0.record()
            for i in range(n): step(bs[i % 4])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gsteps = [GraphedEvalStep(model, b, k=10, static_inputs=True) for b in bs]     # one graph per resident batch
            for g in gsteps: g.replay()
            torch.cuda.synchronize(); e0.record()
            for i in range(n): gsteps[i % 4].replay()
            e1.record(); torch.cuda.synchronize()
            ms_g = e0.elapsed_time(e1) / n
        rows.append(dict(maxlen=L, batch=Bn, ms_per_step=round(ms, 4), users_per_s=round(Bn / ms * 1e3),
                         graph_ms_per_step=round(ms_g, 4), graph_users_per_s=round(Bn / ms_g * 1e3), fused=bool(fused_path),
                         mean_valid=round(float((bs[0]["p_x"] != 0).sum(1).float().mean()), 2)))
        print(json.dumps(rows[-1]), flush=True)
