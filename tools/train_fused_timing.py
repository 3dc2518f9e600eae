"""Train-step time, fused training core vs per-op kernels, eager and as a CUDA graph.
python tools/train_fused_timing.py [decoder] [batches...]"""
import sys; sys.path.insert(0, '.')
import torch
import carca_replication_b200 as cb
from carca_replication_b200 import _native as N, synth
from carca_replication_b200.graph import GraphedTrainStep
decoder = sys.argv[1] if len(sys.argv) > 1 else "ca"
batches = [int(x) for x in sys.argv[2:]] or [256, 1024, 4096]
dev = torch.device("cuda")
import dataclasses, os
shape = dataclasses.replace(synth.BEAUTY, seq_len=int(os.environ.get("CARCA_L", synth.BEAUTY.seq_len)))   # maxlen sweep
table = synth.make_attr_table(shape).to(dev)
L = shape.seq_len
for Bt in batches:
    b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=7).items()}
    for fused in (True, False):
        model = synth.build_model(shape, decoder, p=0.5).to(dev).train()
        model.embeds.set_attr_table(table)
        model.use_fused_train = fused
        optim = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), capturable=True)
        loss_fn = cb.BinaryCrossEntropy()
        def one():
            o_x, o_c = b["o_x"], b["o_c"]
            optim.zero_grad()
            y = model.forward((b["p_x"], None, b["p_c"]), [(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
            loss = loss_fn.forward(y, b["y_true"], cb.get_mask(o_x))
            loss.backward()
            optim.step()
            return loss
        for _ in range(3): one()
        torch.cuda.synchronize()
        n0 = N.lib().carca_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): loss = one()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        launches = (N.lib().carca_launch_count() - n0) // 10
        line = f"L={L} B={Bt} fused={fused} (applies={model._fused_train_applies((b['p_x'], None, b['p_c']), [(b['o_x'][:, :L], None, b['o_c'][:, :L]), (b['o_x'][:, L:], None, b['o_c'][:, L:])])}): eager {ms:.3f} ms/step ({Bt / ms * 1e3:.0f} seqs/s, {launches} launches of ours, loss {loss.item():.4f})"
        del loss          # the autograd graph it keeps alive would break the capture below
        try:
            step = GraphedTrainStep(model, optim, b, loss_fn)
            for _ in range(3):
                step(b)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20): gl = step(b)
            e1.record(); torch.cuda.synchronize()
            msg = e0.elapsed_time(e1) / 20
            line += f"; graph {msg:.3f} ms/step ({Bt / msg * 1e3:.0f} seqs/s, loss {float(gl):.4f})"
        except Exception as ex:  # noqa
            line += f"; graph failed: {type(ex).__name__}: {ex}"
        print(line, flush=True)
