"""Train-step time vs batch size (launch-bound or GPU-bound?).  python tools/train_scaling.py [decoder]"""
import sys, time; sys.path.insert(0, '.')
import torch
import carca_replication_b200 as cb
from carca_replication_b200 import _native as N, synth
decoder = sys.argv[1] if len(sys.argv) > 1 else "ca"
dev = torch.device("cuda")
shape = synth.BEAUTY
table = synth.make_attr_table(shape).to(dev)
for Bt in (256, 1024, 4096):
    model = synth.build_model(shape, decoder, p=0.5).to(dev).train()
    model.embeds.set_attr_table(table)
    optim = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
    loss_fn = cb.BinaryCrossEntropy()
    L = shape.seq_len
    b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=7).items()}
    def one():
        o_x, o_c = b["o_x"], b["o_c"]
        optim.zero_grad()
        y = model.forward((b["p_x"], None, b["p_c"]), [(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = loss_fn.forward(y, b["y_true"], cb.get_mask(o_x))
        loss.backward()
        optim.step()
    for _ in range(3): one()
    torch.cuda.synchronize()
    n0 = N.lib().carca_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): one()
    e1.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={Bt}: {ms:.3f} ms/step device, host issue {t_issue * 100:.3f} ms/step, {Bt / ms * 1e3:.0f} seqs/s, "
          f"{(N.lib().carca_launch_count() - n0) // 10} launches/step of ours")
