"""Phase timing of the first bins of CTA 0 of the fused training kernels (clock64 ticks of thread 0).
Run on the GPU box: python tools/train_phase_times.py [decoder] [B]"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import carca_replication_b200 as cb
from carca_replication_b200 import _native as N, synth
FWD = {0: "bin start", 1: "load_bin", 2: "profile embed + dropout", 3: "save x + LN1", 4: "Q,K,V projections", 5: "save Q,K,V",
       6: "self-attention", 7: "save s + LN2", 8: "FFN (2 proj, passes, save a1)", 9: "save x + final LN",
       10: "decoder K,V proj + save", 11: "target embed", 12: "target Q proj", 13: "cross-attention", 14: "save s + scores"}
BWD = {100: "bin start", 101: "(loop top)", 102: "load s_t, g, d wf, ds", 103: "target embed (recompute)", 104: "target Q proj",
       105: "cross-attention bwd", 106: "d WQ", 107: "d o (proj bwd)", 108: "target embed bwd",
       109: "dec K,V grads + final LN bwd", 110: "(block top)", 111: "FFN bwd", 112: "LN2 bwd", 113: "load Q,K,V",
       114: "self-attention bwd", 115: "load x, LN1, d WQ/WK/WV", 116: "d qn, d x proj bwd, LN1 bwd",
       117: "copy + dropout bwd", 118: "profile embed bwd"}
decoder = sys.argv[1] if len(sys.argv) > 1 else "ca"
Bt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda"); shape = synth.BEAUTY; L = shape.seq_len
model = synth.build_model(shape, decoder, p=0.5).to(dev).train()
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
loss_fn = cb.BinaryCrossEntropy()
b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=7).items()}
def one():
    model.zero_grad()
    y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])])
    loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"])).backward()
for _ in range(3): one()
buf = torch.zeros((4096, 2), dtype=torch.int64, device=dev)
for name, labels in (("forward", FWD), ("backward", BWD)):
    pass
N.lib().carca_train_core_set_ticks(buf.data_ptr(), 2000)
one(); torch.cuda.synchronize()
N.lib().carca_train_core_set_ticks(None, 0)
t = buf.cpu().numpy()
# the forward and the backward kernel both start writing at slot 0: run them separately
def report(labels, first):
    buf.zero_()
    N.lib().carca_train_core_set_ticks(buf.data_ptr(), 2000)
    if first == 0:
        with torch.no_grad():
            pass
    model.zero_grad()
    y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])])
    torch.cuda.synchronize()
    if first == 0:
        t = buf.cpu().numpy().copy()
        N.lib().carca_train_core_set_ticks(None, 0)
        loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"])).backward()
    else:
        buf.zero_()
        loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"])).backward()
        torch.cuda.synchronize()
        t = buf.cpu().numpy().copy()
        N.lib().carca_train_core_set_ticks(None, 0)
    n = int(np.argmax(t[1:, 1] == 0)) + 1 if (t[1:, 1] == 0).any() else len(t)
    t = t[:n]
    bins = int((t[:, 0] == first).sum())
    tot = {}
    for (l0, c0), (l1, c1) in zip(t[:-1], t[1:]):
        tot.setdefault(int(l1), []).append(int(c1 - c0))
    total = int(t[-1, 1] - t[0, 1])
    print(f"{'forward' if first == 0 else 'backward'}: CTA 0 processed {bins} bin(s) in {total} cycles ({total / 1.965e3:.1f} us at 1965 MHz)")
    for l, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {labels.get(l, l):36s} n={len(v):3d} sum={sum(v):8d} ({100 * sum(v) / total:5.1f}%) avg={sum(v) / len(v):8.0f}")
report(FWD, 0)
report(BWD, 100)
