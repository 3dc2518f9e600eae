"""Per-op eval forward (Men shape) a few times, for an ncu launch list.  python tools/men_one.py [B] [train]"""
import sys; sys.path.insert(0, '.')
import torch
import carca_replication_b200 as cb
from carca_replication_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
train = len(sys.argv) > 2
dev = torch.device("cuda"); shape = synth.MEN
model = synth.build_model(shape, "ca", p=0.5).to(dev)
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
L = shape.seq_len
if train:
    model.train()
    b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, B, seed=7).items()}
    lf = cb.BinaryCrossEntropy()
    for _ in range(2):
        model.zero_grad()
        y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])])
        lf.forward(y, b["y_true"], cb.get_mask(b["o_x"])).backward()
else:
    model.eval()
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=7).items()}
    with torch.no_grad():
        for _ in range(2):
            y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
torch.cuda.synchronize()
print("ok", float(y.sum()))
