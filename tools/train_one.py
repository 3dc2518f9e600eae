"""Three eager train steps at the reference batch size (for an ncu launch list).  python tools/train_one.py [B]"""
import sys; sys.path.insert(0, '.')
import torch
import carca_replication_b200 as cb
from carca_replication_b200 import synth
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda"); shape = synth.BEAUTY
model = synth.build_model(shape, "ca", p=0.5).to(dev).train()
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
optim = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
loss_fn = cb.BinaryCrossEntropy(); L = shape.seq_len
b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=7).items()}
for _ in range(3):
    optim.zero_grad()
    y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])])
    loss = loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
    loss.backward()
    optim.step()
torch.cuda.synchronize()
print("loss", loss.item())
