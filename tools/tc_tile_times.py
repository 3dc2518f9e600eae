"""Tile start clocks of one CTA of the tcgen05 fused kernel (dbg_stage = -2 - cta): how many tiles the CTA processed
and how long each took.  Run on the GPU box: python tools/tc_tile_times.py [B] [variant]"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from carca_replication_b200 import fused, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
VAR = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
shape = synth.BEAUTY
model = synth.build_model(shape, "ca").to(dev).eval()
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=1).items()}
o_c = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
prof, tgt = (b["p_x"], None, b["p_c"]), [(b["o_x"], None, o_c)]
with torch.no_grad():
    for _ in range(3):
        fused.forward(model, prof, tgt, variant=VAR)
    for cta in (0, 73, 147):
        dbg = torch.zeros((128, 64), device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fused.forward(model, prof, tgt, variant=VAR, dbg=dbg, dbg_stage=-2 - cta)
        e1.record()
        torch.cuda.synchronize()
        t = dbg.cpu().numpy().view(np.int64).reshape(-1, 2)
        n = int(np.argmax(t[1:, 1] == 0)) + 1 if (t[1:, 1] == 0).any() else len(t)
        c = t[:n, 1]
        d = np.diff(c)
        print(f"CTA {cta}: {n - 1} tiles, total {int(c[-1] - c[0])} cycles ({(c[-1] - c[0]) / 1.965e3:.1f} us), call {e0.elapsed_time(e1) * 1e3:.0f} us; "
              f"tile cycles min {d.min()} mean {d.mean():.0f} max {d.max()}: {d.tolist()}")
