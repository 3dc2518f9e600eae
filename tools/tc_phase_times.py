"""Phase timing of the first user tile of the tcgen05 fused kernel (clock64 ticks written by thread 0
of CTA 0 when dbg_stage == -1).  Run on the GPU box: python tools/tc_phase_times.py [decoder] [B]"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from carca_replication_b200 import fused, synth
LABELS = {0: "tile start", 1: "ids + profile embed", 2: "x store + LN1 + publish", 3: "Q MMA wait + Q lo",
          4: "K MMA wait + K store + publish", 5: "V MMA", 6: "V store + publish", 30: "  window, bits, Q -> TMEM (st_operand)", 31: "  publish (st wait + fences + CTA sync)", 32: "  next-row gather issue", 33: "  scores MMA issue (elected lane of warps 0/1)", 20: "  scores MMA wait",
          21: "  softmax pair + publish", 22: "  PV MMA (head pair)", 7: "O read + LN2 + publish", 8: "ffn_1 MMA",
          9: "LeakyReLU + publish", 10: "ffn_2 MMA", 11: "block out", 12: "final LN + publish",
          13: "dec K,V proj + store", 14: "loop top + candidate finish", 15: "(unused)", 23: "score + sigmoid + store",
          40: "decoder tables visible (CTA sync)", 41: "  row / pair decoder: ids of it+2, gather issue of it+1",
          42: "  row / pair decoder: both heads + sigmoid + store"}
decoder = sys.argv[1] if len(sys.argv) > 1 else "ca"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
VAR = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda"
shape = synth.BEAUTY
model = synth.build_model(shape, decoder).to(dev).eval()
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=1).items()}
# one context row per user, handed over as an expanded [B,T,C] view (what bench.py and the device loader do)
o_c = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
prof, tgt = (b["p_x"], None, b["p_c"]), [(b["o_x"], None, o_c)]
with torch.no_grad():
    for _ in range(3):
        fused.forward(model, prof, tgt, variant=VAR)
    dbg = torch.zeros((128, 64), device=dev)
    fused.forward(model, prof, tgt, variant=VAR, dbg=dbg, dbg_stage=-1)
torch.cuda.synchronize()
t = dbg.cpu().numpy().view(np.int64).reshape(-1, 2)
n = int(np.argmax(t[1:, 1] == 0)) + 1 if (t[1:, 1] == 0).any() else len(t)
t = t[:n]
tot = {}
for (l0, c0), (l1, c1) in zip(t[:-1], t[1:]):
    tot.setdefault(int(l1), []).append(int(c1 - c0))
total = int(t[-1, 1] - t[0, 1])
print(f"first tile: {total} cycles ({total / 1.965e3:.1f} us at 1965 MHz), {n} ticks")
for l, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
    print(f"{LABELS.get(l, l):36s} n={len(v):3d} sum={sum(v):7d} ({100 * sum(v) / total:5.1f}%) avg={sum(v) / len(v):7.0f}")
