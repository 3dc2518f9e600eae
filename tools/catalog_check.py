"""Development tool: tensor-core catalog kernel vs the score-matrix path at a given size.
    python tools/catalog_check.py [B] [n_items] [decoder] [L]"""
import dataclasses
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import threading
import torch  # noqa: E402

DBG = torch.zeros(8 * 4096, dtype=torch.int32).pin_memory()
if os.environ.get("WATCHDOG_S"):        # progress markers of the catalog kernel in mapped host memory (debugging hangs)
    os.environ["CARCA_CAT_DBG"] = str(DBG.data_ptr())


def watchdog():
    time.sleep(float(os.environ.get("WATCHDOG_S", "8")))
    d = DBG.view(-1, 8).numpy()
    live = d[d[:, 0] != 0]
    print("WATCHDOG: CTAs started", len(live), "finished", int((live[:, 7] != 0).sum()))
    stuck = live[live[:, 7] == 0]
    for r in stuck[:24]:
        print("   stuck CTA: groups", r[0] - 1, "alloc", r[1], "producer", r[2], "mma", r[3], "wg0", r[4], "wg1", r[5], "pre-sync", r[6])
    os._exit(3)


if os.environ.get("WATCHDOG_S"):
    threading.Thread(target=watchdog, daemon=True).start()

from carca_replication_b200 import catalog, fused, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_items = int(sys.argv[2]) if len(sys.argv) > 2 else 7001
decoder = sys.argv[3] if len(sys.argv) > 3 else "ca"
L = int(sys.argv[4]) if len(sys.argv) > 4 else 50
shape = dataclasses.replace(synth.BEAUTY, n_items=n_items, n_attrs=200, seq_len=L)
model = synth.build_model(shape, decoder, seed=4).to("cuda").eval()
model.embeds.set_attr_table(synth.make_attr_table(shape, seed=4).to("cuda"))
b = {k: v.to("cuda") for k, v in synth.make_eval_batch(shape, B, seed=4).items()}
prof = (b["p_x"], None, b["p_c"])
pos, ctx = b["o_x"][:, 0].contiguous(), b["o_c"][:, 0].contiguous()
for use_tc in (True, False):
    t0 = time.time()
    r = catalog.catalog_ranks(model, prof, pos, ctx, use_tc=use_tc)
    torch.cuda.synchronize()
    t1 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = catalog.catalog_ranks(model, prof, pos, ctx, use_tc=use_tc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    st = int(fused._plans[model].status.item())
    print(f"B={B} items={n_items} {decoder} L={L} use_tc={use_tc}: first {t1 - t0:.2f}s, {ms:.3f} ms -> "
          f"{B * (n_items - 1) / ms / 1e6:.2f} G scores/s, status {st}, mean rank {r.double().mean().item():.1f}", flush=True)
    if use_tc:
        r_tc = r
print("ranks equal:", float((r_tc == r).double().mean()))
