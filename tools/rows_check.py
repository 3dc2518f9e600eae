"""Development tool: bf16 packed-rows pipeline (and the fp32 paths) against the CPU oracle on seeded cases.
    python tools/rows_check.py [case ...]          (needs a B200)"""
import dataclasses
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import carca_replication_b200 as cb  # noqa: E402
from carca_replication_b200 import fused, synth  # noqa: E402
from oracle import carca_oracle as O  # noqa: E402

MEN_S = dataclasses.replace(synth.MEN, name="men_small", n_items=3000, n_attrs=96)
CASES = {
    "tiny_ca": (synth.TINY, "ca", 9, False),
    "tiny_dot": (synth.TINY, "dot", 9, False),
    "tiny_ca_dense": (synth.TINY, "ca", 5, True),
    "beauty_ca": (dataclasses.replace(synth.BEAUTY, n_items=5000, n_attrs=500), "ca", 64, False),
    "beauty_dot": (dataclasses.replace(synth.BEAUTY, n_items=5000, n_attrs=500), "dot", 64, False),
    "men_ca": (MEN_S, "ca", 48, False),
    "men_dot": (MEN_S, "dot", 48, False),
    "men_ca_dense": (MEN_S, "ca", 12, True),
    "beauty_L200_dense": (dataclasses.replace(synth.BEAUTY, n_items=5000, n_attrs=500, seq_len=200), "ca", 6, True),
    "beauty_L100": (dataclasses.replace(synth.BEAUTY, n_items=5000, n_attrs=500, seq_len=100), "dot", 40, False),
}


def run(name):
    shape, decoder, B, all_valid = CASES[name]
    model = synth.build_model(shape, decoder, p=0.5, seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=5)
    b = synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder)
    dense = (b["p_x"], table.gather_dense(b["p_x"]), b["p_c"], b["o_x"], table.gather_dense(b["o_x"]), b["o_c"])
    with torch.no_grad():
        y_ref = O.carca_forward(sd, cfg, dense[:3], [dense[3:6]], training=False).numpy()
    model = model.cuda().eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=5).cuda())
    d = {k: v.cuda() for k, v in b.items()}
    for dt in ("bf16", "fp32"):
        model.set_eval_dtype(dt)
        model.force_eval_path = "rows_fp32" if dt == "fp32" else None
        with torch.no_grad():
            y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])])
        torch.cuda.synchronize()
        st = int(fused._plans[model].status.item())
        report(name + ":" + dt, y.cpu().numpy(), y_ref, st)


def report(name, y, y_ref, st):
    abs_err = np.abs(y - y_ref).max()
    rel = (np.abs(y - y_ref) / np.maximum(np.abs(y_ref), 1e-12)).max()
    logit = lambda p: np.log(np.clip(p, 1e-30, 1) / np.clip(1 - p, 1e-30, 1))  # noqa: E731
    le = np.abs(logit(y.astype(np.float64)) - logit(y_ref.astype(np.float64)))
    le = le[np.isfinite(le)]
    top = np.mean([len(set(np.argsort(-a)[:10]) & set(np.argsort(-r)[:10])) / 10 for a, r in zip(y, y_ref)])
    hr = np.mean((y_ref[:, 1:] > y_ref[:, :1]).sum(1) < 10), np.mean((y[:, 1:] > y[:, :1]).sum(1) < 10)
    print(f"{name:24s} status {st} nan {int(np.isnan(y).sum())} max|dp| {abs_err:.2e} max rel {rel:.2e} "
          f"logit err max {le.max():.3e} mean {le.mean():.3e} (|logit| max {np.abs(logit(y_ref.astype(np.float64))).max():.1f}) "
          f"top10 overlap {top:.3f} HR ref/ours {hr[0]:.3f}/{hr[1]:.3f}", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        try:
            run(n)
        except Exception as ex:  # noqa: BLE001
            print(f"{n:20s} FAILED {type(ex).__name__}: {ex}", flush=True)
