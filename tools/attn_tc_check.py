"""Development tool: tensor-core attention of the bf16 rows pipeline (rows_attn_tc.cuh) against the CUDA-core attention
kernel it replaces (CARCA_ROWS_ATTN_FFMA=1) on the same batches, plus step times of both.
    python tools/attn_tc_check.py [quick]          (needs a B200)"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from carca_replication_b200 import fused, synth  # noqa: E402

MEN_S = dataclasses.replace(synth.MEN, name="men_small", n_items=3000, n_attrs=96)
BEAUTY_S = dataclasses.replace(synth.BEAUTY, n_items=5000, n_attrs=500)
CASES = [
    ("beauty sparse", BEAUTY_S, "ca", 700, False),
    ("beauty all-valid", BEAUTY_S, "ca", 333, True),
    ("beauty dot L=100 all-valid", dataclasses.replace(BEAUTY_S, seq_len=100), "dot", 77, True),
    ("beauty L=129 sparse", dataclasses.replace(BEAUTY_S, seq_len=129), "ca", 300, False),
    ("men sparse", MEN_S, "ca", 500, False),
    ("men all-valid", MEN_S, "ca", 100, True),
    ("tiny", synth.TINY, "ca", 9, False),
]


def clock(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd(model, d):
    if os.environ.get("POISON"):          # uninitialised scratch must not matter: fill it with NaN patterns first
        for buf in fused._rows_scratch_cache.values():
            buf.fill_(255)
    with torch.no_grad():
        y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])])
    torch.cuda.synchronize()
    return y.clone(), int(fused._plans[model].status.item())


def main():
    dev = torch.device("cuda")
    for name, shape, dec, B, all_valid in CASES:
        try:
            model = synth.build_model(shape, dec, p=0.5, seed=5).to(dev).eval()
            model.embeds.set_attr_table(synth.make_attr_table(shape, seed=5).to(dev))
            model.set_eval_dtype("bf16")
            d = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid).items()}
            os.environ["CARCA_ROWS_ATTN_FFMA"] = "1"
            y0, st0 = fwd(model, d)
            y0b, _ = fwd(model, d)
            del os.environ["CARCA_ROWS_ATTN_FFMA"]
            nn = lambda t: int(torch.isnan(t).sum())  # noqa: E731
            for vs in ("0", "1") if os.environ.get("TRY_VSWAP") else ("0",):
                os.environ["CARCA_ATTN_VSWAP"] = vs
                y1, st1 = fwd(model, d)
                y2, _ = fwd(model, d)
                diff = (y1 - y0b).abs().max().item()
                print(f"{name:28s} vswap {vs}: status ffma {st0} tc {st1} nan ffma {nn(y0)} {nn(y0b)} tc {nn(y1)} {nn(y2)} max|dp| vs ffma "
                      f"{diff:.2e} rerun diff ffma {(y0 - y0b).abs().max().item():.1e} tc {(y1 - y2).abs().max().item():.1e}", flush=True)
            os.environ["CARCA_ATTN_VSWAP"] = "0"
        except Exception as ex:  # noqa: BLE001
            print(f"{name:28s} FAILED {type(ex).__name__}: {str(ex)[:300]}", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        return
    for name, shape, B, all_valid in (("beauty", synth.BEAUTY, 8192, False), ("beauty", synth.BEAUTY, 8192, True),
                                      ("men", synth.MEN, 8192, False), ("men", synth.MEN, 4096, True),
                                      ("beauty L=100", dataclasses.replace(synth.BEAUTY, seq_len=100), 8192, False)):
        model = synth.build_model(shape, "ca", p=0.5).to(dev).eval()
        model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
        model.set_eval_dtype("bf16")
        bs = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=i, all_valid=all_valid).items()} for i in range(3)]
        for b in bs:
            b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
        for ffma in (True, False):
            if ffma:
                os.environ["CARCA_ROWS_ATTN_FFMA"] = "1"
            else:
                os.environ.pop("CARCA_ROWS_ATTN_FFMA", None)
            i = [0]

            def step():
                b = bs[i[0] % 3]
                i[0] += 1
                with torch.no_grad():
                    return model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
            ms = clock(step)
            graphs = []
            for b in bs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g), torch.no_grad():
                    model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
                graphs.append(g)
            j = [0]

            def rep():
                graphs[j[0] % 3].replay()
                j[0] += 1
            msg = clock(rep)
            print(f"{name} B={B} all_valid={int(all_valid)} attention {'ffma' if ffma else 'tc  '}: eager {ms:.3f} ms graph {msg:.3f} ms "
                  f"-> {B / msg * 1e3 / 1e6:.2f} M users/s  status {int(fused._plans[model].status.item())}", flush=True)


if __name__ == "__main__":
    main()
