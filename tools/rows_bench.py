"""Development tool: device-timed throughput of the bf16 packed-rows pipeline next to the fp32 path.
    python tools/rows_bench.py [shape=beauty|men] [B] [decoder] [all_valid 0|1] [L]"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from carca_replication_b200 import synth  # noqa: E402


def clock(fn, n=20, warm=5):
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


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "beauty"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    decoder = sys.argv[3] if len(sys.argv) > 3 else "ca"
    all_valid = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
    shape = synth.SHAPES[name]
    if len(sys.argv) > 5:
        shape = dataclasses.replace(shape, seq_len=int(sys.argv[5]))
    dev = torch.device("cuda")
    model = synth.build_model(shape, decoder, p=0.5).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
    bs = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=i, all_valid=all_valid).items()} for i in range(4)]
    for b in bs:
        b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
    rows = float(sum(((b["p_x"] != 0).sum() + (b["p_x"][:, -1] == 0).sum()).item() for b in bs)) / len(bs)
    out = {}
    only = os.environ.get("ROWS_BENCH_ONLY")
    for dt in ((only,) if only else ("bf16", "fp32", "fp32rows")):
        model.set_eval_dtype("bf16" if dt == "bf16" else "fp32")
        model.force_eval_path = "rows_fp32" if dt == "fp32rows" else None
        i = [0]

        def step():
            b = bs[i[0] % 4]
            i[0] += 1
            with torch.no_grad():
                return model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
        try:
            ms = clock(step, n=20 if (dt == "bf16" or shape.d == 64) else 4, warm=3)
            if os.environ.get("ROWS_BENCH_NOGRAPH"):
                print(f"{name} L={shape.seq_len} B={B} {decoder} all_valid={int(all_valid)} {dt}: eager {ms:.3f} ms", flush=True)
                continue
            g = torch.cuda.CUDAGraph()
            graphs = []
            for b in bs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g), torch.no_grad():
                    model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
                graphs.append(g)
            j = [0]

            def rep():
                graphs[j[0] % 4].replay()
                j[0] += 1
            msg = clock(rep, n=20 if (dt == "bf16" or shape.d == 64) else 4, warm=3)
        except Exception as ex:  # noqa: BLE001
            print(dt, "failed:", type(ex).__name__, str(ex)[:200])
            continue
        out[dt] = (ms, msg)
        print(f"{name} L={shape.seq_len} B={B} {decoder} all_valid={int(all_valid)} rows/batch={rows:.0f} {dt}: eager {ms:.3f} ms "
              f"graph {msg:.3f} ms -> {B / msg * 1e3 / 1e6:.2f} M users/s", flush=True)


if __name__ == "__main__":
    main()
