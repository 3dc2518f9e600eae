"""Development tool: step time of the packed-rows pipeline against the number of concurrent batch slices (fused.ROWS_LANES).
    python tools/lanes_probe.py          (needs a B200)"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from carca_replication_b200 import fused, synth  # noqa: E402
from tools.attn_tc_check import clock  # noqa: E402


def main():
    dev = torch.device("cuda")
    cases = [("beauty", synth.BEAUTY, 8192, False, ("bf16", "rows_fp32", "tc")),
             ("beauty L=100", dataclasses.replace(synth.BEAUTY, seq_len=100), 8192, False, ("bf16", "rows_fp32")),
             ("men", synth.MEN, 8192, False, ("bf16", "rows_fp32")),
             ("beauty all-valid", synth.BEAUTY, 8192, True, ("bf16", "rows_fp32"))]
    for name, shape, B, all_valid, modes in cases:
        model = synth.build_model(shape, "ca", p=0.5).to(dev).eval()
        model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
        bs = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=i, all_valid=all_valid).items()} for i in range(3)]
        for b in bs:
            b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
        for mode in modes:
            model.set_eval_dtype("bf16" if mode == "bf16" else "fp32")
            model.force_eval_path = None if mode == "bf16" else mode
            ref = None
            for lanes in ((1,) if mode == "tc" else (1, 2, 3, 4, 6, 8)):
                fused.ROWS_LANES = lanes
                graphs = []
                with torch.no_grad():
                    y = model.forward((bs[0]["p_x"], None, bs[0]["p_c"]), [(bs[0]["o_x"], None, bs[0]["o_c"])])
                torch.cuda.synchronize()
                ref = y if ref is None else ref
                err = (y - ref).abs().max().item()
                for b in bs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g), torch.no_grad():
                        model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
                    graphs.append(g)
                j = [0]

                def rep():
                    graphs[j[0] % 3].replay()
                    j[0] += 1
                ms = clock(rep, n=20)
                print(f"{name:18s} {mode:9s} lanes {lanes}: {ms:.3f} ms -> {B / ms * 1e3 / 1e6:6.2f} M users/s   max|dp| vs 1 lane {err:.1e} "
                      f"status {int(fused._plans[model].status.item())}", flush=True)
    fused.ROWS_LANES = 1


if __name__ == "__main__":
    main()
