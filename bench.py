#!/usr/bin/env python
"""bench.py — CARCA hot path on B200: eval users/sec (1 positive + 100 negatives per user,
forward + BCE + HR@10/NDCG@10), Beauty-shaped synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one evaluate() batch body (src/train.py:42-51) over `--batch` users per GPU.  `--dtype bf16` (default:
BASELINE configs[1] "bf16/fp32", north_star "bf16 within 1e-2") runs the headline legs through the bf16 packed-rows
pipeline; the fp32 path (3xTF32, within 1e-4) is then the `fp32` section of the line — and vice versa with --dtype fp32.
  value : users/s with the step's inputs already resident in HBM (device-timed, CUDA events)
  e2e   : same metric through the public API from pinned HOST buffers: per step the ids/context
          H2D copies, forward, metrics and the D2H read of the accumulators are inside the timed region
          (the step is replayed through carca_replication_b200.graph.GraphedEvalStep; --e2e-eager: eager calls)
  roofline / ops / kernels : per C-ABI op (and, for the rows pipeline, per kernel) device time measured live with
          CUDA events, against MEASURED_PEAKS.json
  cpu_baseline : the oracle port of the reference timed on this box's host cores (bounded sample)
  train : fwd + BCE + bwd + Adam step seqs/s at the reference's batch size (256 per GPU)
`--impl reference` times the reference's CPU path (oracle port, all host threads) on a bounded sample.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="beauty", choices=["beauty", "men", "tiny"])
    ap.add_argument("--decoder", default="ca", choices=["ca", "dot"])
    ap.add_argument("--batch", type=int, default=8192, help="users per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"],
                    help="arithmetic of the headline legs (BASELINE configs[1] 'bf16/fp32'); the other one is a section")
    ap.add_argument("--train-batch", type=int, default=256)
    ap.add_argument("--cpu-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-catalog", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the bf16 / sweep / Men sections")
    ap.add_argument("--men-batch", type=int, default=8192)
    ap.add_argument("--catalog-users", type=int, default=1024)
    ap.add_argument("--rotate", type=int, default=8, help="distinct input batches cycled through")
    ap.add_argument("--e2e-eager", action="store_true",
                    help="e2e leg: issue the step's calls eagerly instead of replaying GraphedEvalStep")
    ap.add_argument("--no-graph", action="store_true",
                    help="time eager steps instead of CUDA-graph replays of them (device-resident leg)")
    return ap.parse_args()


# --------------------------------------------------------------------------------- helpers
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if (args.impl == "ours") else "gloo"
        if args.impl == "ours":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


# --------------------------------------------------------------------------------- reference arm (CPU)
def dense_batch(table, b):
    """The dense [B, N, A] attribute tensors the reference API consumes (src/data.py:119-131)."""
    return (b["p_x"], table.gather_dense(b["p_x"]), b["p_c"], b["o_x"], table.gather_dense(b["o_x"]), b["o_c"],
            b["y_true"])


def oracle_config(shape, decoder):
    from oracle.carca_oracle import OracleConfig

    return OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder, p_drop=0.0)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_model(shape, decoder, sd, device="cpu"):
    """The UNMODIFIED reference (baseline/_ref, staged by __graft_entry__.build()) built as scripts/training.py:165-172
    does and loaded with our weights (identical state_dict layout), or None when the copy is absent."""
    if not os.path.isdir(os.path.join(REF_DIR, "src")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torch.nn as nn
    from src import carca as R  # noqa: E402  (the reference package)

    emb = R.AllEmbedding(shape.n_items, shape.d, shape.g, shape.n_ctx, shape.n_attrs, R.IdentityEncoding())
    enc = nn.ModuleList([R.SelfAttentionBlock(shape.d, shape.n_heads, 0.5, True) for _ in range(shape.n_blocks)])
    dec = R.CrossAttentionBlock(shape.d, shape.n_heads, 0.5, True) if decoder == "ca" else R.DotProduct()
    model = R.CARCA(d=shape.d, p=0.5, emb=emb, enc=enc, dec=dec)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def reference_eval_body(model, batch):
    """evaluate()'s per-batch body exactly as src/train.py:42-50 runs it (forward, BCE, two sorts, three .item())."""
    from src.carca import BinaryCrossEntropy
    from src.train import compute_HR, compute_NDCG
    from src.utils import get_mask

    p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch
    with torch.no_grad():
        y = model.forward(profile=(p_x, p_a, p_c), targets=[(o_x, o_a, o_c)])
        loss = BinaryCrossEntropy().forward(y, y_true, get_mask(o_x)).item()
        return compute_HR(y, y_true, 10), compute_NDCG(y, y_true, 10), loss


def time_cpu_eval(shape, decoder, sd, table, B, steps, warmup):
    """users/s of the reference's evaluate() body on the host cores, all threads: the real reference when
    baseline/_ref is staged (kind "reference"), else the oracle port (kind "port")."""
    from carca_replication_b200 import synth
    from oracle import carca_oracle as O

    torch.set_num_threads(os.cpu_count())
    batches = [dense_batch(table, synth.make_eval_batch(shape, B, seed=100 + i)) for i in range(2)]
    ref = reference_model(shape, decoder, sd)
    if ref is not None:
        body, kind = (lambda b: reference_eval_body(ref, b)), "reference"
    else:
        cfg = oracle_config(shape, decoder)
        body, kind = (lambda b: O.eval_batch(sd, cfg, b)), "port"
    for i in range(warmup):
        body(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        body(batches[i % 2])
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, kind


def time_reference_cuda(shape, decoder, sd, table, B, steps, dev):
    """Secondary baseline: the same reference code, PyTorch eager on this B200, dense attributes
    resident on the device (what `model.forward` consumes), CUDA-event timed."""
    from carca_replication_b200 import synth

    ref = reference_model(shape, decoder, sd, device=dev)
    if ref is None:
        return {"unavailable": "baseline/_ref not staged"}
    batches = [tuple(t.to(dev) for t in dense_batch(table, synth.make_eval_batch(shape, B, seed=100 + i)))
               for i in range(2)]
    for i in range(3):
        reference_eval_body(ref, batches[i % 2])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        reference_eval_body(ref, batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms * 1e-3), "unit": "users/s", "batch": B, "ms_per_batch": ms,
            "what": "unmodified reference (baseline/_ref) on device='cuda', PyTorch eager, fp32, dense [B,N,A] attributes "
                    "already on the device; evaluate() body incl. its three .item() syncs"}


def run_reference(args):
    from carca_replication_b200 import synth

    rank, world, _ = dist_setup(args)
    if rank != 0:
        return
    shape = synth.SHAPES[args.shape]
    model = synth.build_model(shape, args.decoder, p=0.5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape)
    B = args.cpu_batch
    ups, sec, kind = time_cpu_eval(shape, args.decoder, sd, table, B, args.steps, max(1, min(args.warmup, 2)))
    cores = os.cpu_count()
    sample = f"{args.steps} batches of {B} users (dense [B,N,A] attributes pre-materialised, loader excluded)"
    out = {
        "impl": "reference", "metric": "eval_users_per_sec", "value": ups, "unit": "users/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, shape, B),
        "cpu_baseline": {"value": ups, "unit": "users/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ups, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the same body at the largest batch whose dense attribute tensors (3.9 MB per Beauty user) fit comfortably
    if not args.no_cpu_baseline and shape.name == "beauty":
        try:
            big = 1024
            ups2, sec2, _ = time_cpu_eval(shape, args.decoder, sd, table, big, max(1, args.steps // 4), 1)
            out["large_batch"] = {"value": ups2, "unit": "users/s", "batch": big, "ms_per_step": sec2 * 1e3}
        except (RuntimeError, MemoryError) as ex:
            out["large_batch"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    print(json.dumps(out), flush=True)


def workload_config(args, shape, batch):
    return {"workload": f"{shape.name}-shaped CARCA eval: 1+{shape.n_targets - 1} candidates/user, maxlen "
                        f"{shape.seq_len}, d={shape.d}, g={shape.g}, H={shape.n_heads}, {shape.n_blocks} blocks, "
                        f"A={shape.n_attrs} {shape.attr_kind}, C={shape.n_ctx}, decoder={args.decoder}",
            "users_per_gpu_per_step": batch, "items": shape.n_items, "k": 10}


# --------------------------------------------------------------------------------- our arm (B200)
def run_ours(args):
    import carca_replication_b200 as cb
    from carca_replication_b200 import _native as N
    from carca_replication_b200 import build as B_, ops, synth

    rank, world, local = dist_setup(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    B_.build()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from carca_replication_b200.parallel import bind_host_to_gpu_numa_node
    # N > 1: each rank stays on its GPU's NUMA node, before any pinned host buffer is allocated (N = 1 keeps every
    # core: the cpu_baseline leg of the same process times the reference path on all of them)
    numa = bind_host_to_gpu_numa_node(local) if world > 1 else {"gpu": local, "bound": False}
    shape = synth.SHAPES[args.shape]
    B, K, W = args.batch, args.steps, args.warmup
    model = synth.build_model(shape, args.decoder, p=0.5)
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table_cpu = synth.make_attr_table(shape)
    table = synth.make_attr_table(shape).to(dev)
    model = model.to(dev).eval()
    model.embeds.set_attr_table(table)
    loss_fn = cb.BinaryCrossEntropy()
    from carca_replication_b200 import fused as fused_
    if args.dtype == "bf16" and not fused_.rows_supported(model, shape.seq_len, shape.n_ctx, "bf16"):
        args.dtype = "fp32"                      # (a shape the bf16 pipeline does not cover, e.g. --shape tiny)
    model.set_eval_dtype(args.dtype)

    names = ("p_x", "p_c", "o_x", "o_c", "y_true")
    host = [synth.make_eval_batch(shape, B, seed=1000 * rank + i) for i in range(args.rotate)]
    host = [{k: b[k].pin_memory() for k in names} for b in host]
    devb = [{k: b[k].to(dev) for k in names} for b in host]
    dev_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in names)
    # every candidate of a user carries the positive's context (src/data.py:185): CARCA.forward is given the
    # [B,T,C] tensor as an expanded view of one row per user, which the kernels read per user
    for b in devb:
        b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)

    acc4 = torch.zeros(4, dtype=torch.float64, device=dev)     # hits, ndcg, users, sum of batch losses
    acc = acc4[:3]

    def step(b):
        y = model.forward(profile=(b["p_x"], None, b["p_c"]), targets=[(b["o_x"], None, b["o_c"])])
        # evaluate()'s per-batch reductions (BinaryCrossEntropy over get_mask(o_x), compute_HR, compute_NDCG) as the
        # repo's evaluate() issues them: one launch (carca_eval_metrics)
        ops.eval_metrics_(acc4, y, b["y_true"], b["o_x"], 10)
        return y

    def timed(fn, n):
        import gc

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(world)
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()                  # no collector pause inside the timed region (the e2e loop is host-paced)
        try:
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
        barrier(world)
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.no_grad():
        # ---- value: inputs resident in HBM
        for i in range(max(W, args.rotate)):        # every rotated batch is touched once before the clock starts
            step(devb[i % args.rotate])
        # a step is ~0.5 ms: keep the warm-up going for ~0.3 s so that the SM clocks of a freshly started process
        # have ramped up before the timed steps (a cold start otherwise costs up to 20 % on short runs)
        t_spin = time.perf_counter()
        while time.perf_counter() - t_spin < 0.3:
            for i in range(16):
                step(devb[i % args.rotate])
            torch.cuda.synchronize()
        # The device-resident leg replays each rotated batch's step (the same public-API calls) as a CUDA graph: the
        # eager step costs ~0.2 ms of host time (Python + ctypes + launches) against ~0.43 ms of GPU time, which is
        # fine for one process but makes N processes sharing one host CPU launch-bound.  e2e below stays eager.
        graphs = None
        if not args.no_graph:
            try:
                torch.cuda.synchronize()
                n0 = N.lib().carca_launch_count()
                graphs = []
                for b in devb:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        step(b)
                    graphs.append(g)
                launches_per_step = (N.lib().carca_launch_count() - n0) / len(graphs)
                for g in graphs:
                    g.replay()
                torch.cuda.synchronize()
            except Exception as ex:  # noqa: BLE001 -- the eager steps are always available
                print(f"bench: CUDA-graph capture of the eval step failed ({type(ex).__name__}: {ex}); timing eager "
                      "steps", file=sys.stderr, flush=True)
                graphs = None
                torch.cuda.synchronize()
        launches0 = N.lib().carca_launch_count()
        clocks = ClockSampler(local)
        clocks.__enter__()                      # sampled through both timed regions (value and e2e)
        if graphs is not None:
            ms_dev = timed(lambda i: graphs[i % args.rotate].replay(), K)
            launches = int(round(launches_per_step * K))
        else:
            ms_dev = timed(lambda i: step(devb[i % args.rotate]), K)
            launches = N.lib().carca_launch_count() - launches0
        hr_ndcg = (acc / acc[2].clamp(min=1)).tolist()

        # ---- e2e: pinned host buffers -> H2D -> forward -> metrics -> D2H of the accumulators.
        # Every step's ids/context cross PCIe inside the timed region; the copy of step i+1 is issued on a
        # second stream before step i's kernels (the double buffering a pinned DataLoader with
        # non_blocking copies gives), and the caller reads the metrics back every step.
        copy_stream = torch.cuda.Stream(device=dev)
        # What crosses PCIe per step is the PACKED batch (device_data.PackedEvalLayout): per-user offsets, candidate ids,
        # one context row per user (every candidate carries the positive's context, src/data.py:185) and the (id,
        # context) records of the VALID profile positions only — the windows are left-padded (src/data.py:112-113) and
        # ~86 % padding at Beauty shape; y_true is the constant [1, 0, ..] row (src/data.py:189-190) and is not sent.
        # The dense p_x / p_c the model takes are rebuilt by one kernel inside the step's graph.
        from carca_replication_b200.device_data import PackedEvalLayout, pack_eval_batch, unpack_eval_batch

        T_ = devb[0]["o_x"].shape[1]
        layout = PackedEvalLayout(B, shape.seq_len, T_, shape.n_ctx)
        host_arenas, used = [], []
        for hb in host:                  # the collate step: one pinned arena per host batch
            arena = torch.empty(layout.capacity, dtype=torch.uint8).pin_memory()
            used.append(pack_eval_batch(layout, arena, hb["p_x"], hb["p_c"], hb["o_x"], hb["o_c"]))
            host_arenas.append(arena)
        h2d_bytes = int(sum(used) / len(used))
        NS = 3                                               # input slots: copies run up to two steps ahead
        dev_arenas = [torch.empty(layout.capacity, dtype=torch.uint8, device=dev) for _ in range(NS)]
        for j, a_ in enumerate(dev_arenas):                  # valid contents in every slot before anything runs on it
            a_[:used[j % args.rotate]].copy_(host_arenas[j % args.rotate][:used[j % args.rotate]])
        slots = [unpack_eval_batch(layout, a_) for a_ in dev_arenas]
        ready = [torch.cuda.Event() for _ in range(NS)]
        freed = [torch.cuda.Event() for _ in range(NS)]
        # the step's result (hits, ndcg sum, users, loss sum) is read back every step; the host waits for the
        # read of step i after it has queued step i+1, so the device never idles on the round trip
        stats_host = [torch.zeros(4, dtype=torch.float64).pin_memory() for _ in range(NS)]
        landed = [torch.cuda.Event() for _ in range(NS)]
        results = []
        # the public API for this loop is GraphedEvalStep (carca_replication_b200/graph.py): evaluate()'s per-batch
        # body (window unpack + CARCA.forward + BinaryCrossEntropy + rank metrics + the D2H copy of the accumulators)
        # captured once per input slot and replayed; --e2e-eager issues the same calls eagerly
        ev_steps = None
        stats_dev = torch.zeros(4, dtype=torch.float64, device=dev)
        if not args.e2e_eager:
            from carca_replication_b200.graph import GraphedEvalStep

            ev_steps = [GraphedEvalStep(model, slots[j], k=10, stats=stats_dev, result=stats_host[j], static_inputs=True,
                                        prologue=(lambda j=j: unpack_eval_batch(layout, dev_arenas[j], slots[j])))
                        for j in range(NS)]
        stats = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(NS)]

        def upload(i):
            slot = i % NS
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])          # the step that last used this slot is done
                n_ = used[i % args.rotate]
                dev_arenas[slot][:n_].copy_(host_arenas[i % args.rotate][:n_], non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_run(n):
            cur = torch.cuda.current_stream()
            for e in freed:
                e.record(cur)
            for j in range(min(NS - 1, n)):
                upload(j)
            for i in range(n):
                if i + NS - 1 < n:
                    upload(i + NS - 1)
                cur.wait_event(ready[i % NS])
                if ev_steps is not None:
                    ev_steps[i % NS].replay()                # ends with the D2H copy of the accumulators
                    freed[i % NS].record(cur)
                else:
                    sl = unpack_eval_batch(layout, dev_arenas[i % NS], slots[i % NS])
                    step(sl)
                    freed[i % NS].record(cur)
                    st = stats[i % NS]
                    st.copy_(acc4)
                    stats_host[i % NS].copy_(st, non_blocking=True)
                landed[i % NS].record(cur)
                if i > 1:                                    # the caller consumes step i-2's metrics: two steps stay
                    landed[(i - 2) % NS].synchronize()       # queued behind the one whose result the host waits for
                    results.append(float(stats_host[(i - 2) % NS][0]))   # (its slot is rewritten by step i+1, queued later)
            for j in range(max(0, n - 2), n):
                landed[j % NS].synchronize()
                results.append(float(stats_host[j % NS][0]))

        e2e_run(max(W, args.rotate))                 # every pinned host arena has crossed PCIe once (the first
                                                     # copy out of a pinned buffer runs at ~1/5 of the link rate)
        ms_e2e = timed(lambda _i: e2e_run(K), 1)
        clocks.__exit__(None, None, None)

        # ---- per-op device time (CUDA events around each C-ABI op), rotating inputs
        op_ms, op_ms_per_op_path = op_breakdown(model, loss_fn, devb, acc, shape, args, K)

    users_per_step = B * world
    value = users_per_step * K / (ms_dev / 1e3)
    e2e = users_per_step * K / (ms_e2e / 1e3)
    pk = peaks()
    roof, ops_table = roofline(op_ms, shape, B, args.decoder, pk)
    _, per_op_table = roofline(op_ms_per_op_path, shape, B, args.decoder, pk)
    stage_ms = None
    if args.dtype == "bf16":
        with torch.no_grad():
            stage_ms = rows_stage_times(model, devb, max(K, 10))
        roof = rows_roofline(stage_ms, ms_dev / K, op_ms, shape, B, args.decoder, pk)
    elif roof["kernel"] == "fused_forward":
        # the SAME CUDA-event timing as `value`: kernel time = ms_per_step x the kernel's share of the step (the share
        # from per-op events around the two launches of a step, cross-checked by the committed ncu launch list)
        t_kernel_ms = (ms_dev / K) * roof["share_of_step"]
        fl = fwd_flops_per_user(shape, args.decoder)
        roof["achieved"] = fl * B / (t_kernel_ms * 1e-3) / 1e12
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel_ms"] = t_kernel_ms
        roof["timing"] = "ms_per_step of the timed region x share_of_step"
        cap_path = os.path.join(ROOT, "profiles", "r02", "roofline_capture.json")
        if os.path.exists(cap_path):
            cap = json.load(open(cap_path))
            roof["traffic"] = cap["dram_bytes_per_user"] * B
            roof["traffic_source"] = cap["source"]
            roof["tensor_pipe_active_pct"] = cap.get("tensor_pipe_active_pct")
        else:
            roof["traffic_source"] = "profiles/r01/ncu_fused_tc_v17_raw.csv (3334 B/user)"

    out = {
        "metric": "eval_users_per_sec", "value": value, "unit": "users/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "f32",
        "data": "synthetic", "config": dict(workload_config(args, shape, B),
                                            attrs="device-resident item->attribute table (CSR for multi-hot), "
                                                  "ids + context per step",
                                            l2=f"{args.rotate} distinct input batches rotated "
                                               f"({args.rotate * dev_bytes / 1e6:.0f} MB > 126 MB L2); the folded item "
                                               "table (14.7 MB) and the weights are L2-resident by design",
                                            launch=("each timed step is a CUDA-graph replay of the eager step "
                                                    "(CARCA.forward + loss + metrics on one rotated batch)"
                                                    if graphs is not None else "eager steps")),
        "e2e": {"value": e2e, "unit": "users/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 32,
                "h2d_bytes_per_step_dense_tensors": layout.dense_bytes(),
                "ms_per_step": ms_e2e / K, "launch": "eager calls" if args.e2e_eager else "GraphedEvalStep (one CUDA "
                "graph replay per step)", "api": "device_data.unpack_eval_batch + CARCA.forward + BinaryCrossEntropy + "
                                                  "rank metrics; per step one H2D copy of a pinned PACKED arena (per-user "
                                                  "offsets, candidate ids, one context row per user, (id, context) of "
                                                  "the valid profile positions; issued up to two steps ahead on a second "
                                                  "stream) and one D2H read of the accumulators (consumed by the host two "
                                                  "steps behind)"},
        "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof, "ops": ops_table,
        "ops_per_op_path": per_op_table,
        "peaks": pk, "hr10": hr_ndcg[0], "ndcg10": hr_ndcg[1], "host_numa": numa,
    }
    out["config"]["arithmetic"] = (
        "bf16 GEMM / attention operands and bf16 candidate table on the tensor cores (tcgen05 kind::f16), fp32 accumulation, "
        "softmax, LayerNorm, residual stream, embedding gather and scorer (north_star: bf16 within 1e-2; tests/test_gpu_bf16.py); "
        "the fp32 path (3xTF32, within 1e-4) is the `fp32` section" if args.dtype == "bf16" else
        "fp32 contract (3xTF32 tensor-core products, scores within 1e-4 of the reference); the bf16 pipeline is the `bf16` section")
    if stage_ms is not None:
        out["kernels"] = stage_ms

    # worst case for the row packing: every profile position valid (one user per 64-row bin)
    with torch.no_grad():
        full = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=5000 + i, all_valid=True).items()}
                for i in range(2)]
        for b in full:
            b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
        for i in range(2):
            step(full[i % 2])
        ms_full = timed(lambda i: step(full[i % 2]), max(3, K // 2)) / max(3, K // 2)
    out["all_valid_profiles"] = {"value": users_per_step / (ms_full / 1e3), "unit": "users/s", "ms_per_step": ms_full,
                                 "what": "same step with all 50 profile positions valid (no padding to skip)"}

    def optional(key, fn):
        # the secondary sections must never cost the headline line (N = 1: a Python error is caught and recorded;
        # N > 1: every rank runs the same code, an exception on one rank would hang the others, so no catch there)
        if world > 1:
            out[key] = fn()
            return
        try:
            out[key] = fn()
        except Exception as ex:  # noqa: BLE001
            out[key] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if world > 1 and not args.no_train:
        optional("train", lambda: time_train_dp(shape, args, dev, table, rank, world))
    def in_fp32(fn):          # full-catalog ranking is an fp32 contract (ranks bit-equal to the score-matrix path)
        model.set_eval_dtype("fp32")
        try:
            return fn()
        finally:
            model.set_eval_dtype(args.dtype)

    if world > 1 and not args.no_catalog:
        optional("catalog", lambda: in_fp32(lambda: time_catalog_sharded(model, shape, args, dev, rank, world)))
    if world > 1 and not args.no_extra:
        optional("men", lambda: time_men(args, dev, pk, rank=rank, world=world))
    if world == 1 and rank == 0:
        if not args.no_train:
            optional("train", lambda: time_train(shape, args, dev, table))
        if not args.no_catalog:
            optional("catalog", lambda: in_fp32(lambda: time_catalog(model, shape, args, dev)))
            optional("device_pipeline", lambda: time_device_pipeline(model, shape, args, dev))
            optional("dense_attrs", lambda: in_fp32(lambda: time_dense_attrs(model, table, shape, args, dev, pk)))
        if not args.no_extra:
            other = "fp32" if args.dtype == "bf16" else "bf16"
            optional(other, lambda: time_other_dtype(model, shape, args, dev, devb, step, pk, other))
            optional("sweep", lambda: time_sweep(shape, args, dev, table))
            optional("men", lambda: time_men(args, dev, pk))
        if not args.no_cpu_baseline:
            cb_B = args.cpu_batch
            ups, sec, kind = time_cpu_eval(shape, args.decoder, sd_cpu, table_cpu, cb_B, 3, 1)
            out["cpu_baseline"] = {"value": ups, "unit": "users/s", "cores": os.cpu_count(), "kind": kind,
                                   "sample": f"3 batches of {cb_B} users after 1 warm-up, dense attributes "
                                             "pre-materialised (reference API), loader excluded",
                                   "ms_per_batch": sec * 1e3,
                                   "same_batch_as_ours": {"batch": cb_B, "ours_users_per_s": None}}
            # the same (reference-sized) batch through our path, so that the ratio at EQUAL batch is on the line too
            small = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, cb_B, seed=900 + i).items()} for i in range(4)]
            for b in small:
                b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
            with torch.no_grad():
                out["cpu_baseline"]["same_batch_as_ours"]["ours_users_per_s"] = graphed_rate(step, small, cb_B, 50)
            optional("reference_cuda", lambda: time_reference_cuda(shape, args.decoder, sd_cpu, table_cpu, cb_B, 5, dev))
    if isinstance(out.get("fp32"), dict) and "value" in out["fp32"]:
        out["value_fp32"] = out["fp32"]["value"]          # the same step under the fp32 contract, next to the headline
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def op_breakdown(model, loss_fn, devb, acc, shape, args, K):
    """Average device ms of each op of one eval step, measured with CUDA events (rotating inputs).
    Returns (stages of the path the step really takes, stages of the per-op path for comparison)."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import ops

    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def run(stages, body):
        ev = {s: [] for s in stages}
        body(devb[0])                      # untimed: one-time work of a path (folded tables, plans, first allocations)
        torch.cuda.synchronize()
        for i in range(K):
            t = body(devb[i % len(devb)])
            torch.cuda.synchronize()
            for s, a, c in zip(stages, t[:-1], t[1:]):
                ev[s].append(a.elapsed_time(c))
        return {s: sum(v) / len(v) for s, v in ev.items()}

    def per_op(b):
        with ops.forward_seed():
            t = [mark()]
            p_mask = cb.get_mask(b["p_x"])
            p_e = model.embeds.forward(b["p_x"], None, b["p_c"], p_mask, False)
            t.append(mark())
            for blk in model.encoder:
                p_e = blk.forward(p_e, p_mask)
            t.append(mark())
            p_e = ops.LayerNormFn.apply(p_e, model.norm.weight, model.norm.bias)
            t.append(mark())
            o_mask = cb.get_mask(b["o_x"])
            o_e = model.embeds.forward(b["o_x"], None, b["o_c"], o_mask, True)
            t.append(mark())
            y = model.decoder.forward(o_e, o_mask, p_e, p_mask)
            t.append(mark())
            loss_fn.forward(y, b["y_true"], o_mask)
            t.append(mark())
            ops.rank_metrics_(acc, y, b["y_true"], 10)
            t.append(mark())
        return t

    def fused(b):
        t = [mark()]
        y = model.forward(profile=(b["p_x"], None, b["p_c"]), targets=[(b["o_x"], None, b["o_c"])])
        t.append(mark())
        ops.eval_metrics_(acc4, y, b["y_true"], b["o_x"], 10)
        t.append(mark())
        return t

    acc4 = torch.zeros(4, dtype=torch.float64, device=acc.device)
    modular = run(["embed_profile", "encoder_blocks", "final_norm", "embed_targets", "decoder", "bce",
                   "rank_metrics"], per_op)
    b0 = devb[0]
    if model._fused_eval_applies((b0["p_x"], None, b0["p_c"]), [(b0["o_x"], None, b0["o_c"])]):
        return run(["fused_forward", "eval_metrics"], fused), modular
    return modular, modular


FP32_FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 lanes x FMA at 1965 MHz = 74.4


def roofline(op_ms, shape, B, decoder, pk):
    """Algorithmic bytes / flops per op (all L positions counted) over measured time."""
    L, T, d, g, C, nb = shape.seq_len, shape.n_targets, shape.d, shape.g, shape.n_ctx, shape.n_blocks
    attr_b = 36 if shape.attr_kind == "multihot" else 4 * shape.n_attrs
    pos_b = 4 + 4 * d + attr_b + 4 * C + 4 * d                       # id + item row + attrs + ctx + e written
    enc = nb * (10 * L * d * d + 4 * L * L * d)
    dec = (2 * T * d * d + 4 * L * d * d + 4 * T * L * d + 2 * T * d) if decoder == "ca" else 2 * T * d
    flops = {"encoder_blocks": enc, "decoder": dec, "fused_forward": enc + dec + 2 * (L + T) * C * d}
    bytes_ = {
        "embed_profile": L * pos_b, "embed_targets": T * pos_b, "final_norm": 2 * L * d * 4,
        "bce": T * 12, "rank_metrics": T * 8, "eval_metrics": T * 12,
    }
    table = {}
    for op, ms in op_ms.items():
        row = {"ms": ms}
        if op in flops:
            tf = flops[op] * B / (ms * 1e-3) / 1e12
            row.update(bound="tensor", achieved=tf, unit="TFLOP/s", peak=pk["bf16_tflops_sustained"],
                       frac=tf / pk["bf16_tflops_sustained"], frac_of_fp32_ffma_peak=tf / FP32_FFMA_PEAK_TFLOPS)
        else:
            gb = bytes_[op] * B / (ms * 1e-3) / 1e9
            row.update(bound="hbm", achieved=gb, unit="GB/s", peak=pk["hbm_gbs"], frac=gb / pk["hbm_gbs"])
        table[op] = row
    dom = max(op_ms, key=op_ms.get)
    r = table[dom]
    # dram__bytes_read.sum + dram__bytes_write.sum of one fused_eval_tc launch over 8192 Beauty-shaped users
    # (ncu --set full, profiles/r01/ncu_fused_tc_v12_raw.csv): 27.24 MB read + 0.08 MB written = 3334 B/user.  Algorithmic HBM bytes with
    # one context row per user: ids + context in, scores out = 2.2 KB/user; the item tables and weights are
    # L2 hits (90 % sector hit rate), the scores were still in L2 when the kernel ended.
    traffic = 3334.0 * B if (dom == "fused_forward" and shape.name == "beauty") else None
    roof = {"kernel": dom, "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"],
            "frac": r["frac"], "traffic": traffic, "peak_source": pk["source"],
            "share_of_step": op_ms[dom] / sum(op_ms.values())}
    if "frac_of_fp32_ffma_peak" in r:
        from carca_replication_b200 import fused
        if dom == "fused_forward" and fused.VARIANT != 1:
            roof["note"] = ("encoder on tcgen05 kind::tf32 (fp32-grade via the 3xTF32 split, 3 MMAs per fp32 product), "
                            "cross-attention decoder as an fp32 loop per candidate row. achieved = algorithmic FLOPs "
                            "(all 50 profile positions per user, as the reference executes them) / measured time; the "
                            "kernel runs the encoder on valid positions only (all_valid_profiles is the no-padding "
                            "case). peak is the measured dense bf16 cuBLAS rate; the encoder is bound by the row "
                            "operations between small dependent MMAs, the decoder by shared-memory operand delivery, "
                            "neither by tensor or HBM throughput (profiles/r01/README.md)")
        else:
            roof["note"] = ("fp32 CUDA-core (FFMA) kernel measured against the bf16 tensor peak; against the "
                            f"{FP32_FFMA_PEAK_TFLOPS:.1f} TFLOP/s fp32 FFMA peak the fraction is "
                            f"{r['frac_of_fp32_ffma_peak']:.3f}")
    return roof, table


def graphed_rate(step, batches, B, n, world=1, dev=None):
    """users/s of `step(batch)` replayed as CUDA graphs over the rotated `batches` (CUDA events; max over ranks)."""
    graphs = []
    for b in batches:
        step(b)
    torch.cuda.synchronize()
    for b in batches:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(b)
        graphs.append(g)
    for g in graphs:
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return world * B / (ms * 1e-3)


def fwd_flops_per_user(shape, decoder, rows=None):
    """Algorithmic forward FLOPs per user: all L positions unless `rows` (valid rows per user) is given."""
    L, T, d, C, nb = shape.seq_len, shape.n_targets, shape.d, shape.n_ctx, shape.n_blocks
    n = L if rows is None else rows
    enc = nb * (10 * n * d * d + 4 * n * n * d)
    dec = (2 * T * d * d + 4 * n * d * d + 4 * T * n * d + 2 * T * d) if decoder == "ca" else 2 * T * d
    return enc + dec + 2 * (n + T) * C * d


def eval_step_fn(model, acc4):
    from carca_replication_b200 import ops

    def step(b):
        y = model.forward(profile=(b["p_x"], None, b["p_c"]), targets=[(b["o_x"], None, b["o_c"])])
        ops.eval_metrics_(acc4, y, b["y_true"], b["o_x"], 10)
        return y
    return step


def device_batches(shape, B, dev, n=4, seed=0, all_valid=False):
    from carca_replication_b200 import synth

    out = []
    for i in range(n):
        b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=seed + i, all_valid=all_valid).items()}
        b["o_c"] = b["o_c"][:, :1, :].contiguous().expand(-1, b["o_x"].shape[1], -1)
        out.append(b)
    return out


def time_other_dtype(model, shape, args, dev, devb, step, pk, other):
    """BASELINE configs[1] "bf16/fp32": the same eval step in the arithmetic the headline does NOT use, and the deviation
    between the two on the same batch.  bf16 = packed-rows pipeline (tcgen05 kind::f16 GEMMs / attention / decoder scores,
    fp32 accumulation / softmax / LayerNorm / embedding / scorer); fp32 = the one-kernel 3xTF32 forward (L <= 64) or
    the fp32 flavour of the rows pipeline."""
    B = args.batch
    main = model.eval_dtype
    b0 = devb[0]
    with torch.no_grad():
        y_main = step(b0).clone()
        model.set_eval_dtype(other)
        try:
            y_other = step(b0).clone()
            path = model._fused_eval_mode((b0["p_x"], None, b0["p_c"]), [(b0["o_x"], None, b0["o_c"])])
            rate = graphed_rate(step, devb[:4], B, max(args.steps, 10))
            full = device_batches(shape, B, dev, n=2, seed=5000, all_valid=True)
            rate_full = graphed_rate(step, full, B, max(3, args.steps // 2))
        finally:
            model.set_eval_dtype(main)
    y32, y16 = (y_other, y_main) if other == "fp32" else (y_main, y_other)
    dp = float((y16 - y32).abs().max().item())
    hr32 = float(((y32[:, 1:] > y32[:, :1]).sum(1) < 10).double().mean().item())
    hr16 = float(((y16[:, 1:] > y16[:, :1]).sum(1) < 10).double().mean().item())
    top = float(sum(len(set(a.tolist()) & set(b.tolist())) for a, b in
                    zip(y32[:512].topk(10, dim=1).indices, y16[:512].topk(10, dim=1).indices)) / (10 * min(512, B)))
    fl = fwd_flops_per_user(shape, args.decoder)
    return {"value": rate, "unit": "users/s", "path": path,
            "dtype": ("bf16 operands, fp32 accumulate / softmax / LayerNorm / scorer" if other == "bf16" else
                      "fp32 contract: 3xTF32 tensor-core products, everything else fp32"),
            "max_abs_prob_diff_bf16_vs_fp32": dp, "hr10_fp32": hr32, "hr10_bf16": hr16, "top10_overlap_bf16_vs_fp32": top,
            "algorithmic_tflops": fl * rate / 1e12, "frac_of_bf16_peak": fl * rate / 1e12 / pk["bf16_tflops_sustained"],
            "all_valid_profiles": {"value": rate_full, "unit": "users/s",
                                   "algorithmic_tflops": fl * rate_full / 1e12,
                                   "frac_of_bf16_peak": fl * rate_full / 1e12 / pk["bf16_tflops_sustained"]}}


ROWS_STAGE_NAMES = {1: "rows_pack", 2: "rows_embed_ln", 3: "rows_gemm(Q|K|V)", 4: "rows_attn", 5: "rows_gemm(FFN-1)",
                    6: "rows_gemm(FFN-2+LN)", 7: "rows_ffn_chain", 8: "rows_gemm(decoder K|V)", 9: "rows_decode"}


def rows_stage_times(model, devb, n):
    """Average device ms of every kernel of the bf16 rows pipeline over n eager steps on rotated batches, from CUDA events
    the library records on the launching stream between its launches (carca_rows_set_stage_events)."""
    import ctypes as C

    from carca_replication_b200 import _native as N

    lib = N.lib()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(48)]
    for e in evs:
        e.record()                                # creates the handles
    torch.cuda.synchronize()
    arr = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
    ids = (C.c_int32 * 64)()
    acc = {}
    try:
        lib.carca_rows_set_stage_events(arr, len(evs))
        for i in range(n + 2):
            b = devb[i % len(devb)]
            model.forward(profile=(b["p_x"], None, b["p_c"]), targets=[(b["o_x"], None, b["o_c"])])
            torch.cuda.synchronize()
            k = lib.carca_rows_stage_ids(ids, 64)
            if i < 2:
                continue
            seen = {}
            for j in range(1, k):
                name = ROWS_STAGE_NAMES.get(ids[j], str(ids[j]))
                seen[name] = seen.get(name, 0) + 1
                a = acc.setdefault(name, [0.0, 0])
                a[0] += evs[j - 1].elapsed_time(evs[j])
                a[1] += 1
    finally:
        lib.carca_rows_set_stage_events(None, 0)
    total = sum(a[0] for a in acc.values()) / n
    return {"per_step_ms": {k: a[0] / n for k, a in acc.items()}, "launches_per_step": {k: a[1] / n for k, a in acc.items()},
            "sum_ms": total, "timing": f"CUDA events between the launches of {n} eager steps on the launching stream"}


def rows_roofline(stage_ms, ms_step, op_ms, shape, B, decoder, pk):
    """Roofline line of the bf16 pipeline's dominant kernel.  Its time inside the timed region = ms_per_step x its share
    of the step (shares from the per-kernel events); algorithmic work per DESIGN.md 4c."""
    per = stage_ms["per_step_ms"]
    metrics_ms = op_ms.get("eval_metrics", 0.0)
    total = stage_ms["sum_ms"] + metrics_ms
    # the dominant kernel = the longest single launch of the step (the decoder: 1 launch; attention and the FFN chain
    # are 3 launches of a third of its duration each)
    dom = max(per, key=lambda k_: per[k_] / max(1.0, stage_ms["launches_per_step"][k_]))
    n_launch = max(1.0, stage_ms["launches_per_step"][dom])
    share = per[dom] / total
    t_kernel = ms_step * share / n_launch * 1e-3                     # seconds per launch
    L, T, d, C, H = shape.seq_len, shape.n_targets, shape.d, shape.n_ctx, shape.n_heads
    rows_per_user = shape.mean_valid if hasattr(shape, "mean_valid") else None
    roof = {"kernel": dom, "share_of_step": share, "kernel_ms": t_kernel * 1e3, "launches_per_step": n_launch,
            "timing": "ms_per_step of the timed region x the kernel's share of the step (per-kernel CUDA events)",
            "peak_source": pk["source"]}
    cap_path = os.path.join(ROOT, "profiles", "r02", "roofline_capture_bf16.json")
    cap = json.load(open(cap_path)) if os.path.exists(cap_path) else {}
    if dom == "rows_decode":
        # candidate scoring: per user T candidate ids + T bf16 table rows gathered + T scores written + one context row;
        # per valid profile position its fp32 key row and per-head terms.  The table (7.3 MB bf16) is L2-resident by
        # design, so DRAM traffic is far below the algorithmic bytes: the bound that matters is the gather path.
        keys = cap.get("rows_per_user", 7.1)
        byt = T * (4 + 2 * d + 4) + 4 * C + keys * (4 * d + H * 36)
        gb = byt * B / t_kernel / 1e9
        roof.update(bound="hbm", achieved=gb, peak=pk["hbm_gbs"], unit="GB/s", frac=gb / pk["hbm_gbs"],
                    algorithmic_bytes_per_user=byt)
    else:
        fl = fwd_flops_per_user(shape, decoder)
        tf = fl * B / (ms_step * 1e-3) / 1e12
        roof.update(bound="tensor", achieved=tf, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                    frac=tf / pk["bf16_tflops_sustained"], note="whole-step algorithmic FLOPs over the step time")
    # the whole step by SURVEY 8d's counting rule (all L positions and the full L x L attention counted, fp32 item rows)
    attr_b = 36 if shape.attr_kind == "multihot" else 4 * shape.n_attrs
    step_bytes = (L + T) * (4 + 4 * d + attr_b + 4 * C) + 4 * T
    step_flops = fwd_flops_per_user(shape, decoder)
    roof["whole_step"] = {
        "algorithmic_bytes_per_user": step_bytes, "achieved_gbs": step_bytes * B / (ms_step * 1e-3) / 1e9,
        "frac_of_hbm_peak": step_bytes * B / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"],
        "algorithmic_flops_per_user": step_flops, "achieved_tflops": step_flops * B / (ms_step * 1e-3) / 1e12,
        "frac_of_bf16_peak": step_flops * B / (ms_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
        "note": "SURVEY 8d figures x users per step / ms_per_step; the pipeline executes the valid positions only"}
    k = cap.get("kernels", {}).get(dom)
    roof["traffic"] = k["dram_bytes_per_launch"] if k else None
    if k:
        roof["traffic_source"] = cap.get("source")
        roof["tensor_pipe_active_pct"] = k.get("tensor_pipe_active_pct")
    return roof


def time_sweep(shape, args, dev, table):
    """BASELINE configs[4]: maxlen 50 / 100 / 200 x batch 128 .. 8192 on this GPU, UNFILTERED synthetic users (the
    length law produces users with more than 64 valid positions at L >= 100: nothing is dropped or re-routed on the
    host), fp32 and bf16, plus every position valid at maxlen 100."""
    import dataclasses

    from carca_replication_b200 import synth

    out = {"unit": "users/s", "rows": []}
    n = max(args.steps, 10)
    for L in (50, 100, 200):
        shp = dataclasses.replace(shape, seq_len=L)
        model = synth.build_model(shp, args.decoder, p=0.5).to(dev).eval()
        model.embeds.set_attr_table(table)
        acc4 = torch.zeros(4, dtype=torch.float64, device=dev)
        step = eval_step_fn(model, acc4)
        for B in (128, 1024, 8192):
            bs = device_batches(shp, B, dev, n=4, seed=300 + L)
            longest = int(max((b["p_x"] != 0).sum(1).max().item() for b in bs))
            row = {"maxlen": L, "batch": B, "longest_profile": longest}
            with torch.no_grad():
                for dt in ("fp32", "bf16"):
                    model.set_eval_dtype(dt)
                    row[dt] = graphed_rate(step, bs, B, n)
                    if dt == "fp32":
                        row["fp32_path"] = model._fused_eval_mode((bs[0]["p_x"], None, bs[0]["p_c"]),
                                                                  [(bs[0]["o_x"], None, bs[0]["o_c"])])
                model.set_eval_dtype("fp32")
            out["rows"].append(row)
        if L == shape.seq_len:
            # past the sweep's range: the step is latency-bound at 8192 users (one short wave per kernel), so the rate
            # keeps growing with the batch
            with torch.no_grad():
                for B in (16384, 32768):
                    bs = device_batches(shp, B, dev, n=2, seed=400 + L)
                    row = {"maxlen": L, "batch": B, "beyond_the_sweep_range": True}
                    for dt in ("fp32", "bf16"):
                        model.set_eval_dtype(dt)
                        try:
                            row[dt] = graphed_rate(step, bs, B, max(5, n // 2))
                        except Exception as ex:  # noqa: BLE001
                            row[dt] = f"{type(ex).__name__}: {ex}"[:120]
                    model.set_eval_dtype("fp32")
                    out["rows"].append(row)
        if L == 100:
            bs = device_batches(shp, 2048, dev, n=2, seed=77, all_valid=True)
            with torch.no_grad():
                row = {"maxlen": L, "batch": 2048, "all_valid": True}
                for dt in ("fp32", "bf16"):
                    model.set_eval_dtype(dt)
                    row[dt] = graphed_rate(step, bs, 2048, max(3, n // 3))
                model.set_eval_dtype("fp32")
            out["rows"].append(row)
    return out


def time_men(args, dev, pk, rank=0, world=1):
    """BASELINE configs[2]: Men-shaped CARCA (d = 256, 4 heads, 512-d dense attributes, 110,637 items), eval step
    (forward + loss + HR@10 / NDCG@10), users sharded over the ranks (weak scaling, no data-path collective):
    fp32 (packed rows, 3xTF32 GEMMs) and bf16 (tcgen05 kind::f16 pipeline), Beauty-like sparse profiles and every
    position valid."""
    from carca_replication_b200 import synth

    shape = synth.MEN
    B = args.men_batch
    model = synth.build_model(shape, args.decoder, p=0.5).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
    acc4 = torch.zeros(4, dtype=torch.float64, device=dev)
    step = eval_step_fn(model, acc4)
    n = max(args.steps, 10)
    out = {"unit": "users/s", "batch_per_gpu": B, "n_gpus": world, "scaling": "weak",
           "workload": f"men-shaped eval: d={shape.d}, H={shape.n_heads}, A={shape.n_attrs} dense, {shape.n_items} items, "
                       f"maxlen {shape.seq_len}, 1+{shape.n_targets - 1} candidates, decoder={args.decoder}"}
    sparse = device_batches(shape, B, dev, n=4, seed=1000 * rank + 11)
    full = device_batches(shape, B, dev, n=2, seed=1000 * rank + 22, all_valid=True)
    rows_sparse = float(sum(((b["p_x"] != 0).sum() + (b["p_x"][:, -1] == 0).sum()).item() for b in sparse)) / len(sparse) / B
    with torch.no_grad():
        for dt in ("fp32", "bf16"):
            model.set_eval_dtype(dt)
            r_s = graphed_rate(step, sparse, B, n, world, dev)
            r_f = graphed_rate(step, full, B, max(3, n // 3), world, dev)
            fl_exec_s = fwd_flops_per_user(shape, args.decoder, rows=rows_sparse)
            fl_all = fwd_flops_per_user(shape, args.decoder)
            peak = pk["bf16_tflops_sustained"]
            out[dt] = {"value": r_s, "all_valid_profiles": r_f,
                       "algorithmic_tflops": fl_all * r_s / 1e12, "executed_tflops": fl_exec_s * r_s / 1e12,
                       "all_valid_tflops": fl_all * r_f / 1e12, "all_valid_frac_of_bf16_peak": fl_all * r_f / 1e12 / peak / world}
        model.set_eval_dtype("fp32")
    out["value"] = out["bf16"]["value"]
    if world == 1:
        # the Men-shaped train step (fwd + BCE + bwd + Adam, dropout 0.5): d = 256 is outside the fused training kernels,
        # so this is the per-op path — every projection / input gradient / weight gradient a tcgen05 3xTF32 GEMM
        try:
            out["train"] = men_train_rate(shape, args, dev, model.embeds.attr_table)
        except Exception as ex:  # noqa: BLE001
            out["train"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    out["valid_rows_per_user"] = rows_sparse
    out["per_kernel_evidence"] = "profiles/r02/README.md (ncu: tensor-pipe utilisation of rows_gemm_kernel<256> per epilogue)"
    return out


def men_train_rate(shape, args, dev, table):
    import carca_replication_b200 as cb
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep

    L, Bt = shape.seq_len, args.train_batch
    model = synth.build_model(shape, args.decoder, p=0.5).to(dev).train()
    model.embeds.set_attr_table(table)
    optim = cb.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
    batches = [{k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=177 + i).items()} for i in range(2)]
    step = GraphedTrainStep(model, optim, batches[0])
    n = max(3, args.steps // 2)
    for i in range(2):
        step(batches[i % 2])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        loss = step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": Bt / (ms * 1e-3), "unit": "seqs/s", "batch_per_gpu": Bt, "ms_per_step": ms, "final_loss": float(loss.item()),
            "path": "per-op kernels under autograd (tcgen05 3xTF32 GEMMs, attention fwd / bwd kernels), FusedAdam, one CUDA graph"}


def time_catalog_sharded(model, shape, args, dev, rank, world):
    """BASELINE configs[3] at N GPUs: the item table sharded over the ranks in contiguous id ranges; every rank scores
    ALL `--catalog-users` users against its shard, one all-reduce of the per-user rank counts (catalog.py)."""
    import torch.distributed as dist

    from carca_replication_b200 import catalog, synth

    Bc = args.catalog_users
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, Bc, seed=4242).items()}       # same users on every rank
    prof = (b["p_x"], None, b["p_c"])
    pos, ctx = b["o_x"][:, 0].contiguous(), b["o_c"][:, 0].contiguous()
    ranks = catalog.catalog_ranks(model, prof, pos, ctx)
    ref = catalog.catalog_ranks(model, prof, pos, ctx, shard=(1, shape.n_items), reduce=False)   # whole table, this rank
    same = bool(torch.equal(ranks, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    n = 3
    for _ in range(n):
        ranks = catalog.catalog_ranks(model, prof, pos, ctx)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": Bc / (ms * 1e-3), "unit": "users/s", "users": Bc, "items": shape.n_items - 1, "ms": ms, "n_gpus": world,
            "scores_per_s": Bc * (shape.n_items - 1) / (ms * 1e-3), "sharded_equals_unsharded": same,
            "scaling": "strong (the item table is split; every rank encodes all users)",
            "collective": "one all-reduce of [users] int32 rank counts"}



def time_catalog(model, shape, args, dev):
    """Full-catalog scoring (BASELINE configs[3]) on this GPU's item shard (all items at N = 1): users/s of
    catalog_ranks = encode + score every item + rank count, device-timed."""
    from carca_replication_b200 import catalog, synth

    Bc = args.catalog_users
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, Bc, seed=4242).items()}
    prof = (b["p_x"], None, b["p_c"])
    pos, ctx = b["o_x"][:, 0].contiguous(), b["o_c"][:, 0].contiguous()
    catalog.catalog_ranks(model, prof, pos, ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 3
    for _ in range(n):
        ranks = catalog.catalog_ranks(model, prof, pos, ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    r = ranks.double()
    # the same ranks through the score-matrix path (scores written to HBM, then counted): the round-1 formulation
    ref = catalog.catalog_ranks(model, prof, pos, ctx, use_tc=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        ref = catalog.catalog_ranks(model, prof, pos, ctx, use_tc=False)
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / n
    return {"value": Bc / (ms * 1e-3), "unit": "users/s", "users": Bc, "items": shape.n_items - 1, "ms": ms,
            "scores_per_s": Bc * (shape.n_items - 1) / (ms * 1e-3), "hr10": float((r < 10).double().mean().item()),
            "mean_rank": float(r.mean().item()),
            "path": "packed-rows fp32 encoder + tcgen05 catalog kernel (scores, softmax, sigmoid and rank comparison in "
                    "the GEMM epilogue; no [users, items] score matrix)",
            "score_matrix_path": {"scores_per_s": Bc * (shape.n_items - 1) / (ms_ref * 1e-3), "ms": ms_ref,
                                  "ranks_equal": float((ranks == ref).double().mean().item())}}


def time_dense_attrs(model, table, shape, args, dev, pk):
    """The reference API as it is: dense [B, N, A] attribute tensors per batch (src/carca.py:86, src/data.py:119-131)
    instead of the device-resident item -> attribute table.  These batches are 3.9 MB per Beauty user, so the step is
    bound by reading them (26,028 B per position); it runs on the per-op kernels (feature projection as a
    tcgen05 GEMM over the dense rows).  Reference batch size."""
    from carca_replication_b200 import ops, synth

    B = args.cpu_batch
    bs = []
    for i in range(2):
        b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=7000 + i).items()}
        b["p_a"], b["o_a"] = table.gather_dense(b["p_x"]).to(dev), table.gather_dense(b["o_x"]).to(dev)
        bs.append(b)
    acc4 = torch.zeros(4, dtype=torch.float64, device=dev)

    def step(b):
        y = model.forward(profile=(b["p_x"], b["p_a"], b["p_c"]), targets=[(b["o_x"], b["o_a"], b["o_c"])])
        ops.eval_metrics_(acc4, y, b["y_true"], b["o_x"], 10)
        return y

    n = max(3, args.steps // 2)
    with torch.no_grad():
        for b in bs:
            step(b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            step(bs[i % 2])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    byt = (shape.seq_len + shape.n_targets) * (4 * shape.n_attrs + 8 + 4 * shape.n_ctx)
    gbs = byt * B / (ms * 1e-3) / 1e9
    return {"value": B / (ms * 1e-3), "unit": "users/s", "batch": B, "ms_per_step": ms,
            "bytes_per_user": byt, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
            "what": "CARCA.forward given the reference's dense [B, N, A] attribute tensors (device-resident, fp32), per-op "
                    "kernels + eval metrics; bound: reading the attribute rows"}


def time_device_pipeline(model, shape, args, dev):
    """evaluate() fed by the device-side loader (carca_replication_b200/device_data.py): windows, sampled
    negatives and labels are built on the GPU from a resident interaction log, so no batch crosses PCIe.
    The reference's host loader builds ~500 users/s."""
    import numpy as np

    import carca_replication_b200 as cb
    from carca_replication_b200.device_data import DeviceInteractions, DeviceLoader

    rng = np.random.default_rng(99)
    U = shape.n_users
    lens = np.clip(np.rint(rng.lognormal(1.8, 0.7, size=U)), 4, 300).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    items = rng.integers(1, shape.n_items, size=int(rowptr[-1])).astype(np.int32)
    ctx = rng.random((int(rowptr[-1]), shape.n_ctx), dtype=np.float32)
    log = DeviceInteractions(torch.from_numpy(rowptr), torch.from_numpy(items), torch.from_numpy(ctx)).to(dev)
    loader = DeviceLoader(log, shape.n_items, shape.seq_len, shape.n_targets - 1, "test", batch_size=args.batch)
    for _ in range(3):                  # as in a training run's validation passes: by the third epoch every batch shape
        cb.evaluate(model, loader, dev, 10)   # of the loader (full and trailing) is replayed as a captured graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    hr, ndcg, loss = cb.evaluate(model, loader, dev, 10)
    e1.record()
    torch.cuda.synchronize()
    n = int(loader.users.numel())
    ms = e0.elapsed_time(e1)
    return {"value": n / (ms * 1e-3), "unit": "users/s", "users": n, "ms": ms, "hr10": hr,
            "what": "evaluate(model, DeviceLoader(...)) over every test user of a Beauty-sized log: batch "
                    "construction + forward + loss + metrics on the device, one D2H read (fourth pass over the loader: "
                    "evaluate() replays repeated batch shapes as CUDA graphs)"}


def time_train(shape, args, dev, table):
    """fwd + BCE + bwd + Adam seqs/s (src/train.py:84-97), dropout 0.5, at the reference batch size: the whole
    step body replayed as one CUDA graph (carca_replication_b200/graph.py).  Default path: the fused training
    kernels (csrc/fused_train.cuh, embeddings folded in) + FusedAdam; for comparison the same step with the stock
    torch.optim.Adam, eagerly, on the per-op kernels, and at a large batch (throughput regime)."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import _native as N
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep

    L = shape.seq_len
    loss_fn = cb.BinaryCrossEntropy()

    def setup(Bt, optimizer="fused", fused_kernels=True):
        model = synth.build_model(shape, args.decoder, p=0.5).to(dev).train()
        model.embeds.set_attr_table(table)
        model.use_fused_train = fused_kernels
        if optimizer == "fused":
            optim = cb.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
        else:
            optim = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), capturable=optimizer == "capturable")
        batches = [{k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=77 + i).items()} for i in range(4)]
        return model, optim, batches

    def eager_step(model, optim, b):
        o_x, o_c = b["o_x"], b["o_c"]
        optim.zero_grad()
        y = model.forward(profile=(b["p_x"], None, b["p_c"]),
                          targets=[(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = loss_fn.forward(y, b["y_true"], cb.get_mask(o_x))
        loss.backward()
        optim.step()
        return loss.detach()

    def clock(fn, n):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n0 = N.lib().carca_launch_count()
        e0.record()
        for i in range(n):
            loss = fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, float(loss.item()), (N.lib().carca_launch_count() - n0) / n

    def graphed(Bt, n, **kw):
        model, optim, batches = setup(Bt, **kw)
        step = GraphedTrainStep(model, optim, batches[0])
        ms, loss, _ = clock(lambda i: step(batches[i % 4]), n)
        return ms, loss

    n = max(args.steps, 5)
    Bt, Bl = args.train_batch, 4096
    ms_g, loss_g = graphed(Bt, 4 * n)
    ms_s, _ = graphed(Bt, 4 * n, optimizer="capturable")
    ms_p, _ = graphed(Bt, 2 * n, optimizer="capturable", fused_kernels=False)
    ms_l, _ = graphed(Bl, 2 * n)
    ms_lp, _ = graphed(Bl, n, optimizer="capturable", fused_kernels=False)
    model, optim, batches = setup(Bt, optimizer="stock")
    ms_e, loss_e, launches = clock(lambda i: eager_step(model, optim, batches[i % 4]), n)
    # algorithmic work of one train step per sequence: 3x the forward FLOPs (fwd + 2x in bwd) with
    # 2L targets; all L positions counted as the reference executes them
    d, g, nb, C = shape.d, shape.g, shape.n_blocks, shape.n_ctx
    T = 2 * L
    fwd = 2 * (L + T) * (g + d) * d + nb * (10 * L * d * d + 4 * L * L * d)
    fwd += (2 * T * d * d + 4 * L * d * d + 4 * T * L * d + 2 * T * d) if args.decoder == "ca" else 2 * T * d
    flops = 3.0 * fwd
    return {"value": Bt / (ms_g * 1e-3), "unit": "seqs/s", "batch_per_gpu": Bt, "ms_per_step": ms_g,
            "mode": "whole step (zero_grad, fwd, BCE, bwd, Adam) replayed as one CUDA graph; fused training kernels "
                    "(active positions only, embeddings folded in) + FusedAdam", "final_loss": loss_g,
            "optimizer": "carca_replication_b200.FusedAdam (one launch; torch.optim.Adam update rule)", "dropout": 0.5,
            "algorithmic_tflops": flops * Bt / (ms_g * 1e-3) / 1e12,
            "flops_per_seq": flops,
            "stock_adam": {"value": Bt / (ms_s * 1e-3), "ms_per_step": ms_s,
                           "what": "same graph with torch.optim.Adam(capturable=True)"},
            "per_op_kernels": {"value": Bt / (ms_p * 1e-3), "ms_per_step": ms_p,
                               "what": "same graph on the per-op kernels (use_fused_train=False), stock Adam"},
            "eager": {"value": Bt / (ms_e * 1e-3), "ms_per_step": ms_e, "final_loss": loss_e,
                      "launches_of_ours_per_step": launches, "what": "no graph, stock Adam"},
            "large_batch": {"value": Bl / (ms_l * 1e-3), "batch_per_gpu": Bl, "ms_per_step": ms_l,
                            "algorithmic_tflops": flops * Bl / (ms_l * 1e-3) / 1e12,
                            "per_op_kernels": {"value": Bl / (ms_lp * 1e-3), "ms_per_step": ms_lp}}}


def time_train_dp(shape, args, dev, table, rank, world, peer=True):
    """Data-parallel train step: every rank runs the fused step on its own `--train-batch` users
    (weak scaling), the flat gradient buffer is all-reduced in place over NCCL (one collective per step, plus the
    two BCE partial sums), FusedAdam steps the replicated weights.  Timed on the device, max over ranks; as one
    CUDA graph per rank when NCCL capture works, eagerly otherwise."""
    import torch.distributed as dist

    import carca_replication_b200 as cb
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep
    from carca_replication_b200.parallel import UserDataParallel

    L, Bt = shape.seq_len, args.train_batch
    model = synth.build_model(shape, args.decoder, p=0.5).to(dev).train()
    model.embeds.set_attr_table(table)
    dp = UserDataParallel(model, peer_allreduce=peer)
    optim = cb.FusedAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))
    batches = [{k: v.to(dev) for k, v in synth.make_train_batch(shape, Bt, seed=77 + 10 * rank + i).items()}
               for i in range(4)]

    def eager_step(b):
        o_x, o_c = b["o_x"], b["o_c"]
        optim.zero_grad()
        y = model.forward(profile=(b["p_x"], None, b["p_c"]),
                          targets=[(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = dp.loss_fn.forward(y, b["y_true"], cb.get_mask(o_x))
        loss.backward()
        optim.step()
        return loss.detach()

    def clock(fn, n):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            loss = fn(i)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(loss.item())

    n = max(args.steps, 5)
    ms_e, loss_e = clock(lambda i: eager_step(batches[i % 4]), n)
    peer_on = dp.peer is not None
    out = {"unit": "seqs/s", "batch_per_gpu": Bt, "global_batch": Bt * world, "scaling": "weak",
           "collective": ("one peer-memory all-reduce kernel (csrc/peer.cu, NVLink loads / stores on CUDA-IPC mapped "
                          "buffers, zero-copy: the flat gradient buffer lives in the communication buffer) per step + "
                          "2 floats of BCE sums the same way; no NCCL kernel in the step") if peer_on else
                         "one in-place NCCL all-reduce of the flat gradient buffer per step + 2 floats of BCE sums",
           "eager": {"value": world * Bt / (ms_e * 1e-3), "ms_per_step": ms_e, "final_loss": loss_e}}
    ok = torch.ones(1, device=dev)
    try:
        step = GraphedTrainStep(model, optim, batches[0], loss_fn=dp.loss_fn)
        ms_g, loss_g = clock(lambda i: step(batches[i % 4]), 4 * n)
    except Exception as ex:  # noqa: BLE001 -- NCCL capture is best effort; the eager number stands
        ok.zero_()
        out["graph_error"] = f"{type(ex).__name__}: {ex}"[:200]
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() > 0:
        out.update({"value": world * Bt / (ms_g * 1e-3), "ms_per_step": ms_g, "final_loss": loss_g,
                    "mode": "whole data-parallel step (incl. the gradient all-reduce) replayed as one CUDA graph per rank"})
    else:
        out.update({"value": out["eager"]["value"], "ms_per_step": ms_e, "mode": "eager"})
    if peer_on:
        out["peer_allreduce_timed_out"] = bool(dp.peer.timed_out() or dp._peer_small.timed_out())
    from carca_replication_b200 import ops as _ops

    _ops.FLAT_GRAD_ALLOC = None
    if peer and peer_on:      # the same step with NCCL as the collective, for the comparison
        nccl = time_train_dp(shape, args, dev, table, rank, world, peer=False)
        out["nccl_allreduce"] = {k: nccl[k] for k in ("value", "ms_per_step", "mode") if k in nccl}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
