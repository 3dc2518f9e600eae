"""Development tool: the peer-memory all-reduce (csrc/peer.cu) against NCCL on the same tensors, eager and inside a CUDA
graph, with timings.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from carca_replication_b200.parallel import PeerAllReduce  # noqa: E402


def clock(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    big = 5_430_000
    comm = PeerAllReduce(big + 64, dev)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    say(f"world {world}: available {comm.available} error {comm.error}")
    if not comm.available:
        return
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    for n in (1, 2, 3, 5, 64, 1000, 4097, 123_457, big):
        for rep in range(3):
            x = torch.randn(n, device=dev, generator=g)
            want = x.clone()
            dist.all_reduce(want)
            got = comm.all_reduce_(x.clone())
            torch.cuda.synchronize()
            err = (got - want).abs().max().item()
            # bit-identical on every rank
            ref0 = got.clone()
            dist.broadcast(ref0, src=0)
            same = bool((ref0 == got).all())
            flags = torch.tensor([err, 0.0 if same else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            if rep == 0 or flags[0].item() > 1e-4 or flags[1].item() > 0:
                say(f"n {n:9d} rep {rep}: max |peer - nccl| {flags[0].item():.2e}  ranks bit-equal {flags[1].item() == 0}  timed out {comm.timed_out()}")
    # zero-copy: the tensor lives in the communication buffer
    for n in (5, 4096, big):
        src = torch.randn(n, device=dev, generator=g)
        want = src.clone()
        dist.all_reduce(want)
        xb = comm.buffer(n)
        for rep in range(2):
            xb.copy_(src)
            comm.all_reduce_(xb)
            torch.cuda.synchronize()
            say(f"in place n {n:9d} rep {rep}: max |peer - nccl| {(xb - want).abs().max().item():.2e} timed out {comm.timed_out()}")
    # graph capture + replay
    x = torch.randn(big, device=dev, generator=g)
    src = x.clone()
    gr = torch.cuda.CUDAGraph()
    comm.all_reduce_(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        x.copy_(src)
        comm.all_reduce_(x)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    want = src.clone()
    dist.all_reduce(want)
    say(f"graph replay: max |peer - nccl| {(x - want).abs().max().item():.2e} timed out {comm.timed_out()}")
    for n in (() if os.environ.get("PEER_QUICK") else (256, 65_536, 1_000_000, big)):
        y = torch.randn(n, device=dev, generator=g)
        ms_p = clock(lambda: comm.all_reduce_(y))
        ms_n = clock(lambda: dist.all_reduce(y))
        gp, gn = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(gp):
            for _ in range(10):
                comm.all_reduce_(y)
        y.zero_()
        ms_pg = clock(gp.replay, n=10, warm=3) / 10
        try:
            with torch.cuda.graph(gn):
                for _ in range(10):
                    dist.all_reduce(y)
            ms_ng = clock(gn.replay, n=10, warm=3) / 10
        except Exception as ex:  # noqa: BLE001
            ms_ng = float("nan")
            say("nccl graph capture failed:", str(ex)[:100])
        yb = comm.buffer(n)
        gz = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gz):
            for _ in range(10):
                comm.all_reduce_(yb)
        yb.zero_()
        ms_zg = clock(gz.replay, n=10, warm=3) / 10
        say(f"n {n:9d}: zero-copy in a graph {ms_zg * 1e3:7.1f} us")
        say(f"n {n:9d} ({n * 4 / 1e6:6.2f} MB): eager peer {ms_p * 1e3:7.1f} us nccl {ms_n * 1e3:7.1f} us | in a graph peer {ms_pg * 1e3:7.1f} us "
            f"nccl {ms_ng * 1e3:7.1f} us  ({n * 4 * 2 * (world - 1) / world / (ms_pg * 1e-3) / 1e9:.0f} GB/s bus bandwidth)  timed out {comm.timed_out()}")
    if os.environ.get("PEER_TRAIN"):
        import argparse

        import bench
        from carca_replication_b200 import synth

        args = argparse.Namespace(train_batch=int(os.environ.get("PEER_TRAIN_BATCH", "256")), decoder="ca", steps=10)
        shape = synth.BEAUTY
        table = synth.make_attr_table(shape).to(dev)
        out = bench.time_train_dp(shape, args, dev, table, rank, world)
        say("train dp:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items() if k != "collective"})
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)            # (skip interpreter teardown: the mapped peer buffers go away with the process)


if __name__ == "__main__":
    main()
