"""Item-table-sharded full-catalog ranking over the ranks of a torchrun job (NCCL), checked on rank 0 against
the unsharded computation.   python -m torch.distributed.run --nproc-per-node N tools/catalog_multi.py [users]"""
import os, sys, json; sys.path.insert(0, '.')
import torch, torch.distributed as dist
from carca_replication_b200 import catalog, synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
dev = torch.device("cuda", local)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
shape = synth.BEAUTY
model = synth.build_model(shape, "ca", p=0.5).to(dev).eval()
model.embeds.set_attr_table(synth.make_attr_table(shape).to(dev))
b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=4242).items()}
prof = (b["p_x"], None, b["p_c"]); pos, ctx = b["o_x"][:, 0].contiguous(), b["o_c"][:, 0].contiguous()
ranks = catalog.catalog_ranks(model, prof, pos, ctx)                      # sharded over the group
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize(); e0.record()
for _ in range(3): ranks = catalog.catalog_ranks(model, prof, pos, ctx)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
full = catalog.catalog_ranks(model, prof, pos, ctx, shard=(1, shape.n_items)) if world == 1 else None
if rank == 0:
    # unsharded reference on rank 0: every item scored locally, no all-reduce
    lo_hi = (1, shape.n_items)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    import carca_replication_b200._native as N
    with torch.no_grad():
        for u0 in range(0, B, 2048):
            u1 = min(B, u0 + 2048)
            pr = (b["p_x"][u0:u1], None, b["p_c"][u0:u1])
            y_pos = model.forward(pr, [(pos[u0:u1].unsqueeze(1), None, ctx[u0:u1].unsqueeze(1))])[:, 0].contiguous()
            y = catalog.score_items(model, pr, ctx[u0:u1].contiguous(), *lo_hi)
            N.call("carca_catalog_rank_count", N.i32p(counts[u0:u1]), N.f32p(y), y.stride(0), N.f32p(y_pos),
                   N.i32p(pos[u0:u1].contiguous()), 1, u1 - u0, shape.n_items - 1, N.stream())
    same = bool(torch.equal(counts, ranks))
    print(json.dumps({"world": world, "users": B, "items": shape.n_items - 1, "ms": float(ms.item()),
                      "users_per_s": B / float(ms.item()) * 1e3, "sharded_equals_unsharded": same,
                      "hr10": float((ranks < 10).double().mean().item())}), flush=True)
if world > 1: dist.destroy_process_group()
