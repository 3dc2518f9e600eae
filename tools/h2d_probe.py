import torch, time
dev="cuda"
n=18284544
hs=[torch.empty(n,dtype=torch.uint8).pin_memory() for _ in range(8)]
ds=[torch.empty(n,dtype=torch.uint8,device=dev) for _ in range(2)]
s=torch.cuda.Stream()
for rep in range(3):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        e0.record(s)
        for i in range(40):
            ds[i%2].copy_(hs[i%8],non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/40
    print(f"H2D {n/1e6:.1f} MB: {ms:.3f} ms/copy = {n/ms/1e6:.1f} GB/s")
# big copy
h=torch.empty(1<<30,dtype=torch.uint8).pin_memory(); d=torch.empty(1<<30,dtype=torch.uint8,device=dev)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); d.copy_(h,non_blocking=True); e1.record(); torch.cuda.synchronize()
print("1GiB H2D GB/s", (1<<30)/e0.elapsed_time(e1)/1e6)
