import torch, time
dev="cuda"
big=torch.empty(200<<20,dtype=torch.uint8).pin_memory()
dst=torch.empty(200<<20,dtype=torch.uint8,device=dev)
for mb in (3,6,12,25,50,100,200):
    n=mb<<20
    for _ in range(3): dst[:n].copy_(big[:n],non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    reps=20
    for _ in range(reps): dst[:n].copy_(big[:n],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/reps
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record(); dst[:n].copy_(big[:n],non_blocking=True); e.record(); torch.cuda.synchronize()
    print(f"{mb:4d} MB: {n/dt/1e9:6.1f} GB/s back-to-back ({dt*1e6:7.1f} us each); single copy by events {s.elapsed_time(e)*1e3:7.1f} us = {n/(s.elapsed_time(e)*1e-3)/1e9:5.1f} GB/s")
