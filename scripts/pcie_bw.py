import torch, time
q = torch.empty(30_000_000, dtype=torch.float32).pin_memory()
r = torch.empty(10_000_000, dtype=torch.float32).pin_memory()
dq = torch.empty_like(q, device='cuda'); dr = torch.empty_like(r, device='cuda')
for name, a, b in (("h2d 120MB", dq, q), ("d2h 40MB", r, dr)):
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print(name, f"{dt*1e3:.3f} ms", f"{a.numel()*4/dt/1e9:.1f} GB/s")
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): dq.copy_(q, non_blocking=True)
    with torch.cuda.stream(s2): r.copy_(dr, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print("both directions concurrently", f"{dt*1e3:.3f} ms")
