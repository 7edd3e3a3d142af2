import torch, time
n=200_000_000
h=torch.empty(n,dtype=torch.int32,pin_memory=True); d=torch.empty(n,dtype=torch.int32,device='cuda')
h2=torch.empty(n//2,dtype=torch.int32,pin_memory=True); d2=torch.empty(n//2,dtype=torch.int32,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for _ in range(2): d.copy_(h,non_blocking=True)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print(f"H2D alone: {n*4/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms for 800 MB)")
t=time.perf_counter()
for _ in range(5): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print(f"D2H alone: {n*2/dt/1e9:.1f} GB/s")
t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print(f"both: H2D 800 MB + D2H 400 MB in {dt*1e3:.2f} ms -> H2D {n*4/dt/1e9:.1f} GB/s")
