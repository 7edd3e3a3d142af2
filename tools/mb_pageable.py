import torch, time, numpy as np
d = torch.empty(1_000_000_000, dtype=torch.int32, device="cuda")  # 4 GB
d.fill_(7)
torch.cuda.synchronize()
h_pin = torch.empty_like(d, device="cpu", pin_memory=True)
h_pag = torch.empty_like(d, device="cpu")
h_pag.fill_(0)
for name, h in (("pinned", h_pin), ("pageable", h_pag)):
    for _ in range(2):
        t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"D2H {name}: {4/dt:.1f} GB/s")
a = h_pin.numpy(); b = h_pag.numpy()
t0 = time.perf_counter(); np.copyto(b, a); dt = time.perf_counter() - t0
print(f"host memcpy 1 thread: {4/dt:.1f} GB/s")
for name, h in (("pinned", h_pin), ("pageable", h_pag)):
    t0 = time.perf_counter(); d.copy_(h); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D {name}: {4/dt:.1f} GB/s")
