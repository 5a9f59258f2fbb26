import torch, time
x = torch.empty(99532800 // 4, dtype=torch.float32, device="cuda")
h = torch.empty(99532800 // 4, dtype=torch.float32).pin_memory()
for n in (1, 3, 6):
    chunks_d = x.chunk(n); chunks_h = h.chunk(n)
    for _ in range(3):
        for a, b in zip(chunks_h, chunks_d): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        for a, b in zip(chunks_h, chunks_d): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"D2H 99.5 MB in {n} chunk(s): {dt*1e3:.3f} ms = {99.5328/dt/1e3:.1f} GB/s")
