import time, torch
n = 320 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms for 320 MiB)")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"duplex: h2d 320 MiB + d2h 160 MiB together in {dt * 1e3:.2f} ms")
