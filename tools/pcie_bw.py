import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2): d.copy_(h, non_blocking=True); h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 5 * n / (time.perf_counter() - t) / 1e9)
t = time.perf_counter()
for _ in range(5): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("D2H GB/s", 5 * n / (time.perf_counter() - t) / 1e9)
t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("both directions GB/s each", 5 * n / (time.perf_counter() - t) / 1e9)
