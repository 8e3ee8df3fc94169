import torch, time
n = 4350800 // 4
h = torch.empty(n).pin_memory(); d = torch.empty(n, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, "%.3f ms" % ms, "%.1f GB/s" % (n * 4 / ms / 1e6))
# latency of a tiny copy + sync
t = torch.empty(3, device="cuda"); hp = torch.empty(3).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100):
    hp.copy_(t, non_blocking=True); torch.cuda.synchronize()
print("tiny d2h + sync: %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
