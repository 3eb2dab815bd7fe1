import torch, time
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it
h2d = t(lambda: d.copy_(h, non_blocking=True)); d2h = t(lambda: h.copy_(d, non_blocking=True))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
bi = t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  simultaneous: %.1f + %.1f GB/s" % (n / h2d / 1e9, n / d2h / 1e9, n / bi / 1e9, n / bi / 1e9))
W, H = 3840, 2160
ideal = max(2 * W * H / (n / bi), 4 * W * H / (n / bi))
print("4K pair with map: full-duplex bound %.3f ms -> %.0f Mpix/s; serial H2D+D2H %.3f ms" % (ideal * 1e3, W * H / ideal / 1e6, (2 * W * H / (n / h2d) + 4 * W * H / (n / d2h)) * 1e3))
