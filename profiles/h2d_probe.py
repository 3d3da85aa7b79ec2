import torch, time
x = torch.randn(65536, 784).pin_memory()
d = torch.empty_like(x, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: d.copy_(x, non_blocking=True))
print("H2D one copy 205MB: %.3f ms %.1f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
for nch in (2, 4, 8, 16):
    cs = 65536 // nch
    def f():
        for i in range(nch):
            d[i * cs:(i + 1) * cs].copy_(x[i * cs:(i + 1) * cs], non_blocking=True)
    ms = t(f)
    print("H2D %d chunks: %.3f ms %.1f GB/s" % (nch, ms, x.numel() * 4 / ms / 1e6))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def g():
    h = 32768
    with torch.cuda.stream(s1): d[:h].copy_(x[:h], non_blocking=True)
    with torch.cuda.stream(s2): d[h:].copy_(x[h:], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = t(g)
print("H2D 2 streams: %.3f ms %.1f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
o = torch.empty(65536).pin_memory(); od = torch.randn(65536, device="cuda")
ms = t(lambda: o.copy_(od, non_blocking=True))
print("D2H 256KB: %.4f ms" % ms)
