"""Dev tool: what a read-only pass over the bench input (65536 x 784 fp32, 205 MB) costs with library kernels."""
import torch
dev = torch.device("cuda:0")
x = torch.randn(65536, 784, device=dev)
y = torch.empty_like(x)
big = torch.empty(64 * 1024 * 1024, device=dev)   # 256 MB L2 flush


def timeit(fn, n=20, flush=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush:
            big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


nbytes = x.numel() * 4
for name, fn, traffic in (("sum", lambda: x.sum(), nbytes), ("sum_dim1", lambda: x.sum(dim=1), nbytes),
                          ("absmax", lambda: x.abs().max(), 0), ("amax", lambda: x.amax(), nbytes),
                          ("copy", lambda: y.copy_(x), 2 * nbytes), ("mul_", lambda: y.mul_(1.5), 2 * nbytes),
                          ("zero_", lambda: y.zero_(), nbytes)):
    for flush in (False, True):
        ms = timeit(fn, flush=flush)
        print("PROBE %-9s flush=%d %.4f ms  %.0f GB/s" % (name, flush, ms, traffic / ms / 1e6))
