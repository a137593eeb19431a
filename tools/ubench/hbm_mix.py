import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / 1e3
N = 1 << 28   # 1 GiB of float32 = 268M elements -> use bytes
x = torch.empty(N, dtype=torch.float32, device="cuda")
y = torch.empty(N, dtype=torch.float32, device="cuda")
z = torch.empty(2 * N, dtype=torch.float32, device="cuda")
s = t(lambda: y.copy_(x)); print("copy      %.0f GB/s (r+w)" % (2 * N * 4 / s / 1e9))
s = t(lambda: y.fill_(1.0)); print("fill      %.0f GB/s (w)" % (N * 4 / s / 1e9))
s = t(lambda: x.sum()); print("sum       %.0f GB/s (r)" % (N * 4 / s / 1e9))
s = t(lambda: torch.cat([x, x], out=z)); print("cat 1r:2w %.0f GB/s (r+w, counting x once)" % (3 * N * 4 / s / 1e9))
