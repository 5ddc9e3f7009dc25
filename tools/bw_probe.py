import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")   # 1 GiB
y = torch.empty_like(x)
ms = t(lambda: x.fill_(1)); print("fill   1GiB: %.3f ms  %.0f GB/s (pure write)" % (ms, x.numel() / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy   1GiB: %.3f ms  %.0f GB/s (read+write)" % (ms, 2 * x.numel() / ms / 1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print("sum    1GiB: %.3f ms  %.0f GB/s (pure read)" % (ms, x.numel() / ms / 1e6))
z = torch.empty(308 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ms = t(lambda: z.fill_(1)); print("fill 308MiB: %.3f ms  %.0f GB/s" % (ms, z.numel() / ms / 1e6))
