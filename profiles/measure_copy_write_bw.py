"""Write-only / copy / read bandwidth of the device through torch (used to decide whether ADPCM decode, 89 % stores,
is bounded by a lower write-only rate: it is not -- memset reaches 7.5 TB/s against 6.7 TB/s for a copy)."""
import torch
x = torch.empty(6 * 1024**3 // 4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print("memset  GB/s", x.numel() * 4 / ms / 1e6)
ms = t(lambda: x.fill_(1.5)); print("fill    GB/s", x.numel() * 4 / ms / 1e6)
ms = t(lambda: y.copy_(x)); print("copy    GB/s (r+w)", 2 * x.numel() * 4 / ms / 1e6)
ms = t(lambda: x.sum()); print("read    GB/s", x.numel() * 4 / ms / 1e6)
# 8:1 write:read like ADPCM: y[:] = small expanded
s = torch.empty(x.numel() // 8, dtype=torch.float32, device="cuda")
ms = t(lambda: torch.add(x[: x.numel() // 9], 1.0, out=y[: x.numel() // 9])); print("add r+w GB/s", 2 * (x.numel() // 9) * 4 / ms / 1e6)
