"""HBM probe: pure-write, pure-read and copy bandwidth (torch, CUDA events) -- context for HBM-bound kernels."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
w = t(lambda: a.fill_(1.0)); print(f"write  {4 * n / w / 1e6:.0f} GB/s")
r = t(lambda: a.sum()); print(f"read   {4 * n / r / 1e6:.0f} GB/s")
c = t(lambda: b.copy_(a)); print(f"copy   {8 * n / c / 1e6:.0f} GB/s (read+write)")
