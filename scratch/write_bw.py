"""Pure-write bandwidth of the box (cudaMemset and a torch fill), for the z sweep's write roofline."""
import torch, time
n = 4 * 1024**3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in [("memset(zero_)", lambda: a.zero_()), ("fill_(7)", lambda: a.fill_(7)), ("int64 fill", lambda: a.view(torch.int64).fill_(5))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:16s} {n / ms / 1e6:8.1f} GB/s")
b = torch.empty(n // 2, dtype=torch.uint8, device="cuda"); c = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
for _ in range(3): c.copy_(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): c.copy_(b)
e1.record(); torch.cuda.synchronize()
print(f"copy (r+w)       {n / (e0.elapsed_time(e1) / 10) / 1e6:8.1f} GB/s")
