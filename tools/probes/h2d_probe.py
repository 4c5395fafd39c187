import torch, time
x = torch.empty(1060864000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("pure pinned H2D: %.2f ms per 1.06 GB = %.1f GB/s" % (ms, x.numel() / ms / 1e6))
