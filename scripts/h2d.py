"""Pinned host -> device copy rate for one step's worth of float32 pixels (the floor of the e2e figure)."""
import torch
x = torch.empty((70000, 784), dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"H2D 219.5 MB: {ms:.3f} ms = {x.numel() * 4 / ms / 1e6:.1f} GB/s -> e2e floor {70000 / ms / 1e3:.2f} M images/s before any kernel")
