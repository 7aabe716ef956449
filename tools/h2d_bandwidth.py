"""Pinned host -> device copy bandwidth at the size of one cfg2 batch (30.7 MB) and at 256 MB:
the ceiling of bench.py's end-to-end number (one batch of emissions crosses PCIe per step)."""
import torch
for mb in (30.72, 256.0):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("H2D pinned %.1f MB: %.3f ms, %.1f GB/s -> at most %.0f k utt/s end to end at cfg2" % (mb, ms, mb / ms, 256 / (30.72 / (mb / ms)) ))
