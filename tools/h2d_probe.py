#!/usr/bin/env python3
"""Host->device copy rate of this box from pinned memory (torch) -- what bounds round 1 of the witness-in e2e."""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (128, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("H2D %4d MiB pinned: %.2f ms  %.1f GB/s" % (mb, dt * 1e3, (mb << 20) / dt / 1e9))
    t0 = time.perf_counter()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("D2H %4d MiB pinned: %.2f ms  %.1f GB/s" % (mb, dt * 1e3, (mb << 20) / dt / 1e9))
