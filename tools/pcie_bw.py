#!/usr/bin/env python
"""Host<->device copy bandwidth of the box (pinned and pageable), the ceiling of the host-pointer (e2e) numbers."""
import time
import torch
n = 512 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
hq = torch.empty(n, dtype=torch.uint8)
hq.fill_(1)
def bw(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1):
        d[: n // 2].copy_(hp[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        hp[n // 2:].copy_(d[n // 2:], non_blocking=True)
print("H2D pinned   %.1f GB/s" % bw(lambda: d.copy_(hp, non_blocking=True)))
print("D2H pinned   %.1f GB/s" % bw(lambda: hp.copy_(d, non_blocking=True)))
print("duplex (sum) %.1f GB/s" % bw(duplex))
print("H2D pageable %.1f GB/s" % bw(lambda: d.copy_(hq)))
print("D2H pageable %.1f GB/s" % bw(lambda: hq.copy_(d)))
import numpy as np
a = np.ones(n, np.uint8); b = np.empty_like(a)
t0 = time.perf_counter(); 
for _ in range(3): np.copyto(b, a)
print("host memcpy (1 thread) %.1f GB/s" % (3 * n / (time.perf_counter() - t0) / 1e9))
