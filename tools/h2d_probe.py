#!/usr/bin/env python
"""Raw pinned host->device bandwidth of this box (the ceiling of bench.py's e2e number): one stream, cudaMemcpyAsync of
the bench step's frame bytes, CUDA-event timed."""
import torch
n = 32 * 2 * 1080 * 1920
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunks in (1, 8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 20
    for _ in range(reps):
        for c in range(chunks):
            a, b = c * n // chunks, (c + 1) * n // chunks
            d[a:b].copy_(h[a:b], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("H2D %d chunk(s): %.1f MB in %.3f ms = %.1f GB/s  -> ceiling %.0f pairs/s" % (chunks, n / 1e6, ms, n / ms / 1e6, 32 / ms * 1e3))
