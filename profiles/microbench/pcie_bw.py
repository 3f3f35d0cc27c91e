"""Pinned host <-> device copy bandwidth of this box (context for bench.py's e2e leg): python profiles/microbench/pcie_bw.py"""
import time
import torch
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(f"{name}: {3 * n / (time.perf_counter() - t) / 1e9:.1f} GB/s (4 GiB pinned buffer, 3 copies)")
