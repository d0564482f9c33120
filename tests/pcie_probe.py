"""Tuning aid (not a test): pinned-memory PCIe bandwidth of the GPU box, the bound of bench.py's e2e leg (33 MB D2H + 9.4 MB H2D per step).
    python tests/pcie_probe.py   -> one JSON line"""
import json
import time

import torch

dev = 'cuda:0'
out = {}
for mb in (8, 33, 128):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (('h2d', lambda: d.copy_(h, non_blocking=True)), ('d2h', lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        out[f'{name}_{mb}MB_GBps'] = round(n * 10 / (time.perf_counter() - t) / 1e9, 2)
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    out[f'duplex_{mb}MB_GBps_each'] = round(n * 10 / (time.perf_counter() - t) / 1e9, 2)
print(json.dumps(out))
