"""Aggregate host <-> device copy bandwidth of the box with all GPUs copying at once (pinned host memory, both directions) - the ceiling
of any end-to-end rate that moves B bytes per solve through host memory: solves/s <= bandwidth / B."""
import time

import torch

G = torch.cuda.device_count()
MB = 256
res = {}
for gpus in sorted({1, 2, 4, G}):
    if gpus > G:
        continue
    hin = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(gpus)]
    hout = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(gpus)]
    din = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{g}") for g in range(gpus)]
    dout = [torch.ones(MB << 20, dtype=torch.uint8, device=f"cuda:{g}") for g in range(gpus)]
    s1 = [torch.cuda.Stream(device=g) for g in range(gpus)]
    s2 = [torch.cuda.Stream(device=g) for g in range(gpus)]
    for mode in ("h2d", "d2h", "both"):
        best = 1e9
        for _ in range(4):
            for g in range(gpus):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for g in range(gpus):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1[g]):
                        din[g].copy_(hin[g], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2[g]):
                        hout[g].copy_(dout[g], non_blocking=True)
            for g in range(gpus):
                torch.cuda.synchronize(g)
            best = min(best, time.perf_counter() - t0)
        nbytes = gpus * (MB << 20) * (2 if mode == "both" else 1)
        print(f"{gpus} GPU(s) {mode:5s}: {nbytes / best / 1e9:7.1f} GB/s aggregate ({nbytes / best / 1e9 / gpus:.1f} per GPU)", flush=True)
