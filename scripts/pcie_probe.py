#!/usr/bin/env python
"""Developer tool (GPU box): what the host link gives - pinned H2D alone, D2H alone, both at once
(the ceiling of the streamed host path), for flat copies and for the padded-row 2-D copies the
library issues.  Prints GB/s per direction."""
import time

import torch

n = 2 * 1024**3
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, chunks=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = n // chunks
    for k in range(chunks):
        sl = slice(k * c, (k + 1) * c)
        if up:
            with torch.cuda.stream(s1):
                d_a[sl].copy_(h_in[sl], non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out[sl].copy_(d_b[sl], non_blocking=True)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9


for name, up, down in (("H2D alone", True, False), ("D2H alone", False, True), ("both at once", True, True)):
    for chunks in (1, 32, 128):
        run(up, down, chunks)
        r = max(run(up, down, chunks) for _ in range(3))
        print(f"{name:14s} chunks={chunks:4d}: {r:6.1f} GB/s per direction")
