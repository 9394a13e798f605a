#!/usr/bin/env python
"""H2D / D2H bandwidth from pinned memory with 1, 2 and 4 concurrent copy streams and different chunk sizes (the host-pointer trace
calls move 48 bytes per ray: is one stream of 12.6 MB chunks all the link gives?).  usage (GPU box): python tools/pcie_streams.py"""
import time
import torch

dev = torch.device("cuda:0")
total = 512 << 20
host = torch.empty(total, dtype=torch.uint8).pin_memory()
host.fill_(1)
devb = torch.empty(total, dtype=torch.uint8, device=dev)
for direction in ("h2d", "d2h"):
    for chunk_mb in (4, 12, 64, 512):
        for n_streams in (1, 2, 4):
            chunk = chunk_mb << 20
            streams = [torch.cuda.Stream() for _ in range(n_streams)]
            best = 0.0
            for rep in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for i, off in enumerate(range(0, total, chunk)):
                    with torch.cuda.stream(streams[i % n_streams]):
                        if direction == "h2d":
                            devb[off:off + chunk].copy_(host[off:off + chunk], non_blocking=True)
                        else:
                            host[off:off + chunk].copy_(devb[off:off + chunk], non_blocking=True)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
                best = max(best, total / dt * 1e-9)
            print(direction, "chunk %3d MB" % chunk_mb, "streams", n_streams, "%.1f GB/s" % best, flush=True)
