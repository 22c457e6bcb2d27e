#!/usr/bin/env python
"""PCIe microbenchmark: pinned H2D alone, D2H alone, and both at once (is the link full duplex on this box, and does it
depend on which streams the two copies are issued on?)."""
import json
import sys
import torch
n = 1 << 29
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
streams = [torch.cuda.Stream() for _ in range(NS)]


def run(s1, s2, reps=2, chunk=n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in (s1, s2):
        if s is not None:
            s.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        for a in range(0, n, chunk):
            if s1 is not None:
                with torch.cuda.stream(s1):
                    d_a[a:a + chunk].copy_(h_in[a:a + chunk], non_blocking=True)
            if s2 is not None:
                with torch.cuda.stream(s2):
                    h_out[a:a + chunk].copy_(d_b[a:a + chunk], non_blocking=True)
    for s in (s1, s2):
        if s is not None:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / 1e9 / (e0.elapsed_time(e1) / 1e3)


run(streams[0], streams[1], 1)
res = dict(async_engines=torch.cuda.get_device_properties(0).__repr__(), h2d_alone_gbs=round(run(streams[0], None), 1), d2h_alone_gbs=round(run(None, streams[1]), 1))
res["duplex_matrix_gbs_each_direction(row=h2d stream, col=d2h stream)"] = [[round(run(streams[i], streams[j]), 1) if i != j else None for j in range(NS)] for i in range(NS)]
res["duplex_128MiB_chunks"] = round(run(streams[0], streams[1], 2, 1 << 27), 1)
print(json.dumps(res))
