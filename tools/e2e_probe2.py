#!/usr/bin/env python
"""Follow-up probe: the encoder's H2D pattern (two 256 MiB copies per window from two big pinned buffers, a sync per window)
replayed with torch alone, beside a D2H copy loop in another thread."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

n = 1 << 28
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
big = 1700 << 20
H1 = torch.empty(big, dtype=torch.uint8).pin_memory()
H2 = torch.empty(big, dtype=torch.uint8).pin_memory()
D1 = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
D2 = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
HO = torch.empty(big, dtype=torch.uint8).pin_memory()
DO = torch.empty(big, dtype=torch.uint8, device="cuda")
res = {}


class Loop(threading.Thread):
    def __init__(self, kind):
        super().__init__()
        self.kind, self.stop, self.bytes, self.t = kind, False, 0, 0.0

    def run(self):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        t0 = time.perf_counter()
        k = 0
        while not self.stop:
            with torch.cuda.stream(s):
                if self.kind == "small":
                    h_out.copy_(d_b, non_blocking=True)
                    self.bytes += n
                else:                              # the decoder's pattern: 128 MiB pieces into a big pinned buffer
                    a = (k * (1 << 27)) % (big - (1 << 27))
                    HO[a:a + (1 << 27)].copy_(DO[a:a + (1 << 27)], non_blocking=True)
                    self.bytes += 1 << 27
                    k += 1
            s.synchronize()
        self.t = time.perf_counter() - t0

    def gbs(self):
        return round(self.bytes / 1e9 / self.t, 1)


def enc_pattern(offset):
    s = torch.cuda.Stream()
    t0 = time.perf_counter()
    moved = 0
    a = offset
    while a + n <= big:
        with torch.cuda.stream(s):
            D1[:n].copy_(H1[a:a + n], non_blocking=True)
            D2[:n].copy_(H2[a:a + n], non_blocking=True)
        s.synchronize()
        moved += 2 * n
        a += n
    dt = time.perf_counter() - t0
    return round(moved / 1e9 / dt, 1)


def with_loop(kind, fn):
    lp = Loop(kind)
    lp.start()
    time.sleep(0.05)
    v = fn()
    lp.stop = True
    lp.join()
    return [v, lp.gbs()]


enc_pattern(0)
res["h2d_pattern_alone_gbs"] = enc_pattern(0)
res["h2d_pattern_unaligned_alone_gbs"] = enc_pattern(1237)
res["h2d_pattern_with_small_d2h_loop"] = with_loop("small", lambda: enc_pattern(0))
res["h2d_pattern_unaligned_with_small_d2h_loop"] = with_loop("small", lambda: enc_pattern(1237))
res["h2d_pattern_with_big_d2h_loop"] = with_loop("big", lambda: enc_pattern(0))
res["h2d_pattern_unaligned_with_big_d2h_loop"] = with_loop("big", lambda: enc_pattern(1237))
print(json.dumps(res))
