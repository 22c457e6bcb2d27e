#!/usr/bin/env python
"""Probe 3: the decoder's D2H pattern (pairs of odd-sized, unaligned ~128 MiB copies, several queued at a time, no host sync
in between) replayed with torch beside the encoder's H2D pattern."""
import ctypes as C
import json
import threading
import time
import torch

n = 1 << 28
big = 1700 << 20
H1 = torch.empty(big, dtype=torch.uint8).pin_memory()
H2 = torch.empty(big, dtype=torch.uint8).pin_memory()
D1 = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
D2 = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
HO1 = torch.empty(big, dtype=torch.uint8).pin_memory()
HO2 = torch.empty(big, dtype=torch.uint8).pin_memory()
DO1 = torch.empty(big, dtype=torch.uint8, device="cuda")
DO2 = torch.empty(big, dtype=torch.uint8, device="cuda")
cudart = C.CDLL("libcudart.so.12")
cnt = C.c_int()
cudart.cudaDeviceGetAttribute(C.byref(cnt), 40, 0)      # cudaDevAttrAsyncEngineCount
res = {"async_engine_count": cnt.value}


class D2HLoop(threading.Thread):
    def __init__(self, depth, odd):
        super().__init__()
        self.depth, self.odd, self.stop, self.bytes, self.t = depth, odd, False, 0, 0.0

    def run(self):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        t0 = time.perf_counter()
        evs = []
        a = 0
        piece = (1 << 27) - (12345 if self.odd else 0)
        while not self.stop:
            if a + piece > big:
                a = 0
            with torch.cuda.stream(s):
                HO1[a:a + piece].copy_(DO1[a:a + piece], non_blocking=True)
                HO2[a:a + piece].copy_(DO2[a:a + piece], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            evs.append(ev)
            self.bytes += 2 * piece
            a += piece
            if len(evs) > self.depth:
                evs.pop(0).synchronize()
        s.synchronize()
        self.t = time.perf_counter() - t0

    def gbs(self):
        return round(self.bytes / 1e9 / self.t, 1)


def enc_pattern(offset):
    s = torch.cuda.Stream(priority=-1)
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
    return round(moved / 1e9 / (time.perf_counter() - t0), 1)


def with_loop(depth, odd):
    lp = D2HLoop(depth, odd)
    lp.start()
    time.sleep(0.05)
    v = [enc_pattern(1237) for _ in range(3)]
    lp.stop = True
    lp.join()
    return [v, lp.gbs()]


enc_pattern(0)
res["h2d_alone"] = enc_pattern(1237)
for depth in (0, 2, 8):
    for odd in (False, True):
        res["h2d_with_d2h_depth%d_%s" % (depth, "odd" if odd else "aligned")] = with_loop(depth, odd)
print(json.dumps(res))
