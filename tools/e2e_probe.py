#!/usr/bin/env python
"""Experiments behind the pipelined e2e number: do the library's host->device and device->host phases overlap with copies in
the other direction issued by another host thread?"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from repaq_b200 import codec as K  # noqa: E402
from tools import fqgen  # noqa: E402

n = 1 << 28
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}


class Loop(threading.Thread):
    """copies 256 MiB chunks in one direction, synchronising after each, until stopped; reports GB/s"""

    def __init__(self, h2d):
        super().__init__()
        self.h2d, self.stop, self.bytes, self.t = h2d, False, 0, 0.0

    def run(self):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        t0 = time.perf_counter()
        while not self.stop:
            with torch.cuda.stream(s):
                if self.h2d:
                    d_a.copy_(h_in, non_blocking=True)
                else:
                    h_out.copy_(d_b, non_blocking=True)
            s.synchronize()
            self.bytes += n
        self.t = time.perf_counter() - t0

    def gbs(self):
        return round(self.bytes / 1e9 / self.t, 1)


def with_loop(h2d, fn):
    lp = Loop(h2d)
    lp.start()
    time.sleep(0.05)
    t0 = time.perf_counter()
    fn()
    ms = 1e3 * (time.perf_counter() - t0)
    lp.stop = True
    lp.join()
    return round(ms, 1), lp.gbs()


# 1. two threads, torch only
a, b = Loop(True), Loop(False)
a.start(); b.start(); time.sleep(1.0); a.stop = b.stop = True; a.join(); b.join()
res["two_threads_torch_only_gbs"] = dict(h2d=a.gbs(), d2h=b.gbs())

pairs = int(os.environ.get("PROBE_PAIRS", 4760000))
r1, r2 = fqgen.generate(pairs, seed=2, paired=True, threads=16)
header = K.make_header(r1, r2)
enc, dec = K.Codec(device=0), K.Codec(device=0)
enc.set_header(header); dec.set_header(header)
h1 = torch.from_numpy(r1).pin_memory()
h2 = torch.from_numpy(r2).pin_memory()
state = {}


def do_enc():
    state["eo"] = enc.encode_raw(h1.data_ptr(), h1.numel(), h2.data_ptr(), h2.numel(), 0, False, 1000000, True, (K.NEVER, K.NEVER), 0, 0)


def do_dec():
    eo = state["eo"]
    state["do"] = dec.decode_raw(eo.data, eo.bytes, 0, True, 0)


def timeit(fn):
    t0 = time.perf_counter(); fn(); return round(1e3 * (time.perf_counter() - t0), 1)


for _ in range(2):
    do_enc(); do_dec()
res["encode_alone_ms"] = timeit(do_enc)
res["decode_alone_ms"] = timeit(do_dec)
res["encode_with_d2h_loop"] = with_loop(False, do_enc)
res["encode_with_h2d_loop"] = with_loop(True, do_enc)
res["decode_with_h2d_loop"] = with_loop(True, do_dec)
res["decode_with_d2h_loop"] = with_loop(False, do_dec)
print(json.dumps(res))
