#!/usr/bin/env python
"""Device-resident encode + decode throughput of the other BASELINE.json shapes (parity cases, not the bench line):
BGI-SEQ-shape single end 100 bp (38-42 quality values: as many position streams) and NovaSeq-shape single end 150 bp.
usage: shape_probe.py [reads]   -> one JSON line per shape (GB/s of FASTQ, ms per kernel)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from repaq_b200 import codec as K  # noqa: E402
from tools import fqgen  # noqa: E402


def probe(name, r1, r2=None, steps=3):
    h = K.make_header(r1, r2)
    enc, dec = K.Codec(0), K.Codec(0)
    enc.set_header(h)
    dec.set_header(h)
    d1 = torch.from_numpy(r1).cuda()
    d2 = torch.from_numpy(r2).cuda() if r2 is not None else None
    fq = r1.size + (r2.size if r2 is not None else 0)

    def step():
        eo = enc.encode_raw(d1.data_ptr(), d1.numel(), d2.data_ptr() if d2 is not None else None, d2.numel() if d2 is not None else 0, 1, False, 1000000, True, (K.NEVER, K.NEVER), 0, 1)
        se = enc.stats()
        do = dec.decode_raw(eo.data, eo.bytes, 1, r2 is not None, 1)
        sd = dec.stats()
        return eo, do, se.ms_total, sd.ms_total
    eo, do, _, _ = step()
    out = torch.empty(do.out1_bytes, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(do.out1), C.c_size_t(do.out1_bytes), 3)
    ok = bool(torch.equal(out, d1))
    step()
    e = d = 0.0
    for _ in range(steps):
        _, _, a, b = step()
        e += a
        d += b
    enc.set_profiling(True)
    dec.set_profiling(True)
    step()
    prof = {k: round(v[1], 3) for k, v in sorted(list(enc.profile().items()) + list(dec.profile().items()), key=lambda kv: -kv[1][1]) if v[1] > 0.05}
    print(json.dumps(dict(shape=name, fastq_gb=fq / 1e9, rfq_ratio=eo.bytes / fq, roundtrip_ok=ok, encode_gbs=fq * steps / 1e6 / e, decode_gbs=fq * steps / 1e6 / d,
                          round_trip_gbs=fq * steps / 1e6 / (e + d), kernels_ms=prof)))
    enc.close()
    dec.close()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    which = sys.argv[2] if len(sys.argv) > 2 else "both"
    if which in ("both", "bgi"):
        r1, _ = fqgen.generate(n, seed=5, shape=fqgen.BGI)
        probe("bgi_se_100bp", r1)
    if which in ("both", "nova"):
        r1, _ = fqgen.generate(n, seed=1)
        probe("nova_se_150bp", r1)
