#!/usr/bin/env python
"""A small pass through every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
paired-end NovaSeq-shape encode, decode from host and from device memory (parallel chunk walk forced), compare, and a BGI-shape
encode (dense spans: k_streams7).  usage: compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RPQ_DEBUG_PAR_WALK_MIN", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from repaq_b200 import codec as K  # noqa: E402
from tools import fqgen  # noqa: E402

cd = K.Codec(0)
r1, r2 = fqgen.generate(7000, seed=3, paired=True)
rfq = K.compress(r1, r2, codec=cd)
d1, d2 = K.decompress(rfq, pe_out=True, codec=cd)
assert d1 == bytes(r1) and d2 == bytes(r2)
print(K.compare(rfq, r1, r2, codec=cd).replace("\n", " "))
h, used = K.parse_header(rfq)
cd.set_header(h)
body = torch.from_numpy(np.frombuffer(rfq, dtype=np.uint8)[used:].copy()).cuda()
o = cd.decode_raw(body.data_ptr(), body.numel(), 1, True, 1)
assert (o.out1_bytes, o.out2_bytes) == (r1.size, r2.size) and cd.stats().dec_walk == 2
t1, t2 = torch.from_numpy(r1.copy()).cuda(), torch.from_numpy(r2.copy()).cuda()
e = cd.encode_raw(t1.data_ptr(), t1.numel(), t2.data_ptr(), t2.numel(), 1, False, 1000000, True, (K.NEVER, K.NEVER), 0, 1)
assert e.bytes == len(rfq) - used
b1, _ = fqgen.generate(20000, seed=5, shape=fqgen.BGI)
rb = K.compress(b1, k=100, codec=cd)
assert K.decompress(rb, codec=cd) == bytes(b1)
rb2 = K.compress(b1, k=100, codec=cd)                  # second batch: the dense hint sends every span to k_streams7
assert rb2 == rb
# runs that cross segments, spans and chunks, exception records, Q16 (k_streams4 -> k_streams7 hand-over, the walk back through the text)
from tests import parity  # noqa: E402
adv = parity.adversarial_quality_column(6000)
ra = K.compress(adv, k=100, codec=cd)
assert K.decompress(ra, codec=cd) == adv
# N positions in few reads (sparse staging of the N-position coder, empty spans)
lines = bytes(r1).split(b"\n")
for k in range(1, len(lines) - 1, 4 * 37):
    lines[k] = b"N" + lines[k][1:-1] + b"N"
nn = b"\n".join(lines)
rn = K.compress(nn, k=100, codec=cd)
assert K.decompress(rn, codec=cd) == nn
cd.close()
print("sanitize probe ok")
