#!/usr/bin/env python
"""Encode time per kernel on a quality column of medium density (a few thousand runs per 16384 positions, four values): k_streams4 -> k_streams7.
usage: medium_probe.py [reads] [p_run]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from repaq_b200 import codec as K  # noqa: E402
from tests import parity  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.14
data = parity.medium_density_quality(n_reads=n, p_run=p)
cd = K.Codec(0)
cd.set_header(K.make_header(data, None))
for i in range(3):
    cd.set_profiling(i == 2)
    out, infos, _ = cd.encode(data, None, False, 1000000, True, (K.NEVER, K.NEVER), 0)
prof = {k: round(v[1], 3) for k, v in sorted(cd.profile().items(), key=lambda kv: -kv[1][1]) if v[1] > 0.02}
print("positions %.1f M, rfq/fastq %.3f" % (n * 150 / 1e6, len(out) / len(data)), prof)
