#!/usr/bin/env python
"""Randomised parity on the GPU: many small inputs of random shape (paired / single / interleaved, NovaSeq / BGI quality columns,
variable and long reads, CRLF, N bases with and without the N quality, runs of equal qualities across every kind of boundary,
values outside the header's alphabet, chunk sizes) - the CUDA library's .rfq must equal the oracle's byte for byte, its decode of
that .rfq the oracle's decode.  TEST INFRASTRUCTURE (uses oracle/).
usage: fuzz_parity.py [seconds] [first_seed]"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from repaq_b200 import codec as K  # noqa: E402
from tests import parity  # noqa: E402
from tools import fqgen  # noqa: E402


def mutate_quality(buf, rnd, first_chunk_reads):
    """runs of equal values, values the first chunk does not hold, Q16 at the column start"""
    lines = bytes(buf).split(b"\n")
    nrec = (len(lines) - 1) // 4
    for _ in range(rnd.randint(0, 40)):
        rec = rnd.randrange(nrec)
        cr = lines[4 * rec + 3].endswith(b"\r")                          # a text with CRLF line ends: the line's own bytes only
        q = bytearray(lines[4 * rec + 3][:-1] if cr else lines[4 * rec + 3])
        if not q:
            continue
        kind = rnd.randrange(6)
        a = rnd.randrange(len(q)); z = rnd.randint(a, len(q))
        if kind == 0:
            q[a:z] = bytes([q[a]]) * (z - a)
        elif kind == 1 and rec >= first_chunk_reads:
            q[a:z] = bytes([rnd.choice(b"~}|{")]) * (z - a)              # not in any generated alphabet: exception records
        elif kind == 2:
            q[:] = bytes([q[0]]) * len(q)
        elif kind == 3:
            q[0:2] = bytes([q[0]]) * min(2, len(q))
        elif kind == 5:
            q += bytes([q[-1]]) * rnd.randint(1, 3)                         # a quality line longer than its sequence: cut by the reference
        else:
            q[-3:] = bytes([q[-1]]) * min(3, len(q))
        lines[4 * rec + 3] = bytes(q) + (b"\r" if cr else b"")
    return b"\n".join(lines)


def one(cd, seed):
    rnd = random.Random(seed)
    shape = rnd.choice([fqgen.NOVA, fqgen.NOVA, fqgen.BGI])
    flags = rnd.choice([0, 0, fqgen.VARLEN, fqgen.LONG, fqgen.VARLEN | fqgen.LONG, fqgen.NO_N_EARLY])
    paired = shape == fqgen.NOVA and rnd.random() < 0.6
    n = rnd.choice([300, 900, 3000, 9000])
    k = rnd.choice([100, 100, 300, 1000])
    r1, r2 = fqgen.generate(n, seed=seed, shape=shape, flags=flags, paired=paired)
    b1 = mutate_quality(r1, rnd, 400)
    b2 = mutate_quality(r2, rnd, 400) if paired else None
    interleaved = False
    if paired and rnd.random() < 0.2:                                      # one interleaved file instead of two
        l1, l2 = b1.split(b"\n"), b2.split(b"\n")
        recs = []
        for i in range(0, len(l1) - 1, 4):
            recs += l1[i:i + 4] + l2[i:i + 4]
        b1, b2, interleaved = b"\n".join(recs) + b"\n", None, True
    crlf = rnd.random() < 0.15
    if crlf:
        b1 = b1.replace(b"\n", b"\r\n")
        b2 = b2.replace(b"\n", b"\r\n") if b2 is not None else None
    if rnd.random() < 0.3 and b1.endswith(b"\n"):                         # no line break at the end of the file
        b1 = b1[:-2] if crlf else b1[:-1]
    nl = b"\r\n" if crlf else b"\n"
    tail = rnd.randrange(12)                                               # how the input ends (Q13: how far the reader gets before it stops)
    if tail == 0:
        b1 += nl * rnd.randint(1, 3)                                       # blank lines at the end
    elif tail == 1:
        b1 += (b"" if b1.endswith(b"\n") else nl) + b"@ragged" + nl + b"ACGT"      # a ragged last record, no break
    elif tail == 2:
        b1 += (b"" if b1.endswith(b"\n") else nl) + b"@ragged" + nl + b"ACGT" + nl + b"+" + nl
    elif tail == 3 and b2 is not None:
        cut_at = b2.rfind(b"@", 0, len(b2) - 1)                            # the mate file one record shorter
        if cut_at > 0:
            b2 = b2[:cut_at]
    elif tail == 4:
        cut_at = len(b1) * rnd.randint(40, 95) // 100                       # an empty line in the middle: the input ends there
        p = b1.find(nl, cut_at)
        if p > 0:
            b1 = b1[:p] + nl + b1[p:]
    # (the mutated qualities may give an N base another quality than the N quality, or a base the N quality: the reference's own round
    # trip loses those, so only the reference's bytes are asked for, not the input back)
    parity.check_against_oracle(cd, b1, b2, k=k, interleaved=interleaved, roundtrip=False)
    if not interleaved and rnd.random() < 0.5:
        # compare mode (Repaq::compare / comparePE): the .rfq against the FASTQ it came from, or against one with a byte changed,
        # a record dropped or a record added - the report must be the reference's, word for word
        rfq = O.compress(b1, b2, chunk_bases=max(100, k) * 1000)
        c1, c2 = b1, b2
        how = rnd.randrange(4)
        if how == 1 and len(c1) > 100:
            p = rnd.randrange(len(c1))
            if c1[p] not in b"\r\n@+":
                c1 = c1[:p] + bytes([c1[p] ^ 1 if (c1[p] ^ 1) not in b"\r\n" else c1[p]]) + c1[p + 1:]
        elif how == 2:
            p = c1.rfind(b"@", 0, len(c1) - 1)
            if p > 0:
                c1 = c1[:p]
        elif how == 3 and c2 is None:
            c1 = c1 + (b"" if c1.endswith(b"\n") else b"\n") + b"@extra\nACGT\n+\nFFFF\n"
        if len(rfq):
            got, exp = K.compare(rfq, c1, c2, codec=cd), O.compare(rfq, c1, c2)
            assert got == exp, "compare report differs:\n%s\n%s" % (got, exp)
    return len(b1) + (len(b2) if b2 else 0)


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    lib = os.environ.get("RPQ_FUZZ_LIB")                                    # the emulation build for a dry run without a GPU
    cd = K.Codec(lib_path=lib) if lib else K.Codec(0)
    t0 = time.time(); n = 0; nbytes = 0
    while time.time() - t0 < budget:
        try:
            nbytes += one(cd, seed)
        except Exception as e:                                             # noqa: BLE001
            print("FAIL seed %d: %s" % (seed, str(e)[:2000]))
            sys.exit(1)
        n += 1; seed += 1
    print("fuzz ok: %d inputs, %.1f MB, seeds %d..%d" % (n, nbytes / 1e6, seed - n, seed - 1))
