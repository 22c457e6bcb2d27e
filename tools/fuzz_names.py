#!/usr/bin/env python
"""Randomised parity of the per-read metadata path: read names of every odd shape (FastqMeta::parse, src/fastqmeta.cpp:22-80: colons,
blanks, numbers that are not numbers, signs, overflow), bases other than A/C/G/T after the first chunk, and mate pairs whose overlap
is exact, broken by one base, or longer than a read (RfqCodec::overlap, src/rfqcodec.cpp:1391-1438).  TEST INFRASTRUCTURE (uses oracle/).
usage: fuzz_names.py [seconds] [first_seed]   (RPQ_FUZZ_LIB=<emulation build> for a run without a GPU)"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from repaq_b200 import codec as K  # noqa: E402
from tests import parity  # noqa: E402

COMP = {65: 84, 84: 65, 67: 71, 71: 67}


def revcomp(s):
    return bytes(COMP.get(c, 78) for c in reversed(s))


def field(rnd, kind):
    if kind == 0:
        return str(rnd.randint(0, 99999)).encode()
    if kind == 1:
        return rnd.choice([b"", b"0", b"007", b"+12", b"-5", b" 42", b"4 2", b"12ab", b"ab12", b"2097151", b"2097152", b"99999999999", b"18446744073709551617", b"1.5", b"\t7"])
    return bytes(rnd.choice(b"0123456789ABCxyz -+") for _ in range(rnd.randint(0, 9)))


def name(rnd, base, mode):
    """mode 0: Illumina-like with the fields varied; 1: few or many colons; 2: anything"""
    if mode == 0:
        f = [b"A00250", b"26", b"H3YTWDSXX", field(rnd, rnd.choice([0, 0, 0, 1])), field(rnd, rnd.choice([0, 0, 1])), field(rnd, rnd.choice([0, 0, 1])), field(rnd, rnd.choice([0, 0, 1]))]
        tail = rnd.choice([b" 1:N:0:ACTG", b" 2:N:0:ACTG", b"", b" ", b":extra", b" x y z", b"/1"])
        return b"@" + b":".join(f) + tail
    if mode == 1:
        n = rnd.randint(0, 10)
        return b"@" + b":".join(field(rnd, rnd.choice([0, 1, 2])) for _ in range(n + 1)) + rnd.choice([b"", b" 1", b" :", b": "])
    return b"@" + bytes(rnd.choice(b"abcXYZ0123456789:: _-/.#") for _ in range(rnd.randint(1, 60))) or b"@x"


def one(cd, seed):
    rnd = random.Random(seed)
    n = rnd.choice([200, 700, 1500])
    rl = rnd.choice([36, 100, 150, 151])
    paired = rnd.random() < 0.6
    mode = rnd.choice([0, 0, 1, 2])
    same_names = rnd.random() < 0.5                       # mates share the name up to the "1" / "2"
    dirty_from = rnd.choice([n, n, 400, 5])                # first read that may hold other characters (the header takes the first chunk)
    recs1, recs2 = [], []
    for i in range(n):
        s1 = bytes(rnd.choice(b"ACGT") for _ in range(rl if rnd.random() < 0.9 else rnd.randint(1, rl)))
        q1 = bytes(rnd.choice(b"FFFFFF,:#") for _ in range(len(s1)))
        nm = name(rnd, i, mode if rnd.random() < 0.8 else rnd.choice([0, 1, 2]))
        if len(nm) > 255:
            nm = nm[:255]
        if i >= dirty_from and rnd.random() < 0.1:
            p = rnd.randrange(len(s1))
            s1 = s1[:p] + bytes([rnd.choice(b"NNNnacgtRYKM.-*")]) + s1[p + 1:]
        recs1.append(nm + b"\n" + s1 + b"\n+\n" + q1 + b"\n")
        if paired:
            how = rnd.randrange(6)
            if how <= 2:                                   # overlapping mates: R2 = revcomp of a stretch that ends inside / at / beyond R1's end
                o = rnd.randint(1, len(s1))
                ext = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(0, rl)))
                frag = s1[len(s1) - o:] + ext
                s2 = revcomp(frag[:rl]) if frag else b"A"
                if how == 2 and len(s2) > 3:               # one base off
                    p = rnd.randrange(len(s2)); s2 = s2[:p] + bytes([b"ACGT"[(b"ACGT".index(s2[p]) + 1) % 4] if s2[p] in b"ACGT" else 65]) + s2[p + 1:]
            elif how == 3:                                 # R1 inside R2 (negative overlap in the reference's terms)
                pre = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 40)))
                s2 = revcomp((pre + s1)[:rl])
            else:
                s2 = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, rl)))
            if i >= dirty_from and rnd.random() < 0.1:
                p = rnd.randrange(len(s2)); s2 = s2[:p] + bytes([rnd.choice(b"NNnacgt.")]) + s2[p + 1:]
            q2 = bytes(rnd.choice(b"FFFFFF,:#") for _ in range(len(s2)))
            nm2 = (nm.replace(b" 1:", b" 2:") if same_names else name(rnd, i, mode))[:255]
            recs2.append(nm2 + b"\n" + s2 + b"\n+\n" + q2 + b"\n")
    b1 = b"".join(recs1)
    b2 = b"".join(recs2) if paired else None
    from oracle import oracle as O
    try:
        O.compress(b1, b2, chunk_bases=100000)
    except Exception as oe:                              # noqa: BLE001
        # what the reference refuses (lower-case bases in the first chunk, a coordinate >= 2^21 ...) must be refused in its words
        try:
            K.compress(b1, b2, k=100, codec=cd)
        except K.RepaqError as e:
            key = "2M" if "2M" in str(oe) else str(oe).split("\n")[0][:40]
            assert key in str(e), "refused for another reason: %s | %s" % (e, oe)
            return 0
        raise AssertionError("the oracle refuses this input (%s), the library does not" % oe)
    parity.check_against_oracle(cd, b1, b2, k=100, roundtrip=False)
    return len(b1) + (len(b2) if b2 else 0)


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = os.environ.get("RPQ_FUZZ_LIB")
    cd = K.Codec(lib_path=lib) if lib else K.Codec(0)
    t0 = time.time(); n = 0; nbytes = 0
    while time.time() - t0 < budget:
        try:
            nbytes += one(cd, seed)
        except Exception as e:                           # noqa: BLE001
            print("FAIL seed %d: %s: %s" % (seed, type(e).__name__, str(e)[:1500]))
            sys.exit(1)
        n += 1; seed += 1
    print("name fuzz ok: %d inputs, %.1f MB, seeds %d..%d" % (n, nbytes / 1e6, seed - n, seed - 1))
