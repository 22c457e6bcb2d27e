"""Regenerates tests/golden/*.rfq and manifest.json by running the UNMODIFIED reference binary
(oracle/_ref/repaq, built by oracle/Makefile from /root/reference) on the inputs of cases.py.

    python -m tests.golden.make_golden        (from the repo root; needs /root/reference)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from tests.golden.cases import build_cases  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    O.build()
    assert O.have_ref(), "reference binary missing: make -C oracle ref"
    manifest = {}
    tmp = tempfile.mkdtemp()
    for c in build_cases():
        try:
            rfq = O.ref_compress(tmp, c["r1"], c["r2"], chunk_kb=c["k"], interleaved=c["interleaved"])
        except subprocess.CalledProcessError as e:
            # the reference error_exit()s (stderr message + exit(-1)): the expected outcome for this input
            manifest[c["name"]] = dict(k=c["k"], interleaved=c["interleaved"], in1_sha256=sha(c["r1"]),
                                       in1_len=len(c["r1"]), error=True, returncode=e.returncode)
            print(c["name"], "reference error_exit", e.returncode)
            continue
        pe = c["r2"] is not None
        entry = dict(k=c["k"], interleaved=c["interleaved"], in1_sha256=sha(c["r1"]), in1_len=len(c["r1"]),
                     rfq_sha256=sha(rfq), rfq_len=len(rfq))
        if pe:
            entry.update(in2_sha256=sha(c["r2"]), in2_len=len(c["r2"]))
            try:
                d1, d2 = O.ref_decompress(tmp, rfq, pe_out=True)
                entry.update(dec1_sha256=sha(d1), dec1_len=len(d1), dec2_sha256=sha(d2), dec2_len=len(d2),
                             roundtrip=bool(d1 == c["r1"] and d2 == c["r2"]))
            except subprocess.CalledProcessError:
                # an EMPTY .rfq (no record in the input) reads as a single-end header: `-O` is refused (src/repaq.cpp:340-342)
                assert len(rfq) == 0
                entry.update(pe_decode_error=True, roundtrip=False)
        # single-output decode (for PE this is the interleaved --stdout style output)
        d = O.ref_decompress(tmp, rfq, pe_out=False)
        entry.update(dec_sha256=sha(d), dec_len=len(d))
        if not pe:
            entry["roundtrip"] = bool(d == c["r1"])
        open(os.path.join(HERE, c["name"] + ".rfq"), "wb").write(rfq)
        manifest[c["name"]] = entry
        print(c["name"], len(rfq), "roundtrip" if entry["roundtrip"] else "LOSSY(reference)")
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
