"""Regenerates tests/golden/compare_manifest.json: what the UNMODIFIED reference binary (oracle/_ref/repaq --compare) prints
for the inputs of compare_cases.py.

    python -m tests.golden.make_compare_golden        (from the repo root; needs /root/reference)
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from tests.golden.compare_cases import build_compare_cases  # noqa: E402


def rfq_of(spec):
    name, _, cut = spec.partition(":")
    data = open(os.path.join(HERE, name + ".rfq"), "rb").read()
    return data[:int(cut)] if cut else data


def main():
    O.build()
    assert O.have_ref(), "reference binary missing: make -C oracle ref"
    tmp = tempfile.mkdtemp()
    man = {}
    for c in build_compare_cases():
        man[c["name"]] = O.ref_compare(tmp, rfq_of(c["rfq"]), c["r1"], c["r2"])
        print(c["name"], json.loads(man[c["name"]], strict=False)["result"])
    json.dump(man, open(os.path.join(HERE, "compare_manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
