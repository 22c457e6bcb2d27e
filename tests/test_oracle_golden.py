"""Pins the plain-C oracle (oracle/rfq_oracle.c) to the reference: known-answer vectors from the reference's own
test + SURVEY section 8c, and .rfq / decoded outputs produced by the unmodified reference binary (tests/golden)."""
import hashlib
import json
import os

import pytest

from oracle import oracle as O
from tests.conftest import ROOT, golden_rfq
from tests.golden.cases import KAT_A1, KAT_A2, KAT_NAMES, build_cases

MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
CASES = {c["name"]: c for c in build_cases()}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_fastqmeta_reference_unit_test():
    # src/fastqmeta.cpp:82-110 (the only unit test the reference runs)
    m = O.meta_parse(b"@A00251:28:H3YV7DSXX:40:1101:2356:1000 1:N:0:TAAGTGGC")
    assert (m["name1"], m["lane"], m["tile"], m["x"], m["y"], m["name2"]) == (
        b"@A00251:28:H3YV7DSXX", 40, 1101, 2356, 1000, b" 1:N:0:TAAGTGGC")


@pytest.mark.parametrize("name,expect", KAT_NAMES)
def test_kat_names(name, expect):
    m = O.meta_parse(name)
    got = (int(m["has"]), m["name1"], m["lane"], m["tile"], m["x"], m["y"], m["name2"])
    assert got == expect


def test_kat_md5():
    assert hashlib.md5(O.compress(KAT_A1)).hexdigest() == "0301b44958febc50df4804e1d3747962"
    assert hashlib.md5(O.compress(KAT_A1, KAT_A2)).hexdigest() == "ecd13fb07a97f1bfac927b52b2cb74f5"


def test_generator_is_stable():
    # the fixtures only make sense if tools/fqgen.c still produces the inputs they were made from
    for name, c in CASES.items():
        assert sha(c["r1"]) == MAN[name]["in1_sha256"], name
        if c["r2"] is not None:
            assert sha(c["r2"]) == MAN[name]["in2_sha256"], name


@pytest.mark.parametrize("name", sorted(MAN))
def test_oracle_encode_matches_reference(name):
    c, m = CASES[name], MAN[name]
    if m.get("error"):
        with pytest.raises(O.OracleError):
            O.compress(c["r1"], c["r2"], chunk_bases=max(100, c["k"]) * 1000, interleaved=c["interleaved"])
        return
    mine = O.compress(c["r1"], c["r2"], chunk_bases=max(100, c["k"]) * 1000, interleaved=c["interleaved"])
    assert mine == golden_rfq(name)


@pytest.mark.parametrize("name", sorted(n for n in MAN if not MAN[n].get("error")))
def test_oracle_decode_matches_reference(name):
    m = MAN[name]
    rfq = golden_rfq(name)
    assert sha(rfq) == m["rfq_sha256"]
    d = O.decompress(rfq, pe_out=False)
    assert (len(d), sha(d)) == (m["dec_len"], m["dec_sha256"])
    if m.get("pe_decode_error"):
        with pytest.raises(O.OracleError):
            O.decompress(rfq, pe_out=True)
    if "dec1_sha256" in m:
        d1, d2 = O.decompress(rfq, pe_out=True)
        assert (len(d1), sha(d1)) == (m["dec1_len"], m["dec1_sha256"])
        assert (len(d2), sha(d2)) == (m["dec2_len"], m["dec2_sha256"])


@pytest.mark.skipif(not O.have_ref(), reason="reference binary not built (needs /root/reference)")
def test_oracle_vs_live_reference(tmp_path):
    """Larger differential run against the live reference binary, when it is available."""
    from tools import fqgen
    r1, r2 = fqgen.generate(15000, seed=21, paired=True)
    ref = O.ref_compress(str(tmp_path), bytes(r1), bytes(r2))
    assert O.compress(r1, r2) == ref
    assert O.decompress(ref, pe_out=True) == O.ref_decompress(str(tmp_path), ref, pe_out=True)


RLE_MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "rle_manifest.json")))


@pytest.mark.parametrize("name", sorted(RLE_MAN))
def test_oracle_decodes_run_length_quality_like_the_reference(name):
    """row a10: files whose header selects the quality run-length coder (never written under ALGORITHM_VER 2, still decoded by the
    reference): made by tests/golden/make_rle_golden.py, decoded by the unmodified reference binary"""
    m = RLE_MAN[name]
    rfq = golden_rfq(name)
    assert sha(rfq) == m["rfq_sha256"]
    d = O.decompress(rfq, pe_out=False)
    assert (len(d), sha(d)) == (m["dec_len"], m["dec_sha256"])
    if "dec1_sha256" in m:
        d1, d2 = O.decompress(rfq, pe_out=True)
        assert (sha(d1), sha(d2)) == (m["dec1_sha256"], m["dec2_sha256"])
