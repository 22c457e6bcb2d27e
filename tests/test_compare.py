"""Compare mode (Repaq::compare / comparePE, reference src/repaq.cpp:36-259): the JSON report must be the reference's, byte for byte.
tests/golden/compare_manifest.json holds what the unmodified reference binary prints for the inputs of compare_cases.py.
  CPU: the oracle's restatement, the emulation build of the kernels through the Python binding and through the C++ driver.
  GPU: the CUDA library through the binding and the driver, plus a property at size (a clean round trip passes, one flipped byte
       deep inside is found exactly where it is)."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from repaq_b200 import codec as K
from tests.conftest import ROOT
from tests.golden.compare_cases import build_compare_cases
from tests.golden.make_compare_golden import rfq_of

MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "compare_manifest.json")))


def check(codec, name):
    # (se_fastq_stops_at_empty_line: a blank line between two records - the reference's getLine() swallows it and reads on;
    # so does the line index, k_index_lines)
    c = CASES[name]
    assert K.compare(rfq_of(c["rfq"]), c["r1"], c["r2"], codec=codec) == MAN[name]
CASES = {c["name"]: c for c in build_compare_cases()}
EMU = os.path.join(ROOT, "tests", "emu", "librepaq_emu.so")
EMU_CLI = os.path.join(ROOT, "tests", "emu", "repaq_emu_cli")
GPU_CLI = os.path.join(ROOT, "repaq_b200", "repaq_b200_cli")


def test_manifest_covers_cases():
    assert sorted(MAN) == sorted(CASES)
    verdicts = {json.loads(v, strict=False)["result"] for v in MAN.values()}
    assert verdicts == {"passed", "failed"}


@pytest.mark.parametrize("name", sorted(MAN))
def test_oracle_compare_matches_reference(name):
    c = CASES[name]
    assert O.compare(rfq_of(c["rfq"]), c["r1"], c["r2"]) == MAN[name]


@pytest.fixture(scope="module")
def emu_codec():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    cd = K.Codec(lib_path=EMU)
    yield cd
    cd.close()


@pytest.mark.parametrize("name", sorted(MAN))
def test_emulated_kernels_compare(emu_codec, name):
    check(emu_codec, name)


def run_cli(cli, tmp_path, name):
    c = CASES[name]
    (tmp_path / "x.rfq").write_bytes(rfq_of(c["rfq"]))
    (tmp_path / "a.fq").write_bytes(c["r1"])
    cmd = [cli, "--compare", "-i", str(tmp_path / "a.fq"), "-r", str(tmp_path / "x.rfq"), "-j", str(tmp_path / "r.json")]
    if c["r2"] is not None:
        (tmp_path / "b.fq").write_bytes(c["r2"])
        cmd += ["-I", str(tmp_path / "b.fq")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, check=True)
    assert p.stdout.decode("latin1") == MAN[name]
    assert (tmp_path / "r.json").read_bytes().decode("latin1") == MAN[name]


CLI_NAMES = ["same_kat_pe", "same_nova_pe_k1000", "pe_seq_r1", "pe_fastq_shorter_r2_only", "pe_fastq_longer", "pe_empty_r2", "se_name", "se_fastq_longer", "se_fastq_empty",
             "same_names_numeric_edge_pe"]


@pytest.mark.parametrize("name", CLI_NAMES)
def test_cli_compare_emulated(tmp_path, name):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    run_cli(EMU_CLI, tmp_path, name)


def test_compare_in_batches(emu_codec):
    """rpq_compare in windows, as the driver feeds files beyond 4 GiB: chunks of one call against a FASTQ window that holds more
    reads than they decode to; the consumed offsets line the next call up; the counters add up to the one-call report"""
    c = CASES["same_nova_pe_k1000"]
    rfq = rfq_of(c["rfq"])
    h, used = K.parse_header(rfq, EMU)
    emu_codec.set_header(h)
    body = np.frombuffer(rfq, dtype=np.uint8)[used:]
    _, _, infos, _ = emu_codec.decode(body, split_pairs=True)
    assert len(infos) == 2
    cut = infos[1]["offset"]
    r1, r2 = np.frombuffer(c["r1"], dtype=np.uint8), np.frombuffer(c["r2"], dtype=np.uint8)
    o1 = emu_codec.compare_raw(body.ctypes.data, cut, 0, False, r1.ctypes.data, r1.size, r2.ctypes.data, r2.size, 0, True)
    assert o1.verdict == 0 and o1.rfq_consumed == cut and o1.rfq_reads == infos[0]["reads"] == o1.fastq_reads
    a, b = o1.r1_consumed, o1.r2_consumed
    assert c["r1"][a - 1:a] == b"\n" and c["r1"][a:a + 1] == b"@"
    # a FASTQ window that ends too early while more text exists: the library asks for more instead of reporting a shorter file
    o = emu_codec.compare_raw(body.ctypes.data + cut, body.size - cut, 0, True, r1.ctypes.data + a, 2000, r2.ctypes.data + b, 2000, 0, False)
    assert o.verdict == 7
    o2 = emu_codec.compare_raw(body.ctypes.data + cut, body.size - cut, 0, True, r1.ctypes.data + a, r1.size - a, r2.ctypes.data + b, r2.size - b, 0, True)
    assert o2.verdict == 0
    whole = json.loads(MAN["same_nova_pe_k1000"], strict=False)
    assert o1.rfq_reads + o2.rfq_reads == whole["rfq_reads"] and o1.fastq_bases + o2.fastq_bases == whole["fastq_bases"]


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu_codec():
    cd = K.Codec(device=0)
    yield cd
    cd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MAN))
def test_gpu_compare_golden(gpu_codec, name):
    check(gpu_codec, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLI_NAMES)
def test_cli_compare_gpu(tmp_path, name):
    run_cli(GPU_CLI, tmp_path, name)


@pytest.mark.gpu
def test_gpu_compare_at_size(gpu_codec):
    """0.29 GB of pairs: the GPU's own .rfq passes against its input and equals the oracle's report on a prefix; one changed quality
    byte 300 k pairs in is reported at that pair with the reference's wording"""
    from tools import fqgen
    r1, r2 = fqgen.generate(400000, seed=77, paired=True)
    rfq = K.compress(r1, r2, codec=gpu_codec)
    rep = json.loads(K.compare(rfq, r1, r2, codec=gpu_codec), strict=False)
    n = 2 * (int(np.count_nonzero(r1 == 10)) // 4)            # the generator rounds the request up to whole rows of reads
    assert n >= 800000
    assert rep == dict(result="passed", msg="", fastq_reads=n, rfq_reads=n, fastq_bases=150 * n, rfq_bases=150 * n)
    p1, p2 = fqgen.truncate_reads(r1, 20000), fqgen.truncate_reads(r2, 20000)
    small = K.compress(p1, p2, codec=gpu_codec)
    assert K.compare(small, p1, p2, codec=gpu_codec) == O.compare(small, bytes(p1), bytes(p2))
    nl = np.flatnonzero(r2 == 10)
    bad = r2.copy()
    at = int(nl[4 * 300000 + 3]) - 5                       # inside the quality line of R2 of pair 300000 (0-based)
    bad[at] = ord("!") if bad[at] != ord("!") else ord("#")
    rep = json.loads(K.compare(rfq, r1, bad, codec=gpu_codec), strict=False)
    assert rep["result"] == "failed" and rep["msg"].startswith("The RFQ file and FASTQ file have different quality in the 300001 pair. ")
    assert rep["rfq_reads"] == 600002 == rep["fastq_reads"] and rep["rfq_bases"] == 600002 * 150


@pytest.mark.parametrize("name", ["same_nova_pe_k1000", "same_nova_se_k100", "pe_seq_r1", "pe_name_r2", "pe_strand_r1_chunk1", "pe_fastq_shorter", "pe_fastq_longer", "pe_fastq_shorter_r2_only",
                                  "se_qual_longer", "se_fastq_longer", "se_fastq_shorter", "se_no_final_newline", "same_pe_demoted_mid_k100", "same_nova_pe_300bp_varlen_k100"])
@pytest.mark.parametrize("windows", [("40000", "300000"), ("400000", "5000"), ("1", "1")])
def test_cli_compare_in_batches(tmp_path, monkeypatch, name, windows):
    """the driver's compare loop over batches of chunks and windows of FASTQ text (files beyond 4 GiB), here with windows of a few
    chunks / a few records / less than one of either: the report must not depend on how the files were cut"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    monkeypatch.setenv("RPQ_CLI_RFQ_WINDOW", windows[0])
    monkeypatch.setenv("RPQ_CLI_FQ_WINDOW", windows[1])
    run_cli(EMU_CLI, tmp_path, name)
