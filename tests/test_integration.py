"""The drop-in boundary, proven: the REFERENCE's own program - its main(), cmdline parser, Options::validate, Writer,
reportCompareResult, compiled from its sources where they lie - with Repaq::run routed into librepaq_b200 by
integration/repaq_gpu.cpp (the file INTEGRATION.md tells a maintainer to add; oracle/Makefile target ref_gpu).
  CPU: linked against the emulation build of the kernels (oracle/_ref_gpu/repaq_emu).
  GPU: linked against the CUDA library (oracle/_ref_gpu/repaq).
Every golden vector must come out exactly as the unmodified reference binary made it."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

from oracle import oracle as O
from tests.conftest import ROOT, golden_rfq
from tests.golden.cases import build_cases
from tests.golden.compare_cases import build_compare_cases
from tests.golden.make_compare_golden import rfq_of

MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
CMP_MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "compare_manifest.json")))
CASES = {c["name"]: c for c in build_cases()}
CMP_CASES = {c["name"]: c for c in build_compare_cases()}
OK = sorted(n for n in MAN if not MAN[n].get("error"))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def roundtrip(binary, tmp_path, name):
    c, m = CASES[name], MAN[name]
    (tmp_path / "a.fq").write_bytes(c["r1"])
    cmd = [binary, "-c", "-i", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])]
    if c["r2"] is not None:
        (tmp_path / "b.fq").write_bytes(c["r2"])
        cmd += ["-I", str(tmp_path / "b.fq")]
    if c["interleaved"]:
        cmd += ["--interleaved_in"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert (tmp_path / "o.rfq").read_bytes() == golden_rfq(name)
    subprocess.check_call([binary, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d.fq")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = (tmp_path / "d.fq").read_bytes()
    assert (len(d), sha(d)) == (m["dec_len"], m["dec_sha256"])
    if "dec1_sha256" in m:
        subprocess.check_call([binary, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d1.fq.gz"), "-O", str(tmp_path / "d2.fq")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        d1, d2 = gzip.decompress((tmp_path / "d1.fq.gz").read_bytes()), (tmp_path / "d2.fq").read_bytes()        # the reference's Writer zips by file name
        assert (sha(d1), sha(d2)) == (m["dec1_sha256"], m["dec2_sha256"])
    elif m.get("pe_decode_error"):
        p = subprocess.run([binary, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d1.fq"), "-O", str(tmp_path / "d2.fq")], capture_output=True)
        assert p.returncode != 0 and b"encoded by single-end FASTQ" in p.stderr


def compare(binary, tmp_path, name):
    c = CMP_CASES[name]
    (tmp_path / "x.rfq").write_bytes(rfq_of(c["rfq"]))
    (tmp_path / "a.fq").write_bytes(c["r1"])
    cmd = [binary, "--compare", "-i", str(tmp_path / "a.fq"), "-r", str(tmp_path / "x.rfq"), "-j", str(tmp_path / "r.json")]
    if c["r2"] is not None:
        (tmp_path / "b.fq").write_bytes(c["r2"])
        cmd += ["-I", str(tmp_path / "b.fq")]
    p = subprocess.run(cmd, capture_output=True, check=True)
    assert p.stdout.decode("latin1") == CMP_MAN[name]
    assert (tmp_path / "r.json").read_bytes().decode("latin1") == CMP_MAN[name]


@pytest.fixture(scope="module")
def emu_binary():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
        O.build_ref_gpu()
    if not os.path.exists(O.REF_GPU_EMU_BIN):
        pytest.skip("oracle/_ref_gpu/repaq_emu is built from /root/reference, which is not here")
    return O.REF_GPU_EMU_BIN


@pytest.mark.parametrize("name", OK)
def test_reference_program_on_the_library_emulated(emu_binary, tmp_path, name):
    roundtrip(emu_binary, tmp_path, name)


@pytest.mark.parametrize("name", sorted(CMP_MAN))
def test_reference_program_compare_on_the_library_emulated(emu_binary, tmp_path, name):
    compare(emu_binary, tmp_path, name)


def test_reference_program_rejects_what_the_reference_rejects(emu_binary, tmp_path):
    for name in sorted(n for n in MAN if MAN[n].get("error")):
        c = CASES[name]
        (tmp_path / "a.fq").write_bytes(c["r1"])
        p = subprocess.run([emu_binary, "-c", "-i", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])], capture_output=True)
        assert p.returncode != 0 and b"cannot be larger than 2M" in p.stderr
    # Options::validate is the reference's own
    p = subprocess.run([emu_binary, "-c", "-i", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o.fq")], capture_output=True)
    assert p.returncode != 0 and b"the output should not be a FASTQ file" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", OK)
def test_reference_program_on_the_library_gpu(tmp_path, name):
    if not os.path.exists(O.REF_GPU_BIN):
        pytest.skip("oracle/_ref_gpu/repaq was not built (needs /root/reference at build time)")
    roundtrip(O.REF_GPU_BIN, tmp_path, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CMP_MAN))
def test_reference_program_compare_on_the_library_gpu(tmp_path, name):
    if not os.path.exists(O.REF_GPU_BIN):
        pytest.skip("oracle/_ref_gpu/repaq was not built (needs /root/reference at build time)")
    compare(O.REF_GPU_BIN, tmp_path, name)
