"""The C++ command-line driver (repaq_b200/csrc/cli/main.cpp): `repaq -c` / `repaq -d` compatible invocations.
On the CPU the driver is linked against the emulation build of the kernels (tests/emu); on the GPU box the real
binary repaq_b200/repaq_b200_cli is used."""
import json
import os
import subprocess

import pytest

from tests.conftest import ROOT, golden_rfq
from tests.golden.cases import build_cases

EMU_CLI = os.path.join(ROOT, "tests", "emu", "repaq_emu_cli")
GPU_CLI = os.path.join(ROOT, "repaq_b200", "repaq_b200_cli")
MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
CASES = {c["name"]: c for c in build_cases()}
NAMES = ["kat_se", "kat_pe", "nova_se_k100", "nova_pe_k1000", "bgi_se_k100", "nova_pe_nonl_k100", "nova_pe_nonl_r2only_k100", "nova_se_nonl_k100",
         "nova_interleaved_in_k100", "pe_demoted_lastpair_k100", "nova_pe_crlf_k100"]


def run_cli(cli, tmp_path, name):
    import hashlib
    c, m = CASES[name], MAN[name]
    p1 = tmp_path / "a.fq"
    p1.write_bytes(c["r1"])
    cmd = [cli, "-c", "-i", str(p1), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])]
    if c["r2"] is not None:
        p2 = tmp_path / "b.fq"
        p2.write_bytes(c["r2"])
        cmd += ["-I", str(p2)]
    if c["interleaved"]:
        cmd += ["--interleaved_in"]
    subprocess.check_call(cmd)
    assert (tmp_path / "o.rfq").read_bytes() == golden_rfq(name)
    subprocess.check_call([cli, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d.fq")])
    d = (tmp_path / "d.fq").read_bytes()
    assert (len(d), hashlib.sha256(d).hexdigest()) == (m["dec_len"], m["dec_sha256"])
    if "dec1_sha256" in m:
        subprocess.check_call([cli, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d1.fq"), "-O", str(tmp_path / "d2.fq")])
        d1, d2 = (tmp_path / "d1.fq").read_bytes(), (tmp_path / "d2.fq").read_bytes()
        assert (len(d1), hashlib.sha256(d1).hexdigest()) == (m["dec1_len"], m["dec1_sha256"])
        assert (len(d2), hashlib.sha256(d2).hexdigest()) == (m["dec2_len"], m["dec2_sha256"])


@pytest.mark.parametrize("name", NAMES)
def test_cli_emulated(tmp_path, name):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    run_cli(EMU_CLI, tmp_path, name)


def test_cli_error_strings(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    p = subprocess.run([EMU_CLI, "-d", "-i", str(tmp_path / "missing.rfq"), "-o", str(tmp_path / "x.fq")], capture_output=True)
    assert p.returncode != 0 and b"Failed to open file" in p.stderr
    (tmp_path / "bad.rfq").write_bytes(b"RFQ0.4.0\x01" + bytes(30))
    p = subprocess.run([EMU_CLI, "-d", "-i", str(tmp_path / "bad.rfq"), "-o", str(tmp_path / "x.fq")], capture_output=True)
    assert p.returncode != 0 and b"different version of repaq" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cli_gpu(tmp_path, name):
    run_cli(GPU_CLI, tmp_path, name)


def run_verify(cli, tmp_path):
    for name, lossy in (("nova_pe_k1000", False), ("names_numeric_edge_pe", True), ("nova_se_k100", False)):
        c = CASES[name]
        (tmp_path / "a.fq").write_bytes(c["r1"])
        cmd = [cli, "-c", "-v", "-i", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])]
        if c["r2"] is not None:
            (tmp_path / "b.fq").write_bytes(c["r2"])
            cmd += ["-I", str(tmp_path / "b.fq")]
        p = subprocess.run(cmd, capture_output=True, check=True)
        assert (tmp_path / "o.rfq").read_bytes() == golden_rfq(name)
        assert (b"integrity check failure" in p.stderr) == lossy
        if lossy:
            assert b"expected: \n@a:b:c:0000000001:00000000002:000000003:0000000000004 1:N:0:ACGT\ngot:\n@a:b:c:1:2:3:4 1:N:0:ACGT\n" in p.stderr


def test_cli_verify_emulated(tmp_path):
    """-v: the batch is decoded again and compared with its input on the device; a lossy input (leading zeros in the name's numbers
    are not kept by the format) is reported on stderr in the reference's words and the output is still written"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    run_verify(EMU_CLI, tmp_path)


@pytest.mark.gpu
def test_cli_verify_gpu(tmp_path):
    run_verify(GPU_CLI, tmp_path)


@pytest.mark.parametrize("name", ["nova_pe_k100_npos", "nova_pe_nonl_k100", "nova_pe_nonl_r2only_k100", "nova_se_nonl_k100", "nova_se_k100", "pe_demoted_mid_k100", "one_pair"])
def test_cli_decompress_in_windows(tmp_path, monkeypatch, name):
    """`-d` decodes the .rfq in windows of whole chunks (here: tiny ones, a few chunks each); the last chunk of a window is held back
    because only the end of the file tells whether it is the file's last one (trailing-newline rule, and decompressPE's skip)"""
    import hashlib
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    m = MAN[name]
    (tmp_path / "x.rfq").write_bytes(golden_rfq(name))
    for window in ("30000", "70000", "1"):
        monkeypatch.setenv("RPQ_CLI_RFQ_WINDOW", window)
        subprocess.check_call([EMU_CLI, "-d", "-i", str(tmp_path / "x.rfq"), "-o", str(tmp_path / "d.fq")])
        d = (tmp_path / "d.fq").read_bytes()
        assert (len(d), hashlib.sha256(d).hexdigest()) == (m["dec_len"], m["dec_sha256"]), window
        if "dec1_sha256" in m:
            subprocess.check_call([EMU_CLI, "-d", "-i", str(tmp_path / "x.rfq"), "-o", str(tmp_path / "d1.fq"), "-O", str(tmp_path / "d2.fq")])
            d1, d2 = (tmp_path / "d1.fq").read_bytes(), (tmp_path / "d2.fq").read_bytes()
            assert (len(d1), hashlib.sha256(d1).hexdigest()) == (m["dec1_len"], m["dec1_sha256"]), window
            assert (len(d2), hashlib.sha256(d2).hexdigest()) == (m["dec2_len"], m["dec2_sha256"]), window


@pytest.mark.parametrize("name", ["nova_pe_k100_npos", "nova_pe_nonl_k100", "nova_pe_nonl_r2only_k100", "nova_se_nonl_k100", "nova_se_k100", "nova_pe_crlf_k100", "nova_pe_varlen_k100",
                                  "pe_demoted_lastpair_k100", "nova_interleaved_in_k100", "bgi_se_k100"])
@pytest.mark.parametrize("window", ["300000", "170001", "1"])
def test_cli_compress_in_batches(tmp_path, monkeypatch, name, window):
    """`-c` feeds the library batches of text (3 GiB per file in production, here a few chunks, or less than one so that the window
    has to grow) and continues where the last whole chunk ended: the .rfq must not depend on where the batches were cut"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    monkeypatch.setenv("RPQ_CLI_FQ_WINDOW", window)
    c = CASES[name]
    (tmp_path / "a.fq").write_bytes(c["r1"])
    cmd = [EMU_CLI, "-c", "-i", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])]
    if c["r2"] is not None:
        (tmp_path / "b.fq").write_bytes(c["r2"])
        cmd += ["-I", str(tmp_path / "b.fq")]
    if c["interleaved"]:
        cmd += ["--interleaved_in"]
    subprocess.check_call(cmd)
    assert (tmp_path / "o.rfq").read_bytes() == golden_rfq(name)
