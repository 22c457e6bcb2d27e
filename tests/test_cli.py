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
    for window in ("30000", "70000", "997"):
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
@pytest.mark.parametrize("window", ["300000", "170001", "5003"])
def test_cli_compress_in_batches(tmp_path, monkeypatch, name, window):
    """`-c` streams the text in windows (256 MiB per file in production, here a few chunks, or far less than one, so that the text
    of a call is pieced together from many windows) and continues where the last whole chunk ended: the .rfq must not depend on
    where the windows were cut"""
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


# ---- section 8 row f4: .gz in and out (src/fastqreader.cpp:33,49-52, src/writer.cpp:40-43) and the xz pipe (src/main.cpp:133-178)
def run_gz_xz(cli, tmp_path, name):
    import gzip
    import hashlib
    import shutil
    c, m = CASES[name], MAN[name]
    (tmp_path / "a.fq.gz").write_bytes(gzip.compress(c["r1"], 3))
    cmd = [cli, "-c", "-i", str(tmp_path / "a.fq.gz"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])]
    ref = None
    if c["r2"] is not None:
        (tmp_path / "b.fq.gz").write_bytes(gzip.compress(c["r2"], 3))
        cmd += ["-I", str(tmp_path / "b.fq.gz")]
    subprocess.check_call(cmd)
    got = (tmp_path / "o.rfq").read_bytes()
    assert got == golden_rfq(name)
    from oracle import oracle as O
    if O.have_ref():                                     # the unmodified reference binary on the same .gz files
        rcmd = [O.REF_BIN if x == cli else (str(tmp_path / "ref.rfq") if x == str(tmp_path / "o.rfq") else x) for x in cmd]
        subprocess.check_call(rcmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = (tmp_path / "ref.rfq").read_bytes()
        assert got == ref
    # decode into .gz files
    if "dec1_sha256" in m:
        subprocess.check_call([cli, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d1.fq.gz"), "-O", str(tmp_path / "d2.fastq.gz")])
        d1, d2 = gzip.decompress((tmp_path / "d1.fq.gz").read_bytes()), gzip.decompress((tmp_path / "d2.fastq.gz").read_bytes())
        assert (hashlib.sha256(d1).hexdigest(), hashlib.sha256(d2).hexdigest()) == (m["dec1_sha256"], m["dec2_sha256"])
    else:
        subprocess.check_call([cli, "-d", "-i", str(tmp_path / "o.rfq"), "-o", str(tmp_path / "d.fq.gz")])
        assert hashlib.sha256(gzip.decompress((tmp_path / "d.fq.gz").read_bytes())).hexdigest() == m["dec_sha256"]
    if not shutil.which("xz"):
        return
    # .rfq.xz out: this program again, piped through xz; and back
    xcmd = [x if x != str(tmp_path / "o.rfq") else str(tmp_path / "o.rfq.xz") for x in cmd] + ["-z", "1"]
    subprocess.check_call(xcmd)
    assert subprocess.run(["xz", "-d", "-c", str(tmp_path / "o.rfq.xz")], capture_output=True, check=True).stdout == golden_rfq(name)
    subprocess.check_call([cli, "-d", "-i", str(tmp_path / "o.rfq.xz"), "-o", str(tmp_path / "x.fq")])
    d = (tmp_path / "x.fq").read_bytes()
    assert (len(d), hashlib.sha256(d).hexdigest()) == (m["dec_len"], m["dec_sha256"])
    # stdin / stdout
    with open(tmp_path / "o.rfq", "rb") as f:
        p = subprocess.run([cli, "-d", "--stdin", "--stdout"], stdin=f, capture_output=True, check=True)
    assert hashlib.sha256(p.stdout).hexdigest() == m["dec_sha256"]


@pytest.mark.parametrize("name", ["nova_se_k100", "nova_pe_k100_npos", "nova_pe_nonl_k100", "bgi_se_k100"])
def test_cli_gz_and_xz_emulated(tmp_path, name):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    run_gz_xz(EMU_CLI, tmp_path, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["nova_se_k100", "nova_pe_k100_npos", "nova_pe_nonl_k100", "bgi_se_k100"])
def test_cli_gz_and_xz_gpu(tmp_path, name):
    run_gz_xz(GPU_CLI, tmp_path, name)


def run_empty_and_failures(cli, tmp_path):
    # an input without records: an empty output and exit code 0, like the reference (src/repaq.cpp:530-638)
    (tmp_path / "e.fq").write_bytes(b"")
    subprocess.check_call([cli, "-c", "-i", str(tmp_path / "e.fq"), "-o", str(tmp_path / "e.rfq")])
    assert (tmp_path / "e.rfq").read_bytes() == b""
    subprocess.check_call([cli, "-d", "-i", str(tmp_path / "e.rfq"), "-o", str(tmp_path / "e.out")])
    assert (tmp_path / "e.out").read_bytes() == b""
    p = subprocess.run([cli, "-d", "-i", str(tmp_path / "e.rfq"), "-o", str(tmp_path / "e1.out"), "-O", str(tmp_path / "e2.out")], capture_output=True)
    assert p.returncode != 0 and b"encoded by single-end FASTQ" in p.stderr
    p = subprocess.run([cli, "--compare", "-i", str(tmp_path / "e.fq"), "-r", str(tmp_path / "e.rfq")], capture_output=True, check=True)
    assert json.loads(p.stdout, strict=False) == dict(result="passed", msg="", fastq_reads=0, rfq_reads=0, fastq_bases=0, rfq_bases=0)
    # a run that fails leaves no partial output behind (the reference has no clean-up: SURVEY section 5)
    bad = b"@a:b:c:1:1:3000000:5 1:N:0:A\nACGT\n+\nFFFF\n" * 3
    (tmp_path / "bad.fq").write_bytes(bad)
    p = subprocess.run([cli, "-c", "-i", str(tmp_path / "bad.fq"), "-o", str(tmp_path / "bad.rfq")], capture_output=True)
    assert p.returncode != 0 and b"cannot be larger than 2M" in p.stderr
    assert not (tmp_path / "bad.rfq").exists()
    # option checks in the reference's words (src/options.cpp:36-111)
    p = subprocess.run([cli, "-c", "-i", str(tmp_path / "e.fq"), "-o", str(tmp_path / "x.fq")], capture_output=True)
    assert p.returncode != 0 and b"the output should not be a FASTQ file" in p.stderr
    p = subprocess.run([cli, "-d", "-i", str(tmp_path / "e.fq"), "-o", str(tmp_path / "x.out")], capture_output=True)
    assert p.returncode != 0 and b"the input should not be a FASTQ file" in p.stderr


def test_cli_empty_input_and_failures_emulated(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    run_empty_and_failures(EMU_CLI, tmp_path)


@pytest.mark.gpu
def test_cli_empty_input_and_failures_gpu(tmp_path):
    run_empty_and_failures(GPU_CLI, tmp_path)


def test_cli_pairs_end_with_the_shorter_file(tmp_path, monkeypatch):
    """paired files whose last windows are not the same window, and files of different length (FastqReaderPair::read stops with
    the shorter one, src/fastqreader.cpp:287-299), streamed in small windows"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    monkeypatch.setenv("RPQ_CLI_FQ_WINDOW", "130000")
    for name in ("reader_r2_shorter_pe_k100", "nova_pe_varlen_k100"):
        c = CASES[name]
        (tmp_path / "a.fq").write_bytes(c["r1"])
        (tmp_path / "b.fq").write_bytes(c["r2"])
        subprocess.check_call([EMU_CLI, "-c", "-i", str(tmp_path / "a.fq"), "-I", str(tmp_path / "b.fq"), "-o", str(tmp_path / "o.rfq"), "-k", str(c["k"])])
        assert (tmp_path / "o.rfq").read_bytes() == golden_rfq(name)
        subprocess.check_call([EMU_CLI, "-c", "-i", str(tmp_path / "b.fq"), "-I", str(tmp_path / "a.fq"), "-o", str(tmp_path / "o2.rfq"), "-k", str(c["k"])])
        from oracle import oracle as O
        assert (tmp_path / "o2.rfq").read_bytes() == O.compress(c["r2"], c["r1"], chunk_bases=c["k"] * 1000)
