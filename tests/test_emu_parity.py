"""Kernel-logic parity on the CPU: the SAME kernel sources as the product, compiled under the lock-step SIMT emulator
(tests/emu, TEST INFRASTRUCTURE), must reproduce the reference's golden vectors bit for bit.  The GPU build of the
same code is checked by tests/test_gpu_parity.py (-m gpu)."""
import os
import subprocess

import pytest

from repaq_b200 import codec as K
from tests import parity
from tests.conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu", "librepaq_emu.so")


@pytest.fixture(scope="module")
def codec():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    cd = K.Codec(lib_path=EMU)
    yield cd
    cd.close()


@pytest.mark.parametrize("name", parity.OK_CASES)
def test_encode_golden(codec, name):
    parity.check_encode_golden(codec, name)


@pytest.mark.parametrize("name", parity.ERR_CASES)
def test_encode_error(codec, name):
    parity.check_encode_error(codec, name)


@pytest.mark.parametrize("name", parity.OK_CASES)
def test_decode_golden(codec, name):
    parity.check_decode_golden(codec, name)


def test_seeded_pe_against_oracle(codec):
    from tools import fqgen
    r1, r2 = fqgen.generate(9000, seed=31, paired=True)
    parity.check_against_oracle(codec, r1, r2)


def test_empty_and_ragged_inputs(codec):
    codec.set_header(K.make_header(b"@a\nACGT\n+\nFFFF\n"))
    data, infos, meta = codec.encode(b"")
    assert data == b"" and infos == []
    data, infos, meta = codec.encode(b"@a\nACGT\n+\nFFFF\n")
    assert infos[0]["reads"] == 1 and len(data) == infos[0]["bytes"]
    # trailing incomplete record and trailing blank lines are dropped like the reference's reader does
    good = b"@a\nACGT\n+\nFFFF\n@b\nACGA\n+\nFFF:\n"
    for tail in (b"@c\nAC", b"\n\n", b"@c\nACGT\n+\n"):
        parity.check_against_oracle(codec, good + tail, roundtrip=False)
