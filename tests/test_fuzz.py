"""Randomised parity (tools/fuzz_parity.py): inputs of random shape, pairing, line ends, quality columns and endings against the
oracle, a fixed range of seeds per run.  What the fuzzer found is pinned as a constructed case (test_crlf_on_reader_buffer_edges)."""
import importlib.util
import os

import pytest

from repaq_b200 import codec as K
from tests.conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu", "librepaq_emu.so")
_spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
fuzz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fuzz)

PINNED = []


def test_fuzz_emulated():
    cd = K.Codec(lib_path=EMU)
    try:
        for seed in PINNED + list(range(300000, 300040)):
            fuzz.one(cd, seed)
    finally:
        cd.close()


@pytest.mark.gpu
def test_fuzz_gpu():
    cd = K.Codec(device=0)
    try:
        for seed in PINNED + list(range(400000, 400400)):
            fuzz.one(cd, seed)
    finally:
        cd.close()
