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
_spec2 = importlib.util.spec_from_file_location("fuzz_names", os.path.join(ROOT, "tools", "fuzz_names.py"))
fuzz_names = importlib.util.module_from_spec(_spec2)
_spec2.loader.exec_module(fuzz_names)


def test_fuzz_emulated():
    cd = K.Codec(lib_path=EMU)
    try:
        for seed in PINNED + list(range(300000, 300040)):
            fuzz.one(cd, seed)
        for seed in range(310000, 310060):
            fuzz_names.one(cd, seed)                     # names, other bases, overlapping mates
    finally:
        cd.close()


@pytest.mark.gpu
def test_fuzz_gpu():
    cd = K.Codec(device=0)
    try:
        for seed in PINNED + list(range(400000, 400400)):
            fuzz.one(cd, seed)
        for seed in range(410000, 410400):
            fuzz_names.one(cd, seed)
    finally:
        cd.close()
