"""The C-ABI library loads on a CPU-only box, exports every function include/repaq_b200.h declares, its host-side
entry points (header construction / IO) work, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from oracle import oracle as O
from repaq_b200 import _lib
from repaq_b200 import codec as K
from tests.conftest import ROOT
from tests.golden.cases import KAT_A1, KAT_A2, build_cases


def declared_functions():
    src = open(os.path.join(ROOT, "include", "repaq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpq_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = declared_functions()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n


def test_header_roundtrip_and_matches_oracle():
    for r1, r2 in ((KAT_A1, None), (KAT_A1, KAT_A2)):
        h = K.make_header(r1, r2)
        hb = K.header_bytes(h)
        ref = O.compress(r1, r2)
        assert ref.startswith(hb)
        h2, used = K.parse_header(ref)
        assert used == len(hb) and K.header_bytes(h2) == hb


@pytest.mark.parametrize("case", [c for c in build_cases() if not c["name"].startswith("name")], ids=lambda c: c["name"])
def test_host_make_header_matches_reference_on_goldens(case):
    from tests.conftest import golden_rfq
    import json
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    if man[case["name"]].get("error"):
        pytest.skip("reference rejects this input")
    h = K.make_header(case["r1"], case["r2"], case["interleaved"], max(100, case["k"]) * 1000)
    if man[case["name"]]["rfq_len"] == 0:
        assert h is None                     # no record: the reference leaves an empty output (RPQ_NO_RECORDS)
        return
    assert golden_rfq(case["name"]).startswith(K.header_bytes(h))


def test_header_errors_use_reference_strings():
    bad = b"@r\nACGTacgt\n+\nFFFFFFFF\n"
    with pytest.raises(K.RepaqError) as e:
        K.make_header(bad)
    assert "lowercase bases" in str(e.value)
    with pytest.raises(K.RepaqError) as e:
        K.parse_header(b"RFQ0.4.0\x01" + bytes(20))
    assert "different version of repaq" in str(e.value)
    with pytest.raises(K.RepaqError):
        K.parse_header(b"XYZ0.5.1\x02" + bytes(20))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(K.RepaqError):
        K.Codec(device=0)


def test_product_library_is_cuda_only():
    # the product .so must not contain the emulator or the oracle
    data = open(_lib.LIB_PATH, "rb").read()
    assert b"emu_switch" not in data and b"orc_compress" not in data
    assert b"k_streams" in data            # the CUDA kernels are in there


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/repaq_b200.h must compile as C99 (plain pointers and sizes, no C++ / CUDA / torch types),
    and the ctypes mirrors in repaq_b200/_lib.py must have the sizes the C compiler gives the structs"""
    import ctypes as C
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "repaq_b200.h"\nint main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rpq_header), sizeof(rpq_encode_in), '
                   'sizeof(rpq_chunk_info), sizeof(rpq_encode_out), sizeof(rpq_decode_in), sizeof(rpq_decode_out), sizeof(rpq_compare_in), sizeof(rpq_compare_out), sizeof(rpq_stats)); return 0; }\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    from repaq_b200 import _lib
    mirrors = [_lib.Header, _lib.EncodeIn, _lib.ChunkInfo, _lib.EncodeOut, _lib.DecodeIn, _lib.DecodeOut, _lib.CompareIn, _lib.CompareOut, _lib.Stats]
    assert sizes == [C.sizeof(m) for m in mirrors]
