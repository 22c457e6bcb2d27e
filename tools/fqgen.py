"""Deterministic synthetic FASTQ (SURVEY.md section 8d shapes) via tools/fqgen.c - inputs for tests and bench.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfqgen.so")
NOVA, BGI = 0, 1
NO_N_EARLY, CRLF, VARLEN, LONG = 1, 2, 4, 8
ROW_READS = 300
_lib = None


def build(force=False):
    src = os.path.join(HERE, "fqgen.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fqgen_max_record_bytes.restype = C.c_size_t
        for fn in (L.fqgen_rows, L.fqgen_rows_mt):
            fn.argtypes = [C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t,
                           C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def generate(n_reads, seed=1, shape=NOVA, flags=0, paired=False, threads=None, first_row=0, out1=None, out2=None):
    """n_reads (per file) is rounded up to whole rows of 300 reads.  Returns numpy uint8 arrays (r1, r2|None).
    out1/out2: optional preallocated uint8 arrays (e.g. views of pinned memory) to generate into."""
    L = lib()
    n_rows = (n_reads + ROW_READS - 1) // ROW_READS
    cap = L.fqgen_max_record_bytes(shape, flags) * ROW_READS * n_rows + 1024
    b1 = out1 if out1 is not None else np.empty(cap, dtype=np.uint8)
    b2 = (out2 if out2 is not None else np.empty(cap, dtype=np.uint8)) if paired else None
    n1, n2 = C.c_size_t(), C.c_size_t()
    if threads is not None:
        L.fqgen_set_threads(int(threads))
    rc = L.fqgen_rows_mt(seed, shape, flags, first_row, n_rows, b1.ctypes.data, b1.size, C.byref(n1),
                         b2.ctypes.data if paired else None, b2.size if paired else 0, C.byref(n2))
    assert rc == 0, "fqgen: output buffer too small"
    return b1[:n1.value], (b2[:n2.value] if paired else None)


def bytes_per_read(shape=NOVA, flags=0):
    return lib().fqgen_max_record_bytes(shape, flags)


def truncate_reads(buf, n_reads):
    """first n_reads records of a '\\n'-terminated FASTQ image"""
    nl = np.flatnonzero(buf == 10)
    return buf[: nl[4 * n_reads - 1] + 1]
