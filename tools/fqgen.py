"""Deterministic synthetic FASTQ (SURVEY.md section 8d shapes) via tools/fqgen.c - inputs for tests and bench.py."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

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
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fqgen_max_record_bytes.restype = C.c_size_t
        L.fqgen_rows.argtypes = [C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t,
                                 C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def _rows(seed, shape, flags, first_row, n_rows, paired):
    L = lib()
    cap = L.fqgen_max_record_bytes(shape, flags) * ROW_READS * n_rows + 1024
    b1 = np.empty(cap, dtype=np.uint8)
    b2 = np.empty(cap, dtype=np.uint8) if paired else None
    n1, n2 = C.c_size_t(), C.c_size_t()
    rc = L.fqgen_rows(seed, shape, flags, first_row, n_rows, b1.ctypes.data, cap, C.byref(n1),
                      b2.ctypes.data if paired else None, cap if paired else 0, C.byref(n2))
    assert rc == 0
    return b1[:n1.value], (b2[:n2.value] if paired else None)


def generate(n_reads, seed=1, shape=NOVA, flags=0, paired=False, threads=None, first_row=0):
    """n_reads (per file) is rounded up to whole rows of 300 reads. Returns numpy uint8 arrays (r1, r2|None)."""
    n_rows = (n_reads + ROW_READS - 1) // ROW_READS
    threads = threads or min(32, os.cpu_count() or 1)
    block = max(1, min(2048, (n_rows + threads - 1) // threads))
    jobs = [(first_row + s, min(block, n_rows - s)) for s in range(0, n_rows, block)]
    if len(jobs) == 1:
        parts = [_rows(seed, shape, flags, jobs[0][0], jobs[0][1], paired)]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda j: _rows(seed, shape, flags, j[0], j[1], paired), jobs))
    r1 = np.concatenate([p[0] for p in parts]) if len(parts) > 1 else parts[0][0]
    r2 = (np.concatenate([p[1] for p in parts]) if len(parts) > 1 else parts[0][1]) if paired else None
    return r1, r2


def truncate_reads(buf, n_reads):
    """first n_reads records of a '\\n'-terminated FASTQ image"""
    nl = np.flatnonzero(buf == 10)
    return buf[: nl[4 * n_reads - 1] + 1]
