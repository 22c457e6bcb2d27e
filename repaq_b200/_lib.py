"""ctypes binding of librepaq_b200.so (the C ABI in include/repaq_b200.h).

The library is the CUDA build made by __graft_entry__.build() / repaq_b200/csrc/Makefile.  There is no CPU
implementation: if the shared object is missing, or no CUDA device is present, loading / rpq_create fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librepaq_b200.so")


class Header(C.Structure):
    _fields_ = [("read_length_bytes", C.c_uint8), ("flags", C.c_uint16), ("name2_diff_pos", C.c_uint8),
                ("name2_diff_char", C.c_uint8), ("n_base_qual", C.c_int8), ("overlap_shift", C.c_int8),
                ("support_interleaved", C.c_uint8), ("qual_bins", C.c_uint8), ("qual_buf", C.c_uint8 * 128)]


class EncodeIn(C.Structure):
    _fields_ = [("r1", C.c_void_p), ("r1_len", C.c_uint64), ("r2", C.c_void_p), ("r2_len", C.c_uint64),
                ("mem", C.c_int), ("interleaved", C.c_int), ("chunk_bases", C.c_uint32), ("final", C.c_int),
                ("nobreak_from", C.c_uint64 * 2), ("tail_flags", C.c_uint16), ("out_mem", C.c_int), ("file_offset", C.c_uint64 * 2)]


class ChunkInfo(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("bytes", C.c_uint32), ("msize", C.c_uint32), ("reads", C.c_uint32),
                ("flags", C.c_uint16), ("seq_size", C.c_uint32), ("qual_size", C.c_uint32), ("npos_size", C.c_uint32),
                ("x_size", C.c_uint32), ("y_size", C.c_uint32), ("name1_size", C.c_uint32), ("name2_size", C.c_uint32),
                ("strand_size", C.c_uint32), ("r1_end", C.c_uint64), ("r2_end", C.c_uint64),
                ("out1_bytes", C.c_uint64), ("out2_bytes", C.c_uint64)]


class EncodeOut(C.Structure):
    _fields_ = [("data", C.c_void_p), ("bytes", C.c_uint64), ("n_chunks", C.c_uint32), ("chunks", C.POINTER(ChunkInfo)),
                ("n_reads", C.c_uint64), ("r1_consumed", C.c_uint64), ("r2_consumed", C.c_uint64)]


class DecodeIn(C.Structure):
    _fields_ = [("data", C.c_void_p), ("bytes", C.c_uint64), ("mem", C.c_int), ("split_pairs", C.c_int), ("out_mem", C.c_int)]


class DecodeOut(C.Structure):
    _fields_ = [("out1", C.c_void_p), ("out1_bytes", C.c_uint64), ("out2", C.c_void_p), ("out2_bytes", C.c_uint64),
                ("n_chunks", C.c_uint32), ("chunks", C.POINTER(ChunkInfo)), ("n_reads", C.c_uint64), ("consumed", C.c_uint64)]


class CompareIn(C.Structure):
    _fields_ = [("rfq", C.c_void_p), ("rfq_bytes", C.c_uint64), ("rfq_mem", C.c_int), ("rfq_final", C.c_int),
                ("r1", C.c_void_p), ("r1_len", C.c_uint64), ("r2", C.c_void_p), ("r2_len", C.c_uint64),
                ("fq_mem", C.c_int), ("fq_final", C.c_int), ("fq_offset", C.c_uint64 * 2)]


class CompareOut(C.Structure):
    _fields_ = [("verdict", C.c_int), ("fastq_reads", C.c_uint64), ("rfq_reads", C.c_uint64), ("fastq_bases", C.c_uint64),
                ("rfq_bases", C.c_uint64), ("read_index", C.c_uint64), ("rfq_field", C.c_void_p), ("rfq_field_len", C.c_uint32),
                ("fastq_field", C.c_void_p), ("fastq_field_len", C.c_uint32), ("r1_consumed", C.c_uint64),
                ("r2_consumed", C.c_uint64), ("rfq_consumed", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("launches", C.c_uint32), ("ms_total", C.c_float), ("ms_kernels", C.c_float), ("ms_h2d", C.c_float),
                ("ms_d2h", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("dec_walk", C.c_uint32)]


EXPORTS = ["rpq_make_header", "rpq_header_write", "rpq_header_read", "rpq_create", "rpq_destroy", "rpq_last_error",
           "rpq_set_header", "rpq_stream", "rpq_encode", "rpq_decode", "rpq_compare", "rpq_get_stats", "rpq_set_profiling", "rpq_get_profile",
           "rpq_host_alloc", "rpq_host_free"]

_libs = {}


def load(path=None):
    """Load the shared object (default: the in-tree CUDA build) and declare the prototypes of include/repaq_b200.h."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(nvcc, sm_100a). repaq_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.rpq_make_header.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(Header), C.c_char_p, C.c_size_t]
    L.rpq_header_write.argtypes = [C.POINTER(Header), C.c_void_p, C.c_size_t]
    L.rpq_header_write.restype = C.c_size_t
    L.rpq_header_read.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Header), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    L.rpq_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.rpq_destroy.argtypes = [C.c_void_p]
    L.rpq_last_error.argtypes = [C.c_void_p]
    L.rpq_last_error.restype = C.c_char_p
    L.rpq_set_header.argtypes = [C.c_void_p, C.POINTER(Header)]
    L.rpq_stream.argtypes = [C.c_void_p]
    L.rpq_stream.restype = C.c_void_p
    L.rpq_encode.argtypes = [C.c_void_p, C.POINTER(EncodeIn), C.POINTER(EncodeOut)]
    L.rpq_decode.argtypes = [C.c_void_p, C.POINTER(DecodeIn), C.POINTER(DecodeOut)]
    L.rpq_compare.argtypes = [C.c_void_p, C.POINTER(CompareIn), C.POINTER(CompareOut)]
    L.rpq_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.rpq_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.rpq_get_profile.argtypes = [C.c_void_p]
    L.rpq_get_profile.restype = C.c_char_p
    _libs[path] = L
    return L
