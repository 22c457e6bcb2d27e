"""ctypes wrapper of oracle/_build/librfq_oracle.so (TEST INFRASTRUCTURE ONLY, see rfq_oracle.h)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librfq_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "repaq")


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref/repaq)."""
    src = os.path.join(HERE, "rfq_oracle.c")
    need = force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "rfq_oracle.h")))
    if need:
        subprocess.check_call(["make", "-s", "-C", HERE, LIB_PATH])
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_BIN)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


REF_GPU_BIN = os.path.join(HERE, "_ref_gpu", "repaq")
REF_GPU_EMU_BIN = os.path.join(HERE, "_ref_gpu", "repaq_emu")


def build_ref_gpu():
    """The reference's own main / Options / Writer with Repaq::run routed into librepaq_b200 (integration/repaq_gpu.cpp): built from
    the reference sources where they lie, when they are present; the GPU box uses the prebuilt binaries."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref_gpu"])


class Header(C.Structure):
    _fields_ = [("read_length_bytes", C.c_uint8), ("flags", C.c_uint16), ("name2_diff_pos", C.c_uint8),
                ("name2_diff_char", C.c_char), ("n_base_qual", C.c_int8), ("overlap_shift", C.c_int8),
                ("support_interleaved", C.c_uint8), ("qual_bins", C.c_uint8), ("qual_buf", C.c_uint8 * 256)]


class Meta(C.Structure):
    _fields_ = [("name1_len", C.c_uint32), ("name2_off", C.c_uint32), ("name2_len", C.c_uint32),
                ("lane", C.c_uint8), ("tile", C.c_uint16), ("x", C.c_uint32), ("y", C.c_uint32),
                ("has_lane_tile_xy", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_meta_parse.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(Meta)]
        L.orc_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_compress_with_header.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32,
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_compare.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.orc_compare.restype = C.c_int
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def meta_parse(name: bytes):
    m = Meta()
    lib().orc_meta_parse(name, len(name), C.byref(m))
    n1 = name[:m.name1_len]
    n2 = name[m.name2_off:m.name2_off + m.name2_len]
    return dict(has=bool(m.has_lane_tile_xy), name1=n1, lane=m.lane, tile=m.tile, x=m.x, y=m.y, name2=n2)


def _as_buf(b):
    """bytes / bytearray / numpy uint8 array -> (ctypes pointer-compatible object, length) without copying"""
    if isinstance(b, bytes):
        return b, len(b)
    import numpy as np
    a = np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b)
    return C.cast(a.ctypes.data, C.c_char_p), a.size


def compress(r1, r2=None, chunk_bases=1000000, interleaved=False) -> bytes:
    out = C.c_void_p()
    n = C.c_size_t()
    p1, l1 = _as_buf(r1)
    p2, l2 = (None, 0) if r2 is None else _as_buf(r2)
    rc = lib().orc_compress(p1, l1, p2, l2, int(interleaved), chunk_bases, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    data = C.string_at(out.value, n.value) if n.value else b""
    lib().orc_free(out)
    return data


def compress_with_header(header: bytes, r1, r2=None, chunk_bases=1000000, interleaved=False) -> bytes:
    """the chunks of r1 (/r2) encoded with the given serialised file header - no header in the result"""
    out = C.c_void_p()
    n = C.c_size_t()
    p1, l1 = _as_buf(r1)
    p2, l2 = (None, 0) if r2 is None else _as_buf(r2)
    rc = lib().orc_compress_with_header(header, len(header), p1, l1, p2, l2, int(interleaved), chunk_bases, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    data = C.string_at(out.value, n.value) if n.value else b""
    lib().orc_free(out)
    return data


def decompress(rfq: bytes, pe_out=False):
    o1, o2 = C.c_void_p(), C.c_void_p()
    n1, n2 = C.c_size_t(), C.c_size_t()
    rc = lib().orc_decompress(rfq, len(rfq), int(pe_out), C.byref(o1), C.byref(n1), C.byref(o2), C.byref(n2))
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    a = C.string_at(o1.value, n1.value) if n1.value else b""
    b = C.string_at(o2.value, n2.value) if n2.value else b""
    lib().orc_free(o1)
    lib().orc_free(o2)
    return (a, b) if pe_out else a


def compare(rfq: bytes, r1, r2=None) -> str:
    """Repaq::compare / comparePE: the JSON report the reference prints for `repaq --compare -i r1 [-I r2] -r x.rfq`."""
    L = lib()
    pq, lq = _as_buf(rfq)
    p1, l1 = _as_buf(r1)
    p2, l2 = (None, 0) if r2 is None else _as_buf(r2)
    out = C.c_void_p()
    rc = L.orc_compare(pq, lq, p1, l1, p2, l2, C.byref(out))
    if rc:
        raise RuntimeError(L.orc_last_error().decode())
    txt = C.string_at(out).decode("latin1")
    L.orc_free(out)
    return txt


def have_ref():
    return os.path.exists(REF_BIN)


def ref_compress(tmpdir, r1: bytes, r2: bytes = None, chunk_kb=1000, interleaved=False) -> bytes:
    """Run the unmodified reference binary (oracle/_ref/repaq -c)."""
    p1 = os.path.join(tmpdir, "in1.fq")
    open(p1, "wb").write(r1)
    out = os.path.join(tmpdir, "out.rfq")
    cmd = [REF_BIN, "-c", "-i", p1, "-o", out, "-k", str(chunk_kb)]
    if r2 is not None:
        p2 = os.path.join(tmpdir, "in2.fq")
        open(p2, "wb").write(r2)
        cmd += ["-I", p2]
    if interleaved:
        cmd += ["--interleaved_in"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return open(out, "rb").read()


def ref_decompress(tmpdir, rfq: bytes, pe_out=False):
    p = os.path.join(tmpdir, "in.rfq")
    open(p, "wb").write(rfq)
    o1 = os.path.join(tmpdir, "o1.fq")
    o2 = os.path.join(tmpdir, "o2.fq")
    cmd = [REF_BIN, "-d", "-i", p, "-o", o1] + (["-O", o2] if pe_out else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    a = open(o1, "rb").read()
    return (a, open(o2, "rb").read()) if pe_out else a


def ref_compare(tmpdir, rfq: bytes, r1: bytes, r2: bytes = None) -> str:
    """Run the unmodified reference binary in compare mode; returns what it prints on stdout."""
    p = os.path.join(tmpdir, "cmp.rfq")
    open(p, "wb").write(rfq)
    p1 = os.path.join(tmpdir, "cmp1.fq")
    open(p1, "wb").write(r1)
    cmd = [REF_BIN, "--compare", "-i", p1, "-r", p]
    if r2 is not None:
        p2 = os.path.join(tmpdir, "cmp2.fq")
        open(p2, "wb").write(r2)
        cmd += ["-I", p2]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode("latin1")
