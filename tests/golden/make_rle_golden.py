"""Golden vectors for the quality run-length coder that ALGORITHM_VER 2 never selects: .rfq files whose
header has neither DONT_ENCODE_QUAL nor ENCODE_QUAL_BY_COL cannot be produced by the reference, but its decoder still takes them
(RfqCodec::decodeQualByRunLenCoding, src/rfqcodec.cpp:919-955).  This script rewrites the quality column of a few golden files with
the restated encoder (oracle: orc_rle_encode, after src/rfqcodec.cpp:767-824), lets the UNMODIFIED reference binary decode the
result, and commits the rewritten .rfq with the sha256 of what the reference made of it.

    python -m tests.golden.make_rle_golden        (from the repo root; needs /root/reference)
"""
import ctypes as C
import hashlib
import json
import os
import struct
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from tests import rfqparse  # noqa: E402

SOURCES = ["nova_se_k100", "nova_pe_k100_npos", "bgi_se_k100", "nova_pe_varlen_k100", "kat_pe"]


def column_to_qualities(h, c, total):
    """the chunk's concatenated qualities from its ENCODE_QUAL_BY_COL column (src/rfqcodec.cpp:957-1047)"""
    bins = list(h["qual_buf"])
    major, nq = bins[0], h["nq"]
    normal = [b for b in bins if b != major or b == nq]
    q = bytearray([major]) * total
    col = c["qual"]
    at = 4 * len(normal)
    for k, v in enumerate(normal):
        n, = struct.unpack_from("<I", col, 4 * k)
        s, e, last = at, at + n, -1
        while s < e:
            b0 = col[s]
            if not b0 & 0x80:
                last += b0 + 1; s += 1
            elif not b0 & 0x40:
                last += (((b0 & 0x3F) << 8) | col[s + 1]) + 1; s += 2
            elif not b0 & 0x20:
                run = (b0 & 0x1F) + 1
                for j in range(run):
                    if last + 1 + j < total:
                        q[last + 1 + j] = v
                last += run; s += 1
                continue
            else:
                last += (((b0 & 0x1F) << 24) | (col[s + 1] << 16) | (col[s + 2] << 8) | col[s + 3]) + 1; s += 4
            if last < total:
                q[last] = v
        at = e
    while at + 5 <= len(col):
        pos, = struct.unpack_from("<I", col, at + 1)
        if pos < total:
            q[pos] = col[at]
        at += 5
    return bytes(q)


def read_lengths(h, c):
    n = c["reads"]
    if c["flags"] & 1:
        v = int.from_bytes(c["readlen"], "little")
        return [v] * n
    w = h["rlb"]
    return [int.from_bytes(c["readlen"][w * i:w * i + w], "little") for i in range(n)]


def rewrite(rfq, truncate_column=False):
    h, chunks = rfqparse.parse(rfq)
    oh = O.Header()
    hb = bytearray(rfq[:17 + h["bins"]])
    flags = (h["flags"] & ~(1 << 7)) & ~(1 << 8)
    hb[10], hb[11] = flags & 0xFF, flags >> 8
    L = O.lib()
    L.orc_header_read.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(O.Header)]
    L.orc_header_read.restype = C.c_size_t
    L.orc_rle_encode.argtypes = [C.POINTER(O.Header), C.c_char_p, C.c_uint32, C.c_char_p]
    L.orc_rle_encode.restype = C.c_size_t
    assert L.orc_header_read(bytes(hb) + bytes(8), len(hb) + 8, C.byref(oh))
    out = bytearray(hb)
    at = len(hb)
    for c in chunks:
        total = sum(read_lengths(h, c))
        q = column_to_qualities(h, c, total)
        buf = C.create_string_buffer(len(q) + 16)
        n = L.orc_rle_encode(C.byref(oh), q, len(q), buf)
        col = buf.raw[:n]
        if truncate_column:
            col = col[:max(1, n // 3)]              # the decoder walks a short column again from its first byte (src/rfqcodec.cpp:930-953)
        raw = bytearray(rfq[at:at + c["bytes"]])
        qoff = raw.rindex(c["qual"], 0, len(raw) - (len(c.get("ov", b"")) + len(c.get("npos", b"")))) if c["qual"] else None
        # rebuild: everything before the quality column, the new column, everything after it
        head_len = c["bytes"] - len(c["qual"]) - len(c.get("ov", b"")) - len(c.get("npos", b""))
        assert qoff is None or qoff == head_len, (qoff, head_len)
        new = bytearray(raw[:head_len]) + col + raw[head_len + len(c["qual"]):]
        struct.pack_into("<I", new, 0, (c["msize"] + len(col) - len(c["qual"])) & 0xFFFFFFFF)
        struct.pack_into("<I", new, 14, len(col))
        out += new
        at += c["bytes"]
    return bytes(out)


def main():
    O.build()
    assert O.have_ref()
    man = {}
    tmp = tempfile.mkdtemp()
    for name in SOURCES:
        rfq = open(os.path.join(HERE, name + ".rfq"), "rb").read()
        for suffix, trunc in (("", False), ("_short", True)):
            out = rewrite(rfq, trunc)
            d = O.ref_decompress(tmp, out, pe_out=False)
            entry = dict(source=name, rfq_sha256=hashlib.sha256(out).hexdigest(), rfq_len=len(out), dec_sha256=hashlib.sha256(d).hexdigest(), dec_len=len(d))
            if rfq[10] & (1 << 5):
                d1, d2 = O.ref_decompress(tmp, out, pe_out=True)
                entry.update(dec1_sha256=hashlib.sha256(d1).hexdigest(), dec1_len=len(d1), dec2_sha256=hashlib.sha256(d2).hexdigest(), dec2_len=len(d2))
            if not trunc and name != "bgi_se_k100":
                # the same qualities, coded differently (bgi_se_k100 holds a quality value its header does not list - an exception
                # record in the column coding; the run-length coder has no code for it and writes the major quality: lossy there)
                assert d == O.ref_decompress(tmp, rfq, pe_out=False), name
            key = "rle_" + name + suffix
            open(os.path.join(HERE, key + ".rfq"), "wb").write(out)
            man[key] = entry
            print(key, len(out), len(d))
    json.dump(man, open(os.path.join(HERE, "rle_manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
