"""Host-side mirror of the reference's codec interface, bound to the CUDA library through the C ABI.

* `Codec`      ~ class RfqCodec          (reference src/rfqcodec.h:17-43): set_header / make_header / encode / decode
* `compress`   ~ Repaq::compress, compressPE   (src/repaq.cpp:530-759) on in-memory FASTQ images
* `decompress` ~ Repaq::decompress, decompressPE (src/repaq.cpp:262-413), including the trailing-newline rule
* `compare`    ~ Repaq::compare, comparePE (src/repaq.cpp:36-233) with the report of reportCompareResult (:235-259)

No algorithmic work happens here: Python sizes buffers, passes pointers, and applies the reference's host-side
file-level rules (Q13 flag thresholds, last-newline trimming).
"""
import ctypes as C

import numpy as np

from . import _lib

NEVER = (1 << 64) - 1
NO_RECORDS = 1
MIB = 1 << 20

NO_LINE_BREAK_AT_END = 1 << 10
NO_LINE_BREAK_AT_END_R2 = 1 << 11
PAIRED_END = 1 << 5


class RepaqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


def _buf(x):
    """bytes / bytearray / numpy uint8 -> (address, length, keepalive)"""
    if x is None:
        return None, 0, None
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.uint8)
    else:
        a = np.frombuffer(x, dtype=np.uint8)
    return (a.ctypes.data if a.size else None), int(a.size), a


def make_header(r1, r2=None, interleaved=False, chunk_bases=1000000, lib_path=None):
    """RfqCodec::makeHeader on the records of the first chunk (host code, as in the reference).  None: the input holds no
    record (the reference then writes an empty output, src/repaq.cpp:530-638)."""
    L = _lib.load(lib_path)
    h = _lib.Header()
    err = C.create_string_buffer(512)
    p1, l1, k1 = _buf(r1)
    p2, l2, k2 = _buf(r2)
    if p1 is None:
        return None
    rc = L.rpq_make_header(p1, l1, p2, l2, int(interleaved), chunk_bases, C.byref(h), err, 512)
    if rc == NO_RECORDS:
        return None
    if rc:
        raise RepaqError(rc, err.value.decode(errors="replace"))
    return h


def header_bytes(h, lib_path=None):
    L = _lib.load(lib_path)
    out = C.create_string_buffer(17 + 128)
    n = L.rpq_header_write(C.byref(h), out, len(out))
    return out.raw[:n]


def parse_header(data, lib_path=None):
    L = _lib.load(lib_path)
    h = _lib.Header()
    used = C.c_size_t()
    err = C.create_string_buffer(512)
    p, n, keep = _buf(data)
    rc = L.rpq_header_read(p, n, C.byref(h), C.byref(used), err, 512)
    if rc:
        raise RepaqError(rc, err.value.decode(errors="replace"))
    return h, used.value


class Codec:
    """One GPU context (one CUDA stream). Not thread safe, like the reference's RfqCodec."""

    def __init__(self, device=0, lib_path=None):
        self.L = _lib.load(lib_path)
        self.lib_path = lib_path
        ctx = C.c_void_p()
        rc = self.L.rpq_create(device, C.byref(ctx))
        if rc:
            raise RepaqError(rc, "rpq_create failed: no usable CUDA device (repaq_b200 has no CPU fallback)")
        self.ctx = ctx
        self.header = None

    def close(self):
        if getattr(self, "ctx", None):
            self.L.rpq_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        return RepaqError(rc, self.L.rpq_last_error(self.ctx).decode(errors="replace"))

    def set_header(self, h):
        rc = self.L.rpq_set_header(self.ctx, C.byref(h))
        if rc:
            raise self._err(rc)
        self.header = h

    def stats(self):
        s = _lib.Stats()
        self.L.rpq_get_stats(self.ctx, C.byref(s))
        return s

    def set_profiling(self, on):
        self.L.rpq_set_profiling(self.ctx, int(on))

    def profile(self):
        """{kernel: (launches, total_ms)} since set_profiling(True)"""
        out = {}
        for line in self.L.rpq_get_profile(self.ctx).decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    def encode_raw(self, p1, l1, p2, l2, mem, interleaved, chunk_bases, final, nobreak_from, tail_flags, out_mem):
        ein = _lib.EncodeIn()
        ein.r1, ein.r1_len, ein.r2, ein.r2_len = p1, l1, p2, l2
        ein.mem, ein.interleaved, ein.chunk_bases, ein.final = mem, int(interleaved), chunk_bases, int(final)
        ein.nobreak_from[0], ein.nobreak_from[1] = nobreak_from
        ein.tail_flags, ein.out_mem = tail_flags, out_mem
        out = _lib.EncodeOut()
        rc = self.L.rpq_encode(self.ctx, C.byref(ein), C.byref(out))
        if rc:
            raise self._err(rc)
        return out

    def encode(self, r1, r2=None, interleaved=False, chunk_bases=1000000, final=True, nobreak_from=(NEVER, NEVER), tail_flags=0):
        """Encode every chunk of a FASTQ batch held in host memory -> (serialised chunk bytes, [chunk info dicts])."""
        p1, l1, k1 = _buf(r1)
        p2, l2, k2 = _buf(r2)
        out = self.encode_raw(p1, l1, p2, l2, 0, interleaved, chunk_bases, final, nobreak_from, tail_flags, 0)
        data = C.string_at(out.data, out.bytes) if out.bytes else b""
        infos = [{f: getattr(out.chunks[i], f) for f, _ in _lib.ChunkInfo._fields_} for i in range(out.n_chunks)]
        return data, infos, dict(n_reads=out.n_reads, r1_consumed=out.r1_consumed, r2_consumed=out.r2_consumed)

    def decode_raw(self, p, n, mem, split_pairs, out_mem):
        din = _lib.DecodeIn()
        din.data, din.bytes, din.mem, din.split_pairs, din.out_mem = p, n, mem, int(split_pairs), out_mem
        out = _lib.DecodeOut()
        rc = self.L.rpq_decode(self.ctx, C.byref(din), C.byref(out))
        if rc:
            raise self._err(rc)
        return out

    def decode(self, body, split_pairs=False):
        """Decode the chunks of an .rfq body (bytes after the file header) -> (out1, out2, [chunk info dicts])."""
        p, n, keep = _buf(body)
        out = self.decode_raw(p, n, 0, split_pairs, 0)
        o1 = C.string_at(out.out1, out.out1_bytes) if out.out1_bytes else b""
        o2 = C.string_at(out.out2, out.out2_bytes) if out.out2_bytes else b""
        infos = [{f: getattr(out.chunks[i], f) for f, _ in _lib.ChunkInfo._fields_} for i in range(out.n_chunks)]
        return o1, o2, infos, dict(n_reads=out.n_reads, consumed=out.consumed)


    def compare_raw(self, prfq, nrfq, rfq_mem, rfq_final, p1, l1, p2, l2, fq_mem, fq_final):
        cin = _lib.CompareIn()
        cin.rfq, cin.rfq_bytes, cin.rfq_mem, cin.rfq_final = prfq, nrfq, rfq_mem, int(rfq_final)
        cin.r1, cin.r1_len, cin.r2, cin.r2_len, cin.fq_mem, cin.fq_final = p1, l1, p2, l2, fq_mem, int(fq_final)
        out = _lib.CompareOut()
        rc = self.L.rpq_compare(self.ctx, C.byref(cin), C.byref(out))
        if rc:
            raise self._err(rc)
        return out


CMP_FIELDS = {1: "name", 2: "sequence", 3: "strand", 4: "quality"}


def compare_report(passed, msg, fastq_reads, fastq_bases, rfq_reads, rfq_bases):
    """Repaq::reportCompareResult (reference src/repaq.cpp:235-259): the text printed on stdout / written with -j"""
    return ("{\n\t\"result\":\"%s\",\n\t\"msg\":\"%s\",\n\t\"fastq_reads\":%d,\n\t\"rfq_reads\":%d,\n\t\"fastq_bases\":%d,\n\t\"rfq_bases\":%d\n}\n"
            % ("passed" if passed else "failed", msg, fastq_reads, rfq_reads, fastq_bases, rfq_bases))


def compare_message(out, pe):
    """the `msg` of a finished comparison, worded as Repaq::compare / comparePE do (src/repaq.cpp:71-122, :174-225)"""
    v = out.verdict
    shown = (lambda n: n // 2) if pe else (lambda n: n)
    unit, units = ("pair", "pairs") if pe else ("read", "reads")
    if v == 0:
        return ""
    if v in CMP_FIELDS:
        a = C.string_at(out.rfq_field, out.rfq_field_len).decode("latin1")
        g = C.string_at(out.fastq_field, out.fastq_field_len).decode("latin1")
        return "The RFQ file and FASTQ file have different %s in the %d %s. %s | %s" % (CMP_FIELDS[v], shown(out.rfq_reads), unit, a, g)
    if v == 5:
        return ("The RFQ file has more reads than the FASTQ file. The RFQ file has >= %d %s, while the FASTQ file only has %d %s"
                % (shown(out.rfq_reads), units, shown(out.fastq_reads), units))
    if v == 6:
        return ("The FASTQ file has more reads than the RFQ file. The FASTQ file has >= %d %s, while the RFQ file only has %d %s"
                % (shown(out.fastq_reads), units, shown(out.rfq_reads), units))
    raise RepaqError(-2, "comparison not finished (verdict %d)" % v)


def compare(rfq, r1, r2=None, codec=None, device=0, lib_path=None):
    """.rfq file image against FASTQ image(s), like `repaq --compare -i r1 [-I r2] -r x.rfq`: returns the JSON text the
    reference prints.  Decode, index and the field-wise check all run on the GPU (rpq_compare)."""
    own = codec is None
    codec = codec or Codec(device, lib_path)
    try:
        if len(rfq) == 0:
            # an empty .rfq reads as a default header without chunks (src/rfqheader.cpp:7-43): any header will do for zero chunks
            h, used = _lib.Header(), 0
            h.read_length_bytes, h.flags, h.n_base_qual, h.overlap_shift, h.qual_bins = 1, 1 << 7, ord("#"), -24, 1
            h.qual_buf[0] = ord("F")
        else:
            h, used = parse_header(rfq, codec.lib_path)
        codec.set_header(h)
        body = np.frombuffer(rfq, dtype=np.uint8)[used:]
        pb, nb, kb = _buf(body)
        p1, l1, k1 = _buf(r1)
        p2, l2, k2 = _buf(r2)
        if r2 is not None and p2 is None:                      # an empty mate file is still "paired end"
            k2 = np.zeros(1, dtype=np.uint8)
            p2, l2 = k2.ctypes.data, 0
        out = codec.compare_raw(pb, nb, 0, True, p1, l1, p2, l2, 0, True)
        return compare_report(out.verdict == 0, compare_message(out, r2 is not None), out.fastq_reads, out.fastq_bases, out.rfq_reads, out.rfq_bases)
    finally:
        if own:
            codec.close()


def nobreak_rule(size, last_byte):
    """Q13 (reference src/fastqreader.cpp:31-46): which chunks get NO_LINE_BREAK_AT_END for a file of `size` bytes.

    The reader refills a 1 MiB buffer; it raises the flag when it loads a SHORT buffer not ending in '\\n'.  Returns
    (threshold, tail): a chunk is flagged when the line break ending its last record lies at offset >= threshold;
    `tail` tells whether the post-loop flush chunk is flagged regardless (the zero-length refill of a file whose size
    is an exact multiple of 1 MiB reads one byte before its buffer).
    """
    has_nl = last_byte == 0x0A
    if size % MIB == 0:
        return (NEVER if has_nl else size), True
    if has_nl:
        return NEVER, False
    return (size // MIB) * MIB, False


def compress(r1, r2=None, k=1000, interleaved=False, codec=None, device=0, lib_path=None):
    """FASTQ image(s) -> .rfq file image, like `repaq -c -i r1 [-I r2] [-k k] [--interleaved_in]`."""
    chunk_bases = max(100, k) * 1000                       # src/main.cpp:69
    own = codec is None
    codec = codec or Codec(device, lib_path)
    try:
        if r2 is not None and len(r2) == 0:
            return b""                                         # compressPE reads pairs: no R2 record, no pair, nothing written (src/repaq.cpp:640-759)
        h = make_header(r1, r2, interleaved, chunk_bases, lib_path=codec.lib_path)
        if h is None:
            return b""                                         # no record: nothing is written, not even the header
        codec.set_header(h)
        a1 = np.frombuffer(r1, dtype=np.uint8) if not isinstance(r1, np.ndarray) else r1
        t1, tail1 = nobreak_rule(a1.size, int(a1[-1]) if a1.size else 0)
        tail = NO_LINE_BREAK_AT_END if tail1 else 0
        t2 = NEVER
        if r2 is not None:
            a2 = np.frombuffer(r2, dtype=np.uint8) if not isinstance(r2, np.ndarray) else r2
            t2, tail2 = nobreak_rule(a2.size, int(a2[-1]) if a2.size else 0)
            if tail2:
                tail |= NO_LINE_BREAK_AT_END_R2
        elif interleaved:
            t2 = t1
            if tail1:
                tail |= NO_LINE_BREAK_AT_END_R2
        data, infos, _ = codec.encode(r1, r2, interleaved, chunk_bases, True, (t1, t2), tail)
        return header_bytes(h, codec.lib_path) + data
    finally:
        if own:
            codec.close()


def decompress(rfq, pe_out=False, codec=None, device=0, lib_path=None):
    """.rfq file image -> FASTQ image(s), like `repaq -d -i x.rfq -o out1 [-O out2]`."""
    if len(rfq) == 0:
        # RfqHeader::read on an empty stream leaves the constructor's defaults (single end, no chunks): src/rfqheader.cpp:7-43
        if pe_out:
            raise RepaqError(-2, "The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>")
        return b""
    own = codec is None
    codec = codec or Codec(device, lib_path)
    try:
        h, used = parse_header(rfq, codec.lib_path)
        if pe_out and not (h.flags & PAIRED_END):
            raise RepaqError(-2, "The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>")
        codec.set_header(h)
        body = np.frombuffer(rfq, dtype=np.uint8)[used:]
        o1, o2, infos, _ = codec.decode(body, split_pairs=pe_out)
        if not infos:
            return (b"", b"") if pe_out else b""
        if not pe_out:
            # Repaq::decompress: only a flagged LAST chunk loses its final newline (src/repaq.cpp:300-328)
            if infos[-1]["flags"] & NO_LINE_BREAK_AT_END:
                o1 = o1[:-1]
            return o1
        # Repaq::decompressPE (src/repaq.cpp:363-413), including what its `continue` does to a flagged chunk that is
        # not the last one: the rest of that chunk and the whole peeked chunk are never written.
        out1, out2 = [], []
        a1 = a2 = 0
        i = 0
        while i < len(infos):
            ci = infos[i]
            s1 = o1[a1:a1 + ci["out1_bytes"]]
            s2 = o2[a2:a2 + ci["out2_bytes"]]
            f1 = bool(ci["flags"] & NO_LINE_BREAK_AT_END)
            f2 = bool(ci["flags"] & NO_LINE_BREAK_AT_END_R2)
            last = i == len(infos) - 1
            skip_next = False
            if f1:
                if last:
                    out1.append(s1[:-1])
                else:
                    out1.append(s1)
                    skip_next = True
            else:
                out1.append(s1)
            if not skip_next:
                if f2:
                    if last:
                        out2.append(s2[:-1])
                    else:
                        out2.append(s2)
                        skip_next = True
                else:
                    out2.append(s2)
            a1 += ci["out1_bytes"]
            a2 += ci["out2_bytes"]
            if skip_next:
                i += 1
                if i < len(infos):
                    a1 += infos[i]["out1_bytes"]
                    a2 += infos[i]["out2_bytes"]
            i += 1
        return b"".join(out1), b"".join(out2)
    finally:
        if own:
            codec.close()
