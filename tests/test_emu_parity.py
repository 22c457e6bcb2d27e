"""Kernel-logic parity on the CPU: the SAME kernel sources as the product, compiled under the lock-step SIMT emulator
(tests/emu, TEST INFRASTRUCTURE), must reproduce the reference's golden vectors bit for bit.  The GPU build of the
same code is checked by tests/test_gpu_parity.py (-m gpu)."""
import os
import subprocess

import numpy as np
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


@pytest.mark.parametrize("name", sorted(parity.RLE_MAN))
def test_decode_run_length_quality_golden(codec, name):
    parity.check_decode_rle_golden(codec, name)


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


@pytest.mark.parametrize("target", [128, 4096, 65536, 65536 + 4096, 131072])
def test_crlf_across_indexer_boundaries(codec, target):
    """k_index_lines: a "\r\n" whose '\n' is the first byte of a 16-byte piece / a thread's row / a warp's bulk copy / a
    tile (the '\r' then lives in the previous one) must still count as a CRLF line end"""
    import numpy as np
    from tools import fqgen
    body = bytes(fqgen.generate(700, seed=5, paired=False, flags=fqgen.CRLF)[0])
    nls = np.flatnonzero(np.frombuffer(body, dtype=np.uint8) == 10)
    done = 0
    for want_nl in nls:
        gap = target - int(want_nl)                 # bytes the prefix record must add
        if gap < 15:
            break
        if gap > 700:
            continue
        name = b"@pad" if gap % 2 else b"@padd"
        L = (gap - len(name) - 9) // 2
        if L < 1:
            continue
        rec = name + b"\r\n" + b"A" * L + b"\r\n+\r\n" + b"F" * L + b"\r\n"
        assert len(rec) == gap
        text = rec + body
        assert text[target] == 10 and text[target - 1] == 13
        parity.check_against_oracle(codec, text, k=100, roundtrip=False)
        done += 1
        if done == 2:
            break
    assert done


def _decode_devmem(codec, rfq, split):
    """decode with mem=RPQ_MEM_DEVICE in and out (under emulation 'device' memory is host memory): exercises the device
    chunk walk (k_dec_walk_fast + k_dec_describe, or the exact k_dec_walk)"""
    import ctypes as C
    import numpy as np
    h, used = K.parse_header(rfq, codec.lib_path)
    codec.set_header(h)
    body = np.frombuffer(rfq, dtype=np.uint8)[used:].copy()
    o = codec.decode_raw(body.ctypes.data, body.size, 1, split, 1)
    return (C.string_at(o.out1, o.out1_bytes) if o.out1_bytes else b"", C.string_at(o.out2, o.out2_bytes) if o.out2_bytes else b"", o.n_chunks)


@pytest.mark.parametrize("name", ["nova_pe_k1000", "bgi_se_k100", "names_mixed", "nova_pe_300bp_varlen_k100", "pe_demoted_mid_k100", "nova_se_tile_change_k100"])
def test_device_side_chunk_walk(codec, name):
    from tests.conftest import golden_rfq
    rfq = golden_rfq(name)
    h, used = K.parse_header(rfq, codec.lib_path)
    codec.set_header(h)
    o1, o2, infos, _ = codec.decode(rfq[used:], split_pairs=False)
    d1, d2, n = _decode_devmem(codec, rfq, False)
    assert (d1, n) == (o1, len(infos))


def test_fallback_kernels_forced(monkeypatch):
    """RPQ_DEBUG_FORCE_V1=1 selects the long-read fallback kernels (warp per read, exact sequential chunk walk)"""
    monkeypatch.setenv("RPQ_DEBUG_FORCE_V1", "1")
    cd = K.Codec(lib_path=EMU)
    try:
        for name in ("nova_pe_k1000", "nova_pe_k100_npos", "bgi_se_varlen_k100", "pe_demoted_lastpair_k100", "nova_se_late_quality"):
            parity.check_encode_golden(cd, name)
            parity.check_decode_golden(cd, name)
        from tests.conftest import golden_rfq
        rfq = golden_rfq("nova_pe_k1000")
        assert _decode_devmem(cd, rfq, True)[:2] == K.decompress(rfq, pe_out=True, codec=cd)
    finally:
        cd.close()


def test_stream_coder_generations(monkeypatch):
    """RPQ_DEBUG_NO_STREAMS4=1: k_streams3 codes every span (it otherwise only sees the spans k_streams4 hands over); and inputs
    built to need that hand-over inside otherwise ordinary chunks: runs of hundreds of equal non-major qualities, qualities the
    header alphabet does not hold (exception records), a run over the first two positions (Q16)"""
    import numpy as np
    from tools import fqgen
    monkeypatch.setenv("RPQ_DEBUG_NO_STREAMS4", "1")
    cd = K.Codec(lib_path=EMU)
    try:
        for name in ("nova_pe_k1000", "nova_pe_k100_npos", "bgi_se_varlen_k100", "nova_se_late_quality"):
            parity.check_encode_golden(cd, name)
    finally:
        cd.close()
    monkeypatch.delenv("RPQ_DEBUG_NO_STREAMS4")
    cd = K.Codec(lib_path=EMU)
    try:
        r1, r2 = fqgen.generate(2500, seed=77, paired=True)
        for buf, seed in ((r1, 1), (r2, 2)):
            rnd = np.random.RandomState(seed)
            lines = bytes(buf).split(b"\n")
            nrec = (len(lines) - 1) // 4
            for rec in rnd.choice(np.arange(400, nrec), 60, replace=False):      # after the first chunk: the header alphabet is fixed
                q = bytearray(lines[4 * rec + 3])
                kind = rnd.randint(4)
                if kind == 0: q[:] = b"#" * len(q)                                  # a run longer than a segment
                elif kind == 1: q[10:140] = b"," * 130
                elif kind == 2: q[5:9] = b"5678"                                     # not in the alphabet: exceptions
                else: q[0:3] = b":::"
                lines[4 * rec + 3] = bytes(q)
            if seed == 1:
                q = bytearray(lines[3]); q[0:2] = b",,"; lines[3] = bytes(q)          # Q16: positions 0 and 1 of the chunk
                r1m = b"\n".join(lines)
            else:
                r2m = b"\n".join(lines)
        parity.check_against_oracle(cd, r1m, r2m, k=100)
        parity.check_against_oracle(cd, r1m, k=100)
    finally:
        cd.close()


def _dense_cases(cd, lib_path):
    """inputs whose quality spans hold more runs than k_streams4's list: BGI-shape (38-42 quality values), also with read-length
    variation, with exceptions (values that are not in chunk 0's alphabet), a run over the first two positions of a chunk (Q16) and
    runs of hundreds of equal values inside dense data (those spans go on to k_streams3)"""
    from tools import fqgen
    r1, _ = fqgen.generate(30000, seed=5, shape=fqgen.BGI)
    parity.check_against_oracle(cd, r1)                                         # 3 chunks of 1 M bases
    parity.check_against_oracle(cd, r1, k=100)
    r1, _ = fqgen.generate(12000, seed=6, shape=fqgen.BGI, flags=fqgen.VARLEN)
    parity.check_against_oracle(cd, r1, k=100)
    lines = bytes(r1).split(b"\n")
    rnd = np.random.RandomState(3)
    nrec = (len(lines) - 1) // 4
    for rec in rnd.choice(np.arange(1500, nrec), 80, replace=False):
        q = bytearray(lines[4 * rec + 3])
        kind = rnd.randint(4)
        if kind == 0 and len(q) > 30: q[3:30] = b"\x7e" * 27                     # '~' is in no BGI alphabet: exception records
        elif kind == 1: q[:] = bytes([q[0]]) * len(q)                              # a run longer than a segment: k_streams3's
        elif kind == 2 and len(q) > 12: q[4:12] = bytes([q[4]]) * 8
        else: q[0:2] = bytes([q[0]]) * 2
        lines[4 * rec + 3] = bytes(q)
    q = bytearray(lines[3]); q[0:2] = bytes([q[0]]) * 2; lines[3] = bytes(q)      # Q16: a run over positions 0 and 1 of the chunk
    parity.check_against_oracle(cd, b"\n".join(lines), k=100)
    # 56 quality values (57 KB of per-stream table columns in k_streams7);
    # 42: about what BGI-SEQ columns hold
    r1, _ = fqgen.generate(12000, seed=8, shape=fqgen.BGI)
    for nvalues in (56, 42):
        lines = bytes(r1).split(b"\n")
        rnd = np.random.RandomState(4)
        for rec in range((len(lines) - 1) // 4):
            q = np.frombuffer(lines[4 * rec + 3], dtype=np.uint8)
            lines[4 * rec + 3] = (35 + (q.astype(np.int32) + rnd.randint(0, nvalues, q.size)) % nvalues).astype(np.uint8).tobytes()
        parity.check_against_oracle(cd, b"\n".join(lines), k=100)


def test_dense_quality_spans(monkeypatch):
    """k_streams7 takes the spans with more runs than k_streams4 lists (the default), RPQ_DEBUG_STREAMS5=0 leaves them to k_streams3,
    =2 lets k_streams7 code every quality span of ordinary inputs too: all three must give the reference's bytes"""
    for knob, names in (("1", ("bgi_se_k100", "bgi_se_varlen_k100")), ("0", ("bgi_se_k100",)),
                        ("2", ("nova_pe_k1000", "nova_pe_k100_npos", "bgi_se_varlen_k100", "nova_se_late_quality", "kat_pe", "one_read", "nova_pe_300bp_varlen_k100"))):
        monkeypatch.setenv("RPQ_DEBUG_STREAMS5", knob)
        cd = K.Codec(lib_path=EMU)
        try:
            for name in names:
                parity.check_encode_golden(cd, name)
            if knob != "0":
                _dense_cases(cd, EMU)
        finally:
            cd.close()


def test_pipelined_host_windows(monkeypatch):
    """the pipelined host path (windows over three lanes) must give the same bytes as one batch; tiny windows via
    RPQ_DEBUG_PIPE_WINDOW so that several windows fit a test input"""
    from tools import fqgen
    monkeypatch.setenv("RPQ_DEBUG_PIPE_WINDOW", "1300000")
    cd = K.Codec(lib_path=EMU)
    try:
        for name in ("nova_pe_k100_npos", "bgi_se_k100", "nova_pe_nonl_k100", "pe_demoted_mid_k100", "nova_pe_varlen_k100", "nova_se_k100"):
            parity.check_encode_golden(cd, name)
        r1, r2 = fqgen.generate(30000, seed=71, paired=True)          # 10.7 MB per file, k=100: ~8 windows
        parity.check_against_oracle(cd, r1, r2, k=100)
    finally:
        cd.close()


@pytest.mark.parametrize("paired", [False, True])
def test_short_reads_many_per_span(codec, paired):
    """reads of 3..40 bases: more reads reach into one 16 K-position span than the flat staging table of k_streams3 holds
    (fallback to the warp-per-read staging), and runs / tokens cross read boundaries everywhere"""
    import random
    from tools import fqgen

    def shorten(buf, seed):
        rnd = random.Random(seed)
        lines = bytes(buf).split(b"\n")
        out = []
        for k in range(0, len(lines) - 3, 4):
            n = rnd.randint(3, 40)
            out += [lines[k], lines[k + 1][:n], lines[k + 2], lines[k + 3][:n]]
        return b"\n".join(out) + b"\n"

    r = fqgen.generate(6000, seed=11, paired=paired)
    if paired:
        parity.check_against_oracle(codec, shorten(r[0], 1), shorten(r[1], 2), k=100)
    else:
        parity.check_against_oracle(codec, shorten(r[0], 1), k=100)


def _mutate_bases(buf, seed):
    """lower-case bases, N and IUPAC codes sprinkled over the sequence lines after the first chunk (the reference validates the
    alphabet only while it builds the header, src/rfqcodec.cpp:20-145)"""
    import numpy as np
    rnd = np.random.RandomState(seed)
    lines = bytes(buf).split(b"\n")
    for k in range(1 + 4 * 700, len(lines) - 1, 4):
        if rnd.rand() < 0.3:
            s = bytearray(lines[k])
            for _ in range(rnd.randint(1, 12)):
                j = rnd.randint(len(s))
                kind = rnd.randint(4)
                s[j] = s[j] | 0x20 if kind == 0 else (ord("N") if kind == 1 else (ord("n") if kind == 2 else b"RYKM.-*"[rnd.randint(7)]))
            if rnd.rand() < 0.2:
                s[:] = bytes(s).lower()
            lines[k] = bytes(s)
    return b"\n".join(lines)


def test_unclean_bases(codec, tmp_path):
    """lower case (a plain base only for the reverse-complemented mate), N, IUPAC codes: the packed-word fast paths of k_meta3
    must fall back exactly where the reference's per-character code differs; pinned against the reference binary when present"""
    from oracle import oracle as O
    from tools import fqgen
    r1, r2 = fqgen.generate(3000, seed=21, paired=True)
    m1, m2 = _mutate_bases(r1, 1), _mutate_bases(r2, 2)
    if O.have_ref():
        assert O.compress(m1, m2, chunk_bases=100000) == O.ref_compress(str(tmp_path), m1, m2, chunk_kb=100)
        assert O.compress(m1, None, chunk_bases=100000) == O.ref_compress(str(tmp_path), m1, None, chunk_kb=100)
    parity.check_against_oracle(codec, m1, m2, k=100, roundtrip=False)
    parity.check_against_oracle(codec, m1, k=100, roundtrip=False)


def test_window_cut_inside_the_chunk_closing_record(codec):
    """a batch that is not final and ends in the middle of the quality line of the very record that would close a chunk: the
    unterminated tail is not a record (it belongs to the next batch), so the batch yields no chunk - it used to be indexed as a
    record with a short quality line and rejected"""
    from tools import fqgen
    r1, _ = fqgen.generate(1200, seed=31)
    h = K.make_header(r1, lib_path=codec.lib_path)
    codec.set_header(h)
    nl = np.flatnonzero(r1 == 10)
    closing = 666                                              # -k 100, 150 bp: read 667 closes the chunk
    cut = int(nl[4 * closing + 3]) - 40
    data, infos, st = codec.encode(r1[:cut], chunk_bases=100000, final=False)
    assert infos == [] and st["r1_consumed"] == 0
    # the whole record, its '\n' the last byte of the batch: whether a '\n' that follows belongs to that break (the reference's
    # reader swallows one, src/fastqreader.cpp:113-116) is for the next batch to see, so the line still counts as unterminated
    data, infos, st = codec.encode(r1[:cut + 41], chunk_bases=100000, final=False)
    assert infos == [] and st["r1_consumed"] == 0
    data, infos, st = codec.encode(r1[:cut + 42], chunk_bases=100000, final=False)     # one byte more: one chunk
    assert len(infos) == 1 and st["r1_consumed"] == cut + 41
    whole = K.compress(r1, k=100, codec=codec)
    assert whole[len(K.header_bytes(h, codec.lib_path)):].startswith(data)


def test_blank_lines(codec):
    """the reference's getLine() swallows a '\\n' that directly follows a line break (src/fastqreader.cpp:113-116): a single blank
    line between records (or between the lines of a record) is invisible to it, two in a row leave an empty line that ends the
    input, and so does one at the very start or the very end of the file.  The line index does the same (k_index_lines; texts
    whose breaks differ in length go through k_canon_*)."""
    from oracle import oracle as O
    from tools import fqgen
    r1, _ = fqgen.generate(1500, seed=32)
    b = bytes(r1)
    nl = np.flatnonzero(r1 == 10)
    at = int(nl[4 * 100 - 1]) + 1                             # start of record 100
    mid = int(nl[4 * 700 + 1]) + 1                            # start of the strand line of record 700 (chunk 1 at -k 100)
    cases = [b[:at] + b"\n" + b[at:], b[:at] + b"\n\n" + b[at:], b[:mid] + b"\n" + b[mid:], b"\n" + b, b[:at] + b"\r" + b[at:],
             b[:at] + b"\n" + b[at:mid] + b"\n" + b[mid:] + b"\n"]
    for x in cases:
        exp = O.compress(x, chunk_bases=100000)
        if O.have_ref():
            import tempfile
            assert exp == O.ref_compress(tempfile.mkdtemp(), x, None, chunk_kb=100)
        assert K.compress(x, k=100, codec=codec) == exp
    for tail in (b"\n", b"\n\n"):
        assert K.compress(b + tail, k=100, codec=codec) == O.compress(b + tail, chunk_bases=100000)


def test_lone_cr_and_mixed_line_ends(codec):
    """lone '\\r' line ends are line ends to the reference's reader (src/fastqreader.cpp:100-105), and a file may mix them"""
    from oracle import oracle as O
    from tools import fqgen
    r1, r2 = fqgen.generate(1500, seed=33, paired=True)
    b1, b2 = bytes(r1), bytes(r2)
    cr1, cr2 = b1.replace(b"\n", b"\r"), b2.replace(b"\n", b"\r")
    lines = b1.split(b"\n")[:-1]
    mixed = b"".join(ln + (b"\n", b"\r", b"\r\n")[i % 3] for i, ln in enumerate(lines))
    for x1, x2 in ((cr1, None), (cr1, cr2), (cr1[:-1], None), (mixed, None), (mixed, b2)):
        exp = O.compress(x1, x2, chunk_bases=100000)
        if O.have_ref():
            import tempfile
            assert exp == O.ref_compress(tempfile.mkdtemp(), x1, x2, chunk_kb=100)
        assert K.compress(x1, x2, k=100, codec=codec) == exp


class _HostAsDevice:
    """under emulation 'device' memory is host memory"""
    @staticmethod
    def put(arr):
        return arr, arr.ctypes.data

    @staticmethod
    def get(ptr, n):
        import ctypes as C
        return C.string_at(ptr, n) if n else b""


def test_parallel_chunk_walk(monkeypatch, lib_path=EMU, mem=_HostAsDevice):
    """k_dec_find_heads + k_dec_walk_par (device-resident bodies of 32 MiB and more; here forced for any size): sixteen warps follow
    the chunk chain from headers found by scanning, the pieces must fit together exactly, and the result must be the sequential
    walk's.  Also: a body with bytes after its last whole chunk (the chain stops short, the exact walk takes over)."""
    from tools import fqgen
    monkeypatch.setenv("RPQ_DEBUG_PAR_WALK_MIN", "1")
    cd = K.Codec(lib_path=lib_path)
    try:
        r1, r2 = fqgen.generate(9000, seed=41, paired=True)
        rfq = K.compress(r1, r2, k=100, codec=cd)                      # 27 chunks
        h, used = K.parse_header(rfq, lib_path)
        ref1, ref2, infos, _ = cd.decode(rfq[used:], split_pairs=True)  # host walk
        assert len(infos) == 27
        body = np.frombuffer(rfq, dtype=np.uint8)[used:].copy()
        keep, ptr = mem.put(body)
        o = cd.decode_raw(ptr, body.size, 1, True, 1)
        assert cd.stats().dec_walk == 2                                 # the parallel walk was used and checked out
        assert (mem.get(o.out1, o.out1_bytes), mem.get(o.out2, o.out2_bytes), o.n_chunks) == (ref1, ref2, 27)
        # bytes after the last whole chunk: consumed stops there, through the exact walk
        tail = np.concatenate([body, body[:1000]])
        keep2, ptr2 = mem.put(tail)
        o = cd.decode_raw(ptr2, tail.size, 1, True, 1)
        assert cd.stats().dec_walk == 3 and o.consumed == body.size and o.n_chunks == 27
        assert mem.get(o.out1, o.out1_bytes) == ref1
        # one-warp chain for comparison
        monkeypatch.setenv("RPQ_DEBUG_NO_PAR_WALK", "1")
        cd2 = K.Codec(lib_path=lib_path)
        cd2.set_header(h)
        o = cd2.decode_raw(ptr, body.size, 1, True, 1)
        assert cd2.stats().dec_walk == 1 and mem.get(o.out1, o.out1_bytes) == ref1
        cd2.close()
    finally:
        cd.close()


def test_crlf_on_reader_buffer_edges(codec):
    r"""The reference's reader (1 MiB refills) reads a "\r\n" whose '\n' is the last byte of a buffer, or the first of the next one,
    as a break followed by an empty line and ends its input there: the file is silently truncated (the oracle restates that; both
    facts are checked against the reference binary when it is present).  The line index reproduces it."""
    from oracle import oracle as O
    from tools import fqgen
    MIB = 1 << 20
    r1, _ = fqgen.generate(5000, seed=21, flags=fqgen.CRLF)
    b = bytes(r1)
    assert len(b) > MIB and all(b[k * MIB - 1:k * MIB] != b"\n" and b[k * MIB:k * MIB + 1] != b"\n" for k in range(1, len(b) // MIB + 1))
    whole = O.compress(b, chunk_bases=100000)
    assert K.compress(b, k=100, codec=codec) == whole
    lines = b.split(b"\r\n")
    for target in (MIB - 1, MIB):
        pos = 0
        for ln in lines:                                      # the first line break at or after target - 100 ...
            end = pos + len(ln) + 1
            if end >= target - 100 and end <= target:
                break
            pos = end + 1
        padded = list(lines)
        padded[0] = padded[0] + b"x" * (target - end)        # ... moved onto the edge by a longer first name
        x = b"\r\n".join(padded)
        assert x[target - 1:target + 1] == b"\r\n"
        exp = O.compress(x, chunk_bases=100000)
        assert len(exp) < len(whole) * 2 // 3                 # the reference loses everything after the edge
        if O.have_ref():
            import tempfile
            assert exp == O.ref_compress(tempfile.mkdtemp(), x, None, chunk_kb=100)
        assert K.compress(x, k=100, codec=codec) == exp       # and so does the line index (k_index_lines: rd_may_swallow)
        # the same file without its final line break: the reader has loaded the file's last, short buffer (which does not end in a
        # line feed) by the time it meets the empty line, so the chunk flushed after the loop carries NO_LINE_BREAK_AT_END although
        # its own last record ends in the buffer before (Q13; found by tools/fuzz_parity.py, seed 56338 of its first version)
        y = x[:-2]
        exp = O.compress(y, chunk_bases=100000)
        if O.have_ref():
            assert exp == O.ref_compress(tempfile.mkdtemp(), y, None, chunk_kb=100)
        assert K.compress(y, k=100, codec=codec) == exp


def test_dense_hint_follows_the_data(codec):
    """after a batch whose quality spans were mostly dense the next batch goes to k_streams7 directly, and back to k_streams4 after a
    sparse one: the bytes are the reference's whichever coder takes the spans"""
    from tools import fqgen
    dense, _ = fqgen.generate(24000, seed=5, shape=fqgen.BGI)
    sparse, _ = fqgen.generate(16000, seed=12)
    for data in (dense, dense, sparse, sparse, dense):
        parity.check_against_oracle(codec, data, k=1000)




def test_read_longer_than_the_header_can_store(codec):
    """the header's read length width comes from the first chunk (src/rfqcodec.cpp:48-53): a later read of more than 255 bases would be
    written truncated (the reference writes a file it cannot decode); refused instead of losing data silently"""
    short = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 25, b"F" * 100) for i in range(1200))       # 120 kb: the whole first chunk
    long_ = b"@long\n%s\n+\n%s\n" % (b"ACGT" * 150, b"F" * 600)
    with pytest.raises(K.RepaqError) as e:
        K.compress(short + long_, k=100, codec=codec)
    assert "does not fit the header" in str(e.value)
    assert len(K.compress(long_ + short, k=100, codec=codec)) > 0                                       # two-byte lengths from the start: fine


def test_n_positions_in_few_reads(codec):
    parity.check_n_positions_in_few_reads(codec)


def test_adversarial_quality_columns(codec):
    parity.check_adversarial_quality_columns(codec)


def test_control_bytes_in_names(codec):
    parity.check_control_bytes_in_names(codec)


@pytest.mark.parametrize("reads", ["64", "32"])
def test_small_formatter_tiles(monkeypatch, reads):
    """RPQ_DEBUG_FMT_READS: formatter tiles of 64 / 32 reads instead of 128 (what reads of a few hundred bases get by themselves: a
    tile's records must fit shared memory); more tiles per chunk, more steps decoded twice at tile seams"""
    monkeypatch.setenv("RPQ_DEBUG_FMT_READS", reads)
    cd = K.Codec(lib_path=EMU)
    try:
        for name in ("nova_pe_k1000", "nova_pe_k100_npos", "bgi_se_varlen_k100", "nova_pe_300bp_varlen_k100", "nova_se_late_quality"):
            parity.check_decode_golden(cd, name)
    finally:
        cd.close()


def test_quality_longer_than_sequence(codec):
    parity.check_quality_longer_than_sequence(codec)


def test_medium_density_quality_columns(codec):
    """spans with more runs than k_streams4's list holds in a column of four values: the hand-over to k_streams7 at medium density"""
    parity.check_medium_density(codec)
