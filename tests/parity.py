"""Shared parity checks: the same assertions run against the CUDA library on a GPU (-m gpu) and against the
emulation build of the same kernel sources on the CPU (tests/emu, kernel-logic coverage without a GPU)."""
import hashlib
import json
import os

from oracle import oracle as O
from repaq_b200 import codec as K
from tests import rfqparse
from tests.conftest import ROOT, golden_rfq
from tests.golden.cases import build_cases

MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
CASES = {c["name"]: c for c in build_cases()}
OK_CASES = sorted(n for n in MAN if not MAN[n].get("error"))
ERR_CASES = sorted(n for n in MAN if MAN[n].get("error"))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def explain(got, exp):
    try:
        return "\n".join(rfqparse.diff(got, exp)[:10])
    except Exception as e:  # pragma: no cover
        return "unparsable output: %r" % (e,)


def check_encode_golden(codec, name):
    c = CASES[name]
    got = K.compress(c["r1"], c["r2"], k=c["k"], interleaved=c["interleaved"], codec=codec)
    exp = golden_rfq(name)
    assert got == exp, explain(got, exp)


def check_encode_error(codec, name):
    import pytest
    c = CASES[name]
    with pytest.raises(K.RepaqError) as e:
        K.compress(c["r1"], c["r2"], k=c["k"], interleaved=c["interleaved"], codec=codec)
    assert "cannot be larger than 2M" in str(e.value)        # the reference's error_exit text (src/rfqcodec.cpp:1316)


def check_decode_golden(codec, name):
    m = MAN[name]
    rfq = golden_rfq(name)
    d = K.decompress(rfq, pe_out=False, codec=codec)
    assert (len(d), sha(d)) == (m["dec_len"], m["dec_sha256"])
    if m.get("pe_decode_error"):
        import pytest
        with pytest.raises(K.RepaqError) as e:
            K.decompress(rfq, pe_out=True, codec=codec)
        assert "encoded by single-end FASTQ" in str(e.value)
    if "dec1_sha256" in m:
        d1, d2 = K.decompress(rfq, pe_out=True, codec=codec)
        assert (len(d1), sha(d1)) == (m["dec1_len"], m["dec1_sha256"])
        assert (len(d2), sha(d2)) == (m["dec2_len"], m["dec2_sha256"])


def check_against_oracle(codec, r1, r2=None, k=1000, interleaved=False, roundtrip=True):
    """encode == oracle encode (bit-exact); decode(oracle rfq) == oracle decode; optionally decode restores the input"""
    r1b = bytes(r1)
    r2b = None if r2 is None else bytes(r2)
    exp = O.compress(r1b, r2b, chunk_bases=max(100, k) * 1000, interleaved=interleaved)
    got = K.compress(r1b, r2b, k=k, interleaved=interleaved, codec=codec)
    assert got == exp, explain(got, exp)
    pe = r2 is not None
    dec = K.decompress(exp, pe_out=pe, codec=codec)
    assert dec == O.decompress(exp, pe_out=pe)
    if roundtrip:
        assert dec == ((r1b, r2b) if pe else r1b)
    return len(exp)


RLE_MAN = json.load(open(os.path.join(ROOT, "tests", "golden", "rle_manifest.json")))


def check_decode_rle_golden(codec, name):
    """row a10: the quality run-length coder (decode only), against what the reference binary decodes"""
    m = RLE_MAN[name]
    rfq = golden_rfq(name)
    d = K.decompress(rfq, pe_out=False, codec=codec)
    assert (len(d), sha(d)) == (m["dec_len"], m["dec_sha256"])
    if "dec1_sha256" in m:
        d1, d2 = K.decompress(rfq, pe_out=True, codec=codec)
        assert (sha(d1), sha(d2)) == (m["dec1_sha256"], m["dec2_sha256"])


def check_n_positions_in_few_reads(codec, n_pairs=12000):
    """ENCODE_N_POS headers (an N whose quality is not the N quality): the N-position coder only stages the reads k_meta3 saw a
    character other than a plain base in.  Paired end (overlaps, reverse strand) and single end, N at read ends included."""
    import random
    from tools import fqgen
    r1, r2 = fqgen.generate(n_pairs, seed=5, paired=True)
    rnd = random.Random(3)
    out = []
    for r in (r1, r2):
        lines = bytes(r).split(b"\n")
        for k in range(1, len(lines) - 1, 4):
            if rnd.random() < 0.03:
                s = bytearray(lines[k])
                for j in rnd.sample(range(len(s)), 3):
                    s[j] = ord("N")
                if rnd.random() < 0.3:
                    s[0] = s[-1] = ord("N")
                lines[k] = bytes(s)
        out.append(b"\n".join(lines))
    assert K.make_header(out[0], out[1]).flags & (1 << 9)
    check_against_oracle(codec, out[0], out[1], k=1000)
    check_against_oracle(codec, out[0], None, k=1000)


def adversarial_quality_column(n_reads=16000, rl=150, seed=11, alphabet=b"F,:#5", rare=b"~!", dense=False):
    """single-end FASTQ whose quality column - read boundaries ignored - is a sequence of runs with heavy-tailed lengths (1 .. several
    spans of 16384 positions), so that runs cross segment (64), span (16384) and chunk boundaries in every phase, start at positions
    0 / 1 of chunks (Q16), and hold values the header (made from the first chunk) does not list (exception records)"""
    import numpy as np
    rnd = np.random.RandomState(seed)
    total = n_reads * rl
    col = np.empty(total, dtype=np.uint8)
    pos = 0
    first_chunk = 1000000 // 10            # callers use k=100: the alphabet is that of the first 100 k bases
    while pos < total:
        kind = rnd.randint(100)
        if dense:
            ln = 1 if kind < 90 else int(rnd.randint(2, 70)) if kind < 99 else int(rnd.randint(70, 40000))
        else:
            ln = int(rnd.randint(1, 4)) if kind < 50 else int(rnd.randint(4, 200)) if kind < 95 else int(rnd.randint(200, 40000))
        v = alphabet[rnd.randint(len(alphabet))]
        if pos > first_chunk + 2000 and rnd.randint(400) == 0:
            v = rare[rnd.randint(len(rare))]
        col[pos:pos + ln] = v
        pos += ln
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rnd.randint(0, 4, total)]
    out = []
    for i in range(n_reads):
        out.append(b"@r%d\n" % i + bases[i * rl:(i + 1) * rl].tobytes() + b"\n+\n" + col[i * rl:(i + 1) * rl].tobytes() + b"\n")
    return b"".join(out)


def check_adversarial_quality_columns(codec, n_reads=16000):
    import string
    check_against_oracle(codec, adversarial_quality_column(n_reads), k=100)
    check_against_oracle(codec, adversarial_quality_column(n_reads, rl=100, seed=12, alphabet=b"F"), k=100)           # a single value: no stream but exceptions
    dense = (string.ascii_uppercase + string.digits + "#$%&").encode()
    check_against_oracle(codec, adversarial_quality_column(n_reads, rl=100, seed=13, alphabet=dense, dense=True), k=100)
    check_against_oracle(codec, adversarial_quality_column(n_reads, rl=37, seed=14, alphabet=dense[:12], dense=True), k=100)


def check_control_bytes_in_names(codec):
    """the line index flags the bytes 0x08..0x0F with one compare and then looks at each: tabs, form feeds ... in names and strand
    lines are ordinary characters, only line feeds and carriage returns end lines"""
    from tools import fqgen
    r1, r2 = fqgen.generate(3000, seed=9, paired=True)
    out = []
    for r in (r1, r2):
        lines = bytes(r).split(b"\n")
        for k in range(0, len(lines) - 1, 4):
            if (k // 4) % 3 == 0:
                lines[k] = lines[k] + b"\tx\x0by\x0cz\x08\x0e\x0f"
            if (k // 4) % 5 == 0:
                lines[k + 2] = b"+\t" + lines[k][1:10]
        out.append(b"\n".join(lines))
    check_against_oracle(codec, out[0], out[1], k=100)
    check_against_oracle(codec, out[0].replace(b"\n", b"\r\n"), None, k=100, roundtrip=False)


def check_quality_longer_than_sequence(codec):
    """a quality line longer than its sequence: the reference copies seq.length() quality bytes and drops the rest (src/rfqcodec.cpp:
    332-407) - for the reverse strand of an interleaved pair the first seq.length() bytes, reversed; a shorter one makes it read past
    its string and is refused here"""
    import pytest
    from tools import fqgen
    r1, r2 = fqgen.generate(2400, seed=17, paired=True)
    out = []
    for r in (r1, r2):
        lines = bytes(r).split(b"\n")
        for k in range(3, len(lines) - 1, 4 * 7):
            lines[k] = lines[k] + lines[k][:1 + (k // 4) % 3]
        out.append(b"\n".join(lines))
    check_against_oracle(codec, out[0], out[1], k=100, roundtrip=False)
    check_against_oracle(codec, out[0], None, k=100, roundtrip=False)
    lines = bytes(r1).split(b"\n")
    lines[4 * 900 + 3] = lines[4 * 900 + 3][:-1]
    with pytest.raises(K.RepaqError) as e:
        K.compress(b"\n".join(lines), k=100, codec=codec)
    assert "quality line shorter than the sequence in record 900" in str(e.value)


def medium_density_quality(n_reads=8000, rl=150, seed=21, p_run=0.14):
    """single-end FASTQ whose quality column holds a few thousand runs per 16384 positions: around and above what k_streams4's list holds
    (binned qualities of older instruments look like this)"""
    import numpy as np
    rnd = np.random.RandomState(seed)
    total = n_reads * rl
    col = np.full(total, ord("F"), dtype=np.uint8)
    starts = np.flatnonzero(rnd.random_sample(total) < p_run)
    vals = np.frombuffer(b",:#5", dtype=np.uint8)[rnd.randint(0, 4, starts.size)]
    lens = rnd.randint(1, 4, starts.size)
    for k in range(3):
        sel = lens > k
        idx = np.minimum(starts[sel] + k, total - 1)
        col[idx] = vals[sel]
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rnd.randint(0, 4, total)]
    return b"".join(b"@r%d\n" % i + bases[i * rl:(i + 1) * rl].tobytes() + b"\n+\n" + col[i * rl:(i + 1) * rl].tobytes() + b"\n" for i in range(n_reads))


def check_medium_density(codec):
    for p_run, seed in ((0.14, 21), (0.10, 22), (0.21, 23), (0.30, 24)):    # around the list size: spans on either side of it
        check_against_oracle(codec, medium_density_quality(seed=seed, p_run=p_run), k=1000)
    st = codec.stats()
    return st
