"""Input recipes for the golden vectors.  Every case is either handcrafted bytes or a deterministic
tools/fqgen.c product (optionally mutated), so only the REFERENCE's outputs need to be committed."""
import numpy as np

from tools import fqgen

KAT_A1 = (b"@A00250:26:H3YTWDSXX:1:1101:1000:1000 1:N:0:ACTG\nACGTACGTACGTACGTNCGT\n+\nFFFFFFFF:FFFFFFF#FFF\n"
          b"@A00250:26:H3YTWDSXX:1:1101:1031:1000 1:N:0:ACTG\nGGGGCCCCAAAATTTTACGT\n+\nFF,FFFFFFFFFFFFFFFFF\n")
KAT_A2 = (b"@A00250:26:H3YTWDSXX:1:1101:1000:1000 2:N:0:ACTG\nACGNACGTACGTACGTACGT\n+\nFFF#FFFFFFFFFFFFFFFF\n"
          b"@A00250:26:H3YTWDSXX:1:1101:1031:1000 2:N:0:ACTG\nTTTTTTTTTTACGTAAAATT\n+\nFFFFFFFFFFFFFFFFFF::\n")

# SURVEY.md section 8c KAT-names: name -> (has, name1, lane, tile, x, y, name2) as printed by the reference's FastqMeta::parse
KAT_NAMES = [
    (b"@A00251:28:H3YV7DSXX:40:1101:2356:1000 1:N:0:TAAGTGGC", (1, b"@A00251:28:H3YV7DSXX", 40, 1101, 2356, 1000, b" 1:N:0:TAAGTGGC")),
    (b"@A00251:28:H3YV7DSXX:40:1101:2356:1000", (0, b"@A00251:28:H3YV7DSXX:40:1101:2356:1000", 0, 0, 0, 0, b"")),
    (b"@A00251:28:H3YV7DSXX:4:1101:2356:1000:UMIACGT 1:N:0:TAAG", (1, b"@A00251:28:H3YV7DSXX", 4, 1101, 2356, 1000, b":UMIACGT 1:N:0:TAAG")),
    (b"@V300035135L2C001R0010000001/1", (0, b"@V300035135L2C001R0010000001/1", 0, 0, 0, 0, b"")),
    (b"@SRR123456.1 A00251:28:H3YV7DSXX:4:1101:2356:1000", (0, b"@SRR123456.1 A00251:28:H3YV7DSXX:4:1101:2356:1000", 0, 0, 0, 0, b"")),
    (b"@HWI-ST1234:100:C1234ACXX:1:1101:1234:2000#ACGT/1", (0, b"@HWI-ST1234:100:C1234ACXX:1:1101:1234:2000#ACGT/1", 0, 0, 0, 0, b"")),
    (b"@a:b:c:d:5 xxx", (1, b"@a:b:c:d", 5, 0, 0, 0, b" xxx")),
    (b"@a:b:c:1:2 yy", (1, b"@a:b:c:1", 2, 0, 0, 0, b" yy")),
    (b"@a:b:c:300:70000:2097151:4000000000 1:N:0:1", (1, b"@a:b:c", 44, 4464, 2097151, 4000000000, b" 1:N:0:1")),
    (b"@a:b:c:004:01101:0002356:01000 1:N:0:1", (1, b"@a:b:c", 4, 1101, 2356, 1000, b" 1:N:0:1")),
    (b"@a:b:c:1:2:3: 4", (1, b"@a:b:c", 1, 2, 3, 0, b" 4")),
    (b"@a:b:c:1:2:3:-7 1", (1, b"@a:b:c", 1, 2, 3, 4294967289, b" 1")),
    (b"@a:b 1:2:3:4:5:6:7", (0, b"@a:b 1:2:3:4:5:6:7", 0, 0, 0, 0, b"")),
    (b"@a:b:c:1:2:x3:y4 q", (1, b"@a:b:c", 1, 2, 0, 0, b" q")),
    (b"@a:b:c:1:2: 12:9 q", (1, b"@a:b:c", 1, 0, 0, 0, b" 12:9 q")),
    (b"@NS500713:64:HFKJJBGXY:1:11101:20469:1097 1:N:0:TATAGCCT+GGTCCCGA", (1, b"@NS500713:64:HFKJJBGXY", 1, 11101, 20469, 1097, b" 1:N:0:TATAGCCT+GGTCCCGA")),
]


def records(buf):
    lines = bytes(buf).split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [lines[i:i + 4] for i in range(0, len(lines) - 3, 4)]


def join(recs):
    return b"".join(b"\n".join(r) + b"\n" for r in recs)


def _names_file():
    # one record per KAT name; sequence/quality constant so only the name path varies
    return [(b"name%02d" % i, n + b"\nACGTTGCAACGTTGCAACGT\n+\nFFFFFFFFFFFFFFFFFFFF\n") for i, (n, _) in enumerate(KAT_NAMES)]


def _numeric_edge_names(name2):
    """records whose names all tokenise with lane/tile/x/y (so the header keeps those columns); x and y stay below 2^21"""
    import random
    rng = random.Random(20261017)
    names = [
        b"@a:b:c:1:2:3:4", b"@a:b:c:0000000001:00000000002:000000003:0000000000004", b"@a:b:c:12345678:87654321:1048575:2097151",
        b"@a:b:c:123456789:999999999:5:6", b"@a:b:c:99999999999999999999:18446744073709551616:7:8", b"@a:b:c:+5:-3:9:10",
        b"@a:b:c:\t7:\x0b2:3:4", b"@a:b:c:1x:2:3y:4", b"@a:b:c::::", b"@a:b:c:4294967296:65536:00:000", b"@a:b:c:2147483648:9223372036854775807:1:1",
        b"@a:b:c:-9223372036854775809:-1:2:2", b"@a:b:c:12a4:1 2:3:4"[:13] + b":5:6:7", b"@" + b"X" * 70 + b":b:c:1:2:3:4", b"@a:" + b"Y" * 80 + b":c:11:22:33:44",
    ]
    for pad in range(40, 72):                      # the space that ends the scan walks across bytes 56..80
        names.append(b"@" + b"Z" * pad + b":bb:cc:3:1101:%d:%d" % (1000 + pad, 2000 + pad))
    for _ in range(300):
        parts = [bytes(rng.choice(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-_") for _ in range(rng.randint(1, 24))) for _ in range(3)]
        def num(maxv):
            v = rng.randint(0, maxv)
            return (b"%d" % v).rjust(rng.choice([0, 0, 0, 1, 4, 5, 8, 9, 12]), b"0")
        fields = [num(300), num(70000), num((1 << 21) - 1), num((1 << 21) - 1)]
        names.append(b"@" + b":".join(parts + fields))
    rec = b"\nACGTTGCAACGTTGCAACGT\n+\nFFFFFFFFFFFFFFFFFFFF\n"
    return b"".join(n + name2 + rec for n in names)


def _requalify(r1, n_values, seed, n_with_quality=None, n_rate=0.0):
    """the records of r1 with qualities drawn from `n_values` distinct characters ('!' upwards); every value occurs in the first
    reads.  n_with_quality: that character is given to N bases only, and only it (the N-from-quality header), with N bases
    injected at n_rate."""
    rng = np.random.RandomState(seed)
    alphabet = [33 + k for k in range(n_values + (1 if n_with_quality else 0)) if 33 + k != n_with_quality][:n_values - (1 if n_with_quality else 0)]
    out = []
    first = True
    for rec in records(r1):
        seq = bytearray(rec[1])
        n = len(seq)
        q = np.array(alphabet, dtype=np.uint8)[rng.randint(0, len(alphabet), n)]
        runs = rng.rand(n) < 0.5                                   # half of the positions repeat their predecessor: runs
        for i in range(1, n):
            if runs[i]:
                q[i] = q[i - 1]
        if first:
            q[:len(alphabet)] = alphabet[:n] if n < len(alphabet) else alphabet
            first = False
        if n_with_quality:
            for i in range(n):
                if seq[i] == ord("N") or rng.rand() < n_rate:
                    seq[i] = ord("N")
                    q[i] = n_with_quality
        out.append([rec[0], bytes(seq), rec[2], bytes(bytearray(q))])
    return join(out)


def build_cases():
    """-> list of dict(name, r1, r2|None, k (chunk kilobases), interleaved)"""
    cases = []

    def add(name, r1, r2=None, k=1000, interleaved=False):
        cases.append(dict(name=name, r1=bytes(r1), r2=None if r2 is None else bytes(r2), k=k, interleaved=interleaved))

    add("kat_se", KAT_A1)
    add("kat_pe", KAT_A1, KAT_A2)
    for nm, data in _names_file():
        add(nm.decode(), data)
    # all KAT names in one file: mixed has/no-has => header without lane/tile/x/y
    add("names_mixed", b"".join(d for _, d in _names_file()))

    r1, _ = fqgen.generate(2100, seed=1)
    add("nova_se_k100", r1, k=100)
    r1, _ = fqgen.generate(7200, seed=12)
    add("nova_se_k1000", r1)                       # >= 100 N in chunk 0: N-from-quality path, 2 chunks
    r1, r2 = fqgen.generate(3600, seed=2, paired=True)
    add("nova_pe_k1000", r1, r2)                   # 2 chunks, overlaps +/-, tile change absent
    r1, r2 = fqgen.generate(1500, seed=3, paired=True)
    add("nova_pe_k100_npos", r1, r2, k=100)        # < 100 N in chunk 0 => ENCODE_N_POS
    r1, _ = fqgen.generate(3000, seed=5, shape=fqgen.BGI)
    add("bgi_se_k100", r1, k=100)
    r1, r2 = fqgen.generate(1500, seed=4, paired=True, flags=fqgen.VARLEN)
    add("nova_pe_varlen_k100", r1, r2, k=100)
    r1, r2 = fqgen.generate(900, seed=6, paired=True, flags=fqgen.LONG)
    add("nova_pe_300bp_k100", r1, r2, k=100)
    r1, r2 = fqgen.generate(900, seed=7, paired=True, flags=fqgen.CRLF)
    add("nova_pe_crlf_k100", r1, r2, k=100)
    r1, r2 = fqgen.generate(1500, seed=8, paired=True)
    add("nova_pe_nonl_k100", bytes(r1)[:-1], bytes(r2)[:-1], k=100)
    add("nova_pe_nonl_r2only_k100", bytes(r1), bytes(r2)[:-1], k=100)
    r1, _ = fqgen.generate(1200, seed=9)
    add("nova_se_nonl_k100", bytes(r1)[:-1], k=100)
    r1, _ = fqgen.generate(1200, seed=10, shape=fqgen.BGI, flags=fqgen.VARLEN)
    add("bgi_se_varlen_k100", r1, k=100)

    # interleaved input
    r1, r2 = fqgen.generate(900, seed=11, paired=True)
    il = join([x for pair in zip(records(r1), records(r2)) for x in pair])
    add("nova_interleaved_in_k100", il, k=100, interleaved=True)

    # Q10: R2 with a different tile mid-chunk (chunk 1 of 3) and on the last pair of a chunk
    r1, r2 = fqgen.generate(1200, seed=13, paired=True)
    a, b = records(r1), records(r2)
    b2 = [list(r) for r in b]
    b2[500][0] = b2[500][0].replace(b":1101:", b":1102:")
    add("pe_demoted_mid_k100", join(a), join(b2), k=100)
    b3 = [list(r) for r in b]
    b3[333][0] = b3[333][0].replace(b":1101:", b":1102:")      # k=100 => 334 pairs per chunk; index 333 = last pair of chunk 0
    add("pe_demoted_lastpair_k100", join(a), join(b3), k=100)
    b4 = [list(r) for r in b]
    b4[700][0] = b4[700][0].replace(b" 2:N:0:", b" 2:Y:0:")     # name2 mismatch in a later chunk
    add("pe_name2_mismatch_late_k100", join(a), join(b4), k=100)
    b5 = [list(r) for r in b]
    b5[0][0] = b5[0][0].replace(b" 2:N:0:", b" 2:Y:1:")         # two differing chars in pair 0 => no interleave support
    add("pe_no_support_k100", join(a), join(b5), k=100)

    # Q3: quality value unseen in chunk 0 (exception records), and a non-N base carrying the N quality later
    r1, _ = fqgen.generate(7200 + 900, seed=14)
    recs = [list(r) for r in records(r1)]
    q = bytearray(recs[7000][3]); q[10] = ord('5'); q[11] = ord('5'); q[100] = ord('J'); recs[7000][3] = bytes(q)
    q = bytearray(recs[7100][3]); s = recs[7100][1]
    for i in range(len(q)):
        if s[i:i + 1] != b"N":
            q[i] = ord('#'); break
    recs[7100][3] = bytes(q)
    add("nova_se_late_quality", join(recs))

    # tile changes inside a chunk (tileSame false) and lane column: stitch rows of two tiles
    ra, _ = fqgen.generate(600, seed=15, first_row=1998)           # rows 1998..1999 of tile 1101, then tile 1102
    add("nova_se_tile_change_k100", ra, k=100)

    # read length > 255 mixed with short: 2-byte length column, variable
    r1, r2 = fqgen.generate(600, seed=16, paired=True, flags=fqgen.LONG | fqgen.VARLEN)
    add("nova_pe_300bp_varlen_k100", r1, r2, k=100)

    # name fields at the edges of the SIMD tokeniser / digit parser of k_meta3: 1..13 digits, leading zeros, signs, blanks,
    # letters, empty fields, saturation, separators on 16-byte boundaries, scans ending before / at / after byte 64
    add("names_numeric_edge", _numeric_edge_names(b" 1:N:0:ACGT"))
    add("names_numeric_edge_pe", _numeric_edge_names(b" 1:N:0:ACGT"), _numeric_edge_names(b" 2:N:0:ACGT"))

    # single read, single pair, empty-ish
    add("one_read", join(records(KAT_A1)[:1]))
    add("one_pair", join(records(KAT_A1)[:1]), join(records(KAT_A2)[:1]))
    # strand line that repeats the name (old style) and differing strand lines
    recs = [list(r) for r in records(fqgen.generate(600, seed=17)[0])]
    for i, r in enumerate(recs):
        if i % 2:
            r[2] = b"+" + r[0][1:]
    add("se_strand_varies_k100", join(recs), k=100)

    # quality alphabets around the 64-value edge of RfqHeader::makeQualityTable (src/rfqheader.cpp:203-234): 63 values + the 0xFF bin
    # = 64 bins, the largest column-coded header (63 streams); 64 values = DONT_ENCODE_QUAL, the raw quality column of
    # src/rfqcodec.cpp:612-615 / :903-908; 64 values one of which is the N quality = 64 bins with BOTH flags set (raw wins)
    base, _ = fqgen.generate(2100, seed=41)
    add("qual63_se_k100", _requalify(base, 63, 1), k=100)
    add("qual64_se_k100", _requalify(base, 64, 2), k=100)
    add("qual64_nqual_se_k100", _requalify(base, 64, 3, n_with_quality=ord("#"), n_rate=0.004), k=100)
    add("qual63_nqual_se_k100", _requalify(base, 63, 4, n_with_quality=ord("#"), n_rate=0.004), k=100)
    add("qual65_se_k100", _requalify(base, 65, 5), k=100)
    add("qual90_se_k100", _requalify(base, 90, 6), k=100)
    p1, p2 = fqgen.generate(900, seed=42, paired=True)
    add("qual70_pe_k100", _requalify(p1, 70, 7), _requalify(p2, 70, 8), k=100)       # raw column with reversed R2 qualities, overlaps
    add("qual64_nqual_pe_k100", _requalify(p1, 64, 9, n_with_quality=ord("#"), n_rate=0.004), _requalify(p2, 64, 10, n_with_quality=ord("#"), n_rate=0.004), k=100)

    # the reference's reader (src/fastqreader.cpp:94-156): lone '\r' line ends, blank lines it swallows, a mix of all three breaks,
    # no record at all (an EMPTY output file, src/repaq.cpp:530-638), an empty line that ends the input early
    r1, r2 = fqgen.generate(900, seed=43, paired=True)
    b1, b2 = bytes(r1), bytes(r2)
    add("reader_cr_se_k100", b1.replace(b"\n", b"\r"), k=100)
    add("reader_cr_pe_k100", b1.replace(b"\n", b"\r"), b2.replace(b"\n", b"\r"), k=100)
    lines = b1.split(b"\n")[:-1]
    add("reader_mixed_breaks_se_k100", b"".join(ln + (b"\n", b"\r", b"\r\n")[i % 3] for i, ln in enumerate(lines)), k=100)
    recs = records(b1)
    add("reader_blank_lines_se_k100", b"".join(b"\n".join(r) + (b"\n\n" if i % 7 == 3 else b"\n") for i, r in enumerate(recs)), k=100)
    add("reader_blank_inside_record_pe_k100", b1.replace(b"\n+\n", b"\n\n+\n", 40), b2, k=100)
    add("reader_two_blank_lines_se_k100", join(recs[:500]) + b"\n\n" + join(recs[500:]), k=100)    # the second one is an empty line: input ends
    add("reader_empty_se", b"")
    add("reader_empty_pe", b"", b"")
    add("reader_leading_blank_se", b"\n" + KAT_A1)
    add("reader_r2_shorter_pe_k100", b1, join(records(b2)[:700]), k=100)              # pairs end with the shorter file
    return cases
