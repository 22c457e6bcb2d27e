"""Test helper: split a serialised .rfq image into header + per-chunk column dicts (wire order of
RfqChunk::write, reference src/rfqchunk.cpp:230-312) so that mismatches can be reported per column."""
import struct


def parse(rfq):
    assert rfq[:3] == b"RFQ"
    h = dict(rlb=rfq[9], flags=rfq[10] | (rfq[11] << 8), diff_pos=rfq[12], diff_char=rfq[13], nq=rfq[14], shift=rfq[15], bins=rfq[16])
    h["qual_buf"] = rfq[17:17 + h["bins"]]
    at = 17 + h["bins"]
    chunks = []
    F = h["flags"]
    while at < len(rfq):
        c = {}
        s = at
        c["msize"], c["reads"], c["flags"], c["seq_size"], c["qual_size"] = struct.unpack_from("<IIHII", rfq, at)
        at += 18
        if F & (1 << 9):
            c["npos_size"], = struct.unpack_from("<I", rfq, at)
            at += 4
        n, fl = c["reads"], c["flags"]
        il = bool(fl & (1 << 9))
        xy = n // 2 if il else n

        def take(name, size):
            nonlocal at
            c[name] = rfq[at:at + size]
            at += size
        take("readlen", h["rlb"] * (1 if fl & 1 else n))
        take("n1len", 1 if fl & 2 else n)
        if F & 16:
            take("n2len", 1 if fl & 4 else n)
        take("slen", 1 if fl & 8 else n)
        if F & 1:
            take("lane", 1 if fl & 16 else xy)
        if F & 2:
            take("tile", 2 * (1 if fl & 32 else xy))
        if F & 4:
            c["x_size"], = struct.unpack_from("<I", rfq, at); at += 4
            take("x", c["x_size"])
        if F & 8:
            c["y_size"], = struct.unpack_from("<I", rfq, at); at += 4
            take("y", c["y_size"])

        def arena(lens, len_same, all_same):
            t = sum(lens)
            if len_same and not all_same:
                t *= n
            return t
        take("n1", arena(c["n1len"], fl & 2, fl & 64))
        if F & 16:
            take("n2", arena(c["n2len"], fl & 4, fl & 128))
        take("strand", arena(c["slen"], fl & 8, fl & 256))
        take("seq", c["seq_size"])
        take("qual", c["qual_size"])
        if il and (F & 64):
            take("ov", n // 2)
        if F & (1 << 9):
            take("npos", c["npos_size"])
        c["bytes"] = at - s
        chunks.append(c)
    return h, chunks


def diff(a, b):
    """human-readable first differences between two .rfq images"""
    out = []
    ha, ca = parse(a)
    hb, cb = parse(b)
    for k in ha:
        if ha[k] != hb[k]:
            out.append(f"header.{k}: {ha[k]!r} != {hb[k]!r}")
    if len(ca) != len(cb):
        out.append(f"chunks: {len(ca)} != {len(cb)}")
    for i, (x, y) in enumerate(zip(ca, cb)):
        for k in x:
            if k not in y:
                out.append(f"chunk {i}: column {k} missing on the right")
            elif x[k] != y[k]:
                if isinstance(x[k], (bytes, bytearray)):
                    n = min(len(x[k]), len(y[k]))
                    d = next((j for j in range(n) if x[k][j] != y[k][j]), n)
                    out.append(f"chunk {i}.{k}: len {len(x[k])} vs {len(y[k])}, first diff at {d}: {x[k][max(0,d-4):d+12].hex()} vs {y[k][max(0,d-4):d+12].hex()}")
                else:
                    out.append(f"chunk {i}.{k}: {x[k]} != {y[k]}")
        if len(out) > 30:
            break
    return out
