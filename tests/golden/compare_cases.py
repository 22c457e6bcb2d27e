"""Inputs of the compare-mode golden vectors: (name, golden .rfq it is run against, FASTQ image(s)).  Every FASTQ side is a
deterministic edit of the inputs in cases.py, so only the reference's JSON reports (compare_manifest.json) are committed."""
from tests.golden.cases import build_cases, join, records


def build_compare_cases():
    cs = {c["name"]: c for c in build_cases()}
    out = []

    def add(name, rfq, r1, r2=None):
        out.append(dict(name=name, rfq=rfq, r1=bytes(r1), r2=None if r2 is None else bytes(r2)))

    # untouched inputs: passes, and the lossy cases the reference itself reports as different (leading zeros, late quality)
    for n in ("kat_se", "kat_pe", "nova_pe_k1000", "nova_se_k100", "bgi_se_k100", "nova_pe_nonl_k100", "nova_pe_crlf_k100", "nova_pe_300bp_varlen_k100",
              "se_strand_varies_k100", "names_numeric_edge", "names_numeric_edge_pe", "nova_se_late_quality", "pe_demoted_mid_k100", "nova_interleaved_in_k100"):
        c = cs[n]
        if c["interleaved"]:
            continue                                        # compare mode has no --interleaved_in: checked as single end below
        add("same_" + n, n, c["r1"], c["r2"])
    c = cs["nova_interleaved_in_k100"]
    add("interleaved_as_se", "nova_interleaved_in_k100", c["r1"])

    c = cs["nova_pe_k1000"]
    a, b = records(c["r1"]), records(c["r2"])

    def edit(recs, i, field, fn):
        r = [list(x) for x in recs]
        r[i][field] = fn(r[i][field])
        return join(r)
    add("pe_seq_r1", "nova_pe_k1000", edit(a, 100, 1, lambda s: s[:50] + (b"N" if s[50:51] != b"N" else b"A") + s[51:]), c["r2"])
    add("pe_qual_r2_last_byte", "nova_pe_k1000", c["r1"], edit(b, 200, 3, lambda s: s[:-1] + b"!"))
    add("pe_name_r2", "nova_pe_k1000", c["r1"], edit(b, 3599, 0, lambda s: s + b"x"))
    add("pe_strand_r1_chunk1", "nova_pe_k1000", edit(a, 3400, 2, lambda s: b"+x"), c["r2"])
    add("pe_seq_shorter", "nova_pe_k1000", edit(a, 7, 1, lambda s: s[:-1]), c["r2"])
    add("pe_two_diffs_first_wins", "nova_pe_k1000", edit(a, 900, 3, lambda s: b"#" + s[1:]), edit(b, 12, 0, lambda s: s[:-1]))
    add("pe_name_and_qual_same_read", "nova_pe_k1000", join([[r[0] + b"y", r[1], r[2], b"!" + r[3][1:]] if i == 5 else r for i, r in enumerate(a)]), c["r2"])
    add("pe_fastq_shorter", "nova_pe_k1000", join(a[:1000]), join(b[:1000]))
    add("pe_fastq_shorter_r2_only", "nova_pe_k1000", c["r1"], join(b[:1000]))
    add("pe_fastq_longer", "nova_pe_k1000", c["r1"] + join(a[:3]), c["r2"] + join(b[:3]))
    add("pe_fastq_longer_r1_only", "nova_pe_k1000", c["r1"] + join(a[:3]), c["r2"])
    add("pe_empty_r2", "nova_pe_k1000", c["r1"], b"")
    add("pe_rfq_vs_se_call", "nova_pe_k1000", c["r1"])       # a PE .rfq compared in single-end mode: read 2 is R2, not record 2 of R1
    add("pe_crlf_fastq", "nova_pe_k1000", c["r1"].replace(b"\n", b"\r\n"), c["r2"].replace(b"\n", b"\r\n"))

    c = cs["nova_se_k100"]
    a = records(c["r1"])
    add("se_name", "nova_se_k100", edit(a, 5, 0, lambda s: s + b"x"))
    add("se_strand", "nova_se_k100", edit(a, 7, 2, lambda s: b"+abc"))
    add("se_qual_longer", "nova_se_k100", edit(a, 2099, 3, lambda s: s + b"F"))
    add("se_fastq_longer", "nova_se_k100", c["r1"] + join(a[:1]))
    add("se_fastq_shorter", "nova_se_k100", join(a[:10]))
    add("se_fastq_empty", "nova_se_k100", b"")
    add("se_fastq_stops_at_empty_line", "nova_se_k100", join(a[:50]) + b"\n" + join(a[50:]))
    add("se_no_final_newline", "nova_se_k100", c["r1"][:-1])
    # (a truncated .rfq is not a case: the reference binary dies with SIGSEGV on it)
    return out
