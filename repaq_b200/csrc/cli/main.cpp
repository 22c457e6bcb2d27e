/*
 * repaq_b200_cli - host C++ driver with the reference's command line for the modes on the hot path
 * (`repaq -c`, `repaq -d`, `repaq --compare`; reference src/main.cpp:31-49, src/repaq.cpp:36-413,530-759) on top of the C ABI.
 * File-level rules kept from the reference: header written first (src/repaq.cpp:554-557), Q13 NO_LINE_BREAK thresholds
 * (src/fastqreader.cpp:31-46), trailing-newline trimming on decode (src/repaq.cpp:300-328, 375-413).
 * Compare mode prints the reference's JSON report (src/repaq.cpp:235-259).
 * -v / -f: every batch is decoded again and checked against its input on the GPU (rpq_compare) after it has been written, and the
 * first difference is reported on stderr in the words of completeCheckAndOutput (src/repaq.cpp:430-528); like the reference, the
 * output is written either way.  (-f checks a tenth of the chunks there to save CPU time; here both check everything.)
 * Not implemented here: .gz input/output, the xz pipe (outside the tier's scope).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "repaq_b200.h"

static void error_exit(const std::string& msg) { fprintf(stderr, "ERROR: %s\n", msg.c_str()); exit(-1); }   /* src/util.h:246-249 */

static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }

static std::vector<char> slurp(const std::string& path) {
    FILE* f = path == "/dev/stdin" ? stdin : fopen(path.c_str(), "rb");
    if (!f) error_exit("Failed to open file: " + path);
    std::vector<char> b;
    size_t cap = 1 << 26, n = 0;
    b.resize(cap);
    for (;;) {
        size_t got = fread(b.data() + n, 1, cap - n, f);
        n += got;
        if (got == 0) break;
        if (n == cap) { cap *= 2; b.resize(cap); }
    }
    if (f != stdin) fclose(f);
    b.resize(n);
    return b;
}
static void spill(const std::string& path, const void* p, size_t n, bool append) {
    FILE* f = path == "/dev/stdout" ? stdout : fopen(path.c_str(), append ? "ab" : "wb");
    if (!f) error_exit("Failed to open file: " + path);
    if (n && fwrite(p, 1, n, f) != n) error_exit("Failed to write file: " + path);
    if (f != stdout) fclose(f);
}

/* Q13: the reader raises hasNoLineBreakAtEnd when it loads a SHORT 1 MiB buffer that does not end in '\n' */
static void nobreak_rule(const std::vector<char>& f, uint64_t& from, bool& tail) {
    const uint64_t MiB = 1ull << 20, n = f.size();
    const bool nl = n && f[n - 1] == '\n';
    if (n % MiB == 0) { from = nl ? UINT64_MAX : n; tail = true; }
    else { from = nl ? UINT64_MAX : (n / MiB) * MiB; tail = false; }
}

struct Opt { std::string in1, in2, out1, out2, rfq_compare, json_compare; bool compress = false, decompress = false, compare = false, interleaved = false, to_stdout = false, from_stdin = false, verify = false; int k = 1000; int device = 0; };

static int do_compress(const Opt& o) {
    std::vector<char> r1 = slurp(o.in1), r2;
    const bool two = !o.in2.empty();
    if (two) r2 = slurp(o.in2);
    const uint32_t chunk_bases = (uint32_t)(o.k < 100 ? 100 : o.k) * 1000u;        /* src/main.cpp:69 */
    char err[768];
    rpq_header h;
    const int hrc = rpq_make_header(r1.data(), r1.size(), two ? r2.data() : NULL, r2.size(), o.interleaved, chunk_bases, &h, err, sizeof err);
    if (hrc == RPQ_NO_RECORDS) { spill(o.out1, NULL, 0, false); return 0; }      /* no record: an empty output, as the reference leaves it */
    if (hrc) error_exit(err);
    uint8_t hb[17 + 128];
    const size_t hn = rpq_header_write(&h, hb, sizeof hb);
    spill(o.out1, hb, hn, false);
    rpq_ctx* ctx = NULL;
    if (rpq_create(o.device, &ctx)) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    uint64_t from1, from2 = UINT64_MAX; bool t1, t2 = false;
    nobreak_rule(r1, from1, t1);
    if (two) nobreak_rule(r2, from2, t2); else if (o.interleaved) { from2 = from1; t2 = t1; }
    uint64_t WIN = 3ull << 30;                             /* < 4 GiB of text per file and call */
    if (const char* e = getenv("RPQ_CLI_FQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) WIN = w; }      /* tests: small batches */
    uint64_t a = 0, b = 0;
    rpq_ctx* check = NULL;                                 /* -v / -f: the reference's codec4check */
    for (;;) {
        rpq_encode_in in; memset(&in, 0, sizeof in);
        const uint64_t n1 = r1.size() - a < WIN ? r1.size() - a : WIN, n2 = two ? (r2.size() - b < WIN ? r2.size() - b : WIN) : 0;
        in.r1 = r1.data() + a; in.r1_len = n1; in.r2 = two ? r2.data() + b : NULL; in.r2_len = n2;
        in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.interleaved = o.interleaved; in.chunk_bases = chunk_bases;
        in.final = (a + n1 == r1.size()) && (!two || b + n2 == r2.size());
        in.file_offset[0] = a; in.file_offset[1] = b;
        in.nobreak_from[0] = from1 == UINT64_MAX ? UINT64_MAX : (from1 > a ? from1 - a : 0);
        in.nobreak_from[1] = from2 == UINT64_MAX ? UINT64_MAX : (two ? (from2 > b ? from2 - b : 0) : in.nobreak_from[0]);
        in.tail_flags = (uint16_t)((t1 ? RPQ_NO_LINE_BREAK_AT_END : 0) | (t2 ? RPQ_NO_LINE_BREAK_AT_END_R2 : 0));
        rpq_encode_out res;
        if (rpq_encode(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        spill(o.out1, res.data, res.bytes, true);
        if (o.verify && res.bytes) {
            /* the check of completeCheckAndOutput: decode what was just written, compare it read by read with what it was made from */
            if (!check && (rpq_create(o.device, &check) || rpq_set_header(check, &h))) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
            rpq_compare_in ci; memset(&ci, 0, sizeof ci);
            ci.rfq = res.data; ci.rfq_bytes = res.bytes; ci.rfq_mem = RPQ_MEM_HOST; ci.rfq_final = 1;
            ci.r1 = in.r1; ci.r1_len = in.final ? in.r1_len : res.r1_consumed; ci.r2 = in.r2; ci.r2_len = in.final ? in.r2_len : res.r2_consumed;
            ci.fq_mem = RPQ_MEM_HOST; ci.fq_final = 1;
            rpq_compare_out co;
            if (o.interleaved) fprintf(stderr, "verify: --interleaved_in input is not checked\n");
            else if (rpq_compare(check, &ci, &co)) error_exit(rpq_last_error(check));
            else if (co.verdict == RPQ_CMP_RFQ_MORE || co.verdict == RPQ_CMP_FASTQ_MORE) error_exit("encoding error in chunk, the output will be wrong, quit now!");
            else if (co.verdict != RPQ_CMP_EQUAL)
                fprintf(stderr, "integrity check failure \nexpected: \n%.*s\ngot:\n%.*s\n", (int)co.fastq_field_len, co.fastq_field, (int)co.rfq_field_len, co.rfq_field);
        }
        if (in.final) break;
        if (res.r1_consumed == 0) {                        /* no whole chunk in this batch */
            if (WIN >= (3ull << 30)) error_exit("a chunk does not fit the 3 GiB batch window; lower --chunk");
            WIN = WIN * 2 < (3ull << 30) ? WIN * 2 : (3ull << 30);
            continue;
        }
        a += res.r1_consumed; b += res.r2_consumed;
    }
    if (check) rpq_destroy(check);
    rpq_destroy(ctx);
    return 0;
}

/* Repaq::decompress / decompressPE (src/repaq.cpp:262-413).  The .rfq is decoded in windows of whole chunks (1 GiB of .rfq by
 * default, RPQ_CLI_RFQ_WINDOW=<bytes> for tests), so that device and pinned host memory stay bounded whatever the file size.
 * Whether a chunk is the LAST one of the file (the trailing-newline rule) is only known once the bytes after it have failed to
 * decode as a chunk, so the last chunk of a window that is not the end of the file is held back and decoded again as the first
 * chunk of the next window. */
static int do_decompress(const Opt& o) {
    std::vector<char> rfq = slurp(o.in1);
    char err[768]; rpq_header h; size_t used = 0;
    if (rpq_header_read((const uint8_t*)rfq.data(), rfq.size(), &h, &used, err, sizeof err)) error_exit(err);
    const bool pe = !o.out2.empty();
    if (pe && !(h.flags & RPQ_PAIRED_END)) error_exit("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
    rpq_ctx* ctx = NULL;
    if (rpq_create(o.device, &ctx)) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    uint64_t window = 1ull << 30;
    if (const char* e = getenv("RPQ_CLI_RFQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) window = w; }
    spill(o.out1, NULL, 0, false);
    if (pe) spill(o.out2, NULL, 0, false);
    uint64_t at = used;
    bool skip_first = false;                               /* decompressPE's `continue`: the chunk after a flagged one is never written */
    while (at < rfq.size()) {
        const uint64_t n = rfq.size() - at < window ? rfq.size() - at : window;
        const bool final = at + n == rfq.size();
        rpq_decode_in in; memset(&in, 0, sizeof in);
        in.data = (const uint8_t*)rfq.data() + at; in.bytes = n; in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.split_pairs = pe;
        rpq_decode_out res;
        if (rpq_decode(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        if (!final && res.n_chunks < 2) { window *= 2; continue; }          /* a window must hold a chunk to write and one to hold back */
        if (res.n_chunks == 0) break;                                       /* what is left is not a chunk: the reference stops here too */
        const uint32_t n_keep = final ? res.n_chunks : res.n_chunks - 1;
        uint64_t a1 = 0, a2 = 0;
        for (uint32_t i = 0; i < n_keep; i++) {
            const rpq_chunk_info& c = res.chunks[i];
            const bool last = final && i + 1 == res.n_chunks;
            const bool f1 = (c.flags & RPQ_NO_LINE_BREAK_AT_END) != 0, f2 = (c.flags & RPQ_NO_LINE_BREAK_AT_END_R2) != 0;
            if (skip_first) { skip_first = false; a1 += c.out1_bytes; a2 += c.out2_bytes; continue; }
            if (!pe) {
                /* Repaq::decompress: only a flagged LAST chunk loses its final newline */
                spill(o.out1, res.out1 + a1, (f1 && last && c.out1_bytes) ? c.out1_bytes - 1 : c.out1_bytes, true);
            } else {
                /* Repaq::decompressPE incl. its `continue` (src/repaq.cpp:395,405): after a flagged chunk that is not the last one,
                 * the rest of that chunk's output and the whole next chunk are never written */
                bool skip_next = false;
                if (f1) { if (last) spill(o.out1, res.out1 + a1, c.out1_bytes ? c.out1_bytes - 1 : 0, true); else { spill(o.out1, res.out1 + a1, c.out1_bytes, true); skip_next = true; } }
                else spill(o.out1, res.out1 + a1, c.out1_bytes, true);
                if (!skip_next) {
                    if (f2) { if (last) spill(o.out2, res.out2 + a2, c.out2_bytes ? c.out2_bytes - 1 : 0, true); else { spill(o.out2, res.out2 + a2, c.out2_bytes, true); skip_next = true; } }
                    else spill(o.out2, res.out2 + a2, c.out2_bytes, true);
                }
                skip_first = skip_next;
            }
            a1 += c.out1_bytes; a2 += c.out2_bytes;
        }
        if (final) break;
        at += res.chunks[res.n_chunks - 1].offset;                          /* the held-back chunk starts the next window */
    }
    rpq_destroy(ctx);
    return 0;
}

/* Repaq::compare / comparePE (src/repaq.cpp:36-233): the .rfq is decoded and checked read by read against the FASTQ file(s) on
 * the GPU, in batches of whole chunks against windows of FASTQ text; the report is reportCompareResult's (:235-259). */
static int do_compare(const Opt& o) {
    std::vector<char> rfq = slurp(o.rfq_compare), r1 = slurp(o.in1), r2;
    const bool pe = !o.in2.empty();
    if (pe) r2 = slurp(o.in2);
    char err[768]; rpq_header h; size_t used = 0;
    if (rpq_header_read((const uint8_t*)rfq.data(), rfq.size(), &h, &used, err, sizeof err)) error_exit(err);
    rpq_ctx* ctx = NULL;
    if (rpq_create(o.device, &ctx)) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    uint64_t FQ_WIN = 3ull << 30;                          /* < 4 GiB of text per file and call */
    uint64_t rfq_win = 256ull << 20;                       /* chunks decoded per call: ~2 GB of FASTQ at the usual ratios */
    if (const char* e = getenv("RPQ_CLI_RFQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) rfq_win = w; }      /* tests: small windows */
    if (const char* e = getenv("RPQ_CLI_FQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) FQ_WIN = w; }
    uint64_t at = used, a = 0, b = 0;
    unsigned long long fq_reads = 0, fq_bases = 0, rfq_reads = 0, rfq_bases = 0;
    bool passed = false; std::string msg;
    for (;;) {
        rpq_compare_in in; memset(&in, 0, sizeof in);
        const uint64_t nq = rfq.size() - at < rfq_win ? rfq.size() - at : rfq_win;
        const uint64_t n1 = r1.size() - a < FQ_WIN ? r1.size() - a : FQ_WIN, n2 = pe ? (r2.size() - b < FQ_WIN ? r2.size() - b : FQ_WIN) : 0;
        in.rfq = (const uint8_t*)rfq.data() + at; in.rfq_bytes = nq; in.rfq_mem = RPQ_MEM_HOST; in.rfq_final = at + nq == rfq.size();
        in.r1 = r1.data() + a; in.r1_len = n1; in.r2 = pe ? (r2.empty() ? "" : r2.data() + b) : NULL; in.r2_len = n2; in.fq_mem = RPQ_MEM_HOST;
        in.fq_final = (a + n1 == r1.size()) && (!pe || b + n2 == r2.size());
        in.fq_offset[0] = a; in.fq_offset[1] = b;
        rpq_compare_out res;
        if (rpq_compare(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        if (res.verdict == RPQ_CMP_NEED_FASTQ) {           /* these chunks decode to more reads than the window of text holds: more text, or fewer chunks */
            if (FQ_WIN < (3ull << 30)) { FQ_WIN = FQ_WIN * 2 < (3ull << 30) ? FQ_WIN * 2 : (3ull << 30); continue; }
            if (rfq_win <= (1ull << 20)) error_exit("compare: a batch of chunks does not fit the FASTQ window");
            rfq_win /= 2; continue;
        }
        fq_reads += res.fastq_reads; fq_bases += res.fastq_bases; rfq_reads += res.rfq_reads; rfq_bases += res.rfq_bases;
        const unsigned long long sr = pe ? rfq_reads / 2 : rfq_reads, sf = pe ? fq_reads / 2 : fq_reads;
        const char* unit = pe ? "pair" : "read";
        if (res.verdict >= RPQ_CMP_NAME && res.verdict <= RPQ_CMP_QUALITY) {
            static const char* what[4] = {"name", "sequence", "strand", "quality"};
            msg = std::string("The RFQ file and FASTQ file have different ") + what[res.verdict - RPQ_CMP_NAME] + " in the " + std::to_string(sr) + " " + unit + ". " +
                  std::string(res.rfq_field, res.rfq_field_len) + " | " + std::string(res.fastq_field, res.fastq_field_len);
            break;
        }
        if (res.verdict == RPQ_CMP_RFQ_MORE) {
            msg = "The RFQ file has more reads than the FASTQ file. The RFQ file has >= " + std::to_string(sr) + " " + unit + "s, while the FASTQ file only has " + std::to_string(sf) + " " + unit + "s";
            break;
        }
        if (res.verdict == RPQ_CMP_FASTQ_MORE) {
            msg = "The FASTQ file has more reads than the RFQ file. The FASTQ file has >= " + std::to_string(sf) + " " + unit + "s, while the RFQ file only has " + std::to_string(sr) + " " + unit + "s";
            break;
        }
        /* equal so far */
        a += res.r1_consumed; b += res.r2_consumed; at += res.rfq_consumed;
        if (in.rfq_final) at = rfq.size();                 /* whatever follows the last whole chunk is not a chunk: the reference stops there too */
        else if (res.rfq_consumed == 0) { if (rfq_win >= (2ull << 30)) error_exit("compare: a chunk does not fit the batch window"); rfq_win *= 2; continue; }
        if (at == rfq.size() && in.fq_final) { passed = true; break; }
        if (res.rfq_consumed == 0 && res.r1_consumed == 0 && !in.fq_final) {          /* the chunks are through and this window of text holds no whole record */
            if (FQ_WIN >= (3ull << 30)) error_exit("compare: a record does not fit the FASTQ window");
            FQ_WIN = FQ_WIN * 2 < (3ull << 30) ? FQ_WIN * 2 : (3ull << 30);
        }
    }
    std::string json = "{\n";
    json += passed ? "\t\"result\":\"passed\",\n" : "\t\"result\":\"failed\",\n";
    json += "\t\"msg\":\"" + msg + "\",\n";
    json += "\t\"fastq_reads\":" + std::to_string(fq_reads) + ",\n";
    json += "\t\"rfq_reads\":" + std::to_string(rfq_reads) + ",\n";
    json += "\t\"fastq_bases\":" + std::to_string(fq_bases) + ",\n";
    json += "\t\"rfq_bases\":" + std::to_string(rfq_bases) + "\n";
    json += "}\n";
    if (!o.json_compare.empty()) spill(o.json_compare, json.data(), json.size(), false);
    fwrite(json.data(), 1, json.size(), stdout);
    rpq_destroy(ctx);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 1) { fprintf(stderr, "repaq_b200: repack FASTQ to a smaller binary file (.rfq) on a B200\nversion 0.5.1 (algorithm 2)\n"); return 0; }
    if (argc == 2 && strcmp(argv[1], "--version") == 0) { printf("repaq 0.5.1\n"); return 0; }
    Opt o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* lng) -> std::string {
            std::string pre = std::string("--") + lng + "=";
            if (a.compare(0, pre.size(), pre) == 0) return a.substr(pre.size());
            if (i + 1 >= argc) error_exit("option needs value: " + a);
            return argv[++i];
        };
        if (a == "-i" || a.compare(0, 6, "--in1=") == 0 || a == "--in1") o.in1 = val("in1");
        else if (a == "-I" || a.compare(0, 6, "--in2=") == 0 || a == "--in2") o.in2 = val("in2");
        else if (a == "-o" || a.compare(0, 7, "--out1=") == 0 || a == "--out1") o.out1 = val("out1");
        else if (a == "-O" || a.compare(0, 7, "--out2=") == 0 || a == "--out2") o.out2 = val("out2");
        else if (a == "-k" || a.compare(0, 8, "--chunk=") == 0 || a == "--chunk") o.k = atoi(val("chunk").c_str());
        else if (a == "-r" || a.compare(0, 17, "--rfq_to_compare=") == 0 || a == "--rfq_to_compare") o.rfq_compare = val("rfq_to_compare");
        else if (a == "-j" || a.compare(0, 22, "--json_compare_result=") == 0 || a == "--json_compare_result") o.json_compare = val("json_compare_result");
        else if (a == "-p" || a == "--compare") o.compare = true;
        else if (a == "-v" || a == "--verify" || a == "-f" || a == "--fast_verify") o.verify = true;
        else if (a == "-c" || a == "--compress") o.compress = true;
        else if (a == "-d" || a == "--decompress") o.decompress = true;
        else if (a == "--interleaved_in") o.interleaved = true;
        else if (a == "--stdout") o.to_stdout = true;
        else if (a == "--stdin") o.from_stdin = true;
        else if (a.compare(0, 9, "--device=") == 0) o.device = atoi(a.c_str() + 9);
        else error_exit("unsupported option for the B200 driver: " + a);
    }
    if ((int)o.compress + (int)o.decompress + (int)o.compare > 1) error_exit("repaq can run in compress/decompress/compare mode, you can only choose any one mode.");
    if (o.compare) {
        if (o.in1.empty()) error_exit("Please specify input file by <in1>, or enable --stdin if you want to read STDIN");
        if (o.rfq_compare.empty()) error_exit("In compare mode, you should specify the RFQ file to compare by <rfq_to_compare>");
        return do_compare(o);
    }
    if (!o.decompress) o.compress = true;                                   /* compress is the default mode */
    if (o.in1.empty()) { if (o.from_stdin) o.in1 = "/dev/stdin"; else error_exit("Please specify input file by <in1>, or enable --stdin if you want to read STDIN"); }
    if (o.out1.empty()) { if (o.to_stdout) o.out1 = "/dev/stdout"; else error_exit("Please specify output file by <out1>, or enable --stdout if you want to read STDIN"); }
    if (ends_with(o.in1, ".gz") || ends_with(o.in2, ".gz") || ends_with(o.out1, ".gz") || ends_with(o.in1, ".xz") || ends_with(o.out1, ".xz"))
        error_exit("gz / xz streams are outside this driver (use zcat / xz pipes with --stdin / --stdout)");
    const long long cs = (long long)(o.k < 100 ? 100 : o.k) * 1000;
    if (cs > 500000000) error_exit("chunk size cannot be greater than 500,000 kb");
    if (o.compress) { if (!o.out2.empty()) error_exit("In compress mode, only one RFQ output file is allowed, but you specified <out2>"); return do_compress(o); }
    if (!o.in2.empty()) error_exit("In decompress mode, only one RFQ input file is allowed, but you specified <in2>");
    return do_decompress(o);
}
