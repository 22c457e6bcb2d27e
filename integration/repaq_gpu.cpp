/*
 * repaq_gpu.cpp - the file a maintainer of OpenGene/repaq adds to src/ to run the codec on a B200 through librepaq_b200.so.
 *
 * It is compiled against the reference's own headers (repaq.h, options.h, util.h, writer.h) and defines Repaq::run(): the
 * dispatcher of src/repaq.cpp:12-34, sending each mode to the library instead of to RfqCodec.  Everything above it - main(),
 * cmdline parsing, Options::validate(), the xz re-exec (src/main.cpp) - and everything beside it (Writer with its .gz output,
 * reportCompareResult) is the reference's unchanged code.  In a real tree the maintainer edits the body of Repaq::run in
 * src/repaq.cpp; the reference checkout here is read-only, so oracle/Makefile (target ref_gpu) compiles src/repaq.cpp with
 * -Drun=run_reference_cpu - the stock dispatcher keeps existing under that name - and links this definition in its place.
 *
 * Built into oracle/_ref_gpu/repaq (CUDA library) and oracle/_ref_gpu/repaq_emu (the CPU emulation build of the same kernels,
 * for the GPU-less test container).  tests/test_integration.py runs the golden vectors through both.
 */
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <fstream>
#include <string>
#include <vector>

#include "repaq_b200.h"
#include "repaq.h"
#include "util.h"
#include "writer.h"

namespace {

/* the whole file; `.gz` through zlib like FastqReader (src/fastqreader.cpp:48-66) */
std::vector<char> slurp(const std::string& path) {
    std::vector<char> b;
    size_t n = 0;
    if (ends_with(path, ".gz")) {
        gzFile g = gzopen(path.c_str(), "r");
        if (!g) error_exit("Failed to open file: " + path);
        b.resize(1 << 24);
        for (;;) {
            const int k = gzread(g, b.data() + n, (unsigned)(b.size() - n));
            if (k < 0) error_exit("Error to read gzip file");
            if (k == 0) break;
            n += (size_t)k;
            if (n == b.size()) b.resize(b.size() * 2);
        }
        gzclose(g);
    } else {
        FILE* f = path == "/dev/stdin" ? stdin : fopen(path.c_str(), "rb");
        if (!f) error_exit("Failed to open file: " + path);
        b.resize(1 << 24);
        for (;;) {
            const size_t k = fread(b.data() + n, 1, b.size() - n, f);
            n += k;
            if (k == 0) break;
            if (n == b.size()) b.resize(b.size() * 2);
        }
        if (f != stdin) fclose(f);
    }
    b.resize(n);
    return b;
}

/* Q13: FastqReader raises mHasNoLineBreakAtEnd when it loads a short 1 MiB buffer not ending in '\n' (src/fastqreader.cpp:42-45) */
void nobreak_rule(const std::vector<char>& f, uint64_t& from, bool& tail) {
    const uint64_t MiB = 1 << 20, n = f.size();
    const bool nl = n && f[n - 1] == '\n';
    if (n % MiB == 0) { from = nl ? UINT64_MAX : n; tail = true; }
    else { from = nl ? UINT64_MAX : (n / MiB) * MiB; tail = false; }
}

rpq_ctx* context(const rpq_header& h) {
    rpq_ctx* ctx = NULL;
    if (rpq_create(0, &ctx)) error_exit("no CUDA device: librepaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    return ctx;
}

/* Repaq::compress / compressPE (src/repaq.cpp:530-759) */
void compress_gpu(Options* opt) {
    std::vector<char> r1 = slurp(opt->in1), r2;
    const bool two = !opt->in2.empty();
    if (two) r2 = slurp(opt->in2);
    std::ofstream out(opt->out1, std::ios::out | std::ios::binary);
    char err[768];
    rpq_header h;                                                      /* RfqCodec::makeHeader on the first chunk */
    const int hrc = rpq_make_header(r1.data(), r1.size(), two ? r2.data() : NULL, r2.size(), opt->interleavedInput, (uint32_t)opt->chunkSize, &h, err, sizeof err);
    if (hrc == RPQ_NO_RECORDS) return;                                 /* nothing is written, not even the header */
    if (hrc) error_exit(err);
    uint8_t hb[17 + 128];
    out.write((const char*)hb, (std::streamsize)rpq_header_write(&h, hb, sizeof hb));
    rpq_ctx* ctx = context(h);                                         /* RfqCodec::setHeader */
    uint64_t from1, from2 = UINT64_MAX; bool t1, t2 = false;
    nobreak_rule(r1, from1, t1);
    if (two) nobreak_rule(r2, from2, t2); else if (opt->interleavedInput) { from2 = from1; t2 = t1; }
    const uint64_t WIN = 3ull << 30;                                   /* < 4 GiB of text per file and call */
    uint64_t a = 0, b = 0;
    for (;;) {
        rpq_encode_in in; memset(&in, 0, sizeof in);
        const uint64_t n1 = r1.size() - a < WIN ? r1.size() - a : WIN, n2 = two ? (r2.size() - b < WIN ? r2.size() - b : WIN) : 0;
        in.r1 = r1.data() + a; in.r1_len = n1; in.r2 = two ? r2.data() + b : NULL; in.r2_len = n2;
        in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.interleaved = opt->interleavedInput; in.chunk_bases = (uint32_t)opt->chunkSize;
        in.final = (a + n1 == r1.size()) && (!two || b + n2 == r2.size());
        in.file_offset[0] = a; in.file_offset[1] = b;
        in.nobreak_from[0] = from1 == UINT64_MAX ? UINT64_MAX : (from1 > a ? from1 - a : 0);
        in.nobreak_from[1] = from2 == UINT64_MAX ? UINT64_MAX : (two ? (from2 > b ? from2 - b : 0) : in.nobreak_from[0]);
        in.tail_flags = (uint16_t)((t1 ? RPQ_NO_LINE_BREAK_AT_END : 0) | (t2 ? RPQ_NO_LINE_BREAK_AT_END_R2 : 0));
        rpq_encode_out res;                                            /* every RfqCodec::encodeChunk + RfqChunk::write of the batch */
        if (rpq_encode(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        out.write((const char*)res.data, (std::streamsize)res.bytes);
        if (in.final) break;
        if (res.r1_consumed == 0) error_exit("a chunk does not fit the 3 GiB batch window; lower --chunk");
        a += res.r1_consumed; b += res.r2_consumed;
    }
    out.flush();
    rpq_destroy(ctx);
}

/* Repaq::decompress / decompressPE (src/repaq.cpp:262-413), output through the reference's Writer (.gz by file name) */
void decompress_gpu(Options* opt) {
    std::vector<char> rfq = slurp(opt->in1);
    const bool pe = !opt->out2.empty();
    Writer w1(opt->out1);
    Writer* w2 = pe ? new Writer(opt->out2) : NULL;
    if (rfq.empty()) {                                                 /* RfqHeader::read keeps the constructor's values: single end, no chunks */
        if (pe) error_exit("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
        return;
    }
    char err[768]; rpq_header h; size_t used = 0;
    if (rpq_header_read((const uint8_t*)rfq.data(), rfq.size(), &h, &used, err, sizeof err)) error_exit(err);
    if (pe && !(h.flags & RPQ_PAIRED_END)) error_exit("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
    rpq_ctx* ctx = context(h);
    rpq_decode_in in; memset(&in, 0, sizeof in);
    in.data = (const uint8_t*)rfq.data() + used; in.bytes = rfq.size() - used; in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.split_pairs = pe;
    rpq_decode_out res;                                                /* every RfqChunk::read + RfqCodec::decodeChunk + Read::toString */
    if (rpq_decode(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
    uint64_t a1 = 0, a2 = 0;
    bool skip = false;                                                 /* decompressPE's `continue` (src/repaq.cpp:395,405) */
    for (uint32_t i = 0; i < res.n_chunks; i++) {
        const rpq_chunk_info& c = res.chunks[i];
        const bool last = i + 1 == res.n_chunks;
        const bool f1 = (c.flags & RPQ_NO_LINE_BREAK_AT_END) != 0, f2 = (c.flags & RPQ_NO_LINE_BREAK_AT_END_R2) != 0;
        if (skip) { skip = false; a1 += c.out1_bytes; a2 += c.out2_bytes; continue; }
        if (!pe) w1.write((char*)res.out1 + a1, (f1 && last && c.out1_bytes) ? c.out1_bytes - 1 : c.out1_bytes);
        else {
            bool skip_next = false;
            if (f1 && !last) { w1.write((char*)res.out1 + a1, c.out1_bytes); skip_next = true; }
            else w1.write((char*)res.out1 + a1, (f1 && c.out1_bytes) ? c.out1_bytes - 1 : c.out1_bytes);
            if (!skip_next) {
                if (f2 && !last) { w2->write((char*)res.out2 + a2, c.out2_bytes); skip_next = true; }
                else w2->write((char*)res.out2 + a2, (f2 && c.out2_bytes) ? c.out2_bytes - 1 : c.out2_bytes);
            }
            skip = skip_next;
        }
        a1 += c.out1_bytes; a2 += c.out2_bytes;
    }
    delete w2;
    rpq_destroy(ctx);
}

}  // namespace

/* the dispatcher of src/repaq.cpp:12-34 */
void Repaq::run() {
    /* Repaq::compare / comparePE (src/repaq.cpp:36-233): inside the member, because the report goes through the private
     * reportCompareResult (:235-259) */
    auto compare_gpu = [this]() {
    Options* opt = mOptions;
    std::vector<char> rfq = slurp(opt->rfqCompare), r1 = slurp(opt->in1), r2;
    const bool pe = !opt->in2.empty();
    if (pe) r2 = slurp(opt->in2);
    char err[768]; rpq_header h; size_t used = 0;
    if (rfq.empty()) { memset(&h, 0, sizeof h); h.read_length_bytes = 1; h.flags = RPQ_ENCODE_QUAL_BY_COL; h.n_base_qual = '#'; h.overlap_shift = -24; h.qual_bins = 1; h.qual_buf[0] = 'F'; }
    else if (rpq_header_read((const uint8_t*)rfq.data(), rfq.size(), &h, &used, err, sizeof err)) error_exit(err);
    rpq_ctx* ctx = context(h);
    rpq_compare_in in; memset(&in, 0, sizeof in);
    in.rfq = (const uint8_t*)rfq.data() + used; in.rfq_bytes = rfq.size() - used; in.rfq_mem = RPQ_MEM_HOST; in.rfq_final = 1;
    in.r1 = r1.data(); in.r1_len = r1.size(); in.r2 = pe ? (r2.empty() ? "" : r2.data()) : NULL; in.r2_len = r2.size(); in.fq_mem = RPQ_MEM_HOST; in.fq_final = 1;
    rpq_compare_out res;                                               /* decodeChunk + the four string comparisons, every read */
    if (rpq_compare(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
    const long sr = (long)(pe ? res.rfq_reads / 2 : res.rfq_reads), sf = (long)(pe ? res.fastq_reads / 2 : res.fastq_reads);
    const string unit = pe ? "pair" : "read";
    string msg;
    if (res.verdict >= RPQ_CMP_NAME && res.verdict <= RPQ_CMP_QUALITY) {
        static const char* what[4] = {"name", "sequence", "strand", "quality"};
        msg = string("The RFQ file and FASTQ file have different ") + what[res.verdict - RPQ_CMP_NAME] + " in the " + to_string(sr) + " " + unit + ". " +
              string(res.rfq_field, res.rfq_field_len) + " | " + string(res.fastq_field, res.fastq_field_len);
    } else if (res.verdict == RPQ_CMP_RFQ_MORE)
        msg = "The RFQ file has more reads than the FASTQ file. The RFQ file has >= " + to_string(sr) + " " + unit + "s, while the FASTQ file only has " + to_string(sf) + " " + unit + "s";
    else if (res.verdict == RPQ_CMP_FASTQ_MORE)
        msg = "The FASTQ file has more reads than the RFQ file. The FASTQ file has >= " + to_string(sf) + " " + unit + "s, while the RFQ file only has " + to_string(sr) + " " + unit + "s";
    reportCompareResult(res.verdict == RPQ_CMP_EQUAL, msg, (long)res.fastq_reads, (long)res.fastq_bases, (long)res.rfq_reads, (long)res.rfq_bases);
    rpq_destroy(ctx);
    };
    if (mOptions->mode == REPAQ_COMPRESS) compress_gpu(mOptions);
    else if (mOptions->mode == REPAQ_DECOMPRESS) decompress_gpu(mOptions);
    else if (mOptions->mode == REPAQ_COMPARE) compare_gpu();
    else error_exit("no mode specified, you should specify one of compress/decompress/compare mode");
}
