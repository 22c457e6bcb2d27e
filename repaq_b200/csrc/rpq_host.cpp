/*
 * rpq_host.cpp - the parts of the path that stay host C++ (BASELINE north_star: "host code stays C++ ... RfqHeader"):
 *   RfqCodec::makeHeader + RfqHeader::makeQualityTable  (reference src/rfqcodec.cpp:20-145, src/rfqheader.cpp:130-237)
 *   RfqHeader::write / read                             (src/rfqheader.cpp:19-43, 84-97)
 *   the chunk walk of RfqChunk::read                    (src/rfqchunk.cpp:161-228; arena sizes derived as in :63-109)
 * No CUDA in this file.
 */
#include <string.h>

#include <string>
#include <vector>

#include "rpq_host.h"

namespace rpq {

static void set_err(char* err, size_t cap, const std::string& s) { if (err && cap) { snprintf(err, cap, "%s", s.c_str()); } }

/* ---- FastqMeta::parse, host side (src/fastqmeta.cpp:22-80), same closed form as the device tokeniser */
static int atoi_like(const char* s, int n) {
    int i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    unsigned long long acc = 0; bool sat = false;
    const unsigned long long lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) {
        unsigned d = (unsigned)(s[i] - '0');
        if (!sat) { if (acc > (lim - d) / 10) { sat = true; acc = lim; } else acc = acc * 10 + d; }
    }
    long long v = neg ? (long long)(0ull - acc) : (long long)acc;
    return (int)v;
}

HostMeta host_meta_parse(const char* name, int len) {
    int c[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    int ncol = 0, S = -1;
    for (int i = 0; i < len; i++) {
        if (name[i] == ':') { ncol++; c[ncol] = i; if (ncol == 7) break; }
        else if (name[i] == ' ') { S = i; break; }
    }
    HostMeta m; memset(&m, 0, sizeof m);
    m.name1_len = len; m.name2_off = len;
    auto fld = [&](int a, int b) { return atoi_like(name + a + 1, b - a - 1); };
    int stop = -1;
    if (ncol == 7) { stop = c[7]; m.name1_len = c[3]; m.lane = (uint8_t)fld(c[3], c[4]); m.tile = (uint16_t)fld(c[4], c[5]); m.x = (uint32_t)fld(c[5], c[6]); m.y = (uint32_t)fld(c[6], c[7]); }
    else if (S >= 0 && ncol == 6) { stop = S; m.name1_len = c[3]; m.lane = (uint8_t)fld(c[3], c[4]); m.tile = (uint16_t)fld(c[4], c[5]); m.x = (uint32_t)fld(c[5], c[6]); m.y = (uint32_t)fld(c[6], S); }
    else if (S >= 0 && ncol == 5) { stop = S; m.name1_len = c[3]; m.lane = (uint8_t)fld(c[3], c[4]); m.tile = (uint16_t)fld(c[5], S); }
    else if (S >= 0 && ncol == 4) { stop = S; m.name1_len = c[4]; m.lane = (uint8_t)fld(c[4], S); }
    if (stop > 0) { m.has = 1; m.name2_off = stop; }
    else { m.name1_len = len; m.name2_off = len; m.lane = 0; m.tile = 0; m.x = m.y = 0; }
    return m;
}

/* ---- records of the first chunk, read as FastqReader::getLine reads them (src/fastqreader.cpp:94-156): a line ends at the first
 * '\r' or '\n'; a '\n' right after that byte belongs to the break unless it is the first or the last byte of one of the
 * reader's 1 MiB buffers (the text is taken to start where the file starts) or the last byte of the input; input stops at the
 * first empty line (:180-181) */
struct Rec { const char* name; int name_len; const char* seq; int seq_len; const char* qual; int qual_len; };

struct LineCursor {
    const char* p; uint64_t len, at;
    bool next(const char*& s, int& n) {
        if (at >= len) return false;
        uint64_t e = at;
        while (e < len && p[e] != '\r' && p[e] != '\n') e++;
        s = p + at; n = (int)(e - at);
        e++;
        const uint64_t in_buf = e & ((1ull << 20) - 1);
        if (e + 1 < len && p[e] == '\n' && in_buf != 0 && in_buf != (1ull << 20) - 1) e++;
        at = e;
        return true;
    }
    bool record(Rec& r) {
        const char* st; int sn;
        if (!next(r.name, r.name_len) || !next(r.seq, r.seq_len) || !next(st, sn)) return false;
        if (!r.name_len || !r.seq_len || !sn) return false;
        if (!next(r.qual, r.qual_len) || !r.qual_len) return false;
        return true;
    }
};

int host_make_header(const char* r1, uint64_t l1, const char* r2, uint64_t l2, int interleaved, uint32_t chunk_bases,
                     rpq_header* h, char* err, size_t err_cap) {
    memset(h, 0, sizeof *h);
    h->read_length_bytes = 1; h->n_base_qual = '#'; h->overlap_shift = -24;      /* RfqHeader ctor, src/rfqheader.cpp:7-17 */
    const bool pe = r2 != nullptr || interleaved;
    LineCursor a{r1, l1, 0}, b{r2, l2, 0};
    std::vector<Rec> reads;
    uint64_t total = 0;
    for (;;) {
        Rec x, y;
        if (!a.record(x)) break;
        if (pe) { if (!(r2 ? b.record(y) : a.record(y))) break; reads.push_back(x); reads.push_back(y); total += (uint64_t)x.seq_len + y.seq_len; }
        else { reads.push_back(x); total += x.seq_len; }
        if (total >= chunk_bases) break;
    }
    if (reads.empty()) { set_err(err, err_cap, "the input holds no FASTQ record"); return RPQ_NO_RECORDS; }

    bool has = true; int maxlen = 0;
    bool support = true; int diff_pos = 0; char diff_char = '\0';
    for (size_t i = 0; i < reads.size(); i++) {
        HostMeta m = host_meta_parse(reads[i].name, reads[i].name_len);
        has = has && m.has;
        if (reads[i].seq_len > maxlen) maxlen = reads[i].seq_len;
        if (pe && (i & 1)) {                                     /* src/rfqcodec.cpp:89-114 */
            HostMeta m1 = host_meta_parse(reads[i - 1].name, reads[i - 1].name_len);
            const char* n1 = reads[i - 1].name + m1.name2_off; const int n1l = reads[i - 1].name_len - m1.name2_off;
            const char* n2 = reads[i].name + m.name2_off; const int n2l = reads[i].name_len - m.name2_off;
            if (!has) support = false;
            else if (support) {
                if (i == 1) {
                    if (n1l != n2l) support = false;
                    for (int p = 0; p < n1l; p++) { char c2 = p < n2l ? n2[p] : '\0'; if (n1[p] != c2) { diff_pos = p; diff_char = c2; break; } }
                }
                if (n1l < diff_pos) support = false;
                else {
                    bool same = n1l == n2l;
                    for (int p = 0; same && p < n1l; p++) { char c1 = n1[p]; if (diff_char != '\0' && p == diff_pos) c1 = diff_char; if (c1 != n2[p]) same = false; }
                    if (!same) support = false;
                }
            }
        }
    }
    if (pe && support) { h->support_interleaved = 1; h->name2_diff_pos = (uint8_t)diff_pos; h->name2_diff_char = (uint8_t)diff_char; h->flags |= RPQ_ENCODE_PE_BY_OVERLAP; }

    /* makeQualityTable, src/rfqheader.cpp:130-237 */
    int table[128]; memset(table, 0, sizeof table);
    int ncount = 0; signed char nq = '#';
    for (const Rec& r : reads) {
        for (int i = 0; i < r.seq_len; i++) {
            signed char q = i < r.qual_len ? (signed char)r.qual[i] : 0;
            if (q < 0) { set_err(err, err_cap, "bad quality value: " + std::to_string((int)q)); return RPQ_ERR_QUALITY; }
            table[(int)q]++;
            char base = r.seq[i];
            if (base == 'N') { if (ncount == 0) nq = q; else if (nq != q) { h->flags |= RPQ_ENCODE_N_POS; nq = -1; } ncount++; }
            if (base != 'A' && base != 'T' && base != 'C' && base != 'G' && base != 'N') {
                std::string msg = (base == 'a' || base == 't' || base == 'c') ? "repaq doesn't support FASTQ with lowercase bases (a/t/c/g)"
                                                                               : "repaq only supports FASTQ with uppercase bases (A/T/C/G/N)";
                msg += "\nbut we get:\n" + std::string(r.seq, r.seq_len);
                set_err(err, err_cap, msg); return RPQ_ERR_QUALITY;
            }
            if (q == nq && ncount > 0 && base != 'N') { h->flags |= RPQ_ENCODE_N_POS; nq = -1; }
        }
    }
    if (ncount < 100) { h->flags |= RPQ_ENCODE_N_POS; nq = -1; }
    int bins = 0, max_num = 0, major = 0; bool has_n = false;
    for (int i = 0; i < 128; i++) {
        if (table[i] > 0) { bins++; if (i == (int)nq) has_n = true; }
        if (table[i] > max_num) { max_num = table[i]; major = i; }
    }
    if (bins == 0) { set_err(err, err_cap, "bad quality string, is this a valid FASTQ file?"); return RPQ_ERR_QUALITY; }
    if (bins >= 64) h->flags |= RPQ_DONT_ENCODE_QUAL;
    if (!has_n) bins += 1;
    h->qual_bins = (uint8_t)bins;
    h->qual_buf[0] = (uint8_t)major;
    int cur = 1;
    for (int i = 0; i < 128; i++) { if (i == major) continue; if (table[i] > 0 && cur < 128) h->qual_buf[cur++] = (uint8_t)i; }
    if (!has_n && bins - 1 < 128) h->qual_buf[bins - 1] = (uint8_t)nq;
    if (bins <= 64) h->flags |= RPQ_ENCODE_QUAL_BY_COL;
    h->n_base_qual = nq;

    if (has) h->flags |= RPQ_HAS_LANE | RPQ_HAS_TILE | RPQ_HAS_X | RPQ_HAS_Y | RPQ_HAS_NAME2;
    if (pe) h->flags |= RPQ_PAIRED_END;
    /* Q1 (src/rfqcodec.cpp:48-53): 4 is overwritten by the second, non-else `if` */
    if (maxlen > 65535) h->read_length_bytes = 4;
    if (maxlen > 255) h->read_length_bytes = 2; else h->read_length_bytes = 1;
    return RPQ_OK;
}

size_t host_header_write(const rpq_header* h, uint8_t* out, size_t cap) {
    const size_t n = 17u + h->qual_bins;
    if (cap < n) return 0;
    memcpy(out, "RFQ", 3); memcpy(out + 3, "0.5.1", 5);
    out[8] = 2; out[9] = h->read_length_bytes;
    out[10] = (uint8_t)h->flags; out[11] = (uint8_t)(h->flags >> 8);
    out[12] = h->name2_diff_pos; out[13] = h->name2_diff_char; out[14] = (uint8_t)h->n_base_qual; out[15] = (uint8_t)h->overlap_shift;
    out[16] = h->qual_bins;
    memcpy(out + 17, h->qual_buf, h->qual_bins);
    return n;
}

int host_header_read(const uint8_t* in, size_t len, rpq_header* h, size_t* consumed, char* err, size_t err_cap) {
    memset(h, 0, sizeof *h);
    if (len < 17) { set_err(err, err_cap, "Not a valid repaq file!"); return RPQ_ERR_HEADER; }
    if (in[8] != 2) {
        set_err(err, err_cap, "The data is encoded by different version of repaq, please try repaq v" + std::string((const char*)in + 3, 5) +
                              ". \nSee: https://github.com/OpenGene/repaq/releases");
        return RPQ_ERR_HEADER;
    }
    h->read_length_bytes = in[9];
    h->flags = (uint16_t)(in[10] | (in[11] << 8));
    h->name2_diff_pos = in[12]; h->name2_diff_char = in[13]; h->n_base_qual = (int8_t)in[14]; h->overlap_shift = (int8_t)in[15];
    h->qual_bins = in[16];
    if (h->qual_bins > 128 || len < 17u + h->qual_bins) { set_err(err, err_cap, "Not a valid repaq file!"); return RPQ_ERR_HEADER; }
    memcpy(h->qual_buf, in + 17, h->qual_bins);
    if (in[0] != 'R' || in[1] != 'F' || in[2] != 'Q') { set_err(err, err_cap, "Not a valid repaq file!"); return RPQ_ERR_HEADER; }
    h->support_interleaved = (h->flags & RPQ_ENCODE_PE_BY_OVERLAP) ? 1 : 0;      /* not stored; decode keys off chunk flag + this bit (Q20) */
    if (consumed) *consumed = 17u + h->qual_bins;
    return RPQ_OK;
}

/* normalQualBins / normalQualBuf (src/rfqheader.cpp:308-328), with the reference's uint8-vs-char comparisons */
int host_normal_bins(const rpq_header* h, uint8_t* out) {
    const int major = (int)(signed char)h->qual_buf[0], nq = (int)h->n_base_qual;
    const int bins = (major == nq) ? h->qual_bins : h->qual_bins - 1;
    int cnt = 0;
    for (int i = 0; i < h->qual_bins; i++) {
        const int e = h->qual_buf[i];
        if (e != major || e == nq) { if (cnt < 129) out[cnt] = h->qual_buf[i]; cnt++; if (cnt > bins) break; }
    }
    return bins < 0 ? 0 : bins;
}

/* ---- chunk walk for decode: sizes of the name/strand arenas are derived from the length columns (src/rfqchunk.cpp:63-109) */
static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int host_walk_chunk(const rpq_header* h, const uint8_t* in, uint64_t len, HostChunk* c) {
    memset(c, 0, sizeof *c);
    uint64_t at = 0;
    const uint64_t head = 18u + ((h->flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    if (len < head) return 1;
    c->msize = rd32(in); c->reads = rd32(in + 4); c->flags = (uint16_t)(in[8] | (in[9] << 8));
    c->seq_size = rd32(in + 10); c->qual_size = rd32(in + 14);
    if (h->flags & RPQ_ENCODE_N_POS) c->npos_size = rd32(in + 18);
    at = head;
    if (c->reads == 0) return 2;
    const uint32_t n = c->reads, fl = c->flags;
    const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
    const uint32_t xy = il ? n / 2 : n;
#define TAKE(field, size) do { c->field = (uint32_t)at; at += (uint64_t)(size); if (at > len) return 1; } while (0)
    c->readlen_size = (uint32_t)h->read_length_bytes * ((fl & RPQ_READ_LEN_SAME) ? 1u : n);
    TAKE(off_readlen, c->readlen_size);
    c->n1len_size = (fl & RPQ_NAME1_LEN_SAME) ? 1u : n; TAKE(off_n1len, c->n1len_size);
    if (h->flags & RPQ_HAS_NAME2) { c->n2len_size = (fl & RPQ_NAME2_LEN_SAME) ? 1u : n; TAKE(off_n2len, c->n2len_size); }
    c->slen_size = (fl & RPQ_STRAND_LEN_SAME) ? 1u : n; TAKE(off_slen, c->slen_size);
    if (h->flags & RPQ_HAS_LANE) { c->lane_size = (fl & RPQ_LANE_SAME) ? 1u : xy; TAKE(off_lane, c->lane_size); }
    if (h->flags & RPQ_HAS_TILE) { c->tile_size = 2u * ((fl & RPQ_TILE_SAME) ? 1u : xy); TAKE(off_tile, c->tile_size); }
    if (h->flags & RPQ_HAS_X) { if (at + 4 > len) return 1; c->x_size = rd32(in + at); at += 4; TAKE(off_x, c->x_size); }
    if (h->flags & RPQ_HAS_Y) { if (at + 4 > len) return 1; c->y_size = rd32(in + at); at += 4; TAKE(off_y, c->y_size); }
    auto arena = [&](uint32_t off, uint32_t cnt, bool len_same, bool all_same) -> uint64_t {
        uint64_t t = 0;
        for (uint32_t i = 0; i < cnt; i++) t += in[off + i];
        if (len_same && !all_same) t *= n;
        return t;
    };
    /* an arena of 4 GiB or more (a crafted length byte times a crafted read count) does not fit the chunk whatever follows: it must
     * not wrap to a small 32-bit size that passes the checks */
    const uint64_t a1 = arena(c->off_n1len, c->n1len_size, fl & RPQ_NAME1_LEN_SAME, fl & RPQ_NAME1_SAME);
    if (a1 > len) return 1;
    c->n1_size = (uint32_t)a1; TAKE(off_n1, c->n1_size);
    if (h->flags & RPQ_HAS_NAME2) {
        const uint64_t a2 = arena(c->off_n2len, c->n2len_size, fl & RPQ_NAME2_LEN_SAME, fl & RPQ_NAME2_SAME);
        if (a2 > len) return 1;
        c->n2_size = (uint32_t)a2; TAKE(off_n2, c->n2_size);
    }
    const uint64_t a3 = arena(c->off_slen, c->slen_size, fl & RPQ_STRAND_LEN_SAME, fl & RPQ_STRAND_SAME);
    if (a3 > len) return 1;
    c->strand_size = (uint32_t)a3; TAKE(off_strand, c->strand_size);
    TAKE(off_seq, c->seq_size);
    TAKE(off_qual, c->qual_size);
    if (il && (h->flags & RPQ_ENCODE_PE_BY_OVERLAP)) { c->ov_size = n / 2; TAKE(off_ov, c->ov_size); }
    if (h->flags & RPQ_ENCODE_N_POS) TAKE(off_npos, c->npos_size);
#undef TAKE
    if (at > 0xFFFFFFFFull) return 1;
    c->bytes = (uint32_t)at;
    return 0;
}

}  // namespace rpq
