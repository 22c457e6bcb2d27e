/*
 * rfq_oracle.c - TEST INFRASTRUCTURE ONLY (see rfq_oracle.h for the rules and the parity status).
 *
 * Sequential plain-C restatement of the reference's FASTQ <-> .rfq path.  Citations are to the
 * reference checkout (OpenGene/repaq v0.5.1): src/<file>:<lines>.
 */
#include "rfq_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[512];
const char* orc_last_error(void) { return g_err; }
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return -1; }
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------ byte sink ---- */
typedef struct { uint8_t* p; size_t n, cap; } sink;
static void sk_reserve(sink* s, size_t extra) {
    if (s->n + extra <= s->cap) return;
    size_t c = s->cap ? s->cap : 4096;
    while (c < s->n + extra) c *= 2;
    s->p = (uint8_t*)realloc(s->p, c); s->cap = c;
}
static void sk_put(sink* s, const void* d, size_t n) { sk_reserve(s, n); if (n) memcpy(s->p + s->n, d, n); s->n += n; }
static void sk_u8(sink* s, uint8_t v) { sk_put(s, &v, 1); }
static void sk_u16(sink* s, uint16_t v) { uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; sk_put(s, b, 2); }
static void sk_u32(sink* s, uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; sk_put(s, b, 4); }

/* ------------------------------------------------------------------ FastqMeta ---- */
/* glibc atoi == (int)strtol(s, NULL, 10): skips isspace, optional sign, digits, saturates at LONG_MIN/MAX. */
static int atoi_like(const char* s, uint32_t n) {
    uint32_t i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    int neg = 0;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    unsigned long long acc = 0; int sat = 0;
    const unsigned long long lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) {
        unsigned d = (unsigned)(s[i] - '0');
        if (!sat) {
            if (acc > (lim - d) / 10) { sat = 1; acc = lim; }
            else acc = acc * 10 + d;
        }
    }
    long long v = neg ? (long long)(0ull - acc) : (long long)acc;
    return (int)v;
}

/* src/fastqmeta.cpp:22-80 - positional state machine over ':' counts 4..7 and the first space. */
void orc_meta_parse(const char* str, uint32_t len, orc_meta* out) {
    int colon = 0;
    int last_colon = 0, start_at = 0, end_at = 0;
    uint8_t lane = 0; uint16_t tile = 0; uint32_t x = 0, y = 0;
    for (uint32_t i = 0; i < len; i++) {
        char c = str[i];
        if (c == ':') colon++;
        if ((c == ':' || c == ' ') && colon >= 4 && colon <= 7) {
            /* item = str.substr(last_colon+1, i-last_colon-1); substr clamps, never throws here since last_colon+1 <= len */
            uint32_t from = (uint32_t)last_colon + 1;
            uint32_t cnt = (i >= from) ? i - from : 0;   /* i - last_colon - 1, never negative except i==0,last_colon==0 */
            int val = (from <= len) ? atoi_like(str + from, cnt) : 0;
            switch (colon) {
                case 4: lane = (uint8_t)val; start_at = last_colon + 1; break;
                case 5: tile = (uint16_t)val; break;
                case 6: if (c == ':') x = (uint32_t)val; break;
                case 7: y = (uint32_t)val; break;
            }
            if (c == ' ' && colon == 6) y = (uint32_t)val;
        }
        if (c == ':') last_colon = (int)i;
        if (c == ' ' || (c == ':' && colon == 7)) { end_at = (int)i; break; }
    }
    memset(out, 0, sizeof *out);
    if (start_at > 0 && end_at > 0) {
        out->lane = lane; out->tile = tile; out->x = x; out->y = y;
        out->has_lane_tile_xy = 1;
        out->name1_len = (uint32_t)(start_at - 1);
        out->name2_off = (uint32_t)end_at;
        out->name2_len = len - (uint32_t)end_at;
    } else {
        out->name1_len = len;
        out->name2_off = len;
        out->name2_len = 0;
    }
}

/* ------------------------------------------------------------------ FastqReader ---- */
#define FQ_BUF_SIZE (1 << 20)

typedef struct { char* p; size_t n, cap; } strbuf;
static void sb_append(strbuf* s, const char* d, size_t n) {
    if (s->n + n + 1 > s->cap) { size_t c = s->cap ? s->cap : 256; while (c < s->n + n + 1) c *= 2; s->p = (char*)realloc(s->p, c); s->cap = c; }
    if (n) memcpy(s->p + s->n, d, n);
    s->n += n;
}

struct orc_reader {
    const char* text; size_t size;
    size_t file_pos;          /* bytes handed out by fread so far */
    const char* buf; long buf_len, used;
    int hit_eof;              /* feof(): set once an fread came back short */
    int no_break;
    strbuf line[4];
};

/* src/fastqreader.cpp:31-46 */
static void rd_fill(orc_reader* r) {
    size_t left = r->size - r->file_pos;
    size_t take = left < FQ_BUF_SIZE ? left : FQ_BUF_SIZE;
    r->buf = r->text + r->file_pos; r->buf_len = (long)take; r->file_pos += take;
    if (take < FQ_BUF_SIZE) r->hit_eof = 1;
    r->used = 0;
    if (r->buf_len < FQ_BUF_SIZE) {
        /* reference reads mBuf[mBufDataLen-1]; with mBufDataLen == 0 that is the byte before the heap block
         * (undefined; in practice not '\n').  Restated as "flag set". */
        if (r->buf_len == 0 || r->buf[r->buf_len - 1] != '\n') r->no_break = 1;
    }
}

orc_reader* orc_reader_open(const char* text, size_t size) {
    orc_reader* r = (orc_reader*)calloc(1, sizeof *r);
    r->text = text; r->size = size;
    rd_fill(r);                                   /* init() -> readToBuf(): src/fastqreader.cpp:48-66 */
    return r;
}
void orc_reader_close(orc_reader* r) { if (!r) return; for (int i = 0; i < 4; i++) free(r->line[i].p); free(r); }
int orc_reader_no_line_break_at_end(const orc_reader* r) { return r->no_break; }

/* src/fastqreader.cpp:94-156 */
static void rd_getline(orc_reader* r, strbuf* out) {
    out->n = 0;
    long start = r->used, end = start;
    while (end < r->buf_len && r->buf[end] != '\r' && r->buf[end] != '\n') end++;
    if (end < r->buf_len || r->buf_len < FQ_BUF_SIZE) {
        sb_append(out, r->buf + start, (size_t)(end > start ? end - start : 0));
        end++;
        if (end < r->buf_len - 1 && r->buf[end] == '\n') end++;
        r->used = end;
        return;
    }
    sb_append(out, r->buf + start, (size_t)(r->buf_len - start));
    for (;;) {
        rd_fill(r);
        end = 0;
        while (end < r->buf_len && r->buf[end] != '\r' && r->buf[end] != '\n') end++;
        if (end < r->buf_len || r->buf_len < FQ_BUF_SIZE) {
            sb_append(out, r->buf, (size_t)end);
            end++;
            if (end < r->buf_len - 1 && r->buf[end] == '\n') end++;
            r->used = end;
            return;
        }
        sb_append(out, r->buf, (size_t)r->buf_len);
    }
}

/* src/fastqreader.cpp:166-196 */
int orc_reader_next(orc_reader* r, orc_read* rd) {
    if (r->used >= r->buf_len && r->hit_eof) return 0;
    rd_getline(r, &r->line[0]);
    rd_getline(r, &r->line[1]);
    rd_getline(r, &r->line[2]);
    if (r->line[0].n == 0 || r->line[1].n == 0 || r->line[2].n == 0) return 0;
    rd_getline(r, &r->line[3]);
    if (r->line[3].n == 0) return 0;
    rd->name = r->line[0].p; rd->name_len = (uint32_t)r->line[0].n;
    rd->seq = r->line[1].p; rd->seq_len = (uint32_t)r->line[1].n;
    rd->strand = r->line[2].p; rd->strand_len = (uint32_t)r->line[2].n;
    rd->qual = r->line[3].p; rd->qual_len = (uint32_t)r->line[3].n;
    return 1;
}

/* ------------------------------------------------------------------ RfqHeader ---- */
static int hdr_major(const orc_header* h) { return (int)(signed char)h->qual_buf[0]; }   /* mBit2QualTable[0], a char */

/* src/rfqheader.cpp:308-328.  Comparisons mix uint8 entries with char major / nBaseQual exactly as there. */
static int hdr_normal_bins(const orc_header* h, uint8_t* out) {
    int major = hdr_major(h), nq = (int)h->n_base_qual;
    int bins = (major == nq) ? h->qual_bins : h->qual_bins - 1;
    int cnt = 0;
    for (int i = 0; i < h->qual_bins; i++) {
        int e = (int)h->qual_buf[i];
        if (e != major || e == nq) {
            if (cnt < 256) out[cnt] = h->qual_buf[i];
            cnt++;
            if (cnt > bins) break;
        }
    }
    return bins;
}

/* src/rfqheader.cpp:130-237 */
static int make_quality_table(orc_header* h, const orc_read* reads, size_t n) {
    int table[128]; memset(table, 0, sizeof table);
    int ncount = 0;
    for (size_t r = 0; r < n; r++) {
        const orc_read* rd = &reads[r];
        for (uint32_t i = 0; i < rd->seq_len; i++) {
            signed char q = (i < rd->qual_len) ? (signed char)rd->qual[i] : 0;
            if (q < 0) return fail("bad quality value");
            table[(int)q]++;
            char base = rd->seq[i];
            if (base == 'N') {
                if (ncount == 0) h->n_base_qual = q;
                else if (h->n_base_qual != q) { h->flags |= ORC_ENCODE_N_POS; h->n_base_qual = -1; }
                ncount++;
            }
            if (base != 'A' && base != 'T' && base != 'C' && base != 'G' && base != 'N') {
                if (base == 'a' || base == 't' || base == 'c')
                    return fail("repaq doesn't support FASTQ with lowercase bases (a/t/c/g)");
                return fail("repaq only supports FASTQ with uppercase bases (A/T/C/G/N)");
            }
            if (q == h->n_base_qual && ncount > 0 && base != 'N') { h->flags |= ORC_ENCODE_N_POS; h->n_base_qual = -1; }
        }
    }
    if (ncount < 100) { h->flags |= ORC_ENCODE_N_POS; h->n_base_qual = -1; }

    int bins = 0, max_num = 0, major = 0, has_n = 0;
    for (int i = 0; i < 128; i++) {
        if (table[i] > 0) { bins++; if (i == (int)h->n_base_qual) has_n = 1; }
        if (table[i] > max_num) { max_num = table[i]; major = i; }
    }
    if (bins == 0) return fail("bad quality string, is this a valid FASTQ file?");
    if (bins >= 64) h->flags |= ORC_DONT_ENCODE_QUAL;
    if (!has_n) bins += 1;
    h->qual_bins = (uint8_t)bins;
    h->qual_buf[0] = (uint8_t)major;
    int cur = 1;
    for (int i = 0; i < 128; i++) {
        if (i == major) continue;
        if (table[i] > 0) h->qual_buf[cur++] = (uint8_t)i;
    }
    if (!has_n) h->qual_buf[bins - 1] = (uint8_t)h->n_base_qual;
    if (bins <= 64) h->flags |= ORC_ENCODE_QUAL_BY_COL;
    return 0;
}

/* src/rfqheader.cpp:7-17 (ctor), src/rfqcodec.cpp:20-57 (SE), :59-145 (PE) */
int orc_make_header(const orc_read* reads, size_t n, int is_pe, orc_header* h) {
    memset(h, 0, sizeof *h);
    h->read_length_bytes = 1; h->n_base_qual = '#'; h->overlap_shift = -24;
    if (n == 0) return fail("no reads");
    int has = 1; uint32_t maxlen = 0;
    int support = 1, diff_pos = 0; char diff_char = '\0';
    for (size_t i = 0; i < n; i++) {
        orc_meta m; orc_meta_parse(reads[i].name, reads[i].name_len, &m);
        has &= m.has_lane_tile_xy;
        if (reads[i].seq_len > maxlen) maxlen = reads[i].seq_len;
        if (is_pe && (i & 1)) {
            orc_meta m1; orc_meta_parse(reads[i - 1].name, reads[i - 1].name_len, &m1);
            const char* n1 = reads[i - 1].name + m1.name2_off; const char* n2 = reads[i].name + m.name2_off;
            if (!has) support = 0;
            else if (support) {
                if (i == 1) {
                    if (m1.name2_len != m.name2_len) support = 0;
                    for (uint32_t p = 0; p < m1.name2_len; p++) {
                        /* reference indexes meta2.namePart2[p] unchecked; beyond its end that reads the NUL */
                        char c2 = p < m.name2_len ? n2[p] : '\0';
                        if (n1[p] != c2) { diff_pos = (int)p; diff_char = c2; break; }
                    }
                }
                if ((int)m1.name2_len < diff_pos) support = 0;
                else {
                    /* one, and just one, differing character */
                    int same = m1.name2_len == m.name2_len;
                    for (uint32_t p = 0; same && p < m1.name2_len; p++) {
                        char c1 = n1[p];
                        if (diff_char != '\0' && (int)p == diff_pos) c1 = diff_char;
                        if (c1 != n2[p]) same = 0;
                    }
                    if (!same) support = 0;
                }
            }
        }
    }
    if (is_pe && support) {
        h->support_interleaved = 1;
        h->name2_diff_pos = (uint8_t)diff_pos; h->name2_diff_char = diff_char;
        h->flags |= ORC_ENCODE_PE_BY_OVERLAP;
    }
    if (make_quality_table(h, reads, n)) return -1;
    if (has) h->flags |= ORC_HAS_LANE | ORC_HAS_TILE | ORC_HAS_X | ORC_HAS_Y | ORC_HAS_NAME2;
    if (is_pe) h->flags |= ORC_PAIRED_END;
    /* Q1: the second test is `if`, not `else if`, so 4 never survives (src/rfqcodec.cpp:48-53) */
    if (maxlen > 65535) h->read_length_bytes = 4;
    if (maxlen > 255) h->read_length_bytes = 2; else h->read_length_bytes = 1;
    return 0;
}

/* src/rfqheader.cpp:84-97 */
size_t orc_header_write(const orc_header* h, uint8_t* out) {
    size_t n = 0;
    memcpy(out, "RFQ", 3); n += 3;
    memcpy(out + n, "0.5.1", 5); n += 5;
    out[n++] = 2;
    out[n++] = h->read_length_bytes;
    out[n++] = (uint8_t)h->flags; out[n++] = (uint8_t)(h->flags >> 8);
    out[n++] = h->name2_diff_pos;
    out[n++] = (uint8_t)h->name2_diff_char;
    out[n++] = (uint8_t)h->n_base_qual;
    out[n++] = (uint8_t)h->overlap_shift;
    out[n++] = h->qual_bins;
    memcpy(out + n, h->qual_buf, h->qual_bins); n += h->qual_bins;
    return n;
}

/* src/rfqheader.cpp:19-43 */
size_t orc_header_read(const uint8_t* in, size_t len, orc_header* h) {
    memset(h, 0, sizeof *h);
    if (len < 17) { fail("Not a valid repaq file!"); return 0; }
    if (in[8] != 2) { fail("The data is encoded by different version of repaq"); return 0; }
    h->read_length_bytes = in[9];
    h->flags = (uint16_t)(in[10] | (in[11] << 8));
    h->name2_diff_pos = in[12];
    h->name2_diff_char = (char)in[13];
    h->n_base_qual = (signed char)in[14];
    h->overlap_shift = (signed char)in[15];
    h->qual_bins = in[16];
    if (len < 17u + h->qual_bins) { fail("truncated header"); return 0; }
    memcpy(h->qual_buf, in + 17, h->qual_bins);
    if (in[0] != 'R' || in[1] != 'F' || in[2] != 'Q') { fail("Not a valid repaq file!"); return 0; }
    /* mSupportInterleaved is not stored; decode keys off chunk flag + OVERLAP bit (Q20) */
    h->support_interleaved = (h->flags & ORC_ENCODE_PE_BY_OVERLAP) ? 1 : 0;
    return 17u + h->qual_bins;
}

/* ------------------------------------------------------------------ sub-coders ---- */
static char complement(char b) {      /* src/read.cpp:92-113 */
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 'N';
    }
}

/* src/rfqcodec.cpp:1391-1438: smallest forward overlap, else smallest backward overlap, else 0 */
static int find_overlap(const char* r1, int len1, const char* r2, int len2) {
    int minlen = len1 < len2 ? len1 : len2;
    for (int o = 12; o <= minlen; o++)
        if (memcmp(r1 + len1 - o, r2, (size_t)o) == 0) return o;
    for (int o = 12; o <= minlen; o++)
        if (memcmp(r2 + len2 - o, r1, (size_t)o) == 0) return -o;
    return 0;
}

/* src/rfqcodec.cpp:625-710.  Positions of value q in data[0..n); sets mask[i] on every hit. */
static uint32_t encode_positions(const uint8_t* data, uint8_t q, uint8_t* out, uint32_t n, uint8_t* mask) {
    uint32_t w = 0; long last = -1; long cur = 0;
    while (cur < (long)n) {
        while (data[cur] != q) { cur++; if (cur >= (long)n) return w; }
        if (mask) mask[cur] = 1;
        if (cur - last == 1 && cur > 1) {
            uint32_t run = 1;
            while (cur + run != n && run < 32 && data[cur + run] == q) run++;
            if (mask) memset(mask + cur, 1, run);
            out[w++] = (uint8_t)(0xC0 | (run - 1));
            cur += run; last = cur - 1;
            continue;
        }
        uint32_t d = (uint32_t)(cur - last) - 1;            /* distance - 1 */
        if (d < 128) out[w++] = (uint8_t)d;
        else if (d < (1u << 14)) { out[w++] = (uint8_t)(0x80 | (d >> 8)); out[w++] = (uint8_t)d; }
        else { out[w++] = (uint8_t)(0xE0 | (d >> 24)); out[w++] = (uint8_t)(d >> 16); out[w++] = (uint8_t)(d >> 8); out[w++] = (uint8_t)d; }
        last = cur; cur++;
    }
    return w;
}

/* src/rfqcodec.cpp:957-1007 */
static void decode_positions(const uint8_t* buf, uint32_t len, char q, char* dst, size_t dst_len) {
    uint32_t c = 0; long last = -1;
    while (c < len) {
        uint8_t b0 = buf[c];
        if (!(b0 & 0x80)) { last += b0 + 1; if ((size_t)last < dst_len) dst[last] = q; c += 1; }
        else if (!(b0 & 0x40)) { long d = (((long)(b0 & 0x3F)) << 8 | buf[c + 1]) + 1; last += d; if ((size_t)last < dst_len) dst[last] = q; c += 2; }
        else if (!(b0 & 0x20)) { int run = (b0 & 0x1F) + 1; for (int i = 1; i <= run; i++) if ((size_t)(last + i) < dst_len) dst[last + i] = q; last += run; c += 1; }
        else { long d = ((((long)(b0 & 0x1F)) << 24) | ((long)buf[c + 1] << 16) | ((long)buf[c + 2] << 8) | buf[c + 3]) + 1; last += d; if ((size_t)last < dst_len) dst[last] = q; c += 4; }
    }
}

/* src/rfqcodec.cpp:1262-1330 */
static int encode_coords(const uint32_t* data, uint32_t num, sink* s) {
    uint32_t last = 1000; uint8_t repeat = 0;
    for (uint32_t i = 0; i < num; i++) {
        uint32_t v = data[i];
        if (repeat > 0 && (v != last || repeat == 32)) { sk_u8(s, (uint8_t)(0xC0 | (repeat - 1))); repeat = 0; }
        if (v == last) { repeat++; continue; }
        int diff = (int)(v - last);
        last = v;
        if (diff > 0 && diff <= 64) { sk_u8(s, (uint8_t)(0x80 | (diff - 1))); continue; }
        if (v <= 32767) { sk_u8(s, (uint8_t)(v >> 8)); sk_u8(s, (uint8_t)v); }
        else if (v < (1u << 21)) { sk_u8(s, (uint8_t)(0xE0 | (v >> 16))); sk_u8(s, (uint8_t)(v >> 8)); sk_u8(s, (uint8_t)v); }
        else return fail("The X/Y coordinate cannot be larger than 2M");
    }
    if (repeat > 0) sk_u8(s, (uint8_t)(0xC0 | (repeat - 1)));
    return 0;
}

/* src/rfqcodec.cpp:1332-1389 */
static void decode_coords(const uint8_t* buf, uint32_t len, uint32_t* data, uint32_t num) {
    uint32_t last = 1000, c = 0, d = 0;
    while (c < len) {
        uint32_t b0 = buf[c++];
        if (!(b0 & 0x80)) { uint32_t v = (b0 << 8) | (c < len ? buf[c] : 0); c++; if (d < num) data[d] = v; d++; last = v; }
        else if (!(b0 & 0x40)) { uint32_t v = last + (b0 & 0x3F) + 1; if (d < num) data[d] = v; d++; last = v; }
        else if (!(b0 & 0x20)) { uint32_t rep = (b0 & 0x1F) + 1; for (uint32_t i = 0; i < rep; i++) { if (d < num) data[d] = last; d++; } }
        else { uint32_t v = (b0 & 0x1F) << 16; v |= (uint32_t)(c < len ? buf[c] : 0) << 8; c++; v |= (c < len ? buf[c] : 0); c++; if (d < num) data[d] = v; d++; last = v; }
    }
}

/* ------------------------------------------------------------------ encodeChunk ---- */
static int bytes_eq(const char* a, uint32_t la, const char* b, uint32_t lb) { return la == lb && (la == 0 || memcmp(a, b, la) == 0); }

/* src/rfqcodec.cpp:163-586 + RfqChunk::calcTotalBufSize (src/rfqchunk.cpp:141-159) + RfqChunk::write (:230-312) */
int orc_encode_chunk(const orc_header* h, const orc_read* reads, size_t n, int is_pe, uint16_t extra_flags,
                     uint8_t** out, size_t* out_len) {
    *out = NULL; *out_len = 0;
    if (n == 0) return 0;
    const uint32_t s = (uint32_t)n;
    orc_meta* meta = (orc_meta*)malloc(sizeof(orc_meta) * n);
    for (size_t i = 0; i < n; i++) orc_meta_parse(reads[i].name, reads[i].name_len, &meta[i]);

    /* pass 1: :220-276 */
    int read_len_same = 1, n1len_same = 1, n2len_same = 1, slen_same = 1, strand_same = 1, lane_same = 1, tile_same = 1, n1_same = 1, n2_same = 1;
    const orc_meta* m0 = &meta[0]; const orc_read* r0 = &reads[0];
    const char* name20 = r0->name + m0->name2_off;
    uint32_t tot_len = 0, tot_n1 = 0, tot_n2 = 0, tot_strand = 0;
    int can_il = is_pe && h->support_interleaved;
    const int encode_overlap = can_il && (h->flags & ORC_ENCODE_PE_BY_OVERLAP);    /* decided before the loop (:213) */
    const char* last_n2 = NULL; uint32_t last_n2_len = 0;
    uint32_t last_x = 0, last_y = 0; uint16_t last_tile = 0; uint8_t last_lane = 0;
    for (uint32_t i = 0; i < s; i++) {
        const orc_read* r = &reads[i]; const orc_meta* m = &meta[i];
        const char* n2 = r->name + m->name2_off;
        read_len_same &= r0->seq_len == r->seq_len;
        n1len_same &= m0->name1_len == m->name1_len;
        n2len_same &= m0->name2_len == m->name2_len;
        slen_same &= r0->strand_len == r->strand_len;
        strand_same &= bytes_eq(r0->strand, r0->strand_len, r->strand, r->strand_len);
        lane_same &= m0->lane == m->lane;
        tile_same &= m0->tile == m->tile;
        n1_same &= bytes_eq(r0->name, m0->name1_len, r->name, m->name1_len);
        int eq0 = bytes_eq(name20, m0->name2_len, n2, m->name2_len);
        if (!can_il) n2_same &= eq0;
        else if (i & 1) {
            /* R1's name2 with the header's diff char substituted must equal R2's name2 (:237-245) */
            int same = last_n2_len == m->name2_len;
            for (uint32_t p = 0; same && p < last_n2_len; p++) {
                char c = last_n2[p];
                if (h->name2_diff_char != '\0' && p == h->name2_diff_pos) c = h->name2_diff_char;
                if (c != n2[p]) same = 0;
            }
            if (!same) { can_il = 0; n2_same &= eq0; }
        } else { last_n2 = n2; last_n2_len = m->name2_len; n2_same &= eq0; }
        if (can_il) {
            if (i & 1) { can_il &= last_lane == m->lane; can_il &= last_tile == m->tile; can_il &= last_x == m->x; can_il &= last_y == m->y; }
            else { last_lane = m->lane; last_tile = m->tile; last_x = m->x; last_y = m->y; }
        }
        tot_len += r->seq_len; tot_n1 += m->name1_len; tot_n2 += m->name2_len; tot_strand += r->strand_len;
    }
    const uint32_t xy_num = can_il ? s / 2 : s;
    uint32_t* xs = (uint32_t*)malloc(4 * (size_t)s); uint32_t* ys = (uint32_t*)malloc(4 * (size_t)s);
    uint8_t* lanes = (uint8_t*)malloc(s); uint16_t* tiles = (uint16_t*)malloc(2 * (size_t)s);
    for (uint32_t p = 0; p < xy_num; p++) {               /* :279-287 */
        uint32_t src = can_il ? p * 2 : p;
        lanes[p] = meta[src].lane; tiles[p] = meta[src].tile; xs[p] = meta[src].x; ys[p] = meta[src].y;
    }

    /* pass 2: :332-407 (R2 reverse-complemented in place there; here a scratch copy) */
    char* seq_cat = (char*)malloc(tot_len + 1); uint8_t* qual_cat = (uint8_t*)malloc(tot_len + 1);
    signed char* overlaps = (signed char*)calloc(s / 2 + 1, 1);
    uint32_t seq_copied = 0, qual_copied = 0;
    uint32_t maxlen = 0; for (uint32_t i = 0; i < s; i++) if (reads[i].seq_len > maxlen) maxlen = reads[i].seq_len;
    char* rc = (char*)malloc(maxlen + 1);
    for (uint32_t i = 0; i < s; i++) {
        const orc_read* r = &reads[i]; const uint32_t rlen = r->seq_len;
        const char* sq = r->seq; const char* ql = r->qual;
        int ov = 0; int reversed = 0;
        if (can_il && (i & 1)) {
            for (uint32_t k = 0; k < rlen; k++) rc[k] = complement(r->seq[rlen - 1 - k]);
            sq = rc; reversed = 1;
            if (encode_overlap) {
                ov = find_overlap(reads[i - 1].seq, (int)reads[i - 1].seq_len, rc, (int)rlen);
                if (ov + h->overlap_shift > 127) ov = 0;
                if (ov + h->overlap_shift < -127) ov = 0;
                overlaps[i / 2] = (signed char)(ov + h->overlap_shift);
            }
        }
        if (ov == 0) { memcpy(seq_cat + seq_copied, sq, rlen); seq_copied += rlen; }
        else if (ov > 0) { memcpy(seq_cat + seq_copied, sq + ov, rlen - (uint32_t)ov); seq_copied += rlen - (uint32_t)ov; }
        else { memcpy(seq_cat + seq_copied, sq, rlen - (uint32_t)(-ov)); seq_copied += rlen - (uint32_t)(-ov); }
        if (reversed) for (uint32_t k = 0; k < rlen; k++) qual_cat[qual_copied + k] = (uint8_t)ql[rlen - 1 - k];
        else memcpy(qual_cat + qual_copied, ql, rlen);
        qual_copied += rlen;
    }
    free(rc);

    /* 2-bit pack: :588-604, continuous across reads (Q14) */
    uint32_t seq_bytes = (seq_copied + 3) / 4;
    uint8_t* seq_enc = (uint8_t*)calloc(seq_bytes + 1, 1);
    for (uint32_t i = 0; i < seq_copied; i++) {
        uint8_t v = 0;
        switch (seq_cat[i]) { case 'G': v = 0; break; case 'A': v = 1; break; case 'T': v = 2; break; case 'C': v = 3; break; default: break; }
        seq_enc[i >> 2] |= (uint8_t)(v << ((i & 3) * 2));
    }

    /* quality: :611-621, :712-765 */
    sink qual = {0, 0, 0};
    if (h->flags & ORC_DONT_ENCODE_QUAL) sk_put(&qual, qual_cat, qual_copied);
    else if (h->flags & ORC_ENCODE_QUAL_BY_COL) {
        uint8_t nb_buf[256]; int nb = hdr_normal_bins(h, nb_buf);
        uint8_t* mask = (uint8_t*)calloc(qual_copied + 1, 1);
        uint8_t* scratch = (uint8_t*)malloc((size_t)qual_copied * 4 + 16);
        size_t table_at = qual.n;
        for (int b = 0; b < nb; b++) sk_u32(&qual, 0);
        for (int b = 0; b < nb; b++) {
            uint32_t len = encode_positions(qual_cat, nb_buf[b], scratch, qual_copied, mask);
            sk_put(&qual, scratch, len);
            uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
            memcpy(qual.p + table_at + 4 * (size_t)b, le, 4);
        }
        int major = hdr_major(h);
        for (uint32_t i = 0; i < qual_copied; i++)
            if (!mask[i] && (int)(signed char)qual_cat[i] != major) { sk_u8(&qual, qual_cat[i]); sk_u32(&qual, i); }
        free(mask); free(scratch);
    } else { free(meta); return fail("quality run-length coding is unreachable under ALGORITHM_VER 2 (Q7)"); }

    /* N positions over the compacted bases: :420-426 */
    sink npos = {0, 0, 0};
    if (h->flags & ORC_ENCODE_N_POS) {
        uint8_t* scratch = (uint8_t*)malloc((size_t)seq_copied * 4 + 16);
        uint32_t len = encode_positions((const uint8_t*)seq_cat, 'N', scratch, seq_copied, NULL);
        sk_put(&npos, scratch, len);
        free(scratch);
    }

    uint16_t flags = 0;
    if (can_il) flags |= ORC_PE_INTERLEAVED;
    if (read_len_same) flags |= ORC_READ_LEN_SAME;
    if (n1len_same) flags |= ORC_NAME1_LEN_SAME;
    if (n2len_same) flags |= ORC_NAME2_LEN_SAME;
    if (slen_same) flags |= ORC_STRAND_LEN_SAME;
    if (strand_same) flags |= ORC_STRAND_SAME;
    if (lane_same) flags |= ORC_LANE_SAME;
    if (tile_same) flags |= ORC_TILE_SAME;
    if (n1_same) flags |= ORC_NAME1_SAME;
    if (n2_same) flags |= ORC_NAME2_SAME;

    sink xbuf = {0, 0, 0}, ybuf = {0, 0, 0};
    int rc_err = 0;
    if (h->flags & ORC_HAS_X) rc_err |= encode_coords(xs, xy_num, &xbuf);
    if (h->flags & ORC_HAS_Y) rc_err |= encode_coords(ys, xy_num, &ybuf);
    if (rc_err) { free(meta); return -1; }

    const uint32_t rl_bytes = h->read_length_bytes;
    const uint32_t readlen_size = read_len_same ? rl_bytes : rl_bytes * s;
    const uint32_t n1len_size = n1len_same ? 1 : s, n2len_size = n2len_same ? 1 : s, slen_size = slen_same ? 1 : s;
    const uint32_t n1_size = n1_same ? m0->name1_len : tot_n1;
    const uint32_t n2_size = n2_same ? m0->name2_len : tot_n2;
    const uint32_t strand_size = strand_same ? r0->strand_len : tot_strand;
    /* Q2: the tile byte count lands in mLaneBufSize, mTileBufSize stays 0 (:503-515) */
    const uint32_t tile_bytes_q2 = tile_same ? 2 : 2 * xy_num;
    uint32_t msize = 18 + readlen_size + n1len_size + n2len_size + slen_size + tile_bytes_q2 + n1_size + n2_size + strand_size
                   + seq_bytes + (uint32_t)qual.n;
    if ((flags & ORC_PE_INTERLEAVED) && (h->flags & ORC_ENCODE_PE_BY_OVERLAP)) msize += s / 2;
    if (h->flags & ORC_ENCODE_N_POS) msize += 4 + (uint32_t)npos.n;
    if (h->flags & ORC_HAS_X) msize += 4 + (uint32_t)xbuf.n;
    if (h->flags & ORC_HAS_Y) msize += 4 + (uint32_t)ybuf.n;
    flags |= extra_flags;

    /* RfqChunk::write, src/rfqchunk.cpp:230-312 */
    sink o = {0, 0, 0};
    sk_u32(&o, msize); sk_u32(&o, s); sk_u16(&o, flags); sk_u32(&o, seq_bytes); sk_u32(&o, (uint32_t)qual.n);
    if (h->flags & ORC_ENCODE_N_POS) sk_u32(&o, (uint32_t)npos.n);
    for (uint32_t i = 0; i < (read_len_same ? 1u : s); i++) {          /* host-endian memcpy of the int (Q15) */
        uint32_t v = reads[i].seq_len;
        if (rl_bytes == 1) sk_u8(&o, (uint8_t)v); else if (rl_bytes == 2) sk_u16(&o, (uint16_t)v); else sk_u32(&o, v);
    }
    for (uint32_t i = 0; i < n1len_size; i++) sk_u8(&o, (uint8_t)meta[i].name1_len);
    if (h->flags & ORC_HAS_NAME2) for (uint32_t i = 0; i < n2len_size; i++) sk_u8(&o, (uint8_t)meta[i].name2_len);
    for (uint32_t i = 0; i < slen_size; i++) sk_u8(&o, (uint8_t)reads[i].strand_len);
    if (h->flags & ORC_HAS_LANE) { if (lane_same) sk_u8(&o, m0->lane); else sk_put(&o, lanes, xy_num); }
    if (h->flags & ORC_HAS_TILE) { if (tile_same) sk_u16(&o, m0->tile); else for (uint32_t i = 0; i < xy_num; i++) sk_u16(&o, tiles[i]); }
    if (h->flags & ORC_HAS_X) { sk_u32(&o, (uint32_t)xbuf.n); sk_put(&o, xbuf.p, xbuf.n); }
    if (h->flags & ORC_HAS_Y) { sk_u32(&o, (uint32_t)ybuf.n); sk_put(&o, ybuf.p, ybuf.n); }
    if (n1_same) sk_put(&o, r0->name, m0->name1_len);
    else for (uint32_t i = 0; i < s; i++) sk_put(&o, reads[i].name, meta[i].name1_len);
    if (h->flags & ORC_HAS_NAME2) {
        if (n2_same) sk_put(&o, name20, m0->name2_len);
        else for (uint32_t i = 0; i < s; i++) sk_put(&o, reads[i].name + meta[i].name2_off, meta[i].name2_len);
    }
    if (strand_same) sk_put(&o, r0->strand, r0->strand_len);
    else for (uint32_t i = 0; i < s; i++) sk_put(&o, reads[i].strand, reads[i].strand_len);
    sk_put(&o, seq_enc, seq_bytes);
    sk_put(&o, qual.p, qual.n);
    if ((flags & ORC_PE_INTERLEAVED) && (h->flags & ORC_ENCODE_PE_BY_OVERLAP)) sk_put(&o, overlaps, s / 2);
    if (h->flags & ORC_ENCODE_N_POS) sk_put(&o, npos.p, npos.n);

    free(meta); free(xs); free(ys); free(lanes); free(tiles); free(seq_cat); free(qual_cat); free(overlaps); free(seq_enc);
    free(qual.p); free(npos.p); free(xbuf.p); free(ybuf.p);
    *out = o.p; *out_len = o.n;
    return 0;
}

/* ------------------------------------------------------------------ decodeChunk ---- */
typedef struct { const uint8_t* p; size_t n, at; int bad; } src_t;
static const uint8_t* take(src_t* s, size_t n) { if (s->at + n > s->n) { s->bad = 1; return NULL; } const uint8_t* r = s->p + s->at; s->at += n; return r; }
static uint32_t take_u32(src_t* s) { const uint8_t* b = take(s, 4); return b ? (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24) : 0; }
static uint16_t take_u16(src_t* s) { const uint8_t* b = take(s, 2); return b ? (uint16_t)(b[0] | (b[1] << 8)) : 0; }

void orc_decoded_free(orc_decoded* d) { free(d->text); free(d->read_end); memset(d, 0, sizeof *d); }

static size_t put_dec(char* p, uint32_t v) { char t[12]; int n = 0; do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v); for (int i = 0; i < n; i++) p[i] = t[n - 1 - i]; return (size_t)n; }

/* RfqChunk::read (src/rfqchunk.cpp:161-228, arena sizes derived :63-109) + RfqCodec::decodeChunk (src/rfqcodec.cpp:1049-1260) */
/* ---- the quality run-length coder that ALGORITHM_VER 2 never selects (Q7: makeQualityTable always sets DONT_ENCODE_QUAL or
 * ENCODE_QUAL_BY_COL), restated so that a header with neither flag decodes as the reference decodes it. */
/* RfqHeader::makeQualBitTable / computeNormalQualBits: src/rfqheader.cpp:103-128 (tables zeroed by the constructor's memset, :8) */
static void rle_tables(const orc_header* h, signed char bit2qual[256], signed char qual2bit[256], int* nq_bits) {
    memset(bit2qual, 0, 256); memset(qual2bit, 0, 256);
    for (int i = 0; i < h->qual_bins; i++) {
        const uint8_t q = h->qual_buf[i];
        const int bit = i > 0 ? 2 * i - 1 : 0;
        qual2bit[q] = (signed char)bit; bit2qual[(uint8_t)bit] = (signed char)q;
    }
    int mx = h->qual_bins * 2 - 3; if (mx < 1) mx = 1;
    *nq_bits = mx >= 64 ? 1 : mx >= 32 ? 2 : mx >= 16 ? 3 : mx >= 8 ? 4 : mx >= 4 ? 5 : mx >= 2 ? 6 : 7;
}
/* RfqCodec::decodeQualByRunLenCoding: src/rfqcodec.cpp:919-955.  The outer `while (decoded < len)` walks the column again from its
 * first byte when it runs out before len positions are filled; an empty column never terminates there (reported as an error). */
static int rle_decode(const orc_header* h, const uint8_t* col, uint32_t col_size, char* qual, uint32_t len) {
    signed char b2q[256], q2b[256]; int nq_bits;
    rle_tables(h, b2q, q2b, &nq_bits);
    const int mq_bits = 7;                                       /* majorQualNumBits(): src/rfqheader.cpp:255-257 */
    uint8_t nq_mask = 0;
    for (int b = 0; b < 8 - nq_bits; b++) nq_mask |= (uint8_t)(1 << b);
    if (len && !col_size) { fail("quality run-length column is empty"); return -1; }
    uint32_t decoded = 0;
    while (decoded < len) {
        for (uint32_t i = 0; i < col_size; i++) {
            const uint8_t e = col[i];
            signed char q; uint8_t num;
            if ((e & 1) == 0) { q = 0; num = (uint8_t)(e >> (8 - mq_bits)); }
            else { q = (signed char)(e & nq_mask); num = (uint8_t)(e >> (8 - nq_bits)); }
            num = (uint8_t)(num + 1);
            const char v = (char)b2q[(uint8_t)q];
            for (uint32_t f = decoded; f < decoded + num && f < len; f++) qual[f] = v;
            decoded += num;
            if (decoded >= len) break;
        }
    }
    return 0;
}
/* RfqCodec::encodeQualRunLenCoding: src/rfqcodec.cpp:767-824 (test-vector construction only: no header the reference makes selects it) */
size_t orc_rle_encode(const orc_header* h, const uint8_t* qual, uint32_t len, uint8_t* out) {
    signed char b2q[256], q2b[256]; int nq_bits;
    rle_tables(h, b2q, q2b, &nq_bits);
    const int mq_bits = 7, mq_max = 1 << mq_bits, nq_max = 1 << nq_bits;
    const char mq = (char)b2q[0];
    size_t n = 0;
    if (!len) return 0;
    char cur = (char)qual[0]; uint32_t first = 0;
    for (uint32_t i = 1; i <= len; i++) {
        int restart = i == len;
        if (!restart) {
            const char q = (char)qual[i];
            if (q != cur) restart = 1;
            else if (cur == mq && (int)(i - first) >= mq_max) restart = 1;
            else if (cur != mq && (int)(i - first) >= nq_max) restart = 1;
        }
        if (restart) {
            const uint8_t num = (uint8_t)(i - first - 1);
            const uint8_t bit = (uint8_t)q2b[(uint8_t)cur];
            out[n++] = (uint8_t)(bit | (uint8_t)(num << (8 - (cur == mq ? mq_bits : nq_bits))));
            first = i;
            if (i < len) cur = (char)qual[i];
        }
    }
    return n;
}

size_t orc_decode_chunk(const orc_header* h, const uint8_t* in, size_t len, orc_decoded* out) {
    memset(out, 0, sizeof *out);
    src_t s = {in, len, 0, 0};
    (void)take_u32(&s);                              /* mSize: read, never used */
    uint32_t reads = take_u32(&s);
    uint16_t flags = take_u16(&s);
    uint32_t seq_size = take_u32(&s), qual_size = take_u32(&s), npos_size = 0;
    if (h->flags & ORC_ENCODE_N_POS) npos_size = take_u32(&s);
    if (s.bad || reads == 0) return 0;
    const uint32_t rlb = h->read_length_bytes;
    if (rlb != 1 && rlb != 2 && rlb != 4) { fail("header incorrect: read length bytes should be 1/2/4"); return 0; }
    const int il = (flags & ORC_PE_INTERLEAVED) != 0;
    const int ov_on = il && (h->flags & ORC_ENCODE_PE_BY_OVERLAP);

    const uint8_t* readlen = take(&s, (size_t)rlb * ((flags & ORC_READ_LEN_SAME) ? 1 : reads));
    const uint8_t* n1len = take(&s, (flags & ORC_NAME1_LEN_SAME) ? 1 : reads);
    const uint8_t* n2len = NULL;
    if (h->flags & ORC_HAS_NAME2) n2len = take(&s, (flags & ORC_NAME2_LEN_SAME) ? 1 : reads);
    const uint8_t* slen = take(&s, (flags & ORC_STRAND_LEN_SAME) ? 1 : reads);
    if (s.bad) return 0;
    const uint32_t xy_num = il ? reads / 2 : reads;
    const uint8_t* lanes = NULL; const uint8_t* tiles = NULL;
    if (h->flags & ORC_HAS_LANE) lanes = take(&s, (flags & ORC_LANE_SAME) ? 1 : xy_num);
    if (h->flags & ORC_HAS_TILE) tiles = take(&s, 2 * (size_t)((flags & ORC_TILE_SAME) ? 1 : xy_num));
    uint32_t xsize = 0, ysize = 0; const uint8_t* xb = NULL; const uint8_t* yb = NULL;
    if (h->flags & ORC_HAS_X) { xsize = take_u32(&s); xb = take(&s, xsize); }
    if (h->flags & ORC_HAS_Y) { ysize = take_u32(&s); yb = take(&s, ysize); }
    if (s.bad) return 0;
#define ARENA(lenarr, lensame, allsame, dst)                                         \
    do { uint64_t t_ = 0; uint32_t c_ = (flags & (lensame)) ? 1 : reads;             \
         for (uint32_t i_ = 0; i_ < c_; i_++) t_ += (lenarr)[i_];                    \
         if ((flags & (lensame)) && !(flags & (allsame))) t_ *= reads;               \
         dst = (uint32_t)t_; } while (0)
    uint32_t n1_size, n2_size = 0, strand_size;
    ARENA(n1len, ORC_NAME1_LEN_SAME, ORC_NAME1_SAME, n1_size);
    if (n2len) ARENA(n2len, ORC_NAME2_LEN_SAME, ORC_NAME2_SAME, n2_size);
    ARENA(slen, ORC_STRAND_LEN_SAME, ORC_STRAND_SAME, strand_size);
    const char* n1 = (const char*)take(&s, n1_size);
    const char* n2 = NULL; if (h->flags & ORC_HAS_NAME2) n2 = (const char*)take(&s, n2_size);
    const char* strand = (const char*)take(&s, strand_size);
    const uint8_t* seqb = take(&s, seq_size);
    const uint8_t* qualb = take(&s, qual_size);
    const signed char* ovb = NULL; if (ov_on) ovb = (const signed char*)take(&s, reads / 2);
    const uint8_t* nposb = NULL; if (h->flags & ORC_ENCODE_N_POS) nposb = take(&s, npos_size);
    if (s.bad) { fail("truncated chunk"); return 0; }

    /* read-length table :1058-1086 */
    uint32_t* rl = (uint32_t*)malloc(4 * (size_t)reads);
    uint64_t total = 0;
    for (uint32_t i = 0; i < reads; i++) {
        uint32_t k = (flags & ORC_READ_LEN_SAME) ? 0 : i;
        uint32_t v = rlb == 1 ? readlen[k] : rlb == 2 ? (uint32_t)(readlen[2 * k] | (readlen[2 * k + 1] << 8))
                   : (uint32_t)readlen[4 * k] | ((uint32_t)readlen[4 * k + 1] << 8) | ((uint32_t)readlen[4 * k + 2] << 16) | ((uint32_t)readlen[4 * k + 3] << 24);
        rl[i] = v; total += v;
    }
    const uint32_t L = (uint32_t)total;
    char* seq = (char*)malloc((size_t)L + 1); memset(seq, 'N', L);
    char* qual = (char*)malloc((size_t)L + 1); memset(qual, hdr_major(h), L);

    /* decodeSeqQual :826-917 */
    if (L) {
        static const char B[4] = {'G', 'A', 'T', 'C'};
        uint32_t d = 0;
        for (uint32_t i = 0; i < seq_size && d < L; i++)
            for (int b = 0; b < 4 && d < L; b++) seq[d++] = B[(seqb[i] >> (2 * b)) & 3];
        if (h->flags & ORC_ENCODE_N_POS) decode_positions(nposb, npos_size, 'N', seq, L);
        if (ov_on) {
            char* dst = (char*)malloc((size_t)L + 1);
            uint32_t sp = 0, dp = 0;
            for (uint32_t r = 0; r < reads; r++) {
                uint32_t rlen = rl[r];
                int o = (r & 1) ? (int)ovb[r / 2] - (int)h->overlap_shift : 0;
                if (o == 0) { memcpy(dst + dp, seq + sp, rlen); sp += rlen; }
                else if (o > 0) { memcpy(dst + dp, seq + sp - o, (size_t)o); memcpy(dst + dp + o, seq + sp, rlen - (uint32_t)o); sp += rlen - (uint32_t)o; }
                else { uint32_t k = rlen - (uint32_t)(-o); memcpy(dst + dp, seq + sp, k); memcpy(dst + dp + k, seq + sp - rl[r - 1], (size_t)(-o)); sp += k; }
                dp += rlen;
            }
            free(seq); seq = dst;
        }
        if (h->flags & ORC_DONT_ENCODE_QUAL) { for (uint32_t i = 0; i < qual_size && i < L; i++) qual[i] = (char)qualb[i]; }
        else if (h->flags & ORC_ENCODE_QUAL_BY_COL) {
            uint8_t nb_buf[256]; int nb = hdr_normal_bins(h, nb_buf);
            uint32_t c = 4 * (uint32_t)nb;
            for (int b = 0; b < nb; b++) {
                const uint8_t* t = qualb + 4 * b;
                uint32_t sl = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                decode_positions(qualb + c, sl, (char)nb_buf[b], qual, L);
                c += sl;
            }
            while (c < qual_size) {                                  /* exceptions :1034-1043 */
                char q = (char)qualb[c]; const uint8_t* t = qualb + c + 1;
                uint32_t pos = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                c += 5;
                if (pos < L) qual[pos] = q;
            }
        } else if (rle_decode(h, qualb, qual_size, qual, L)) { free(seq); free(qual); return 0; }
    }
    if (!(h->flags & ORC_ENCODE_N_POS))                              /* :1093-1100 */
        for (uint32_t i = 0; i < L; i++) if (qual[i] == (char)h->n_base_qual) seq[i] = 'N';

    uint32_t* xs = (uint32_t*)calloc(xy_num + 1, 4); uint32_t* ys = (uint32_t*)calloc(xy_num + 1, 4);
    if (h->flags & ORC_HAS_X) decode_coords(xb, xsize, xs, xy_num);
    if (h->flags & ORC_HAS_Y) decode_coords(yb, ysize, ys, xy_num);

    /* per read :1141-1254, then Read::toString (src/read.cpp:170-172) */
    size_t cap = (size_t)L * 2 + (size_t)reads * 64 + n1_size + n2_size + strand_size + 16;
    cap += (size_t)reads * ((size_t)n1len[0] + (n2len ? n2len[0] : 0) + slen[0] + 48);
    char* text = (char*)malloc(cap); size_t w = 0;
    out->read_end = (size_t*)malloc(sizeof(size_t) * reads);
    const char* c1 = n1; const char* c2 = n2; const char* cs = strand; uint32_t cur = 0;
    for (uint32_t r = 0; r < reads; r++) {
        uint32_t rlen = rl[r];
        uint32_t l1 = (flags & (ORC_NAME1_SAME | ORC_NAME1_LEN_SAME)) ? n1len[0] : n1len[r];
        if (flags & ORC_NAME1_SAME) memcpy(text + w, n1, l1); else { memcpy(text + w, c1, l1); c1 += l1; }
        w += l1;
        uint32_t xy = il ? r / 2 : r;
        if (h->flags & ORC_HAS_LANE) { text[w++] = ':'; w += put_dec(text + w, (flags & ORC_LANE_SAME) ? lanes[0] : lanes[xy]); }
        if (h->flags & ORC_HAS_TILE) { uint32_t k = (flags & ORC_TILE_SAME) ? 0 : xy; text[w++] = ':'; w += put_dec(text + w, (uint32_t)(tiles[2 * k] | (tiles[2 * k + 1] << 8))); }
        if (h->flags & ORC_HAS_X) { text[w++] = ':'; w += put_dec(text + w, xs[xy]); }
        if (h->flags & ORC_HAS_Y) { text[w++] = ':'; w += put_dec(text + w, ys[xy]); }
        if (h->flags & ORC_HAS_NAME2) {
            uint32_t l2 = (flags & (ORC_NAME2_SAME | ORC_NAME2_LEN_SAME)) ? n2len[0] : n2len[r];
            if (flags & ORC_NAME2_SAME) {
                memcpy(text + w, n2, l2);
                if (il && (r & 1) && h->name2_diff_char != '\0' && h->name2_diff_pos < l2) text[w + h->name2_diff_pos] = h->name2_diff_char;
            } else { memcpy(text + w, c2, l2); c2 += l2; }
            w += l2;
        }
        text[w++] = '\n';
        if (il && (r & 1)) { for (uint32_t k = 0; k < rlen; k++) text[w + k] = complement(seq[cur + rlen - 1 - k]); }
        else memcpy(text + w, seq + cur, rlen);
        w += rlen; text[w++] = '\n';
        uint32_t ls = (flags & (ORC_STRAND_SAME | ORC_STRAND_LEN_SAME)) ? slen[0] : slen[r];
        if (flags & ORC_STRAND_SAME) memcpy(text + w, strand, ls); else { memcpy(text + w, cs, ls); cs += ls; }
        w += ls; text[w++] = '\n';
        if (il && (r & 1)) { for (uint32_t k = 0; k < rlen; k++) text[w + k] = qual[cur + rlen - 1 - k]; }
        else memcpy(text + w, qual + cur, rlen);
        w += rlen; text[w++] = '\n';
        cur += rlen;
        out->read_end[r] = w;
    }
    free(rl); free(seq); free(qual); free(xs); free(ys);
    out->n_reads = reads; out->flags = flags; out->text = text; out->text_len = w;
    return s.at;
}

/* ------------------------------------------------------------------ Repaq driver ---- */
typedef struct { orc_read* v; char** own; size_t n, cap; } readvec;
static void rv_push(readvec* rv, const orc_read* r) {
    if (rv->n == rv->cap) { rv->cap = rv->cap ? rv->cap * 2 : 1024; rv->v = (orc_read*)realloc(rv->v, rv->cap * sizeof(orc_read)); rv->own = (char**)realloc(rv->own, rv->cap * sizeof(char*)); }
    size_t tot = (size_t)r->name_len + r->seq_len + r->strand_len + r->qual_len;
    char* b = (char*)malloc(tot + 1); char* p = b;
    orc_read c = *r;
    memcpy(p, r->name, r->name_len); c.name = p; p += r->name_len;
    memcpy(p, r->seq, r->seq_len); c.seq = p; p += r->seq_len;
    memcpy(p, r->strand, r->strand_len); c.strand = p; p += r->strand_len;
    memcpy(p, r->qual, r->qual_len); c.qual = p;
    rv->v[rv->n] = c; rv->own[rv->n] = b; rv->n++;
}
static void rv_clear(readvec* rv) { for (size_t i = 0; i < rv->n; i++) free(rv->own[i]); rv->n = 0; }

/* src/repaq.cpp:530-638 (compress), :640-759 (compressPE); FastqReaderPair::read src/fastqreader.cpp:287-299 */
static int compress_impl(const char* r1, size_t l1, const char* r2, size_t l2, int interleaved, uint32_t chunk_bases,
                         const orc_header* fixed, uint8_t** out, size_t* out_len) {
    const int is_pe = (r2 != NULL) || interleaved;
    orc_reader* a = orc_reader_open(r1, l1);
    orc_reader* b = (r2 != NULL) ? orc_reader_open(r2, l2) : NULL;
    sink o = {0, 0, 0};
    readvec rv = {0, 0, 0, 0};
    orc_header h; int have_header = 0;
    if (fixed) { h = *fixed; have_header = 1; }       /* a later part of a sharded file: the header of chunk 0 is given, only chunks are written */
    uint32_t total = 0; int rc = 0;
    for (;;) {
        orc_read x, y; int got = orc_reader_next(a, &x);
        int flush = 0;
        if (is_pe) {
            /* FastqReaderPair::read always pulls both mates, even when the left one is already exhausted */
            if (got) rv_push(&rv, &x);              /* copy first: the next read reuses the line buffers */
            int got2 = b ? orc_reader_next(b, &y) : orc_reader_next(a, &y);
            if (got && !got2) { free(rv.own[rv.n - 1]); rv.n--; }
            got = got && got2;
            if (got) { rv_push(&rv, &y); total += rv.v[rv.n - 2].seq_len + y.seq_len; }
        } else if (got) { rv_push(&rv, &x); total += x.seq_len; }
        if (got && total >= chunk_bases) flush = 1;
        if (!got && rv.n > 0) flush = 1;
        if (flush) {
            if (!have_header) {
                if (orc_make_header(rv.v, rv.n, is_pe, &h)) { rc = -1; break; }
                uint8_t hb[17 + 256]; size_t hn = orc_header_write(&h, hb); sk_put(&o, hb, hn);
                have_header = 1;
            }
            uint16_t extra = 0;
            if (orc_reader_no_line_break_at_end(a)) extra |= ORC_NO_LINE_BREAK_AT_END;
            if (is_pe) { if (b ? orc_reader_no_line_break_at_end(b) : orc_reader_no_line_break_at_end(a)) extra |= ORC_NO_LINE_BREAK_AT_END_R2; }
            uint8_t* cb; size_t cn;
            if (orc_encode_chunk(&h, rv.v, rv.n, is_pe, extra, &cb, &cn)) { rc = -1; break; }
            sk_put(&o, cb, cn); free(cb);
            rv_clear(&rv); total = 0;
        }
        if (!got) break;
    }
    rv_clear(&rv); free(rv.v); free(rv.own);
    orc_reader_close(a); orc_reader_close(b);
    if (rc) { free(o.p); return rc; }
    *out = o.p; *out_len = o.n;
    return 0;
}

int orc_compress(const char* r1, size_t l1, const char* r2, size_t l2, int interleaved, uint32_t chunk_bases,
                 uint8_t** out, size_t* out_len) {
    return compress_impl(r1, l1, r2, l2, interleaved, chunk_bases, NULL, out, out_len);
}

/* The chunks of a text encoded with a GIVEN file header (serialised, as RfqHeader::write leaves it): what Repaq::compress writes
 * after the header for these records when its first chunk - the one the header is made from, src/repaq.cpp:554-566 - was
 * somewhere else.  Checker for chunk-sharded encodes: every rank encodes with rank 0's header. */
int orc_compress_with_header(const uint8_t* header, size_t header_len, const char* r1, size_t l1, const char* r2, size_t l2, int interleaved,
                             uint32_t chunk_bases, uint8_t** out, size_t* out_len) {
    orc_header h;
    if (!orc_header_read(header, header_len, &h)) return -1;
    return compress_impl(r1, l1, r2, l2, interleaved, chunk_bases, &h, out, out_len);
}

/* src/repaq.cpp:262-333 (decompress), :335-413 (decompressPE) */
int orc_decompress(const uint8_t* rfq, size_t len, int pe_out, char** out1, size_t* l1, char** out2, size_t* l2) {
    if (len == 0) {
        /* RfqHeader::read on an empty stream: every ifs.read() fails and the constructor's values stay (src/rfqheader.cpp:7-43) -
         * "RFQ", ALGORITHM_VER, flags 0: a single-end file without chunks (what `repaq -c` leaves for an input without records) */
        if (pe_out) return fail("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
        *out1 = NULL; *l1 = 0; *out2 = NULL; *l2 = 0;
        return 0;
    }
    orc_header h; size_t at = orc_header_read(rfq, len, &h);
    if (!at) return -1;
    if (pe_out && !(h.flags & ORC_PAIRED_END)) return fail("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
    sink a = {0, 0, 0}, b = {0, 0, 0};
    while (at < len) {
        orc_decoded d; size_t used = orc_decode_chunk(&h, rfq + at, len - at, &d);
        if (!used) break;
        at += used;
        /* is there a further chunk with reads? (peek: src/repaq.cpp:304-314 / :378-388) */
        orc_decoded peek; size_t peek_used = 0; int last = 1;
        const int f1 = (d.flags & ORC_NO_LINE_BREAK_AT_END) != 0, f2 = (d.flags & ORC_NO_LINE_BREAK_AT_END_R2) != 0;
        if (f1 || (pe_out && f2)) {
            if (at < len) { peek_used = orc_decode_chunk(&h, rfq + at, len - at, &peek); if (peek_used) { last = 0; orc_decoded_free(&peek); } }
        }
        if (!pe_out) {
            if (f1 && last) { sk_put(&a, d.text, d.text_len ? d.text_len - 1 : 0); orc_decoded_free(&d); break; }
            sk_put(&a, d.text, d.text_len);
        } else {
            sink s1 = {0, 0, 0}, s2 = {0, 0, 0};
            size_t prev = 0;
            for (size_t r = 0; r < d.n_reads; r++) { sk_put((r & 1) ? &s2 : &s1, d.text + prev, d.read_end[r] - prev); prev = d.read_end[r]; }
            int skip_rest = 0;
            if (f1) {
                if (last) sk_put(&a, s1.p, s1.n ? s1.n - 1 : 0);
                else { sk_put(&a, s1.p, s1.n); skip_rest = 1; }   /* `continue`: R2 of this chunk and the peeked chunk are dropped */
            } else sk_put(&a, s1.p, s1.n);
            if (!skip_rest) {
                if (f2) {
                    if (last) sk_put(&b, s2.p, s2.n ? s2.n - 1 : 0);
                    else { sk_put(&b, s2.p, s2.n); skip_rest = 1; }
                } else sk_put(&b, s2.p, s2.n);
            }
            free(s1.p); free(s2.p);
            if ((f1 || f2) && !last) at += peek_used;   /* the peeked chunk was consumed from the stream and lost */
        }
        orc_decoded_free(&d);
    }
    *out1 = (char*)a.p; *l1 = a.n;
    if (out2) { *out2 = (char*)b.p; *l2 = b.n; } else free(b.p);
    return 0;
}

/* ------------------------------------------------------------------ compare mode ---- */
static void sb_put(strbuf* b, const char* s, size_t n) {
    if (b->n + n + 1 > b->cap) { b->cap = (b->n + n + 1) * 2; b->p = (char*)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0;
}
static void sb_str(strbuf* b, const char* s) { sb_put(b, s, strlen(s)); }
static void sb_num(strbuf* b, long v) { char t[32]; snprintf(t, sizeof t, "%ld", v); sb_str(b, t); }

/* Repaq::reportCompareResult (src/repaq.cpp:235-259) */
static char* cmp_report(int passed, const strbuf* msg, long fq_reads, long fq_bases, long rfq_reads, long rfq_bases) {
    strbuf j = {0, 0, 0};
    sb_str(&j, "{\n");
    sb_str(&j, passed ? "\t\"result\":\"passed\",\n" : "\t\"result\":\"failed\",\n");
    sb_str(&j, "\t\"msg\":\""); if (msg->n) sb_put(&j, msg->p, msg->n); sb_str(&j, "\",\n");
    sb_str(&j, "\t\"fastq_reads\":"); sb_num(&j, fq_reads); sb_str(&j, ",\n");
    sb_str(&j, "\t\"rfq_reads\":"); sb_num(&j, rfq_reads); sb_str(&j, ",\n");
    sb_str(&j, "\t\"fastq_bases\":"); sb_num(&j, fq_bases); sb_str(&j, ",\n");
    sb_str(&j, "\t\"rfq_bases\":"); sb_num(&j, rfq_bases); sb_str(&j, "\n");
    sb_str(&j, "}\n");
    return j.p;
}

/* Repaq::compare (src/repaq.cpp:36-130) and comparePE (:132-233): every chunk of the .rfq is decoded and its reads are
 * checked name, sequence, strand, quality against the reads of the FASTQ file(s); the first difference ends the run.
 * r2 == NULL: single end.  *json = the text the reference prints (malloc'd). */
int orc_compare(const uint8_t* rfq, size_t len, const char* r1, size_t l1, const char* r2, size_t l2, char** json) {
    orc_header h; size_t at = orc_header_read(rfq, len, &h);
    if (!at) return -1;
    const int pe = r2 != NULL;
    orc_reader* a = orc_reader_open(r1, l1);
    orc_reader* b = pe ? orc_reader_open(r2, l2) : NULL;
    long fq_reads = 0, fq_bases = 0, rfq_reads = 0, rfq_bases = 0;
    strbuf msg = {0, 0, 0};
    int done = 0, passed = 0;
    orc_read left, right; int have_pair = 0;
    char* keep[4] = {0, 0, 0, 0};                      /* the left mate's fields: the reader reuses its line buffers */
    const char* unit = pe ? " pair. " : " read. ";
    while (!done && at < len) {
        orc_decoded d; size_t used = orc_decode_chunk(&h, rfq + at, len - at, &d);
        if (!used) break;
        at += used;
        size_t prev = 0;
        for (size_t r = 0; r < d.n_reads && !done; r++) {
            /* the decoded read: four lines of d.text[prev, read_end[r]) */
            const char* f[4]; size_t fl[4]; size_t p = prev;
            for (int k = 0; k < 4; k++) { f[k] = d.text + p; size_t q = p; while (d.text[q] != '\n') q++; fl[k] = q - p; p = q + 1; }
            prev = d.read_end[r];
            rfq_bases += (long)fl[1];
            rfq_reads++;
            orc_read fq; int got = 1;
            if (!pe) got = orc_reader_next(a, &fq);
            else {
                if (!have_pair) {
                    /* FastqReaderPair::read pulls both mates, and fails when either is missing */
                    int g1 = orc_reader_next(a, &left);
                    if (g1) {
                        const char* src[4] = {left.name, left.seq, left.strand, left.qual}; const uint32_t sl[4] = {left.name_len, left.seq_len, left.strand_len, left.qual_len};
                        for (int k = 0; k < 4; k++) { free(keep[k]); keep[k] = (char*)malloc(sl[k] + 1); memcpy(keep[k], src[k], sl[k]); }
                        left.name = keep[0]; left.seq = keep[1]; left.strand = keep[2]; left.qual = keep[3];
                    }
                    int g2 = orc_reader_next(b, &right);
                    got = g1 && g2;
                    have_pair = got;
                }
                if (got) fq = (rfq_reads % 2 == 1) ? left : right;
            }
            const long shown = pe ? rfq_reads / 2 : rfq_reads;
            if (!got) {
                sb_str(&msg, "The RFQ file has more reads than the FASTQ file. The RFQ file has >= "); sb_num(&msg, shown);
                sb_str(&msg, pe ? " pairs, while the FASTQ file only has " : " reads, while the FASTQ file only has ");
                sb_num(&msg, pe ? fq_reads / 2 : fq_reads); sb_str(&msg, pe ? " pairs" : " reads");
                done = 1; break;
            }
            fq_reads++; fq_bases += (long)fq.seq_len;
            const char* g[4] = {fq.name, fq.seq, fq.strand, fq.qual}; const uint32_t gl[4] = {fq.name_len, fq.seq_len, fq.strand_len, fq.qual_len};
            static const char* what[4] = {"name", "sequence", "strand", "quality"};
            for (int k = 0; k < 4; k++) {
                if (fl[k] != gl[k] || memcmp(f[k], g[k], fl[k]) != 0) {
                    sb_str(&msg, "The RFQ file and FASTQ file have different "); sb_str(&msg, what[k]); sb_str(&msg, " in the "); sb_num(&msg, shown); sb_str(&msg, unit);
                    sb_put(&msg, f[k], fl[k]); sb_str(&msg, " | "); sb_put(&msg, g[k], gl[k]);
                    done = 1; break;
                }
            }
            if (pe && rfq_reads % 2 == 0) have_pair = 0;
        }
        orc_decoded_free(&d);
    }
    if (!done) {
        orc_read x, y; int more;
        if (!pe) more = orc_reader_next(a, &x);
        else { int g1 = orc_reader_next(a, &x); int g2 = orc_reader_next(b, &y); more = g1 && g2; }
        if (more) {
            fq_reads++;
            sb_str(&msg, "The FASTQ file has more reads than the RFQ file. The FASTQ file has >= "); sb_num(&msg, pe ? fq_reads / 2 : fq_reads);
            sb_str(&msg, pe ? " pairs, while the RFQ file only has " : " reads, while the RFQ file only has ");
            sb_num(&msg, pe ? rfq_reads / 2 : rfq_reads); sb_str(&msg, pe ? " pairs" : " reads");
        } else passed = 1;
    }
    *json = cmp_report(passed, &msg, fq_reads, fq_bases, rfq_reads, rfq_bases);
    free(msg.p);
    for (int k = 0; k < 4; k++) free(keep[k]);
    orc_reader_close(a); if (b) orc_reader_close(b);
    return 0;
}
