/*
 * fqgen - deterministic synthetic FASTQ generators for the workloads BASELINE.json names
 * (SURVEY.md section 8d).  Not part of the codec: it only makes inputs for tests and bench.py.
 *
 * Generation unit is a "row" (ROW_READS reads sharing one y coordinate, x strictly increasing),
 * seeded by (seed, global row index), so any row range can be produced independently and in
 * parallel and the bytes only depend on (seed, shape, flags, row).
 *
 * Shapes
 *   FQ_NOVA  NovaSeq-like 150 bp, Illumina names, 4 quality bins F : , #  ('#' only on N),
 *            paired: fragment N(350,90) clipped to [40,900], R2 = revcomp of the fragment tail,
 *            adapter read-through below read length, 0.2 % substitution errors.
 *   FQ_BGI   BGI-like 100 bp single end, names without colons, ~40 quality values, rare N.
 *
 * Build: gcc -O2 -fPIC -shared -o tools/libfqgen.so tools/fqgen.c -lm
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

#define ROW_READS 300
#define ROWS_PER_TILE 2000

#define FQ_NOVA 0
#define FQ_BGI 1

#define FQ_FLAG_NO_N_EARLY 1u   /* no N in the first 8000 reads: header takes the N_POS path */
#define FQ_FLAG_CRLF 2u         /* \r\n line ends */
#define FQ_FLAG_VARLEN 4u       /* variable read lengths (trimmed reads) */
#define FQ_FLAG_LONG 8u         /* 300 bp reads (2-byte length column) */

typedef struct { uint64_t s; } rng_t;

static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t rng_next(rng_t* r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32); }
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rng_gauss(rng_t* r) {
    double u1 = rng_unit(r), u2 = rng_unit(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}

static char* put_uint(char* p, uint64_t v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
static char* put_str(char* p, const char* s) { size_t n = strlen(s); memcpy(p, s, n); return p + n; }
static char* put_pad(char* p, uint64_t v, int width) {
    for (int i = width - 1; i >= 0; i--) { p[i] = (char)('0' + v % 10); v /= 10; }
    return p + width;
}

/* NovaSeq-like quality string: 2-state Markov chain, low-quality runs of geometric mean 3,
 * more frequent towards the read tail. */
static void nova_qual(rng_t* r, char* q, int len) {
    int bad = 0; char badq = ':';
    for (int i = 0; i < len; i++) {
        uint32_t u = (uint32_t)(rng_next(r) >> 40);           /* 24 bits */
        if (!bad) {
            double f = (double)i / (double)len;
            uint32_t thr = (uint32_t)(16777216.0 * 0.017 * (1.0 + 4.0 * f * f));
            if (u < thr) { bad = 1; badq = (rng_next(r) & 0xFF) < 150 ? ':' : ','; }
        } else {
            if (u < 16777216u / 3) bad = 0;
            else if ((u & 0xFF) < 40) badq = (badq == ':') ? ',' : ':';
        }
        q[i] = bad ? badq : 'F';
    }
}

/* sequence a read of `len` from `frag` (flen bases) + adapter read-through + random tail */
static void read_from_fragment(rng_t* r, const char* frag, int flen, const char* adapter, char* seq, int len) {
    int alen = (int)strlen(adapter);
    for (int i = 0; i < len; i++) {
        if (i < flen) seq[i] = frag[i];
        else if (i - flen < alen) seq[i] = adapter[i - flen];
        else seq[i] = BASES[rng_below(r, 4)];
    }
}

static void add_errors_and_n(rng_t* r, char* seq, char* qual, int len, int allow_n) {
    for (int i = 0; i < len; i++) {
        uint64_t u = rng_next(r);
        if ((u & 0xFFFF) < 131) {                            /* 0.2 % substitution */
            char c = BASES[(u >> 16) & 3];
            if (c == seq[i]) c = BASES[((u >> 16) + 1) & 3];
            seq[i] = c;
        }
        if (allow_n && ((u >> 32) & 0xFFFFF) < 210) {        /* ~2e-4 N, always with '#' */
            seq[i] = 'N'; qual[i] = '#';
        }
    }
}

static const char* ADAPTER1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACACTGTTCCATCTCGTATGCCGTCTTCTGCTTG";
static const char* ADAPTER2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT";

static char* emit_record(char* p, const char* name, size_t nlen, const char* seq, const char* qual, int len, int crlf) {
    memcpy(p, name, nlen); p += nlen; if (crlf) *p++ = '\r'; *p++ = '\n';
    memcpy(p, seq, (size_t)len); p += len; if (crlf) *p++ = '\r'; *p++ = '\n';
    *p++ = '+'; if (crlf) *p++ = '\r'; *p++ = '\n';
    memcpy(p, qual, (size_t)len); p += len; if (crlf) *p++ = '\r'; *p++ = '\n';
    return p;
}

/* upper bound of bytes per record for sizing buffers */
size_t fqgen_max_record_bytes(int shape, uint32_t flags) {
    (void)shape;
    int len = (flags & FQ_FLAG_LONG) ? 300 : 150;
    return (size_t)(2 * len + 96);
}

int fqgen_row_reads(void) { return ROW_READS; }

/*
 * Generate rows [first_row, first_row + n_rows).  out2 may be NULL (single end).
 * Returns 0, or -1 if a buffer is too small.  *len1 / *len2 receive the bytes written.
 */
int fqgen_rows(uint64_t seed, int shape, uint32_t flags, uint64_t first_row, uint64_t n_rows,
               char* out1, size_t cap1, size_t* len1, char* out2, size_t cap2, size_t* len2) {
    char* p1 = out1; char* p2 = out2;
    const int crlf = (flags & FQ_FLAG_CRLF) != 0;
    const int base_len = (shape == FQ_BGI) ? 100 : ((flags & FQ_FLAG_LONG) ? 300 : 150);
    char frag[1100], s1[512], s2[512], q1[512], q2[512], tmp[512], name[160];
    const size_t maxrec = fqgen_max_record_bytes(shape, flags);

    for (uint64_t g = first_row; g < first_row + n_rows; g++) {
        rng_t r; r.s = mix64(seed * 0x100000001B3ull + g);
        uint64_t tile_idx = g / ROWS_PER_TILE, row_in_tile = g % ROWS_PER_TILE;
        uint32_t t624 = (uint32_t)(tile_idx % 624);
        uint32_t tile = (1 + t624 / 312) * 1000 + (1 + (t624 / 78) % 4) * 100 + (1 + t624 % 78);
        uint32_t lane = 1 + (uint32_t)((tile_idx / 624) % 4);
        uint32_t y = 1000 + 16 * (uint32_t)row_in_tile + (uint32_t)(mix64(seed ^ (g * 31)) % 3);
        uint32_t x = 1000 + rng_below(&r, 100);
        for (int k = 0; k < ROW_READS; k++) {
            uint64_t ridx = g * ROW_READS + (uint64_t)k;
            if ((size_t)(p1 - out1) + maxrec > cap1) return -1;
            if (out2 && (size_t)(p2 - out2) + maxrec > cap2) return -1;
            int allow_n = !((flags & FQ_FLAG_NO_N_EARLY) && ridx < 8000);
            if (shape == FQ_BGI) {
                int len = base_len;
                if (flags & FQ_FLAG_VARLEN) len = 30 + (int)rng_below(&r, 71);
                for (int i = 0; i < len; i++) s1[i] = BASES[rng_below(&r, 4)];
                for (int i = 0; i < len; i++) {
                    int q = (int)lrint(35.0 + 4.5 * rng_gauss(&r));
                    if (q < 2) q = 2;
                    if (q > 41) q = 41;
                    q1[i] = (char)(33 + q);
                }
                /* < 100 N in total: one N per ~200k reads */
                if (allow_n && rng_below(&r, 200000) == 0) { int pos = (int)rng_below(&r, (uint32_t)len); s1[pos] = 'N'; q1[pos] = '"'; }
                char* n = name;
                n = put_str(n, "@V300035135L2C");
                n = put_pad(n, 1 + (ridx / 10000000ull / 80) % 999, 3);
                *n++ = 'R';
                n = put_pad(n, 1 + (ridx / 10000000ull) % 80, 3);
                n = put_pad(n, ridx % 10000000ull, 7);
                n = put_str(n, "/1");
                p1 = emit_record(p1, name, (size_t)(n - name), s1, q1, len, crlf);
                continue;
            }
            /* NovaSeq shape */
            x += 9 + rng_below(&r, 172);
            int len1 = base_len, len2 = base_len;
            if (flags & FQ_FLAG_VARLEN) { len1 = 35 + (int)rng_below(&r, (uint32_t)(base_len - 34)); len2 = 35 + (int)rng_below(&r, (uint32_t)(base_len - 34)); }
            int flen = (int)lrint(350.0 + 90.0 * rng_gauss(&r));
            if (flen < 40) flen = 40;
            if (flen > 900) flen = 900;
            if (flags & FQ_FLAG_LONG) flen += 150;
            for (int i = 0; i < flen; i++) frag[i] = BASES[rng_below(&r, 4)];
            read_from_fragment(&r, frag, flen, ADAPTER1, s1, len1);
            nova_qual(&r, q1, len1);
            add_errors_and_n(&r, s1, q1, len1, allow_n);
            char* n = name;
            n = put_str(n, "@A00250:26:H3YTWDSXX:");
            n = put_uint(n, lane); *n++ = ':';
            n = put_uint(n, tile); *n++ = ':';
            n = put_uint(n, x); *n++ = ':';
            n = put_uint(n, y);
            char* mate = n;
            n = put_str(n, " 1:N:0:ACTGTTCC");
            p1 = emit_record(p1, name, (size_t)(n - name), s1, q1, len1, crlf);
            if (out2) {
                /* R2 reads the reverse strand from the fragment's far end */
                int take = flen < 512 ? flen : 512;
                for (int i = 0; i < take; i++) tmp[i] = comp(frag[flen - 1 - i]);
                read_from_fragment(&r, tmp, take, ADAPTER2, s2, len2);
                nova_qual(&r, q2, len2);
                add_errors_and_n(&r, s2, q2, len2, allow_n);
                mate[1] = '2';
                p2 = emit_record(p2, name, (size_t)(n - name), s2, q2, len2, crlf);
            }
        }
    }
    *len1 = (size_t)(p1 - out1);
    if (len2) *len2 = out2 ? (size_t)(p2 - out2) : 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Parallel front end (OpenMP): rows [first_row, first_row+n_rows) generated block-wise by all cores into scratch
 * buffers, then packed into out1/out2 in row order.  Output bytes are identical to fqgen_rows().
 */
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fqgen_rows_mt(uint64_t seed, int shape, uint32_t flags, uint64_t first_row, uint64_t n_rows,
                  char* out1, size_t cap1, size_t* len1, char* out2, size_t cap2, size_t* len2) {
    const uint64_t BLK = 32;
    const uint64_t nblk = (n_rows + BLK - 1) / BLK;
    const size_t blk_cap = fqgen_max_record_bytes(shape, flags) * ROW_READS * BLK + 64;
    char** b1 = (char**)calloc(nblk, sizeof(char*)); char** b2 = (char**)calloc(nblk, sizeof(char*));
    size_t* n1 = (size_t*)calloc(nblk + 1, sizeof(size_t)); size_t* n2 = (size_t*)calloc(nblk + 1, sizeof(size_t));
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (long long k = 0; k < (long long)nblk; k++) {
        uint64_t r0 = first_row + (uint64_t)k * BLK, nr = n_rows - (uint64_t)k * BLK < BLK ? n_rows - (uint64_t)k * BLK : BLK;
        b1[k] = (char*)malloc(blk_cap); if (out2) b2[k] = (char*)malloc(blk_cap);
        size_t a = 0, b = 0;
        if (fqgen_rows(seed, shape, flags, r0, nr, b1[k], blk_cap, &a, out2 ? b2[k] : NULL, out2 ? blk_cap : 0, &b)) {
#pragma omp atomic write
            rc = -1;
        }
        n1[k] = a; n2[k] = b;
    }
    size_t t1 = 0, t2 = 0;
    for (uint64_t k = 0; k < nblk; k++) { size_t a = n1[k], b = n2[k]; n1[k] = t1; n2[k] = t2; t1 += a; t2 += b; }
    n1[nblk] = t1; n2[nblk] = t2;
    if (t1 > cap1 || (out2 && t2 > cap2)) rc = -1;
    if (!rc) {
#pragma omp parallel for schedule(dynamic, 1)
        for (long long k = 0; k < (long long)nblk; k++) {
            memcpy(out1 + n1[k], b1[k], n1[k + 1] - n1[k]);
            if (out2) memcpy(out2 + n2[k], b2[k], n2[k + 1] - n2[k]);
        }
    }
    for (uint64_t k = 0; k < nblk; k++) { free(b1[k]); free(b2[k]); }
    free(b1); free(b2);
    *len1 = t1; if (len2) *len2 = out2 ? t2 : 0;
    free(n1); free(n2);
    return rc;
}

void fqgen_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
