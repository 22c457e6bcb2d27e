/*
 * rpq_common.cuh - device-side data model shared by the encode and decode kernels.
 *
 * Vocabulary follows the reference: a FASTQ *record* is 4 lines (name, sequence, strand, quality); a *read* is one
 * record; paired-end reads are interleaved R1,R2,R1,R2 inside a *chunk* exactly as RfqCodec::encodeChunk(pairs)
 * does (reference src/rfqcodec.cpp:147-161); a chunk is the unit of the .rfq container (src/rfqchunk.h:52-113).
 */
#pragma once
#include "rpq_rt.h"
#include "../../include/repaq_b200.h"

namespace rpq {

typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned int u32;
typedef unsigned long long u64;

constexpr u32 NONE32 = 0xFFFFFFFFu;

/* bin LUT classes */
constexpr u8 LUT_SKIP = 0xFF;   /* the major quality: no stream, it is the decoder's fill value */
constexpr u8 LUT_EXC = 0xFE;    /* not in the header alphabet: 5-byte exception record */
constexpr int MAX_BINS = 64;    /* ENCODE_QUAL_BY_COL needs mQualBins <= 64 (src/rfqheader.cpp:233-234) */

/* RfqHeader as the kernels see it (kernel parameter, < 400 bytes) */
struct HeaderDev {
    u16 flags;
    u8 read_length_bytes;
    u8 name2_diff_pos;
    u8 name2_diff_char;
    signed char n_base_qual;
    signed char overlap_shift;
    u8 support_interleaved;
    u8 major;                 /* majorQual() = mQualBuf[0] */
    u8 nb;                    /* normalQualBins() */
    u8 normal_bins[MAX_BINS + 1];  /* normalQualBuf() order = stream order */
    u8 lut[256];              /* quality byte -> stream index | LUT_SKIP | LUT_EXC */
    u8 rle_b2q[128];          /* mBit2QualTable (src/rfqheader.cpp:103-115): code of the quality run-length coder -> quality */
    u8 rle_nq_bits;           /* mNormalQualNumBits (src/rfqheader.cpp:117-128) */
};

/* error bits raised by kernels (per batch) */
constexpr u32 ERRBIT_EMPTY_LINE = 1u << 0;     /* informational: input ends at the first record with an empty line */
constexpr u32 ERRBIT_QUAL_LEN = 1u << 1;       /* quality length != sequence length */
constexpr u32 ERRBIT_NAME_LEN = 1u << 2;       /* name or strand line longer than 255 bytes (README.md:129, Q18) */
constexpr u32 ERRBIT_COORD = 1u << 3;          /* X/Y >= 2^21 */
constexpr u32 ERRBIT_READ_LEN = 1u << 4;       /* read longer than 65535 (Q1) */
constexpr u32 ERRBIT_INTERNAL = 1u << 5;
constexpr u32 ERRBIT_RFQ = 1u << 6;            /* malformed stream on decode */
constexpr u32 INFOBIT_NEED_NAMES = 1u << 16;   /* not an error: some chunk stores per-read name1 / name2 / strand bytes */

/* one FASTQ image in HBM + its line index */
struct TextDev {
    const u8* text;
    u64 len;
    const u32* nl;     /* nl[j] = offset of the '\n' that ends line j (a virtual one at len (+crlf) if the file lacks it) */
    u32 n_lines;
    u32 crlf;          /* 1: every line break has two bytes ("\r\n") */
    /* a text with line breaks of both lengths is indexed, then copied line by line into a text of plain '\n' breaks that all later
     * kernels work on (k_canon_*); these are the original and its index, for offsets that are reported to the caller */
    const u8* otext;
    const u32* onl;
};

__device__ __forceinline__ u32 line_start(const TextDev& t, u32 j) { return j == 0 ? 0u : t.nl[j - 1] + 1u; }
__device__ __forceinline__ u32 line_end(const TextDev& t, u32 j) { return t.nl[j] - t.crlf; }
/* offset, in the text the caller passed, of the first break character after line j / of the last byte of that break */
__device__ __forceinline__ u32 caller_break_last(const TextDev& t, u32 j) { return t.onl ? t.onl[j] : t.nl[j]; }
__device__ __forceinline__ u32 caller_break_first(const TextDev& t, u32 j) {
    if (!t.onl) return t.nl[j] - t.crlf;
    const u32 e = t.onl[j], start = j ? t.onl[j - 1] + 1u : 0u;
    const bool two = t.otext[e] == '\n' && e > start && (t.otext[e - 1] == '\n' || t.otext[e - 1] == '\r');
    return e - (two ? 1u : 0u);
}

/* result of FastqMeta::parse for one read, 16 bytes */
struct ReadMeta {
    u32 x, y;
    u16 tile;
    u8 lane;
    u8 has;          /* hasLaneTileXY */
    u8 name1_len;    /* namePart1 = name[0, name1_len) */
    u8 name2_off;    /* namePart2 = name[name2_off, name_len) */
    u8 name_len;
    u8 strand_len;
};

/* per-chunk accumulator + layout */
struct ChunkDev {
    /* set by the cutter */
    u32 first;            /* first read (interleaved index) */
    u32 count;            /* reads */
    /* accumulated by k_meta (atomics) */
    u32 and_bits;         /* chunk-flag candidates, bit cleared when a read differs from read 0 */
    u32 fA, fB;           /* first odd read (chunk-relative) failing the name2 / the lane-tile-x-y pair check (Q10) */
    u32 last_odd_neq;     /* 1 + largest odd chunk-relative index whose name2 != read0.name2; 0 = none */
    u32 even_neq;         /* some even read has name2 != read0.name2 */
    /* decided by k_chunk_finish */
    u32 flags;            /* mFlags */
    u32 interleaved;
    u32 xy_num;
    u32 total_len;        /* sum of read lengths = quality positions */
    u32 seq_kept;         /* bases after overlap elision */
    u32 tot_n1, tot_n2, tot_strand;
    /* column sizes */
    u32 readlen_size, n1len_size, n2len_size, slen_size, lane_size, tile_size;
    u32 x_size, y_size, n1_size, n2_size, strand_size, seq_size, qual_size, ov_size, npos_size;
    u32 msize;
    /* column offsets inside the serialised chunk */
    u32 off_readlen, off_n1len, off_n2len, off_slen, off_lane, off_tile, off_x, off_y, off_n1, off_n2, off_strand, off_seq,
        off_qual, off_ov, off_npos;
    u32 bytes;            /* serialised size */
    u32 r1_end, r2_end;   /* text offsets of the first break character after the last record's quality line (Q13 compares them with nobreak_from) */
    u32 pad0;             /* explicit: the table is copied to the host word by word (no uninitialised padding) */
    u64 out_offset;
};

struct EncBatchDev {
    TextDev t[2];
    u32 is_pe;            /* reads alternate R1,R2 */
    u32 two_files;        /* mates come from t[0] / t[1] (else both from t[0], interleaved records) */
    u32 n_reads;
    u32 n_chunks;
    const u32* chunk_first;   /* [n_chunks+1] */
    u32* rlen;            /* [n_reads] */
    uint4* loc;           /* [n_reads] text offsets of the read's 4 line starts: name, sequence, strand, quality */
    u32* pk;              /* [n_reads][pkw] 2-bit packed bases of every read as stored (R2: forward); written by k_meta2 */
    u32* pk_rc;           /* [n_reads/2][pkw] 2-bit packed reverse complement of every R2 */
    u32 pkw;              /* words per read in pk / pk_rc */
    ReadMeta* meta;       /* [n_reads] */
    ReadMeta* meta0;      /* [n_chunks] FastqMeta of each chunk's first read */
    short* ov;            /* [n_reads/2] clamped overlap of each pair (valid when the chunk ends up interleaved) */
    u32* seqoff;          /* [n_reads] chunk-relative offset of the read's kept bases */
    u32* qualoff;         /* [n_reads] chunk-relative offset of the read's qualities */
    u32* n1off; u32* n2off; u32* soff;
    ChunkDev* chunks;
    u32* err;             /* error bits */
    u32 reach[2];         /* final batches: how far the reference's reader has read in each file when its loop ends (k_cut_ends; Q13) */
    u8* unclean;          /* [n_reads] 1: the read holds a character that is not a plain base (k_meta3); nullptr: not known */
    u32 uniform_reads_per_chunk;   /* != 0: chunk c = reads [c*u, (c+1)*u) */
};

/* which file / record a read lives in */
__device__ __forceinline__ void read_locus(const EncBatchDev& b, u32 i, u32& file, u32& rec) {
    if (b.is_pe && b.two_files) { file = i & 1u; rec = i >> 1; }
    else { file = 0; rec = i; }
}

__device__ __forceinline__ u32 chunk_of_read(const EncBatchDev& b, u32 i) {
    if (b.uniform_reads_per_chunk) { u32 c = i / b.uniform_reads_per_chunk; return c < b.n_chunks ? c : b.n_chunks - 1; }
    u32 lo = 0, hi = b.n_chunks;          /* largest c with chunk_first[c] <= i */
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (b.chunk_first[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

/* ---------------------------------------------------------------- small warp helpers ---- */
/* exact per-byte equality of two words, SIMD in a register: bit 7 of every byte of the result is set iff the bytes are equal
 * (no carries between bytes: (x & 0x7f) + 0x7f <= 0xfe) */
__device__ __forceinline__ u32 eq_bytes(u32 a, u32 b) {
    const u32 x = a ^ b;
    return ~((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x)) & 0x80808080u;
}

__device__ __forceinline__ u32 warp_excl_scan(u32 v, int lane, u32& total) {
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}
__device__ __forceinline__ u32 warp_max(u32 v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) { u32 t = __shfl_xor_sync(0xffffffffu, v, d); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ u32 warp_min(u32 v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) { u32 t = __shfl_xor_sync(0xffffffffu, v, d); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

/* src/read.cpp:92-113 */
__device__ __forceinline__ u8 complement_base(u8 b) {
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 'N';
    }
}
/* src/rfqcodec.cpp:593-599 */
__device__ __forceinline__ u32 base_code(u8 c) { return c == 'A' ? 1u : c == 'T' ? 2u : c == 'C' ? 3u : 0u; }

}  // namespace rpq
