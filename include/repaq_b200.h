/*
 * repaq_b200.h - C ABI of librepaq_b200.so: the FASTQ <-> .rfq chunk codec of OpenGene/repaq v0.5.1
 * (ALGORITHM_VER 2) as hand-written sm_100a CUDA kernels.  Output is byte-identical to the reference.
 *
 * The reference has no plugin/FFI layer; the seam this library replaces is the public surface of
 * `class RfqCodec` (reference src/rfqcodec.h:17-43), whose only callers are Repaq::compress / compressPE /
 * decompress / decompressPE / compare* (reference src/repaq.cpp:37,131,263,336,438,484,531,538,641,648).
 * INTEGRATION.md shows the few lines a maintainer of the reference adds to call it.
 *
 * Conventions: plain pointers and sizes only; 0 = success, negative = error (rpq_last_error() has the text; the
 * strings of errors the reference also raises are the reference's own error_exit() strings).  A context is bound
 * to one GPU and one host caller thread.  There is NO CPU fallback: without a CUDA device rpq_create() fails.
 */
#ifndef REPAQ_B200_H
#define REPAQ_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPQ_OK 0
#define RPQ_NO_RECORDS 1           /* rpq_make_header: the input holds no record; the reference then writes an EMPTY output file and
                                     exits 0 (src/repaq.cpp:530-638: nothing is flushed, not even the header) */
#define RPQ_ERR_CUDA (-1)          /* CUDA runtime / launch failure, or no device */
#define RPQ_ERR_ARG (-2)           /* bad argument / header not set */
#define RPQ_ERR_NOMEM (-3)
#define RPQ_ERR_FASTQ (-4)         /* input outside the supported FASTQ domain */
#define RPQ_ERR_COORD (-5)         /* "The X/Y coordinate cannot be larger than 2M" (src/rfqcodec.cpp:1316) */
#define RPQ_ERR_HEADER (-6)        /* bad .rfq header (src/rfqheader.cpp:23-25,40-42; src/rfqcodec.cpp:1065) */
#define RPQ_ERR_RFQ (-7)           /* an .rfq the decoder cannot take (too many reads for one call, an empty run-length column ...); columns that merely
                                      disagree with each other are decoded with the reference's tolerances (positions past the chunk ignored, sizes
                                      clamped to the chunk: SURVEY Q20), not rejected */
#define RPQ_ERR_QUALITY (-8)       /* header construction: bad quality / base characters (src/rfqheader.cpp:141,155-166,204) */

/* RfqHeader flag bits (src/rfqheader.h:24-42) */
#define RPQ_HAS_LANE (1 << 0)
#define RPQ_HAS_TILE (1 << 1)
#define RPQ_HAS_X (1 << 2)
#define RPQ_HAS_Y (1 << 3)
#define RPQ_HAS_NAME2 (1 << 4)
#define RPQ_PAIRED_END (1 << 5)
#define RPQ_ENCODE_PE_BY_OVERLAP (1 << 6)
#define RPQ_ENCODE_QUAL_BY_COL (1 << 7)
#define RPQ_DONT_ENCODE_QUAL (1 << 8)
#define RPQ_ENCODE_N_POS (1 << 9)

/* RfqChunk flag bits (src/rfqchunk.h:25-50) */
#define RPQ_READ_LEN_SAME (1 << 0)
#define RPQ_NAME1_LEN_SAME (1 << 1)
#define RPQ_NAME2_LEN_SAME (1 << 2)
#define RPQ_STRAND_LEN_SAME (1 << 3)
#define RPQ_LANE_SAME (1 << 4)
#define RPQ_TILE_SAME (1 << 5)
#define RPQ_NAME1_SAME (1 << 6)
#define RPQ_NAME2_SAME (1 << 7)
#define RPQ_STRAND_SAME (1 << 8)
#define RPQ_PE_INTERLEAVED (1 << 9)
#define RPQ_NO_LINE_BREAK_AT_END (1 << 10)
#define RPQ_NO_LINE_BREAK_AT_END_R2 (1 << 11)

#define RPQ_MEM_HOST 0
#define RPQ_MEM_DEVICE 1

/* POD mirror of RfqHeader (src/rfqheader.h:44-109): the serialised fields plus mSupportInterleaved. */
typedef struct rpq_header {
    uint8_t read_length_bytes;     /* mReadLengthBytes: 1 or 2 (4 can never be produced, src/rfqcodec.cpp:48-53) */
    uint16_t flags;                /* mFlags */
    uint8_t name2_diff_pos;        /* mName2DiffPos */
    uint8_t name2_diff_char;       /* mName2DiffChar */
    int8_t n_base_qual;            /* mNBaseQual (-1: N positions are stored explicitly) */
    int8_t overlap_shift;          /* mOverlapShift (-24) */
    uint8_t support_interleaved;   /* mSupportInterleaved: not serialised (src/rfqheader.cpp:84-97) */
    uint8_t qual_bins;             /* mQualBins */
    uint8_t qual_buf[128];         /* mQualBuf: major quality first, the rest ascending */
} rpq_header;

typedef struct rpq_ctx rpq_ctx;

/* ---- header: RfqCodec::makeHeader (src/rfqcodec.cpp:20-145) + RfqHeader::makeQualityTable (src/rfqheader.cpp:130-237).
 * Host code, as in the reference; it reads only the records of the FIRST chunk of the FASTQ image(s).
 * r2 == NULL: single end, or interleaved pairs in r1 when interleaved != 0. */
int rpq_make_header(const char* r1, uint64_t r1_len, const char* r2, uint64_t r2_len, int interleaved,
                    uint32_t chunk_bases, rpq_header* out, char* err, size_t err_cap);
/* RfqHeader::write (src/rfqheader.cpp:84-97); returns bytes written (17 + qual_bins) or 0 if cap is too small */
size_t rpq_header_write(const rpq_header* h, uint8_t* out, size_t cap);
/* RfqHeader::read (src/rfqheader.cpp:19-43); *consumed = header bytes */
int rpq_header_read(const uint8_t* in, size_t len, rpq_header* out, size_t* consumed, char* err, size_t err_cap);

/* ---- context */
int rpq_create(int device, rpq_ctx** out);
void rpq_destroy(rpq_ctx* ctx);
const char* rpq_last_error(const rpq_ctx* ctx);
/* RfqCodec::setHeader (src/rfqcodec.cpp:16-18) */
int rpq_set_header(rpq_ctx* ctx, const rpq_header* h);
/* the CUDA stream all work of this context is issued on (a cudaStream_t), for callers that time with events */
void* rpq_stream(rpq_ctx* ctx);

/* ---- encode: Repaq::compress/compressPE's chunk loop (src/repaq.cpp:546-553,656-663) + RfqCodec::encodeChunk
 * (src/rfqcodec.cpp:147-586) + RfqChunk::write (src/rfqchunk.cpp:230-312) for every chunk of a FASTQ batch. */
typedef struct rpq_encode_in {
    const char* r1; uint64_t r1_len;      /* FASTQ text, line breaks as the reference's reader takes them ('\n', '\r', "\r\n"); < 4 GiB per call */
    const char* r2; uint64_t r2_len;      /* mate file, or NULL */
    int mem;                              /* RPQ_MEM_HOST (pageable or pinned) or RPQ_MEM_DEVICE */
    int interleaved;                      /* r1 holds R1,R2,R1,R2,... (--interleaved_in) */
    uint32_t chunk_bases;                 /* Options::chunkSize = max(100,k)*1000 (src/main.cpp:69) */
    int final;                            /* 1: also flush the trailing partial chunk (end of input) */
    /* Q13 (src/fastqreader.cpp:31-46, src/repaq.cpp:571-572,683-692): a chunk gets NO_LINE_BREAK_AT_END{,_R2} when the
     * line break that ends its last record lies at text offset >= nobreak_from[0|1]; UINT64_MAX = never */
    uint64_t nobreak_from[2];
    uint16_t tail_flags;                  /* OR-ed into the trailing partial chunk only */
    int out_mem;                          /* where the serialised chunks are wanted: RPQ_MEM_HOST or RPQ_MEM_DEVICE */
    /* offsets of r1 / r2 inside their files (0: the text starts where the file starts).  The reference's reader refills a 1 MiB
     * buffer and does not take a '\n' that is the first or the last byte of a buffer as the second byte of a line break
     * (src/fastqreader.cpp:113-116): a "\r\n" there reads as a line end followed by an EMPTY line, where the reader ends its input
     * (:180-181).  The line index reproduces that, so it needs to know where the buffers lie. */
    uint64_t file_offset[2];
} rpq_encode_in;

/* scalar fields of RfqChunk (src/rfqchunk.h:52-113) for one encoded / indexed chunk */
typedef struct rpq_chunk_info {
    uint64_t offset;        /* of the serialised chunk in the stream */
    uint32_t bytes;         /* its serialised length (NOT mSize, which the reference computes wrongly: Q2) */
    uint32_t msize;         /* mSize as the reference writes it */
    uint32_t reads;         /* mReads */
    uint16_t flags;         /* mFlags */
    uint32_t seq_size, qual_size, npos_size, x_size, y_size;
    uint32_t name1_size, name2_size, strand_size;
    uint64_t r1_end, r2_end;   /* encode: text offset of the first line-break character after the chunk's last record (what Q13 compares
                                  with nobreak_from); to continue after a batch use rpq_encode_out.r?_consumed, not these */
    uint64_t out1_bytes, out2_bytes;   /* decode: FASTQ bytes this chunk decodes to (R1 / R2 stream) */
} rpq_chunk_info;

typedef struct rpq_encode_out {
    const uint8_t* data;      /* serialised chunks back to back, owned by ctx, valid until the next call on ctx */
    uint64_t bytes;
    uint32_t n_chunks;
    const rpq_chunk_info* chunks;   /* host memory, owned by ctx */
    uint64_t n_reads;         /* records encoded (both mates counted) */
    uint64_t r1_consumed, r2_consumed;   /* text bytes covered by the emitted chunks; feed the rest again */
} rpq_encode_out;

int rpq_encode(rpq_ctx* ctx, const rpq_encode_in* in, rpq_encode_out* out);

/* ---- decode: RfqChunk::read (src/rfqchunk.cpp:161-228) + RfqCodec::decodeChunk (src/rfqcodec.cpp:1049-1260) +
 * Read::toString (src/read.cpp:170-172) for every chunk of an .rfq body (the bytes after the file header). */
typedef struct rpq_decode_in {
    const uint8_t* data; uint64_t bytes;  /* whole chunks; trailing partial chunk is reported via consumed */
    int mem;                              /* RPQ_MEM_HOST or RPQ_MEM_DEVICE */
    int split_pairs;                      /* 1: reads alternate between two outputs (decompressPE), 0: one output */
    int out_mem;
} rpq_decode_in;

typedef struct rpq_decode_out {
    const char* out1; uint64_t out1_bytes;   /* FASTQ text, every record ends with '\n' */
    const char* out2; uint64_t out2_bytes;   /* only when split_pairs */
    uint32_t n_chunks;
    const rpq_chunk_info* chunks;            /* host memory: flags (for the trailing-newline rule) and out?_bytes */
    uint64_t n_reads;
    uint64_t consumed;                       /* bytes of whole chunks decoded */
} rpq_decode_out;

int rpq_decode(rpq_ctx* ctx, const rpq_decode_in* in, rpq_decode_out* out);

/* ---- compare: Repaq::compare / comparePE (src/repaq.cpp:36-233), also the check behind `-c ... -v` (completeCheckAndOutput,
 * src/repaq.cpp:430-528): every chunk of an .rfq body is decoded and its reads are checked one by one - name, sequence, strand,
 * quality - against the reads of the FASTQ file(s); the first difference ends the run.  Decoded text and FASTQ index stay in
 * device memory.  Large files are compared in batches: whole chunks on one side, a window of FASTQ text that holds at least as
 * many reads on the other (r?_consumed tells where the next window starts); the caller adds the counters up. */
#define RPQ_CMP_EQUAL 0          /* every decoded read equals its FASTQ read */
#define RPQ_CMP_NAME 1           /* "... have different name in the N read" (src/repaq.cpp:86) */
#define RPQ_CMP_SEQUENCE 2
#define RPQ_CMP_STRAND 3
#define RPQ_CMP_QUALITY 4
#define RPQ_CMP_RFQ_MORE 5       /* "The RFQ file has more reads than the FASTQ file." (:75); needs fq_final */
#define RPQ_CMP_FASTQ_MORE 6     /* "The FASTQ file has more reads than the RFQ file." (:120); needs rfq_final */
#define RPQ_CMP_NEED_FASTQ 7     /* the FASTQ window ended before the chunks did and fq_final is 0: call again with more text */

typedef struct rpq_compare_in {
    const uint8_t* rfq; uint64_t rfq_bytes;   /* whole chunks of the .rfq body (may be empty) */
    int rfq_mem;                              /* RPQ_MEM_HOST or RPQ_MEM_DEVICE */
    int rfq_final;                            /* 1: no chunks follow these */
    const char* r1; uint64_t r1_len;          /* FASTQ text the decoded reads are checked against, < 4 GiB per file and call */
    const char* r2; uint64_t r2_len;          /* mate file (comparePE), or NULL */
    int fq_mem;
    int fq_final;                             /* 1: the FASTQ text ends here */
    uint64_t fq_offset[2];                    /* offsets of r1 / r2 inside their files (see rpq_encode_in.file_offset) */
} rpq_compare_in;

typedef struct rpq_compare_out {
    int verdict;                              /* RPQ_CMP_* */
    /* the counters of reportCompareResult (src/repaq.cpp:235-259) for this call: reads and bases up to and including the read
     * that ended the run (both mates counted) */
    uint64_t fastq_reads, rfq_reads, fastq_bases, rfq_bases;
    uint64_t read_index;                      /* 0-based index (mates interleaved) of the read that ended the run */
    const char* rfq_field; uint32_t rfq_field_len;       /* verdicts 1..4: the two strings of the message, host memory owned */
    const char* fastq_field; uint32_t fastq_field_len;   /* by ctx, NUL terminated, valid until the next call */
    uint64_t r1_consumed, r2_consumed, rfq_consumed;     /* RPQ_CMP_EQUAL: text / body bytes covered by the reads compared */
} rpq_compare_out;

int rpq_compare(rpq_ctx* ctx, const rpq_compare_in* in, rpq_compare_out* out);

/* ---- page-locked host memory for callers that stream files through the library (the host<->device copies of rpq_encode /
 * rpq_decode / rpq_compare run at full link speed from such buffers, and overlap with kernels).  No counterpart in the reference,
 * whose reader refills a 1 MiB heap buffer (src/fastqreader.cpp:5,31-46). */
void* rpq_host_alloc(size_t bytes);
void rpq_host_free(void* p);

/* ---- instrumentation for bench.py: kernels launched and device milliseconds (CUDA events on the context's
 * stream) of the last rpq_encode / rpq_decode call, split by stage. */
typedef struct rpq_stats {
    uint32_t launches;
    float ms_total;          /* first kernel start to last kernel end, copies included if any */
    float ms_kernels;        /* same window with host<->device copies excluded */
    float ms_h2d, ms_d2h;
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t dec_walk;       /* rpq_decode of a device-resident body: 1 chain on mSize by one warp, 2 by several warps (k_dec_walk_par),
                                3 exact sequential walk (forced, or the fast chain did not check out); 0 host walk / no decode */
} rpq_stats;
int rpq_get_stats(const rpq_ctx* ctx, rpq_stats* out);
/* per-kernel device time: after rpq_set_profiling(ctx, 1) every launch is bracketed by a CUDA event pair;
 * rpq_get_profile() returns "kernel launches total_ms" lines accumulated since then (owned by ctx) */
int rpq_set_profiling(rpq_ctx* ctx, int on);
const char* rpq_get_profile(rpq_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
