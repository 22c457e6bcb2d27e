/*
 * rfq_oracle - TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the FASTQ <-> .rfq chunk codec of OpenGene/repaq v0.5.1
 * (ALGORITHM_VER 2), written from the behaviour of the reference sources; every function cites the
 * reference file:line it follows (paths relative to the reference checkout).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  Nothing under repaq_b200/ links, loads or calls it; the product path is the CUDA library and
 * fails loudly without it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement byte-for-byte against
 *  (1) the known-answer vectors of SURVEY.md section 8c (FastqMeta::test, KAT-names, KAT-SE, KAT-PE) and
 *  (2) .rfq files / decoded FASTQ produced by the unmodified reference binary (oracle/_ref/repaq, built by
 *      oracle/Makefile from the reference sources), committed under tests/golden/ with the generating script
 *      tests/golden/make_golden.py.
 */
#ifndef RFQ_ORACLE_H
#define RFQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* header flag bits: src/rfqheader.h:24-42 */
#define ORC_HAS_LANE (1 << 0)
#define ORC_HAS_TILE (1 << 1)
#define ORC_HAS_X (1 << 2)
#define ORC_HAS_Y (1 << 3)
#define ORC_HAS_NAME2 (1 << 4)
#define ORC_PAIRED_END (1 << 5)
#define ORC_ENCODE_PE_BY_OVERLAP (1 << 6)
#define ORC_ENCODE_QUAL_BY_COL (1 << 7)
#define ORC_DONT_ENCODE_QUAL (1 << 8)
#define ORC_ENCODE_N_POS (1 << 9)

/* chunk flag bits: src/rfqchunk.h:25-50 */
#define ORC_READ_LEN_SAME (1 << 0)
#define ORC_NAME1_LEN_SAME (1 << 1)
#define ORC_NAME2_LEN_SAME (1 << 2)
#define ORC_STRAND_LEN_SAME (1 << 3)
#define ORC_LANE_SAME (1 << 4)
#define ORC_TILE_SAME (1 << 5)
#define ORC_NAME1_SAME (1 << 6)
#define ORC_NAME2_SAME (1 << 7)
#define ORC_STRAND_SAME (1 << 8)
#define ORC_PE_INTERLEAVED (1 << 9)
#define ORC_NO_LINE_BREAK_AT_END (1 << 10)
#define ORC_NO_LINE_BREAK_AT_END_R2 (1 << 11)

/* mirror of RfqHeader's serialised + derived state: src/rfqheader.h:44-109 */
typedef struct {
    uint8_t read_length_bytes;
    uint16_t flags;
    uint8_t name2_diff_pos;
    char name2_diff_char;
    signed char n_base_qual;
    signed char overlap_shift;
    uint8_t support_interleaved; /* not serialised: src/rfqheader.cpp:84-97 */
    uint8_t qual_bins;
    uint8_t qual_buf[256];
} orc_header;

/* one FASTQ record; fields are views (not NUL terminated) */
typedef struct {
    const char* name;   uint32_t name_len;
    const char* seq;    uint32_t seq_len;
    const char* strand; uint32_t strand_len;
    const char* qual;   uint32_t qual_len;
} orc_read;

/* FastqMeta: src/fastqmeta.h:22-35 */
typedef struct {
    uint32_t name1_len;  /* name1 = name[0 .. name1_len) */
    uint32_t name2_off;  /* name2 = name[name2_off .. name_len) */
    uint32_t name2_len;
    uint8_t lane;
    uint16_t tile;
    uint32_t x, y;
    int has_lane_tile_xy;
} orc_meta;

const char* orc_last_error(void);

void orc_meta_parse(const char* name, uint32_t len, orc_meta* out);

/* FastqReader restated over an in-memory file image (keeps the 1 MiB refill cadence, Q13). */
typedef struct orc_reader orc_reader;
orc_reader* orc_reader_open(const char* text, size_t size);
void orc_reader_close(orc_reader*);
/* returns 1 and fills *r (views into reader-owned storage valid until close), 0 at end of input */
int orc_reader_next(orc_reader*, orc_read* r);
int orc_reader_no_line_break_at_end(const orc_reader*);

/* RfqCodec::makeHeader: reads are R1,R2,R1,R2,... when is_pe */
int orc_make_header(const orc_read* reads, size_t n, int is_pe, orc_header* h);
size_t orc_header_write(const orc_header* h, uint8_t* out /* >= 17+256 */);
/* returns bytes consumed, 0 on error */
size_t orc_header_read(const uint8_t* in, size_t len, orc_header* h);

/* RfqCodec::encodeChunk + RfqChunk::write: serialised chunk, malloc'd into *out. extra_flags are OR-ed into
 * mFlags after mSize is computed (the NO_LINE_BREAK bits, src/repaq.cpp:571-572). */
int orc_encode_chunk(const orc_header* h, const orc_read* reads, size_t n, int is_pe, uint16_t extra_flags,
                     uint8_t** out, size_t* out_len);

/* decoded chunk = n reads as 4 strings each */
typedef struct {
    size_t n_reads;
    uint16_t flags;
    char* text;          /* concatenated Read::toString() of all reads, in chunk order */
    size_t text_len;
    size_t* read_end;    /* read_end[i] = end offset of read i in text */
} orc_decoded;
void orc_decoded_free(orc_decoded*);
/* RfqChunk::read + RfqCodec::decodeChunk.  Returns bytes consumed (0 on error / end). */
size_t orc_decode_chunk(const orc_header* h, const uint8_t* in, size_t len, orc_decoded* out);

/* Repaq::compress / compressPE (src/repaq.cpp:530-759). r2 == NULL: single end (or interleaved_in when
 * interleaved != 0). chunk_bases = Options::chunkSize. */
int orc_compress(const char* r1, size_t l1, const char* r2, size_t l2, int interleaved, uint32_t chunk_bases,
                 uint8_t** out, size_t* out_len);
/* the chunks only, encoded with a given serialised file header (chunk-sharded encodes: every part uses the header of chunk 0) */
int orc_compress_with_header(const uint8_t* header, size_t header_len, const char* r1, size_t l1, const char* r2, size_t l2, int interleaved,
                             uint32_t chunk_bases, uint8_t** out, size_t* out_len);
/* Repaq::decompress (pe_out == 0) / decompressPE (pe_out != 0) (src/repaq.cpp:262-413) */
int orc_decompress(const uint8_t* rfq, size_t len, int pe_out, char** out1, size_t* l1, char** out2, size_t* l2);

/* Repaq::compare / comparePE (src/repaq.cpp:36-233) with the report of reportCompareResult (:235-259): *json is the
 * text the reference prints on stdout (malloc'd).  r2 == NULL: single end. */
int orc_compare(const uint8_t* rfq, size_t len, const char* r1, size_t l1, const char* r2, size_t l2, char** json);

void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
