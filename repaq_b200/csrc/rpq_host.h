/* rpq_host.h - host-only helpers (header construction / IO, chunk walk); see rpq_host.cpp */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/repaq_b200.h"

namespace rpq {

struct HostMeta { int name1_len, name2_off; uint8_t lane; uint16_t tile; uint32_t x, y; int has; };
HostMeta host_meta_parse(const char* name, int len);

int host_make_header(const char* r1, uint64_t l1, const char* r2, uint64_t l2, int interleaved, uint32_t chunk_bases,
                     rpq_header* h, char* err, size_t err_cap);
size_t host_header_write(const rpq_header* h, uint8_t* out, size_t cap);
int host_header_read(const uint8_t* in, size_t len, rpq_header* h, size_t* consumed, char* err, size_t err_cap);
int host_normal_bins(const rpq_header* h, uint8_t* out);

/* scalar fields + column offsets of one serialised chunk */
struct HostChunk {
    uint32_t msize, reads; uint16_t flags;
    uint32_t seq_size, qual_size, npos_size, x_size, y_size;
    uint32_t readlen_size, n1len_size, n2len_size, slen_size, lane_size, tile_size, n1_size, n2_size, strand_size, ov_size;
    uint32_t off_readlen, off_n1len, off_n2len, off_slen, off_lane, off_tile, off_x, off_y, off_n1, off_n2, off_strand, off_seq, off_qual, off_ov, off_npos;
    uint32_t bytes;
};
/* 0 ok, 1 truncated, 2 chunk with zero reads (end marker behaviour of the reference's decode loops) */
int host_walk_chunk(const rpq_header* h, const uint8_t* in, uint64_t len, HostChunk* c);

}  // namespace rpq
