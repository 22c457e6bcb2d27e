/*
 * rpq_compare.cuh - the check of compare mode: Repaq::compare / comparePE (reference src/repaq.cpp:36-233), which is also the
 * check behind `repaq -c ... -v` (completeCheckAndOutput, src/repaq.cpp:430-528).  The reference decodes chunk after chunk and
 * compares every decoded read with the next read of the FASTQ file, name, sequence, strand, quality, stopping at the first
 * difference.  Here the chunks are decoded by the decode kernels into FASTQ text in HBM, the FASTQ file(s) are indexed by the
 * encode path's line indexer, and one warp per read compares the four fields of both sides; the first difference in the
 * reference's order (read, then field) is the minimum of a 64-bit key.
 */
#pragma once
#include "rpq_decode4.cuh"
#include "rpq_index.cuh"

namespace rpq {

struct CmpDev {
    u64 first_diff;       /* min over reads of 4 * read + field (0 name, 1 sequence, 2 strand, 3 quality); ~0 = none */
    u64 rfq_bases;        /* k_cmp_sums: bases of the decoded reads [0, n) */
    u64 fq_bases;         /* bases of the FASTQ reads [0, n) */
    u32 loc_rfq[8];       /* k_cmp_locate: offset / length of the four fields of one decoded read (and the output stream in [8]) */
    u32 loc_fq[8];
    u64 rfq_base;         /* absolute offset of that decoded record inside its output stream */
    u32 rfq_stream, fq_file;
};

/* where the four fields of decoded read i lie: record text = name \n sequence \n strand \n quality \n (src/read.cpp:170-172) */
__device__ __forceinline__ void cmp_decoded_fields(const DecBatchDev& d, const HeaderDev& h, u32 i, u32& stream, u64& base, u32 off[4], u32 len[4]) {
    const u32 c = d.read_chunk[i];
    const DecChunk& ck = d.chunks[c];
    const u32 r = i - ck.read_base;
    stream = d.split_pairs ? (r & 1u) : 0u;
    base = ck.out_off[stream] + d.outoff[i];
    const u32 rl = d.rlen[i], ol = d.olen[i];
    const u8* in = d.body + ck.in_off;
    const u32 ls = (ck.flags & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
    const u32 nlen = ol - (2u * rl + ls + 4u);
    (void)h;
    off[0] = 0; len[0] = nlen;
    off[1] = nlen + 1; len[1] = rl;
    off[2] = off[1] + rl + 1; len[2] = ls;
    off[3] = off[2] + ls + 1; len[3] = rl;
}

/* the four fields of FASTQ read i as FastqReader::getLine delivers them (line break and '\r' stripped, src/fastqreader.cpp:94-156) */
__device__ __forceinline__ void cmp_fastq_fields(const EncBatchDev& e, u32 i, u32& file, u32 off[4], u32 len[4]) {
    u32 rec; read_locus(e, i, file, rec);
    const TextDev& t = e.t[file];
    const uint4 lc = e.loc[i];
    const u32 brk = 1u + t.crlf;
    off[0] = lc.x; len[0] = lc.y - brk - lc.x;
    off[1] = lc.y; len[1] = lc.z - brk - lc.y;
    off[2] = lc.z; len[2] = lc.w - brk - lc.z;
    off[3] = lc.w; len[3] = line_end(t, 4 * rec + 3) - lc.w;
}

constexpr int CMP_WARPS = 8;
/* warp per read */
__global__ void __launch_bounds__(32 * CMP_WARPS) k_compare(DecBatchDev d, HeaderDev h, EncBatchDev e, u32 n, CmpDev* res) {
    const int lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    u32 stream, file, ro[4], rl[4], fo[4], fl[4]; u64 base;
    cmp_decoded_fields(d, h, i, stream, base, ro, rl);
    cmp_fastq_fields(e, i, file, fo, fl);
    const u8* a = d.out[stream] + base;
    const u8* g = e.t[file].text;
    u32 field = 4;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        bool differ = rl[k] != fl[k];
        if (!differ) { for (u32 p = lane; p < rl[k]; p += 32) if (a[ro[k] + p] != g[fo[k] + p]) differ = true; }
        if (__any_sync(0xffffffffu, differ)) field = (u32)k;
    }
    if (field < 4 && lane == 0) atomicMin(&res->first_diff, 4ull * i + field);
}

/* bases of the first n decoded reads and of the first n FASTQ reads (the counters of the reference's report) */
__global__ void __launch_bounds__(256) k_cmp_sums(const u32* __restrict__ rfq_rlen, const u32* __restrict__ fq_rlen, u32 n_rfq, u32 n_fq, CmpDev* res) {
    u64 a = 0, g = 0;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n_rfq || k < n_fq; k += gridDim.x * blockDim.x) {
        if (k < n_rfq) a += rfq_rlen[k];
        if (k < n_fq) g += fq_rlen[k];
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); g += __shfl_xor_sync(0xffffffffu, g, s); }
    if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(&res->rfq_bases, a); if (g) atomicAdd(&res->fq_bases, g); }
}

/* field offsets of one read on both sides, for the message of a failed comparison */
__global__ void k_cmp_locate(DecBatchDev d, HeaderDev h, EncBatchDev e, u32 i, CmpDev* res) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 stream, file, ro[4], rl[4], fo[4], fl[4]; u64 base;
    cmp_decoded_fields(d, h, i, stream, base, ro, rl);
    cmp_fastq_fields(e, i, file, fo, fl);
    for (int k = 0; k < 4; k++) { res->loc_rfq[k] = ro[k]; res->loc_rfq[4 + k] = rl[k]; res->loc_fq[k] = fo[k]; res->loc_fq[4 + k] = fl[k]; }
    res->rfq_base = base; res->rfq_stream = stream; res->fq_file = file;
}

}  // namespace rpq
