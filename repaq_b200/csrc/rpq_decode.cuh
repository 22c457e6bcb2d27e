/*
 * rpq_decode.cuh - the decode kernels: RfqCodec::decodeChunk (reference src/rfqcodec.cpp:1049-1260) and
 * Read::toString (src/read.cpp:170-172) for all chunks of an .rfq body at once, writing FASTQ text directly.
 *
 *   k_dec_walk     chunk boundaries on the device (RfqChunk::read, src/rfqchunk.cpp:161-228; arena sizes :63-109)
 *   k_dec_coords   X / Y varint streams -> values                              (src/rfqcodec.cpp:1332-1389)
 *   k_dec_reads    read-length table (:1058-1086), per-read scans, output sizes of every record
 *   k_dec_streams  quality position streams + exceptions -> quality plane; N positions -> bitmap (:957-1047, :856-858)
 *   k_dec_format   one warp per read: name rebuild (:1157-1231), 2-bit unpack (:833-853), overlap expansion (:860-901),
 *                  N restore (:1093-1100), reverse complement of interleaved R2 (:1248-1252), record text
 */
#pragma once
#include "rpq_common.cuh"

namespace rpq {

struct DecChunk {
    u64 in_off;              /* chunk start inside the body */
    u32 reads, flags;
    u32 seq_size, qual_size, npos_size, x_size, y_size;
    u32 off_readlen, off_n1len, off_n2len, off_slen, off_lane, off_tile, off_x, off_y, off_n1, off_n2, off_strand, off_seq, off_qual, off_ov, off_npos;
    u32 bytes;
    u32 read_base;           /* reads before this chunk */
    /* device results */
    u32 total_len, seq_kept;
    u32 out_bytes[2];
    u64 plane_off;           /* offset of the chunk's quality plane (prefix of total_len) */
    u64 nmap_off;            /* offset (in u32 words) of the chunk's N bitmap */
    u64 out_off[2];
};

struct DecBatchDev {
    const u8* body;
    u64 body_len;
    DecChunk* chunks;
    u32 n_chunks;
    u32 n_reads;
    u32 split_pairs;
    u32* rlen; u32* qualoff; u32* seqoff; u32* n1off; u32* n2off; u32* soff;
    u64* outoff;             /* per read: offset inside its output stream */
    u32* olen;               /* per read: bytes of its record text */
    u32* read_chunk;         /* per read: its chunk */
    u32* xs; u32* ys;        /* [n_reads] decoded coordinates, indexed read_base + xy */
    u8* plane;               /* quality plane, all chunks */
    u32* nmap;               /* N bitmap over compacted positions */
    u8* out[2];
    u32* err;
    u64* totals;             /* [0] plane bytes, [1] nmap words, [2] out1 bytes, [3] out2 bytes, [4] = {u32 longest record text, u32 longest read} */
};

__device__ __forceinline__ u32 ld32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }

/* ------------------------------------------------------------------ chunk walk on the device ---- */
/* One warp.  Fills chunks[], *n_out, *consumed.  Same derivation as the host walk in rpq_host.cpp. */
__global__ void k_dec_walk(const u8* body, u64 len, HeaderDev h, DecChunk* chunks, u32 cap, u32* n_out, u64* consumed, u32* err) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    u64 at = 0; u32 n = 0, read_base = 0;
    const u32 head = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    while (at + head <= len && n < cap) {
        const u8* in = body + at;
        DecChunk c; memset(&c, 0, sizeof c);
        c.in_off = at; c.reads = ld32(in + 4); c.flags = (u32)in[8] | ((u32)in[9] << 8);
        c.seq_size = ld32(in + 10); c.qual_size = ld32(in + 14);
        if (h.flags & RPQ_ENCODE_N_POS) c.npos_size = ld32(in + 18);
        if (c.reads == 0) break;
        const u32 nr = c.reads, fl = c.flags;
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const u32 xy = il ? nr / 2 : nr;
        const u64 left = len - at;
        u64 o = head; bool bad = false;
        auto take = [&](u32& field, u64 size) { field = (u32)o; o += size; if (o > left) bad = true; };
        take(c.off_readlen, (u64)h.read_length_bytes * ((fl & RPQ_READ_LEN_SAME) ? 1u : nr));
        const u32 n1cnt = (fl & RPQ_NAME1_LEN_SAME) ? 1u : nr; take(c.off_n1len, n1cnt);
        u32 n2cnt = 0; if (h.flags & RPQ_HAS_NAME2) { n2cnt = (fl & RPQ_NAME2_LEN_SAME) ? 1u : nr; take(c.off_n2len, n2cnt); }
        const u32 scnt = (fl & RPQ_STRAND_LEN_SAME) ? 1u : nr; take(c.off_slen, scnt);
        if (h.flags & RPQ_HAS_LANE) take(c.off_lane, (fl & RPQ_LANE_SAME) ? 1u : xy);
        if (h.flags & RPQ_HAS_TILE) take(c.off_tile, 2ull * ((fl & RPQ_TILE_SAME) ? 1u : xy));
        if (!bad && (h.flags & RPQ_HAS_X)) { if (o + 4 > left) bad = true; else { c.x_size = ld32(in + o); o += 4; take(c.off_x, c.x_size); } }
        if (!bad && (h.flags & RPQ_HAS_Y)) { if (o + 4 > left) bad = true; else { c.y_size = ld32(in + o); o += 4; take(c.off_y, c.y_size); } }
        if (bad) break;
        auto arena = [&](u32 off, u32 cnt, bool len_same, bool all_same) -> u64 {
            u32 s = 0;
            for (u32 k = lane; k < cnt; k += 32) s += in[off + k];
            u64 t = warp_sum(s);
            if (len_same && !all_same) t *= nr;
            return t;
        };
        take(c.off_n1, arena(c.off_n1len, n1cnt, fl & RPQ_NAME1_LEN_SAME, fl & RPQ_NAME1_SAME));
        if (h.flags & RPQ_HAS_NAME2) take(c.off_n2, arena(c.off_n2len, n2cnt, fl & RPQ_NAME2_LEN_SAME, fl & RPQ_NAME2_SAME));
        take(c.off_strand, arena(c.off_slen, scnt, fl & RPQ_STRAND_LEN_SAME, fl & RPQ_STRAND_SAME));
        take(c.off_seq, c.seq_size);
        take(c.off_qual, c.qual_size);
        if (il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) take(c.off_ov, nr / 2);
        if (h.flags & RPQ_ENCODE_N_POS) take(c.off_npos, c.npos_size);
        if (bad) break;
        c.bytes = (u32)o; c.read_base = read_base;
        if (lane == 0) chunks[n] = c;
        n++; read_base += nr; at += o;
    }
    if (lane == 0) { n_out[0] = n; n_out[1] = read_base; *consumed = at; (void)err; }      /* chunks, reads, body bytes */
}

/* ------------------------------------------------------------------ coordinates ---- */
/* decodeCoords (src/rfqcodec.cpp:1332-1389): one thread per (chunk, column); the streams are ~1 byte per read */
__global__ void k_dec_coords(DecBatchDev b, HeaderDev h) {
    const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * b.n_chunks) return;
    const u32 c = id >> 1, col = id & 1;
    if (!(h.flags & (col ? RPQ_HAS_Y : RPQ_HAS_X))) return;
    const DecChunk& ck = b.chunks[c];
    const u8* buf = b.body + ck.in_off + (col ? ck.off_y : ck.off_x);
    const u32 len = col ? ck.y_size : ck.x_size;
    const u32 num = (ck.flags & RPQ_PE_INTERLEAVED) ? ck.reads / 2 : ck.reads;
    u32* data = (col ? b.ys : b.xs) + ck.read_base;
    u32 last = 1000, cns = 0, d = 0;
    while (cns < len) {
        const u32 b0 = buf[cns++];
        if (!(b0 & 0x80)) { const u32 v = (b0 << 8) | (cns < len ? buf[cns] : 0u); cns++; if (d < num) data[d] = v; d++; last = v; }
        else if (!(b0 & 0x40)) { const u32 v = last + (b0 & 0x3F) + 1; if (d < num) data[d] = v; d++; last = v; }
        else if (!(b0 & 0x20)) { const u32 rep = (b0 & 0x1F) + 1; for (u32 i = 0; i < rep; i++) { if (d < num) data[d] = last; d++; } }
        else { u32 v = (b0 & 0x1F) << 16; v |= (u32)(cns < len ? buf[cns] : 0u) << 8; cns++; v |= (cns < len ? buf[cns] : 0u); cns++; if (d < num) data[d] = v; d++; last = v; }
    }
    for (; d < num; d++) data[d] = 0;             /* memset(xBuf, 0): values the stream does not cover stay 0 */
}

/* ------------------------------------------------------------------ per-read tables ---- */
__device__ __forceinline__ u32 dec_digits(u32 v) {
    return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u : v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}

constexpr int DR_THREADS = 256;
struct Scan7 { u32 v[7]; };

__device__ __forceinline__ u32 dec_rlen(const DecBatchDev& b, const HeaderDev& h, const DecChunk& ck, u32 r) {
    const u8* p = b.body + ck.in_off + ck.off_readlen;
    const u32 k = (ck.flags & RPQ_READ_LEN_SAME) ? 0u : r;
    if (h.read_length_bytes == 1) return p[k];
    if (h.read_length_bytes == 2) return (u32)p[2 * k] | ((u32)p[2 * k + 1] << 8);
    return ld32(p + 4 * k);
}

/* one CTA per chunk: lengths, exclusive scans (quality, compacted bases, name parts, output text per stream).  Seven running sums
 * per read - but which of them can differ from read to read is a property of the chunk (its flags), so only those are scanned:
 * a NovaSeq chunk (same read length, same name parts and strand) scans three.  One barrier per 256 reads: the warp totals go
 * through a double-buffered table and every thread keeps the carries itself. */
__global__ void __launch_bounds__(DR_THREADS) k_dec_reads(DecBatchDev b, HeaderDev h) {
    constexpr int NW = DR_THREADS / 32;
    __shared__ u32 s_warp[2][NW][8];
    const u32 c = blockIdx.x;
    DecChunk& ck = b.chunks[c];
    const u8* in = b.body + ck.in_off;
    const u32 n = ck.reads, fl = ck.flags;
    const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
    const bool ov_on = il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 maxrec = 0, maxrl = 0;
    /* components: 0 read length, 1 kept bases, 2 name1, 3 name2, 4 strand, 5 / 6 record text of output stream 0 / 1 */
    u32 need = 1u << 5;
    if (!(fl & RPQ_READ_LEN_SAME)) need |= 1u << 0;
    if (ov_on) need |= 1u << 1;
    if (!(fl & RPQ_NAME1_SAME)) need |= 1u << 2;
    if ((h.flags & RPQ_HAS_NAME2) && !(fl & RPQ_NAME2_SAME)) need |= 1u << 3;
    if (!(fl & RPQ_STRAND_SAME)) need |= 1u << 4;
    if (b.split_pairs) need |= 1u << 6;
    const u32 rl_same = (fl & RPQ_READ_LEN_SAME) ? dec_rlen(b, h, ck, 0) : 0u;
    u32 carry[7];
#pragma unroll
    for (int k = 0; k < 7; k++) carry[k] = 0;
    u32 it = 0;
    for (u32 base = 0; base < n; base += DR_THREADS, it ^= 1u) {
        const u32 r = base + tid;
        u32 v[7];
#pragma unroll
        for (int k = 0; k < 7; k++) v[k] = 0;
        if (r < n) {
            const u32 rl = (fl & RPQ_READ_LEN_SAME) ? rl_same : dec_rlen(b, h, ck, r);
            u32 kept = rl;
            if (ov_on && (r & 1u)) { int o = (int)(signed char)in[ck.off_ov + (r >> 1)] - (int)h.overlap_shift; u32 a = (u32)(o < 0 ? -o : o); kept = a <= rl ? rl - a : 0; }
            const u32 l1 = (fl & (RPQ_NAME1_SAME | RPQ_NAME1_LEN_SAME)) ? in[ck.off_n1len] : in[ck.off_n1len + r];
            u32 l2 = 0;
            if (h.flags & RPQ_HAS_NAME2) l2 = (fl & (RPQ_NAME2_SAME | RPQ_NAME2_LEN_SAME)) ? in[ck.off_n2len] : in[ck.off_n2len + r];
            const u32 ls = (fl & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
            const u32 xy = il ? r >> 1 : r;
            u32 name = l1 + l2;
            if (h.flags & RPQ_HAS_LANE) name += 1 + dec_digits((fl & RPQ_LANE_SAME) ? in[ck.off_lane] : in[ck.off_lane + xy]);
            if (h.flags & RPQ_HAS_TILE) { const u32 k = (fl & RPQ_TILE_SAME) ? 0u : xy; name += 1 + dec_digits((u32)in[ck.off_tile + 2 * k] | ((u32)in[ck.off_tile + 2 * k + 1] << 8)); }
            if (h.flags & RPQ_HAS_X) name += 1 + dec_digits(b.xs[ck.read_base + xy]);
            if (h.flags & RPQ_HAS_Y) name += 1 + dec_digits(b.ys[ck.read_base + xy]);
            const u32 text = name + 1 + rl + 1 + ls + 1 + rl + 1;          /* Read::toString */
            v[0] = rl; v[1] = kept;
            v[2] = (fl & RPQ_NAME1_SAME) ? 0u : l1; v[3] = (fl & RPQ_NAME2_SAME) ? 0u : l2; v[4] = (fl & RPQ_STRAND_SAME) ? 0u : ls;
            const u32 stream = b.split_pairs ? (r & 1u) : 0u;
            v[5 + stream] = text;
            b.rlen[ck.read_base + r] = rl;
            b.olen[ck.read_base + r] = text;
            b.read_chunk[ck.read_base + r] = c;
            if (text > maxrec) maxrec = text;
            if (rl > maxrl) maxrl = rl;
        }
        u32 inc[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            inc[k] = v[k];
            if ((need >> k) & 1u) {                              /* uniform over the CTA */
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc[k], d); if (lane >= d) inc[k] += t; }
                if (lane == 31) s_warp[it][warp][k] = inc[k];
            }
        }
        __syncthreads();
        u32 pre[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            pre[k] = carry[k];
            if ((need >> k) & 1u) {
                u32 tot = 0;
#pragma unroll
                for (int q = 0; q < NW; q++) { const u32 t = s_warp[it][q][k]; if (q < warp) pre[k] += t; tot += t; }
                carry[k] += tot;
            }
        }
        if (r < n) {
            const u32 i = ck.read_base + r;
            const u32 qo = (need & 1u) ? pre[0] + inc[0] - v[0] : r * rl_same;
            b.qualoff[i] = qo;
            b.seqoff[i] = (need & 2u) ? pre[1] + inc[1] - v[1] : qo;
            b.n1off[i] = pre[2] + inc[2] - v[2];
            b.n2off[i] = pre[3] + inc[3] - v[3];
            b.soff[i] = pre[4] + inc[4] - v[4];
            const u32 stream = b.split_pairs ? (r & 1u) : 0u;
            b.outoff[i] = pre[5 + stream] + inc[5 + stream] - v[5 + stream];
        }
    }
    if (tid == 0) {
        const u32 total = (need & 1u) ? carry[0] : n * rl_same;
        ck.total_len = total; ck.seq_kept = (need & 2u) ? carry[1] : total; ck.out_bytes[0] = carry[5]; ck.out_bytes[1] = carry[6];
    }
    maxrec = warp_max(maxrec); maxrl = warp_max(maxrl);
    if (lane == 0) {
        u32* mx = reinterpret_cast<u32*>(b.totals + 4);
        if (maxrec > ((volatile u32*)mx)[0]) atomicMax(mx, maxrec);
        if (maxrl > ((volatile u32*)mx)[1]) atomicMax(mx + 1, maxrl);
    }
}

/* prefix over chunks: plane, N bitmap and output offsets; totals for the host */
__global__ void __launch_bounds__(256) k_dec_offsets(DecBatchDev b) {
    __shared__ u64 s_w[8][4];
    __shared__ u64 s_carry[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) s_carry[tid] = 0;
    __syncthreads();
    for (u32 base = 0; base < b.n_chunks; base += 256) {
        const u32 c = base + tid;
        u64 v[4] = {0, 0, 0, 0};
        if (c < b.n_chunks) { const DecChunk& ck = b.chunks[c]; v[0] = ck.total_len; v[1] = (ck.seq_kept + 31) / 32 + 1; v[2] = ck.out_bytes[0]; v[3] = ck.out_bytes[1]; }
        u64 inc[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) for (int k = 0; k < 4; k++) { const u64 t = __shfl_up_sync(0xffffffffu, inc[k], d); if (lane >= d) inc[k] += t; }
        if (lane == 31) for (int k = 0; k < 4; k++) s_w[warp][k] = inc[k];
        __syncthreads();
        u64 pre[4];
        for (int k = 0; k < 4; k++) { pre[k] = s_carry[k]; for (int q = 0; q < warp; q++) pre[k] += s_w[q][k]; }
        if (c < b.n_chunks) { DecChunk& ck = b.chunks[c]; ck.plane_off = pre[0] + inc[0] - v[0]; ck.nmap_off = pre[1] + inc[1] - v[1]; ck.out_off[0] = pre[2] + inc[2] - v[2]; ck.out_off[1] = pre[3] + inc[3] - v[3]; }
        __syncthreads();
        if (tid < 4) { u64 t = s_carry[tid]; for (int q = 0; q < 8; q++) t += s_w[q][tid]; s_carry[tid] = t; }
        __syncthreads();
    }
    if (tid < 4) b.totals[tid] = s_carry[tid];
}

/* ------------------------------------------------------------------ position streams ---- */
/*
 * decodeSingleQualByCol (src/rfqcodec.cpp:957-1007) with a warp per (chunk, stream): 32 stream bytes per step.
 * Token length depends on the first byte only (0xxxxxxx:1  10xxxxxx:2  110xxxxx:1  111xxxxx:4), so the heads of a
 * step follow from its entry offset by visiting just the bytes that look like multi-byte heads; every head's position
 * is an exclusive warp sum of the advances before it.
 * A warp per (chunk, stream); DS_WARPS chunks per CTA (same stream index: similar lengths), blockIdx.y = stream (quality bins, then
 * exceptions, then N positions).
 */
constexpr int DS_WARPS = 4;
/* stream indices by decreasing length in chunk `c0`: the CTAs of the longest stream are scheduled first (no long tail) */
__global__ void k_dec_stream_order(DecBatchDev b, HeaderDev h, u32 c0, u32 n_qstreams, u32 n_streams, u32* __restrict__ order) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const DecChunk& ck = b.chunks[c0];
    const u8* qcol = b.body + ck.in_off + ck.off_qual;
    u32 len[MAX_BINS + 3];
    u64 sum = 4ull * h.nb;
    for (u32 st = 0; st < n_streams; st++) {
        u32 l = 0;
        if (st < n_qstreams) {
            if (!(h.flags & RPQ_DONT_ENCODE_QUAL) && 4ull * h.nb <= ck.qual_size) {
                if (st < h.nb) { l = ld32(qcol + 4 * st); sum += l; }
                else l = ck.qual_size > sum ? (u32)(ck.qual_size - sum) : 0u;
            }
        } else l = ck.npos_size;
        len[st] = l;
    }
    for (u32 st = 0; st < n_streams; st++) {
        u32 rank = 0;
        for (u32 o = 0; o < n_streams; o++) if (len[o] > len[st] || (len[o] == len[st] && o < st)) rank++;
        order[rank] = st;
    }
}

__global__ void __launch_bounds__(32 * DS_WARPS) k_dec_streams(DecBatchDev b, HeaderDev h, u32 n_qstreams, u32 chunk_base, u32 chunk_end, const u32* __restrict__ order) {
    const u32 c = chunk_base + blockIdx.x * DS_WARPS + (threadIdx.x >> 5), st = order[blockIdx.y];
    if (c >= chunk_end) return;
    const int lane = threadIdx.x & 31;
    const DecChunk& ck = b.chunks[c];
    const u8* in = b.body + ck.in_off;
    const u8* stream; u32 slen; u8 q; bool is_npos = false;
    u32 dst_len;
    if (st < n_qstreams) {
        if (h.flags & RPQ_DONT_ENCODE_QUAL) return;
        const u8* qcol = in + ck.off_qual;
        u32 off = 4u * h.nb;
        if ((u64)off > ck.qual_size) return;
        for (u32 k = 0; k < st && k < h.nb; k++) off += ld32(qcol + 4 * k);
        if (st == h.nb) {
            /* exceptions: {q, u32 LE pos} until the end of the column (src/rfqcodec.cpp:1034-1043) */
            u8* plane = b.plane + ck.plane_off;
            for (u64 p = (u64)off + 5ull * lane; p + 5 <= ck.qual_size; p += 160) { const u32 pos = ld32(qcol + p + 1); if (pos < ck.total_len) plane[pos] = qcol[p]; }
            return;
        }
        stream = qcol + off; slen = ld32(qcol + 4 * st); q = h.normal_bins[st];
        if ((u64)off + slen > ck.qual_size) slen = ck.qual_size > off ? ck.qual_size - off : 0;
        dst_len = ck.total_len;
    } else {
        if (!(h.flags & RPQ_ENCODE_N_POS)) return;
        stream = in + ck.off_npos; slen = ck.npos_size; q = 'N'; is_npos = true;
        dst_len = ck.total_len;      /* the reference marks N in a buffer of seqLen bytes, compacted coordinates */
    }
    u8* plane = b.plane + ck.plane_off;
    u32* nmap = b.nmap + ck.nmap_off;
    const u32 nmap_bits = ((ck.seq_kept + 31) / 32 + 1) * 32;
    /* 128 stream bytes per step, four consecutive bytes per lane (aligned word loads, the next step's words already in
     * flight).  Where the tokens of a lane start depends on how many payload bytes spill in from the lane before (0..3): every
     * lane tabulates its exit spill for the four possible entries; most tables are constant, so the chain resolves in a round
     * or two.  Each lane then decodes its (at most four) tokens; one warp scan of the per-lane advances places them. */
    const uintptr_t sa = reinterpret_cast<uintptr_t>(stream);
    const u32* A = reinterpret_cast<const u32*>(sa & ~(uintptr_t)3);
    const u32 sh = 8u * (u32)(sa & 3u);
    const u32* Aend = reinterpret_cast<const u32*>((reinterpret_cast<uintptr_t>(b.body + b.body_len) + 3u) & ~(uintptr_t)3);
    auto ldw = [&](u32 k) -> u32 { const u32* w = A + k; return w < Aend ? *w : 0u; };
    const u32 lim_pos = is_npos ? (nmap_bits < dst_len ? nmap_bits : dst_len) : dst_len;      /* positions >= this are ignored (Q20) */
    u32 next = 0;                                         /* 1 + the position of the last element so far (`last` starts at -1); positions
                                                             of a chunk fit 32 bits, a corrupt stream that wraps stays below lim_pos */
    u32 skip = 0;                                         /* payload bytes at the start of the step that belong to the previous token */
    u32 cura = slen ? ldw((u32)lane) : 0u;
    for (u32 base = 0; base < slen; base += 128) {
        const u32 nexta = base + 128 < slen + 8 ? ldw((base >> 2) + 32u + (u32)lane) : 0u;
        u32 a1 = __shfl_down_sync(0xffffffffu, cura, 1), a2 = __shfl_down_sync(0xffffffffu, cura, 2);
        const u32 n0 = __shfl_sync(0xffffffffu, nexta, 0), n1 = __shfl_sync(0xffffffffu, nexta, 1);
        if (lane == 31) { a1 = n0; a2 = n1; } else if (lane == 30) a2 = n0;
        u64 B = (u64)__funnelshift_r(cura, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
        const u32 p0 = base + 4u * (u32)lane;
        const u32 left = p0 < slen ? slen - p0 : 0u;       /* stream bytes from this lane's first byte on */
        if (left < 8u) B = left ? B & ((1ull << (8u * left)) - 1ull) : 0ull;
        const u32 nv = left < 4u ? left : 4u;
        const u32 cur = (u32)B;
        /* token length by first byte: 0xxxxxxx 1, 10xxxxxx 2, 110xxxxx 1, 111xxxxx 4 */
        u32 L[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const u32 b0 = (cur >> (8 * k)) & 0xFFu; L[k] = !(b0 & 0x80u) ? 1u : !(b0 & 0x40u) ? 2u : !(b0 & 0x20u) ? 1u : 4u; }
        /* exit spill if the first token of the lane starts at byte r */
        const u32 e3 = L[3] - 1u;
        const u32 e2 = L[2] == 1u ? e3 : L[2] - 2u;
        const u32 e1 = L[1] == 1u ? e2 : (L[1] == 2u ? e3 : 1u);
        const u32 e0 = L[0] == 1u ? e1 : (L[0] == 2u ? e2 : 0u);
        const u32 f = e0 | (e1 << 2) | (e2 << 4) | (e3 << 6);
        const bool is_const = f == e0 * 0x55u;
        u32 r_in = lane == 0 ? skip : 4u;                    /* 4 = not known yet */
        for (;;) {
            const u32 mine = r_in < 4u ? (f >> (2u * r_in)) & 3u : (is_const ? e0 : 4u);
            const u32 got = __shfl_up_sync(0xffffffffu, mine, 1);
            if (r_in == 4u && lane > 0) r_in = got;
            if (__ballot_sync(0xffffffffu, r_in == 4u) == 0u) break;
        }
        skip = __shfl_sync(0xffffffffu, (f >> (2u * r_in)) & 3u, 31);
        /* the lane's tokens */
        u32 adv[4], run[4]; u32 lane_adv = 0; u32 next_head = r_in;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            adv[k] = 0; run[k] = 0;
            if ((u32)k == next_head && (u32)k < nv) {
                const u32 t = (u32)(B >> (8 * k));            /* the token's bytes, first byte lowest */
                const u32 b0 = t & 0xFFu;
                if (!(b0 & 0x80u)) adv[k] = b0 + 1u;
                else if (!(b0 & 0x40u)) adv[k] = (((b0 & 0x3Fu) << 8) | ((t >> 8) & 0xFFu)) + 1u;
                else if (!(b0 & 0x20u)) { run[k] = (b0 & 0x1Fu) + 1u; adv[k] = run[k]; }
                else adv[k] = (((b0 & 0x1Fu) << 24) | (((t >> 8) & 0xFFu) << 16) | (((t >> 16) & 0xFFu) << 8) | ((t >> 24) & 0xFFu)) + 1u;
                lane_adv += adv[k];
                next_head = (u32)k + L[k];
            }
        }
        u32 tot; const u32 ex = warp_excl_scan(lane_adv, lane, tot);
        u32 acc = next + ex;                                            /* 1 + the position before this lane's first token */
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (adv[k]) {
                const u32 end1 = acc + adv[k];                          /* 1 + the position of the token's last element */
                acc = end1;
                const u32 first = run[k] ? end1 - run[k] : end1 - 1u;
                const u32 stop = end1 < lim_pos ? end1 : lim_pos;
                if (first < stop) {
                    const u32 n = stop - first;                         /* 1..32 positions */
                    if (is_npos) { for (u32 j = 0; j < n; j++) { const u32 pos = first + j; atomicOr(&nmap[pos >> 5], 1u << (pos & 31)); } }
                    else {
                        u8* d = plane + first;
                        d[0] = q; if (n > 1) d[1] = q; if (n > 2) d[2] = q; if (n > 3) d[3] = q;
                        if (n > 4u) {                                  /* the rest of a long run: bytes up to a word boundary, words, bytes */
                            u8* w = d + 4; u8* const we = d + n;
                            while (w < we && (reinterpret_cast<uintptr_t>(w) & 3u)) *w++ = q;
                            const u32 q4 = 0x01010101u * q;
                            for (; w + 4 <= we; w += 4) *reinterpret_cast<u32*>(w) = q4;
                            while (w < we) *w++ = q;
                        }
                    }
                }
            }
        }
        next += tot;
        cura = nexta;
    }
}

/*
 * decodeQualByRunLenCoding (src/rfqcodec.cpp:919-955): the coder no header made under ALGORITHM_VER 2 selects (Q7), kept so that a
 * header with neither DONT_ENCODE_QUAL nor ENCODE_QUAL_BY_COL decodes as the reference decodes it.  Every column byte is a run:
 * bit 0 clear = the major quality, run length in the upper 7 bits; else code = low (8 - nq_bits) bits, run in the upper nq_bits;
 * lengths count from 1.  A column that runs out before the chunk's positions are filled is walked again from its first byte.
 * One CTA per chunk: total run length, then a scan of the runs and the fills.  Works on the plane of the long-read path.
 */
__global__ void __launch_bounds__(256) k_dec_rle(DecBatchDev b, HeaderDev h) {
    __shared__ u32 s_w[8];
    __shared__ u32 s_carry;
    const DecChunk& ck = b.chunks[blockIdx.x];
    const u8* col = b.body + ck.in_off + ck.off_qual;
    const u32 n = ck.qual_size, len = ck.total_len;
    u8* plane = b.plane + ck.plane_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!len) return;
    if (!n) { if (tid == 0) atomicOr(b.err, ERRBIT_RFQ); return; }       /* the reference never returns from this one */
    const u32 nq_mask = (1u << (8u - h.rle_nq_bits)) - 1u;
    auto run_of = [&](u32 e) -> u32 { return ((e & 1u) ? (e >> (8u - h.rle_nq_bits)) : (e >> 1)) + 1u; };
    /* ---- total positions of one walk over the column */
    u32 mine = 0;
    for (u32 k = tid; k < n; k += 256) mine += run_of(col[k]);
    mine = warp_sum(mine);
    if (lane == 0) s_w[warp] = mine;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    u32 T = 0;
    for (int q = 0; q < 8; q++) T += s_w[q];
    __syncthreads();
    /* ---- runs -> positions */
    for (u32 base = 0; base < n; base += 256) {
        const u32 k = base + (u32)tid;
        const u32 e = k < n ? col[k] : 0u;
        const u32 run = k < n ? run_of(e) : 0u;
        u32 wt; const u32 ex = warp_excl_scan(run, lane, wt);
        if (lane == 0) s_w[warp] = wt;
        __syncthreads();
        u32 start = s_carry + ex;
        for (int q = 0; q < warp; q++) start += s_w[q];
        if (run) {
            const u8 v = h.rle_b2q[(e & 1u) ? (e & nq_mask) & 127u : 0u];
            for (u64 p0 = start; p0 < len; p0 += T)                        /* this walk of the column, and every later one */
                for (u32 f = 0; f < run && p0 + f < len; f++) plane[p0 + f] = v;
        }
        __syncthreads();
        if (tid == 0) { u32 t = s_carry; for (int q = 0; q < 8; q++) t += s_w[q]; s_carry = t; }
        __syncthreads();
        if (s_carry >= len) break;                                         /* like the reference: stop once len positions are decoded */
    }
}

/* ------------------------------------------------------------------ record formatter ---- */
constexpr int FMT_WARPS = 8;

__device__ __forceinline__ u32 put_dec(u8* p, u32 v) {
    const u32 n = dec_digits(v);
    for (int i = (int)n - 1; i >= 0; i--) { p[i] = (u8)('0' + v % 10u); v /= 10u; }
    return n;
}

__global__ void __launch_bounds__(32 * FMT_WARPS) k_dec_format(DecBatchDev b, HeaderDev h, const u32* __restrict__ read_chunk_hint) {
    __shared__ u8 s_num[FMT_WARPS][48];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 i = blockIdx.x * FMT_WARPS + w;
    if (i >= b.n_reads) return;
    /* chunk of read i: binary search over read_base */
    u32 lo = 0, hi = b.n_chunks;
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (b.chunks[mid].read_base <= i) lo = mid; else hi = mid; }
    (void)read_chunk_hint;
    const DecChunk& ck = b.chunks[lo];
    const u8* in = b.body + ck.in_off;
    const u32 r = i - ck.read_base, fl = ck.flags;
    const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
    const bool ov_on = il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP);
    const bool odd = (r & 1u) != 0;
    const u32 stream = b.split_pairs ? (r & 1u) : 0u;
    u8* out = b.out[stream] + ck.out_off[stream] + b.outoff[i];
    const u32 rl = b.rlen[i];
    const u32 xy = il ? r >> 1 : r;

    /* ---- name */
    const u32 l1 = (fl & (RPQ_NAME1_SAME | RPQ_NAME1_LEN_SAME)) ? in[ck.off_n1len] : in[ck.off_n1len + r];
    const u8* n1 = in + ck.off_n1 + ((fl & RPQ_NAME1_SAME) ? 0u : b.n1off[i]);
    for (u32 k = lane; k < l1; k += 32) out[k] = n1[k];
    u32 w_at = l1;
    u32 numlen = 0;
    if (lane == 0) {
        u8* p = s_num[w];
        if (h.flags & RPQ_HAS_LANE) { p[numlen++] = ':'; numlen += put_dec(p + numlen, (fl & RPQ_LANE_SAME) ? in[ck.off_lane] : in[ck.off_lane + xy]); }
        if (h.flags & RPQ_HAS_TILE) { const u32 k = (fl & RPQ_TILE_SAME) ? 0u : xy; p[numlen++] = ':'; numlen += put_dec(p + numlen, (u32)in[ck.off_tile + 2 * k] | ((u32)in[ck.off_tile + 2 * k + 1] << 8)); }
        if (h.flags & RPQ_HAS_X) { p[numlen++] = ':'; numlen += put_dec(p + numlen, b.xs[ck.read_base + xy]); }
        if (h.flags & RPQ_HAS_Y) { p[numlen++] = ':'; numlen += put_dec(p + numlen, b.ys[ck.read_base + xy]); }
    }
    numlen = __shfl_sync(0xffffffffu, numlen, 0);
    __syncwarp();
    for (u32 k = lane; k < numlen; k += 32) out[w_at + k] = s_num[w][k];
    w_at += numlen;
    if (h.flags & RPQ_HAS_NAME2) {
        const u32 l2 = (fl & (RPQ_NAME2_SAME | RPQ_NAME2_LEN_SAME)) ? in[ck.off_n2len] : in[ck.off_n2len + r];
        const u8* n2 = in + ck.off_n2 + ((fl & RPQ_NAME2_SAME) ? 0u : b.n2off[i]);
        const bool subst = (fl & RPQ_NAME2_SAME) && il && odd && h.name2_diff_char != 0;
        for (u32 k = lane; k < l2; k += 32) { u8 ch = n2[k]; if (subst && k == h.name2_diff_pos) ch = h.name2_diff_char; out[w_at + k] = ch; }
        w_at += l2;
    }
    if (lane == 0) out[w_at] = '\n';
    w_at += 1;

    /* ---- sequence */
    const u8* seqb = in + ck.off_seq;
    const u8* plane = (h.flags & RPQ_DONT_ENCODE_QUAL) ? (in + ck.off_qual) : (b.plane + ck.plane_off);
    const u32 plane_len = (h.flags & RPQ_DONT_ENCODE_QUAL) ? (ck.qual_size < ck.total_len ? ck.qual_size : ck.total_len) : ck.total_len;
    const u32* nmap = b.nmap + ck.nmap_off;
    const u32 so = b.seqoff[i], qo = b.qualoff[i];
    int o = 0; u32 prev_rl = 0;
    if (ov_on && odd) { o = (int)(signed char)in[ck.off_ov + (r >> 1)] - (int)h.overlap_shift; prev_rl = b.rlen[i - 1]; }
    const bool rc = il && odd;
    const bool npos_mode = (h.flags & RPQ_ENCODE_N_POS) != 0;
    const u8 nq = (u8)h.n_base_qual;
    const u32 unpacked = ck.seq_size * 4u < ck.total_len ? ck.seq_size * 4u : ck.total_len;   /* decoded bases; the rest of the buffer stays 'N' */
    for (u32 jo = lane; jo < rl; jo += 32) {
        const u32 j = rc ? rl - 1 - jo : jo;            /* position inside the read before the final reverse complement */
        long long ci;
        if (o == 0) ci = (long long)so + j;
        else if (o > 0) ci = j < (u32)o ? (long long)so - o + j : (long long)so + j - o;
        else { const u32 k = rl - (u32)(-o); ci = j < k ? (long long)so + j : (long long)so - prev_rl + (j - k); }
        u8 base = 'N';
        if (ci >= 0 && (u64)ci < unpacked) {
            const u32 code = (seqb[ci >> 2] >> (2 * (ci & 3))) & 3u;
            base = code == 0 ? 'G' : code == 1 ? 'A' : code == 2 ? 'T' : 'C';
        }
        if (npos_mode) { if (ci >= 0 && (u64)ci < ck.total_len && ((nmap[ci >> 5] >> (ci & 31)) & 1u)) base = 'N'; }
        else { const u32 qp = qo + j; if (qp < plane_len ? plane[qp] == nq : h.major == nq) base = 'N'; }
        out[w_at + jo] = rc ? complement_base(base) : base;
    }
    if (lane == 0) out[w_at + rl] = '\n';
    w_at += rl + 1;

    /* ---- strand */
    const u32 ls = (fl & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
    const u8* sp = in + ck.off_strand + ((fl & RPQ_STRAND_SAME) ? 0u : b.soff[i]);
    for (u32 k = lane; k < ls; k += 32) out[w_at + k] = sp[k];
    if (lane == 0) out[w_at + ls] = '\n';
    w_at += ls + 1;

    /* ---- quality */
    for (u32 jo = lane; jo < rl; jo += 32) {
        const u32 j = rc ? rl - 1 - jo : jo;
        const u32 qp = qo + j;
        out[w_at + jo] = qp < plane_len ? plane[qp] : h.major;
    }
    if (lane == 0) out[w_at + rl] = '\n';
}

}  // namespace rpq
