/*
 * rpq_decode2.cuh - k_dec_format2: the record formatter, second generation: one THREAD per read, quality plane of the
 * CTA's reads staged in shared memory, records assembled in shared memory and written out with aligned 16-byte stores
 * (v1 k_dec_format used a warp per read and byte stores to global memory: 1180 warp instructions per read,
 * profiles/r01_v1_ncu_full_k_dec_format.csv).
 *
 * Same semantics as k_dec_format: name rebuild (reference src/rfqcodec.cpp:1157-1231), 2-bit unpack (:833-853),
 * overlap expansion (:860-901), N restore (:856-858 or :1093-1100), reverse complement of interleaved R2 (:1248-1252),
 * Read::toString (src/read.cpp:170-172).
 */
#pragma once
#include "rpq_decode.cuh"

namespace rpq {

struct Fmt2Cfg {
    u32 reads_per_cta;     /* = blockDim.x */
    u32 plane_cap;         /* shared bytes for the staged quality plane (incl. 16 bytes of alignment slack) */
    u32 out_cap;           /* shared bytes per output stream staging (incl. 16 bytes of alignment slack) */
};

__global__ void __launch_bounds__(128) k_dec_format2(DecBatchDev b, HeaderDev h, Fmt2Cfg cfg) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u64 s_start[2], s_end[2];          /* absolute byte range of the CTA's records in each output stream */
    __shared__ u64 s_q0, s_q1;                    /* absolute byte range of the CTA's reads in the quality plane */
    const int tid = threadIdx.x;
    const u32 G = cfg.reads_per_cta;
    const u32 i0 = blockIdx.x * G;
    const u32 n_here = b.n_reads - i0 < G ? b.n_reads - i0 : G;
    u8* s_plane = dyn;
    u8* s_out[2] = {dyn + cfg.plane_cap, dyn + cfg.plane_cap + cfg.out_cap};
    const u32 nstreams = b.split_pairs ? 2u : 1u;

    /* ---- ranges */
    if (tid < 2) { s_start[tid] = ~0ull; s_end[tid] = 0; }
    __syncthreads();
    const bool active = tid < (int)n_here;
    const u32 i = i0 + tid;
    u32 c = 0, r = 0, rl = 0, stream = 0, olen = 0;
    u64 oabs = 0, qabs = 0;
    if (active) {
        c = b.read_chunk[i];
        const DecChunk& ck = b.chunks[c];
        r = i - ck.read_base; rl = b.rlen[i]; olen = b.olen[i];
        stream = b.split_pairs ? (r & 1u) : 0u;
        oabs = ck.out_off[stream] + b.outoff[i];
        qabs = ck.plane_off + b.qualoff[i];
        /* the first / last read of each stream inside the CTA delimit the range (records are contiguous in read order) */
        if ((u32)tid < nstreams) s_start[stream] = oabs;
        if ((u32)tid + nstreams >= n_here) s_end[stream] = oabs + olen;
        if (tid == 0) s_q0 = qabs;
        if ((u32)tid == n_here - 1) s_q1 = qabs + rl;
    }
    __syncthreads();
    const bool raw_qual = (h.flags & RPQ_DONT_ENCODE_QUAL) != 0;       /* host never selects this kernel then */
    (void)raw_qual;
    /* ---- stage the quality plane [q0, q1): aligned 16-byte loads, shared offset = absolute offset - (q0 & ~15) */
    const u64 q0 = s_q0, q1 = s_q1, qa = q0 & ~15ull;
    {
        const u32 nvec = (u32)((q1 - qa + 15) >> 4);
        const uint4* src = reinterpret_cast<const uint4*>(b.plane + qa);
        uint4* dst = reinterpret_cast<uint4*>(s_plane);
        for (u32 k = tid; k < nvec; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    if (active) {
        const DecChunk& ck = b.chunks[c];
        const u8* in = b.body + ck.in_off;
        const u32 fl = ck.flags;
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const bool ov_on = il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP);
        const bool odd = (r & 1u) != 0;
        const u32 xy = il ? r >> 1 : r;
        u8* o = s_out[stream] + (u32)(oabs - (s_start[stream] & ~15ull));
        const u8* q = s_plane + (u32)(qabs - qa);
        u32 w_at = 0;
        /* name */
        const u32 l1 = (fl & (RPQ_NAME1_SAME | RPQ_NAME1_LEN_SAME)) ? in[ck.off_n1len] : in[ck.off_n1len + r];
        const u8* n1 = in + ck.off_n1 + ((fl & RPQ_NAME1_SAME) ? 0u : b.n1off[i]);
        for (u32 k = 0; k < l1; k++) o[w_at + k] = n1[k];
        w_at += l1;
        if (h.flags & RPQ_HAS_LANE) { o[w_at++] = ':'; w_at += put_dec(o + w_at, (fl & RPQ_LANE_SAME) ? in[ck.off_lane] : in[ck.off_lane + xy]); }
        if (h.flags & RPQ_HAS_TILE) { const u32 k = (fl & RPQ_TILE_SAME) ? 0u : xy; o[w_at++] = ':'; w_at += put_dec(o + w_at, (u32)in[ck.off_tile + 2 * k] | ((u32)in[ck.off_tile + 2 * k + 1] << 8)); }
        if (h.flags & RPQ_HAS_X) { o[w_at++] = ':'; w_at += put_dec(o + w_at, b.xs[ck.read_base + xy]); }
        if (h.flags & RPQ_HAS_Y) { o[w_at++] = ':'; w_at += put_dec(o + w_at, b.ys[ck.read_base + xy]); }
        if (h.flags & RPQ_HAS_NAME2) {
            const u32 l2 = (fl & (RPQ_NAME2_SAME | RPQ_NAME2_LEN_SAME)) ? in[ck.off_n2len] : in[ck.off_n2len + r];
            const u8* n2 = in + ck.off_n2 + ((fl & RPQ_NAME2_SAME) ? 0u : b.n2off[i]);
            for (u32 k = 0; k < l2; k++) o[w_at + k] = n2[k];
            if ((fl & RPQ_NAME2_SAME) && il && odd && h.name2_diff_char != 0 && h.name2_diff_pos < l2) o[w_at + h.name2_diff_pos] = h.name2_diff_char;
            w_at += l2;
        }
        o[w_at++] = '\n';
        /* sequence */
        const u8* seqb = in + ck.off_seq;
        const u32* nmap = b.nmap + ck.nmap_off;
        const u32 so = b.seqoff[i];
        int ov = 0; u32 prev_rl = 0;
        if (ov_on && odd) { ov = (int)(signed char)in[ck.off_ov + (r >> 1)] - (int)h.overlap_shift; prev_rl = b.rlen[i - 1]; }
        const bool rc = il && odd;
        const bool npos_mode = (h.flags & RPQ_ENCODE_N_POS) != 0;
        const u8 nq = (u8)h.n_base_qual;
        const u32 unpacked = ck.seq_size * 4u < ck.total_len ? ck.seq_size * 4u : ck.total_len;
        long long cur_byte = -1; u32 cur_val = 0;          /* cache of the packed byte in use */
        long long cur_nw = -1; u32 cur_nv = 0;
        for (u32 jo = 0; jo < rl; jo++) {
            const u32 j = rc ? rl - 1 - jo : jo;
            long long ci;
            if (ov == 0) ci = (long long)so + j;
            else if (ov > 0) ci = j < (u32)ov ? (long long)so - ov + j : (long long)so + j - ov;
            else { const u32 k = rl - (u32)(-ov); ci = j < k ? (long long)so + j : (long long)so - prev_rl + (j - k); }
            u8 base = 'N';
            if (ci >= 0 && (u64)ci < unpacked) {
                const long long bi = ci >> 2;
                if (bi != cur_byte) { cur_byte = bi; cur_val = seqb[bi]; }
                const u32 code = (cur_val >> (2 * (ci & 3))) & 3u;
                base = code == 0 ? 'G' : code == 1 ? 'A' : code == 2 ? 'T' : 'C';
            }
            if (npos_mode) {
                if (ci >= 0 && (u64)ci < ck.total_len) { const long long wi = ci >> 5; if (wi != cur_nw) { cur_nw = wi; cur_nv = nmap[wi]; } if ((cur_nv >> (ci & 31)) & 1u) base = 'N'; }
            } else if (q[j] == nq) base = 'N';
            o[w_at + jo] = rc ? complement_base(base) : base;
        }
        w_at += rl;
        o[w_at++] = '\n';
        /* strand */
        const u32 ls = (fl & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
        const u8* sp = in + ck.off_strand + ((fl & RPQ_STRAND_SAME) ? 0u : b.soff[i]);
        for (u32 k = 0; k < ls; k++) o[w_at + k] = sp[k];
        w_at += ls;
        o[w_at++] = '\n';
        /* quality */
        if (rc) for (u32 jo = 0; jo < rl; jo++) o[w_at + jo] = q[rl - 1 - jo];
        else for (u32 jo = 0; jo < rl; jo++) o[w_at + jo] = q[jo];
        w_at += rl;
        o[w_at] = '\n';
    }
    __syncthreads();
    /* ---- records to global memory: shared offset = absolute offset - (start & ~15), so 16-byte pieces line up */
    for (u32 s = 0; s < nstreams; s++) {
        const u64 a = s_start[s], e = s_end[s];
        if (a == ~0ull || e <= a) continue;
        const u64 base = a & ~15ull;
        u8* g = b.out[s];
        const u8* sm = s_out[s];
        /* head bytes up to the first 16-byte boundary, whole vectors, tail bytes */
        const u64 v0 = (a + 15) & ~15ull, v1 = e & ~15ull;
        if (v0 >= v1) { for (u64 p = a + tid; p < e; p += blockDim.x) g[p] = sm[p - base]; continue; }
        for (u64 p = a + tid; p < v0; p += blockDim.x) g[p] = sm[p - base];
        const u32 nvec = (u32)((v1 - v0) >> 4);
        uint4* gd = reinterpret_cast<uint4*>(g + v0);
        const uint4* sd = reinterpret_cast<const uint4*>(sm + (v0 - base));
        for (u32 k = tid; k < nvec; k += blockDim.x) gd[k] = sd[k];
        for (u64 p = v1 + tid; p < e; p += blockDim.x) g[p] = sm[p - base];
    }
}

}  // namespace rpq

namespace rpq {

/*
 * k_dec_coords2: decodeCoords (reference src/rfqcodec.cpp:1332-1389) with a WARP per (chunk, column), 32 stream bytes per
 * step (v1 ran one thread per stream and was latency bound: 1 ms for 3.4 KB streams).  Token length depends on the first
 * byte only (0xxxxxxx: 2 bytes absolute, 10xxxxxx: 1 byte delta, 110xxxxx: 1 byte repeat, 111xxxxx: 3 bytes absolute);
 * values are a segmented inclusive scan (absolute tokens reset, deltas add), output slots an exclusive scan of counts.
 * grid (n_chunks, 2), 32 threads.
 */
__global__ void __launch_bounds__(32) k_dec_coords2(DecBatchDev b, HeaderDev h) {
    const u32 c = blockIdx.x, col = blockIdx.y;
    const int lane = threadIdx.x;
    if (!(h.flags & (col ? RPQ_HAS_Y : RPQ_HAS_X))) return;
    const DecChunk& ck = b.chunks[c];
    const u8* buf = b.body + ck.in_off + (col ? ck.off_y : ck.off_x);
    const u32 len = col ? ck.y_size : ck.x_size;
    const u32 num = (ck.flags & RPQ_PE_INTERLEAVED) ? ck.reads / 2 : ck.reads;
    u32* data = (col ? b.ys : b.xs) + ck.read_base;
    u32 last = 1000, d = 0, skip = 0;
    for (u32 base = 0; base < len; base += 32) {
        const u32 p = base + lane;
        const bool valid = p < len;
        const u32 b0 = valid ? buf[p] : 0xC0u;
        const u32 tlen = !(b0 & 0x80) ? 2u : ((b0 & 0xE0) == 0xE0) ? 3u : 1u;
        u32 multi = __ballot_sync(0xffffffffu, valid && tlen > 1);
        const u32 is3 = __ballot_sync(0xffffffffu, valid && tlen == 3);
        u32 skipped = skip >= 32 ? 0xffffffffu : ((1u << skip) - 1u);
        u32 next_skip = skip > 32 ? skip - 32 : 0;
        while (multi) {
            const int i = __ffs((int)multi) - 1;
            multi &= multi - 1;
            if ((skipped >> i) & 1u) continue;
            const u32 hi = (u32)i + (((is3 >> i) & 1u) ? 2u : 1u);
            for (u32 k = (u32)i + 1; k <= hi && k < 32; k++) skipped |= 1u << k;
            if (hi >= 32) next_skip = hi - 31;
        }
        const bool head = valid && !((skipped >> lane) & 1u);
        u32 reset = 0, val = 0, cnt = 0;
        if (head) {
            if (!(b0 & 0x80)) { reset = 1; val = (b0 << 8) | (p + 1 < len ? buf[p + 1] : 0u); cnt = 1; }
            else if (!(b0 & 0x40)) { val = (b0 & 0x3F) + 1; cnt = 1; }
            else if (!(b0 & 0x20)) { cnt = (b0 & 0x1F) + 1; }
            else { reset = 1; val = ((b0 & 0x1F) << 16) | ((u32)(p + 1 < len ? buf[p + 1] : 0u) << 8) | (u32)(p + 2 < len ? buf[p + 2] : 0u); cnt = 1; }
        }
        /* segmented inclusive scan of (reset, val) */
        u32 r = reset, v = val;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const u32 pr = __shfl_up_sync(0xffffffffu, r, s), pv = __shfl_up_sync(0xffffffffu, v, s);
            if (lane >= s && !r) { v += pv; r = pr; }
        }
        const u32 value = r ? v : last + v;
        u32 tot; const u32 ex = warp_excl_scan(cnt, lane, tot);
        if (head) { for (u32 k = 0; k < cnt; k++) if (d + ex + k < num) data[d + ex + k] = value; }
        last = __shfl_sync(0xffffffffu, value, 31);
        d += tot;
        skip = next_skip;
    }
    for (u32 k = d + lane; k < num; k += 32) data[k] = 0;      /* memset(xBuf, 0): values the stream does not cover stay 0 */
}

}  // namespace rpq

namespace rpq {

/*
 * Two-step chunk walk for .rfq bodies that are already in HBM (v1 k_dec_walk needed ~4 dependent loads per chunk):
 *  k_dec_walk_fast  one warp follows the chain with ONE load per chunk: mSize, which the reference computes wrongly but
 *                   deterministically (Q2: true size = mSize + lane bytes - name2/tile bytes counted although absent);
 *  k_dec_describe   a warp per chunk derives every column offset exactly as RfqChunk::read does (reference
 *                   src/rfqchunk.cpp:161-228, :63-109) and checks the true size against the one assumed by the chain;
 *                   any mismatch makes the host fall back to the exact sequential walk (k_dec_walk).
 */
__global__ void k_dec_walk_fast(const u8* body, u64 len, HeaderDev h, DecChunk* chunks, u32 cap, u32* n_out, u64* consumed) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 at = 0; u32 n = 0, read_base = 0;
    const u32 head = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    while (at + head <= len && n < cap) {
        const u8* in = body + at;
        const u32 ms = ld32(in), reads = ld32(in + 4), fl = (u32)in[8] | ((u32)in[9] << 8);
        if (reads == 0) break;
        const u32 xy = (fl & RPQ_PE_INTERLEAVED) ? reads / 2 : reads;
        long long bytes = (long long)ms;
        if (h.flags & RPQ_HAS_LANE) bytes += (fl & RPQ_LANE_SAME) ? 1 : xy;
        if (!(h.flags & RPQ_HAS_NAME2)) bytes -= (fl & RPQ_NAME2_LEN_SAME) ? 1 : reads;
        if (!(h.flags & RPQ_HAS_TILE)) bytes -= 2ll * ((fl & RPQ_TILE_SAME) ? 1 : xy);
        if (bytes < (long long)head || at + (u64)bytes > len) break;
        DecChunk c; memset(&c, 0, sizeof c);
        c.in_off = at; c.reads = reads; c.flags = fl; c.bytes = (u32)bytes; c.read_base = read_base;
        chunks[n] = c;
        n++; read_base += reads; at += (u64)bytes;
    }
    *n_out = n; *consumed = at;
}

__global__ void __launch_bounds__(128) k_dec_describe(const u8* body, HeaderDev h, DecChunk* chunks, u32 n_chunks, u32* mismatch) {
    const int lane = threadIdx.x & 31;
    const u32 ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= n_chunks) return;
    DecChunk c = chunks[ci];
    const u8* in = body + c.in_off;
    const u64 left = c.bytes;
    const u32 head = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    c.seq_size = ld32(in + 10); c.qual_size = ld32(in + 14);
    if (h.flags & RPQ_ENCODE_N_POS) c.npos_size = ld32(in + 18);
    const u32 nr = c.reads, fl = c.flags;
    const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
    const u32 xy = il ? nr / 2 : nr;
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
    if (!bad) {
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
    }
    if (bad || o != left) { if (lane == 0) atomicOr(mismatch, 1u); return; }
    if (lane == 0) chunks[ci] = c;
}

}  // namespace rpq
