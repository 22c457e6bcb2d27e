/*
 * rpq_decode2.cuh - second-generation decode helpers: the CTA-per-stream coordinate decoder and the two-step device chunk
 * walk.
 */
#pragma once
#include "rpq_decode.cuh"

namespace rpq {

}  // namespace rpq

namespace rpq {

/*
 * k_dec_coords3: decodeCoords (reference src/rfqcodec.cpp:1332-1389) with a CTA per (chunk, column), one stream byte per thread
 * and DC3_THREADS bytes per step.  v1 ran a thread per stream (latency bound: 1 ms for 3.4 KB streams), v2 a warp per stream
 * whose step still resolved the token heads one after the other (0.39 ms, the time of its longest warp whatever the batch).
 * Token length depends on the first byte only (0xxxxxxx: 2 bytes absolute, 10xxxxxx: 1 byte delta, 110xxxxx: 1 byte repeat,
 * 111xxxxx: 3 bytes absolute), so "how many payload bytes are still to be skipped" is a 3-state machine; every byte is a map of
 * that state (three 2-bit fields), and an inclusive scan of the maps under composition tells every thread whether its byte is a
 * head.  Values are a segmented inclusive scan (absolute tokens reset, deltas add), output slots an exclusive scan of counts;
 * both scans and the state cross warps through shared memory.  grid (n_chunks, 2).
 */
constexpr int DC3_THREADS = 256;
constexpr u32 DC3_IDENT = 0x24u;                      /* state s -> s */
/* first f, then g */
__device__ __forceinline__ u32 dc3_compose(u32 f, u32 g) {
    return ((g >> (2u * (f & 3u))) & 3u) | (((g >> (2u * ((f >> 2) & 3u))) & 3u) << 2) | (((g >> (2u * ((f >> 4) & 3u))) & 3u) << 4);
}

__global__ void __launch_bounds__(DC3_THREADS) k_dec_coords3(DecBatchDev b, HeaderDev h) {
    constexpr int NW = DC3_THREADS / 32;
    __shared__ u32 s_map[NW], s_r[NW], s_v[NW], s_c[NW];
    const u32 c = blockIdx.x, col = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(h.flags & (col ? RPQ_HAS_Y : RPQ_HAS_X))) return;
    const DecChunk& ck = b.chunks[c];
    const u8* buf = b.body + ck.in_off + (col ? ck.off_y : ck.off_x);
    const u32 len = col ? ck.y_size : ck.x_size;
    const u32 num = (ck.flags & RPQ_PE_INTERLEAVED) ? ck.reads / 2 : ck.reads;
    u32* data = (col ? b.ys : b.xs) + ck.read_base;
    u32 last = 1000, d = 0, state = 0;                 /* state: payload bytes of the previous step's last token still to come */
    u32 nb = (u32)tid < len ? buf[tid] : 0u;
    for (u32 base = 0; base < len; base += DC3_THREADS) {
        const u32 p = base + (u32)tid;
        const bool valid = p < len;
        const u32 b0 = nb;
        nb = p + DC3_THREADS < len ? buf[p + DC3_THREADS] : 0u;          /* the next step's byte is in flight during this one */
        const u32 cont = !(b0 & 0x80u) ? 1u : ((b0 & 0xE0u) == 0xE0u ? 2u : 0u);
        /* this byte as a map: state 0 (a head) -> cont, 1 -> 0, 2 -> 1 */
        u32 inc = valid ? (cont | 0x10u) : DC3_IDENT;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const u32 prev = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc = dc3_compose(prev, inc); }
        if (lane == 31) s_map[warp] = inc;
        u32 exm = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exm = DC3_IDENT;
        __syncthreads();
        u32 st = state, st_all = state;
#pragma unroll
        for (int q = 0; q < NW; q++) { const u32 m = s_map[q]; st_all = (m >> (2u * st_all)) & 3u; if (q < warp) st = st_all; }
        state = st_all;
        const bool head = valid && ((exm >> (2u * st)) & 3u) == 0u;
        u32 reset = 0, val = 0, cnt = 0;
        if (head) {
            if (!(b0 & 0x80u)) { reset = 1; val = (b0 << 8) | (p + 1 < len ? buf[p + 1] : 0u); cnt = 1; }
            else if (!(b0 & 0x40u)) { val = (b0 & 0x3Fu) + 1; cnt = 1; }
            else if (!(b0 & 0x20u)) { cnt = (b0 & 0x1Fu) + 1; }
            else { reset = 1; val = ((b0 & 0x1Fu) << 16) | ((u32)(p + 1 < len ? buf[p + 1] : 0u) << 8) | (u32)(p + 2 < len ? buf[p + 2] : 0u); cnt = 1; }
        }
        /* segmented inclusive scan of (reset, val), inclusive scan of cnt: inside the warp ... */
        u32 r = reset, v = val, ci = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const u32 pr = __shfl_up_sync(0xffffffffu, r, s), pv = __shfl_up_sync(0xffffffffu, v, s), pc = __shfl_up_sync(0xffffffffu, ci, s);
            if (lane >= s) { if (!r) { v += pv; r = pr; } ci += pc; }
        }
        if (lane == 31) { s_r[warp] = r; s_v[warp] = v; s_c[warp] = ci; }
        __syncthreads();
        /* ... and across the warps before this one (and all of them, for the next step) */
        u32 wr = 0, wv = 0, wc = 0, ar = 0, av = 0, ac = 0;
#pragma unroll
        for (int q = 0; q < NW; q++) {
            const u32 qr = s_r[q], qv = s_v[q], qc = s_c[q];
            if (qr) { ar = 1; av = qv; } else av += qv;
            ac += qc;
            if (q + 1 == warp) { wr = ar; wv = av; wc = ac; }
        }
        if (!r) { v += wv; r = wr; }
        const u32 value = r ? v : last + v;
        if (head) { const u32 o = d + wc + ci - cnt; for (u32 k = 0; k < cnt; k++) if (o + k < num) data[o + k] = value; }
        last = ar ? av : last + av;
        d += ac;
    }
    for (u32 k = d + (u32)tid; k < num; k += DC3_THREADS) data[k] = 0;      /* memset(xBuf, 0): values the stream does not cover stay 0 */
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
    /* all 32 lanes follow the chain redundantly (broadcast loads); lane 0 records, and every lane warms L2 around the place
     * the header after the next one is expected (chunks of one file are nearly the same size) */
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
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
#ifndef RPQ_EMU
        {
            /* 16 KiB around the expected place of the header after the next one, four lines per lane: the size of the chunks of one
             * file varies by a few KiB, and a header that is not in L2 when the chain arrives costs a DRAM round trip */
            const long long pf = (long long)at + 3ll * bytes + ((long long)lane - 16) * 512;
#pragma unroll
            for (int k = 0; k < 4; k++) { const long long q = pf + 128 * k; if (q >= 0 && (u64)q + 128 <= len) asm volatile("prefetch.global.L2 [%0];" ::"l"(body + q)); }
        }
#endif
        if (lane == 0) {
            DecChunk c; memset(&c, 0, sizeof c);
            c.in_off = at; c.reads = reads; c.flags = fl; c.bytes = (u32)bytes; c.read_base = read_base;
            chunks[n] = c;
        }
        n++; read_base += reads; at += (u64)bytes;
    }
    if (lane == 0) { n_out[0] = n; n_out[1] = read_base; *consumed = at; }                 /* chunks, reads, body bytes */
}

/*
 * The same chain followed by several warps at once.  The chain itself cannot be entered in the middle - but a place where a chunk
 * header starts can be recognised: k_dec_find_heads scans forward from W-1 evenly spaced anchors for the first offset whose header
 * fields are consistent with the container AND whose mSize leads to another such header (or to the end of the body).  Warp w of
 * k_dec_walk_par then follows the chain from its start to the start of the next warp.  A start that is not on the real chain is
 * harmless: the warp before it does not land on it, the stitch test fails and the host takes the sequential walk, exactly as it
 * does when k_dec_describe finds a chunk whose columns do not add up.  1430 hops of ~0.5 us become 16 x 90.
 */
constexpr int WP_WARPS = 16;
constexpr u32 WP_SLAB = 1u << 16;                 /* chunks a warp can record */
constexpr u32 WP_NONE = 0xFFFFFFFFu;
struct WalkSeg { u64 in_off; u32 reads, flags, bytes, read_base; };

/* the walk's one-load header test on the bytes `in` found at body offset p (global or staged in shared memory): 0 if they cannot be
 * a chunk header, else the chunk's size */
__device__ __forceinline__ u64 wp_header_bytes(const u8* in, u64 len, const HeaderDev& h, u64 p, u32 head, u32& reads, u32& fl) {
    if (p + head > len) return 0;
    const u32 ms = ld32(in);
    reads = ld32(in + 4); fl = (u32)in[8] | ((u32)in[9] << 8);
    if (reads == 0) return 0;
    const u32 xy = (fl & RPQ_PE_INTERLEAVED) ? reads / 2 : reads;
    long long bytes = (long long)ms;
    if (h.flags & RPQ_HAS_LANE) bytes += (fl & RPQ_LANE_SAME) ? 1 : xy;
    if (!(h.flags & RPQ_HAS_NAME2)) bytes -= (fl & RPQ_NAME2_LEN_SAME) ? 1 : reads;
    if (!(h.flags & RPQ_HAS_TILE)) bytes -= 2ll * ((fl & RPQ_TILE_SAME) ? 1 : xy);
    if (bytes < (long long)head || p + (u64)bytes > len) return 0;
    return (u64)bytes;
}

/* grid (WP_FIND_CTAS, WP_WARPS - 1): the CTAs of row j look for the start of warp j + 1, the first header at or after its anchor,
 * each in its own 16 KiB of a 1 MiB window, all at once; starts[] is preset to ~0 and takes the minimum.  A CTA stages its bytes
 * in shared memory with coalesced 16-byte loads (a test per candidate straight from global memory was a chain of 64 dependent
 * loads per thread: 0.18 ms) and only follows a candidate's mSize into global memory.  No header in the window (chunks of
 * several MB): that warp sits out and the one before it walks on. */
constexpr u32 WP_FIND_CTAS = 64, WP_FIND_SPAN = 16384;
__global__ void __launch_bounds__(256) k_dec_find_heads(const u8* body, u64 len, HeaderDev h, unsigned long long* starts) {
    __shared__ uint4 s_buf4[(WP_FIND_SPAN + 64) / 16];
    u8* s_buf = reinterpret_cast<u8*>(s_buf4);
    const u32 w = blockIdx.y + 1;
    const u32 head = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    const u64 anchor = len / WP_WARPS * w;
    const u64 lo = anchor + (u64)blockIdx.x * WP_FIND_SPAN;
    if (lo >= len) return;
    /* stage [lo - phase, lo - phase + SPAN + 64) where the source address is 16-byte aligned; bytes past the body read as 0 */
    const u32 phase = (u32)(reinterpret_cast<uintptr_t>(body + lo) & 15u);
    const u8* src = body + lo - phase;                      /* lo >= len / 16 > 15: still inside the body */
    const u64 avail = len - (lo - phase);
    for (u32 k = threadIdx.x; k < (WP_FIND_SPAN + 64) / 16; k += blockDim.x) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (16ull * k + 16 <= avail) v = *reinterpret_cast<const uint4*>(src + 16 * k);
        else for (u32 q = 0; q < 16 && 16ull * k + q < avail; q++) reinterpret_cast<u8*>(&v)[q] = src[16 * k + q];
        *reinterpret_cast<uint4*>(s_buf + 16 * k) = v;
    }
    __syncthreads();
    u64 best = ~0ull;
    for (u32 k = threadIdx.x; k < WP_FIND_SPAN; k += blockDim.x) {
        const u64 p = lo + k;
        if (p + head > len) break;
        const u8* in = s_buf + phase + k;                    /* at most 15 + 16383 + 22 bytes into the buffer */
        /* cheap rejections first: 12 flag bits, a read count and column sizes that fit the size the header claims */
        if (in[9] & 0xF0u) continue;
        u32 reads, fl;
        const u64 bytes = wp_header_bytes(in, len, h, p, head, reads, fl);
        if (!bytes || reads > (1u << 26)) continue;
        const u64 seq = ld32(in + 10), qual = ld32(in + 14);
        if (seq + qual + head > bytes || seq > 4ull * reads * 65536ull) continue;
        /* a real header leads to another one, four times over (or to the end of the body).  Two hops are not enough: the position
         * streams of a file with ~40 quality values are dense small bytes, of which a 0.6 GB body holds a handful of places that
         * pass the tests above and point at another such place */
        bool chain = true;
        u64 q = p + bytes;
        for (int hop = 0; hop < 4 && chain && q != len; hop++) {
            u32 r2, f2;
            const u64 b2 = wp_header_bytes(body + q, len, h, q, head, r2, f2);
            if (!b2 || (body[q + 9] & 0xF0u) || r2 > (1u << 26)) chain = false;
            q += b2;
        }
        if (!chain) continue;
        best = p;
        break;                                      /* this thread's candidates only grow */
    }
    if (best != ~0ull) atomicMin(&starts[w], (unsigned long long)best);
}

/* one CTA of WP_WARPS warps.  out: u32[0] chunks, u32[1] reads, u64[1] body bytes covered (as the sequential walks), mismatch |= 1
 * when the pieces do not fit together */
__global__ void __launch_bounds__(32 * WP_WARPS) k_dec_walk_par(const u8* body, u64 len, HeaderDev h, const unsigned long long* __restrict__ starts, WalkSeg* slabs,
                                                                DecChunk* chunks, u32 cap, u32* n_out, u64* consumed, u32* mismatch) {
    __shared__ u64 s_from[WP_WARPS], s_stop[WP_WARPS], s_end[WP_WARPS];
    __shared__ u32 s_n[WP_WARPS], s_reads[WP_WARPS], s_use[WP_WARPS];
    __shared__ u32 s_nbase[WP_WARPS], s_rbase[WP_WARPS];
    __shared__ u32 s_bad;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 head = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    if (threadIdx.x == 0) s_bad = 0;
    /* this warp's piece of the chain: from its start to the next start that exists */
    u64 from = w == 0 ? 0ull : starts[w];
    u64 stop = len;
    for (int j = WP_WARPS - 1; j > w; j--) { const u64 sj = starts[j]; if (sj != ~0ull) stop = sj; }
    const bool active = from != ~0ull;
    WalkSeg* slab = slabs + (size_t)w * WP_SLAB;
    u64 at = active ? from : 0ull; u32 n = 0, reads_sum = 0;
    bool overflow = false;
    while (active && at < stop) {
        u32 reads, fl;
        const u64 bytes = wp_header_bytes(body + at, len, h, at, head, reads, fl);
        if (!bytes) break;                              /* the chain ends here: what follows is not a chunk */
        if (n >= WP_SLAB) { overflow = true; break; }
#ifndef RPQ_EMU
        {
            const long long pf = (long long)at + 3ll * (long long)bytes + ((long long)lane - 16) * 512;
#pragma unroll
            for (int k = 0; k < 4; k++) { const long long q = pf + 128 * k; if (q >= 0 && (u64)q + 128 <= len) asm volatile("prefetch.global.L2 [%0];" ::"l"(body + q)); }
        }
#endif
        if (lane == 0) { WalkSeg g; g.in_off = at; g.reads = reads; g.flags = fl; g.bytes = (u32)bytes; g.read_base = reads_sum; slab[n] = g; }
        n++; reads_sum += reads; at += bytes;
    }
    if (lane == 0) { s_from[w] = from; s_stop[w] = stop; s_end[w] = at; s_n[w] = n; s_reads[w] = reads_sum; s_use[w] = 0; if (overflow) s_bad = 1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* stitch: every piece must end exactly where the next one starts; a piece that stops short ends the chain (as the
         * sequential walk would), one that overshoots was aimed at a start that is not a chunk header */
        u32 nb = 0, rb = 0; u64 end = 0; bool open = true;
        for (int j = 0; j < WP_WARPS && open; j++) {
            if (s_from[j] == ~0ull) continue;
            s_use[j] = 1; s_nbase[j] = nb; s_rbase[j] = rb;
            nb += s_n[j]; rb += s_reads[j]; end = s_end[j];
            if (s_end[j] > s_stop[j]) { s_bad = 1; open = false; }
            else if (s_end[j] < s_stop[j]) open = false;
        }
        if (nb > cap) s_bad = 1;
        n_out[0] = nb; n_out[1] = rb; *consumed = end;
        if (s_bad) atomicOr(mismatch, 1u);
    }
    __syncthreads();
    if (s_bad || !s_use[w]) return;
    const u32 nb = s_nbase[w], rb = s_rbase[w];
    for (u32 k = lane; k < n; k += 32) {
        const WalkSeg g = slab[k];
        DecChunk c; memset(&c, 0, sizeof c);
        c.in_off = g.in_off; c.reads = g.reads; c.flags = g.flags; c.bytes = g.bytes; c.read_base = rb + g.read_base;
        chunks[nb + k] = c;
    }
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
