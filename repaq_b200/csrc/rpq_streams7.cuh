/*
 * rpq_streams7.cuh - k_streams7: the position-stream coder of the quality column (reference src/rfqcodec.cpp:625-765) for DENSE
 * spans: quality columns with ~40 values (BGI-SEQ, older Illumina), where nearly every position starts a run and emits one token
 * byte into one of ~40 streams.  It replaces k_streams6 (a warp walks 32 positions per step and sorts them into streams with
 * __match_any_sync, twice: 356 thread instructions per position, 83 % of the issue slots busy - instruction bound at 1.5 % of the
 * HBM roofline, profiles/README.md r02) and takes the spans k_streams4 hands over for their long runs as well.
 *
 * A first attempt kept bit planes of a thread's 64 positions and walked the streams (occupancy mask of each value, tokens from bit
 * operations): loop-free only while no run is longer than one position, and a warp's 32 segments never all are - 2.7 G warp
 * instructions per 200 M positions with 11 active lanes (profiles/README.md r02, the table of builds), slower than k_streams6.  What is dense
 * in these columns is the POSITIONS, so the thread walks positions, not streams:
 *
 *   - a thread owns 64 consecutive positions of the span, their bytes in 16 registers, their "equals the previous position" mask E
 *     in two; the walk is fully unrolled (no index arithmetic, every lane does the same 64 steps: no divergence on dense data);
 *   - its state per stream is ONE 32-bit cell of a [stream][thread] table in shared memory - a column per thread, so lanes never
 *     share a bank whatever streams they touch.  Count pass: {last occurrence, bytes, the start that has no predecessor in the
 *     segment}; write pass: {byte offset, last occurrence};
 *   - between the passes a warp per stream scans the 256 cells (predecessor across segments, size of the tokens that waited for
 *     it, byte offsets, SpanDir);
 *   - a token belongs to the position that heads it (src/rfqcodec.cpp:648-700): a run start its distance token, position 1 of a run
 *     that starts at position 0 a zero byte (Q16), every 32nd position of a run after that a length token (the run's end from E,
 *     from the neighbours' E beyond the segment), a value that is not in the header a 5-byte exception record (:750-758).
 *
 * Same contract as k_streams3 / 4 (SpanDir, slots, the stream's first distance token of a span left to k_layout).  The start of a
 * run that crosses into the span is found by a warp, 32 positions per step: in the staged halo, beyond it in the text - no span is
 * handed to another coder.  Quality streams only (mode 0).  16.5 KB + 1 KB per stream of shared memory.
 */
#pragma once
#include "rpq_streams4.cuh"

namespace rpq {

constexpr int S7_THREADS = ST_SPAN / 64;                  /* 256: a thread per 64 positions */
constexpr u32 S7_NOLAST = 0x7FFFu;                        /* 15-bit "no position" (span-relative positions are < 16384) */
constexpr u32 S7_NOFIRST = 127u;                          /* 7-bit "no start waits for its predecessor" */

__host__ __device__ inline size_t streams7_smem(u32 nstreams) {
    size_t tables = (size_t)nstreams * S7_THREADS * sizeof(u32);
    const size_t stage_tables = 2 * (SQ_CAP + 1) * sizeof(u32);
    if (tables < stage_tables) tables = stage_tables;
    return (size_t)ST_SPAN + 2 * ST_HALO + 16 + tables;
}

/* what a walk needs to know about the thread's segment */
struct S7Seg {
    u64 E;                /* bit j: position s + j equals position s + j - 1 (0 for position 0 and beyond the column) */
    u64 Enext;            /* E of the next segment (of the 64 positions after the span for the last one) */
    u64 V;                /* bit j: position s + j is below hi */
    const u8* row;        /* the segment's bytes (shared memory, 64-byte aligned) */
    u32 s, lo;            /* the segment's first position, the span's first position */
    u32 nh;               /* the next position of the run position s lies in that heads a token (Q16 byte or length token) */
    bool q16;             /* that run starts at position 0 of the column */
};

/*
 * One pass over the thread's positions, 16 (one 128-bit shared load) per iteration, those unrolled.  cell = the thread's column of
 * the table (stride S7_THREADS).
 * !WRITE: cells {last occurrence in the segment (15 bits, span relative) | bytes << 15 | start without predecessor << 25}; exc = bytes
 * of exception records; starts = run starts of values other than the major one.
 * WRITE: cells {offset of the stream's next byte inside the span's slot (17 bits) | last occurrence << 17}; exc = offset of the next
 * exception record.
 * Which positions head a token of their run (src/rfqcodec.cpp:677-700): position 1 of a run that starts at position 0 (a zero byte,
 * Q16), then every 32nd - kept as "the next head" nh, set at every run start whatever the value.
 */
template <bool WRITE>
__device__ __forceinline__ void s7_walk(const S7Seg& g, const u8* s_lut, u32* cell, u8* slot, u32& exc, u32& starts) {
    u32 nh = g.nh;
    bool q16 = g.q16;
#pragma unroll 1
    for (u32 piece = 0; piece < 4u; piece++) {
        const u32 v16 = (u32)(g.V >> (16u * piece)) & 0xFFFFu;
        if (!v16) break;
        /* 64 bits of E from the piece's first position on: enough for every length token headed in the piece (32 positions) */
        const u64 ewin = piece ? (g.E >> (16u * piece)) | (g.Enext << (64u - 16u * piece)) : g.E;
        const uint4 q = *reinterpret_cast<const uint4*>(g.row + 16u * piece);
        const u32 ewl = (u32)ewin, ewh = (u32)(ewin >> 32);
        const bool col0 = g.s == 0 && piece == 0;                   /* the piece starts the column: positions 0 and 1 are special */
#pragma unroll
        for (int jj = 0; jj < 16; jj++) {
            if (!((v16 >> jj) & 1u)) break;
            const u32 word = jj < 4 ? q.x : jj < 8 ? q.y : jj < 12 ? q.z : q.w;
            const u32 v = (word >> (8 * (jj & 3))) & 0xFFu;
            const bool cont = ((ewl >> jj) & 1u) != 0;
            const u32 j = 16u * piece + (u32)jj;
            const u32 p = g.s + j;
            const bool p_is_0 = jj == 0 && col0;
            const bool head = cont && p == nh;
            const bool zero = jj == 1 && col0 && head && q16;       /* Q16 */
            if (!cont) { nh = p + 1u; if (jj == 0) q16 = col0; else q16 = false; }
            else if (head) nh = zero ? 2u : p + 32u;
            const u32 l = s_lut[v];
            if (l == LUT_SKIP) continue;                             /* the major quality: the decoder's fill value, no token */
            if (!WRITE && !cont) starts++;
            if (l == LUT_EXC) {
                if (WRITE) { u8* o = slot + exc; o[0] = (u8)v; o[1] = (u8)p; o[2] = (u8)(p >> 8); o[3] = (u8)(p >> 16); o[4] = (u8)(p >> 24); }
                exc += 5u;
                continue;
            }
            u32* c = cell + l * S7_THREADS;
            u32 ent = *c;
            if (!WRITE) {
                u32 add = head ? 1u : 0u;
                if (!cont) {
                    if ((ent & 0x7FFFu) != S7_NOLAST || p_is_0) add = 1u;           /* a predecessor less than 64 positions back: one byte */
                    else ent = (ent & ~(0x7Fu << 25)) | (j << 25);                 /* sized by the scan */
                }
                ent = ((ent & ~0x7FFFu) | (p - g.lo)) + (add << 15);
            } else {
                u32 at = ent & 0x1FFFFu;
                const u32 lastb = ent >> 17;
                u8* o = slot + at;
                if (!cont) {
                    if (lastb != S7_NOLAST) {
                        const u32 dm = (p - g.lo) - lastb - 1u;
                        if (dm < 128u) { o[0] = (u8)dm; at += 1u; }
                        else if (dm < (1u << 14)) { o[0] = (u8)(0x80u | (dm >> 8)); o[1] = (u8)dm; at += 2u; }
                        else { o[0] = (u8)(0xE0u | (dm >> 24)); o[1] = (u8)(dm >> 16); o[2] = (u8)(dm >> 8); o[3] = (u8)dm; at += 4u; }
                    } else if (p_is_0) { o[0] = 0; at += 1u; }
                    /* else: the stream's first token of the span, k_layout's */
                } else if (head) {
                    if (zero) o[0] = 0;
                    else {
                        /* the positions of the run after this one, as far as the token counts them (31) */
                        const u32 more = (u32)(__ffs((int)~__funnelshift_r(ewl, ewh, jj + 1)) - 1);      /* no zero among 32: 0 - 1, clipped */
                        o[0] = (u8)(0xC0u | (more < 31u ? more : 31u));
                    }
                    at += 1u;
                }
                ent = at | ((p - g.lo) << 17);
            }
            *c = ent;
        }
    }
}

__global__ void __launch_bounds__(S7_THREADS) k_streams7(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk, const u32* __restrict__ list) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ u64 s_eq[S7_THREADS + 2];
    __shared__ u32 s_total[MAX_BINS + 2], s_base[MAX_BINS + 2];
    __shared__ u64 s_slot;
    __shared__ u32 s_tmp, s_cross_p, s_bytes, s_runs;
    const u32 span = list ? list[blockIdx.x] : blockIdx.x;          /* the spans k_streams4 passed on, or all of them */
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    const u32 n = ck.total_len;
    const u32 lo = (span - job.span_first[c]) * ST_SPAN;
    const u32 hi = lo + ST_SPAN < n ? lo + ST_SPAN : n;
    const u32 sm_lo = lo >= ST_HALO ? lo - ST_HALO : 0, sm_hi = hi + ST_HALO < n ? hi + ST_HALO : n;
    const u32 nstreams = job.nstreams, nb = nstreams - 1u;         /* streams 0..nb-1: the header's values; nb: exception records */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u8* sm = dyn;
    u32* T = reinterpret_cast<u32*>(dyn + ST_SPAN + 2 * ST_HALO + 16);         /* [nstreams][S7_THREADS]; the staging tables alias it */

    s_lut[tid] = h.lut[tid];
    if (tid == 0) { s_cross_p = lo; s_runs = 0; }
    for (u32 k = tid; k < 16; k += S7_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO + 16)) sm[sm_hi - sm_lo + k] = h.major;
    if (!stage_quality_flat(b, ck, sm_lo, sm_hi, sm, job.span_read0[span], T, T + SQ_CAP + 1, &s_tmp))
        stage_quality_words(b, ck, sm_lo, sm_hi, sm, job.span_read0[span]);
    __syncthreads();

    /* ---- the thread's 64 positions: which of them equal their predecessor */
    const u32 s = lo + (u32)tid * 64u;
    const u8* row = sm + ((s < hi ? s : lo) - sm_lo);               /* 64-byte aligned: lo - sm_lo is 0 or 64 */
    u32 El = 0, Eh = 0;
    u64 V = 0;
    if (s < hi) {
        V = hi - s >= 64u ? ~0ull : (1ull << (hi - s)) - 1ull;
        u32 prevw = s > sm_lo ? (u32)row[-1] << 24 : 0u;            /* the byte before the segment in the top byte */
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 v = *reinterpret_cast<const uint4*>(row + 16 * k);
            const u32 e0 = eq_bytes(v.x, __funnelshift_l(prevw, v.x, 8)), e1 = eq_bytes(v.y, __funnelshift_l(v.x, v.y, 8));
            const u32 e2 = eq_bytes(v.z, __funnelshift_l(v.y, v.z, 8)), e3 = eq_bytes(v.w, __funnelshift_l(v.z, v.w, 8));
            const u32 bits = ((((e0 >> 4) | e1) * 0x00204081u) >> 24) | (((((e2 >> 4) | e3) * 0x00204081u) >> 16) & 0xFF00u);
            if (k & 2) Eh |= bits << (16 * (k & 1)); else El |= bits << (16 * (k & 1));
            prevw = v.w;
        }
        if (s == 0) El &= ~1u;                                      /* position 0 has no predecessor */
    }
    const u64 E = (((u64)Eh << 32) | El) & V;                        /* V: below hi; a span that ends before lo + 16384 ends the column */
    s_eq[tid] = E;
    /* the table starts empty */
    for (u32 st = 0; st < nb; st++) T[st * S7_THREADS + tid] = S7_NOLAST | (S7_NOFIRST << 25);
    if (warp == S7_THREADS / 32 - 1) {
        /* the 64 positions after the span, for runs that leave it (their length tokens are clipped to 32 positions) */
        const u32 q0 = lo + (u32)ST_SPAN + (u32)lane, q1 = q0 + 32u;
        const bool full = hi == lo + (u32)ST_SPAN;
        const u32 m0 = __ballot_sync(0xffffffffu, full && q0 < n && sm[q0 - sm_lo] == sm[q0 - 1u - sm_lo]);
        const u32 m1 = __ballot_sync(0xffffffffu, full && q1 < n && sm[q1 - sm_lo] == sm[q1 - 1u - sm_lo]);
        if (lane == 0) { s_eq[S7_THREADS] = ((u64)m1 << 32) | m0; s_eq[S7_THREADS + 1] = 0; }
    }
    if (warp == 0 && lo > 0) {
        /* a run crosses into the span: where it starts matters if its value has a stream (the heads of its length tokens are
         * counted from there).  32 positions per step backwards, from the staged halo, beyond it (rare) from the text */
        const bool cross = __shfl_sync(0xffffffffu, (u32)(E & 1ull), 0) != 0;
        const u8 v0 = sm[lo - sm_lo];
        if (cross && s_lut[v0] < LUT_EXC) {
            u32 q = lo;                                              /* the run is known to hold [q, lo] */
            for (;;) {
                const bool ok = q >= 1u + (u32)lane;
                const u32 cand = q - 1u - (u32)lane;
                const bool same = ok && (cand >= sm_lo ? sm[cand - sm_lo] : stream_byte_slow(b, ck, 0, cand)) == v0;
                const u32 m = __ballot_sync(0xffffffffu, same);
                const u32 take = m == 0xffffffffu ? 32u : (u32)(__ffs((int)~m) - 1);
                q -= take;
                if (take < 32u) break;
            }
            if (lane == 0) s_cross_p = q;
        }
    }
    __syncthreads();
    S7Seg g; g.E = E; g.Enext = s_eq[tid + 1]; g.V = V; g.row = row; g.s = s; g.lo = lo; g.nh = 0; g.q16 = false;
    if (E & 1ull) {
        /* the run the first position lies in starts at the last position before it that differs from its predecessor; its token
         * heads: position 1 and then 2, 34, ... if it starts at position 0 of the column, else every 32nd from its second position */
        u32 p0 = s_cross_p;
        for (u32 ww = (u32)tid; ww-- > 0;) { const u64 z = ~s_eq[ww]; if (z) { p0 = lo + 64u * ww + 63u - (u32)__clzll((long long)z); break; } }
        g.q16 = p0 == 0;
        const u32 h0 = p0 == 0 ? 2u : p0 + 1u;
        g.nh = h0 >= s ? h0 : h0 + ((s - h0 + 31u) / 32u) * 32u;
        if (p0 == 0 && s <= 1u) g.nh = 1u;
    }

    /* ---- pass 1: count */
    u32 exc = 0, starts = 0;
    s7_walk<false>(g, s_lut, T + tid, nullptr, exc, starts);
    T[nb * S7_THREADS + tid] = S7_NOLAST | (exc << 15) | (S7_NOFIRST << 25);      /* exception records: 5 bytes per position (:750-758) */
    if (!list) {
        /* coding every span of the batch (the last batch was mostly dense): count the spans k_streams4 would have coded itself, so
         * that the host sends the next batch there again when the data change */
        const u32 ws = warp_sum(starts);
        if (lane == 0 && ws) atomicAdd(&s_runs, ws);
    }
    __syncthreads();

    /* ---- per stream (a warp each, a lane per 8 segments): the predecessor of every segment, the size of the tokens that waited for
     * it, byte offsets, the directory.  Entries become {offset of the segment's bytes (17 bits), last occurrence before it (15)} */
    for (u32 st = warp; st < nstreams; st += S7_THREADS / 32) {
        u32* row = T + st * S7_THREADS + 8u * (u32)lane;
        u32 ent[8]; u32 lane_last = S7_NOLAST;
#pragma unroll
        for (int k = 0; k < 8; k++) { ent[k] = row[k]; const u32 l = ent[k] & 0x7FFFu; if (l != S7_NOLAST) lane_last = l; }
        u32 incl = lane_last;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d && incl == S7_NOLAST) incl = t; }
        u32 prev = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) prev = S7_NOLAST;
        u32 cntv[8], lastb[8]; u32 lane_sum = 0, span_first = NONE32;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 cv = (ent[k] >> 15) & 0x3FFu;
            const u32 f = ent[k] >> 25;
            if (f != S7_NOFIRST) {
                const u32 p = (8u * (u32)lane + (u32)k) * 64u + f;              /* span-relative */
                if (prev != S7_NOLAST) cv += distance_len(p - prev - 1u);
                else span_first = lo + p;                                       /* the stream's first token of the span: k_layout sizes it */
            }
            lastb[k] = prev;
            const u32 l = ent[k] & 0x7FFFu;
            if (l != S7_NOLAST) prev = l;
            cntv[k] = cv; lane_sum += cv;
        }
        u32 tot; u32 run = warp_excl_scan(lane_sum, lane, tot);
#pragma unroll
        for (int k = 0; k < 8; k++) { row[k] = run | (lastb[k] << 17); run += cntv[k]; }
        const u32 sf = warp_min(span_first);                                    /* at most one lane has it */
        const u32 sl = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            s_total[st] = tot;
            SpanDir d; d.bytes = tot; d.slot_off = 0; d.firstpos = st == nb ? NONE32 : sf; d.lastpos = (st == nb || sl == S7_NOLAST) ? NONE32 : lo + sl;
            d.dst = 0; d.first_tok = 0; d.first_len = 0; d.pad = 0;
            job.dir[(size_t)span * nstreams + st] = d;
        }
    }
    __syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (u32 st = 0; st < nstreams; st++) { s_base[st] = acc; acc += s_total[st]; }
        const u64 at = atomicAdd(job.slot_cursor, (u64)acc);
        job.span_slot[span] = at;
        if (at + acc > job.slot_cap) { atomicOr(job.overflow, 1u); s_slot = ~0ull; } else s_slot = at;
        s_bytes = acc;
        if (!list && s_runs <= (u32)RL_CAP) atomicAdd(job.dense_count, 1u);
    }
    __syncthreads();
    if (s_slot == ~0ull || s_bytes == 0) return;
    for (u32 st = tid; st < nstreams; st += S7_THREADS) job.dir[(size_t)span * nstreams + st].slot_off = s_base[st];

    /* ---- pass 2: the bytes.  The cells' offsets become offsets inside the slot */
    for (u32 st = 0; st < nstreams; st++) T[st * S7_THREADS + tid] += s_base[st];
    exc = T[nb * S7_THREADS + tid] & 0x1FFFFu;
    s7_walk<true>(g, s_lut, T + tid, job.slots + s_slot, exc, starts);
}

}  // namespace rpq
