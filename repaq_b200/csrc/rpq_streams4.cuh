/*
 * rpq_streams4.cuh - k_streams4: the position-stream coder (reference src/rfqcodec.cpp:625-765 and :420-426), fourth
 * generation.  Same contract as k_streams3 (SpanDir / slots / deferred first token), same staging and the same bit masks, but
 * the runs are not coded by the thread that owns their 64 positions:
 *
 *   k_streams3 walks "lowest set bit = next run" per thread.  NovaSeq-like data has 2.9 runs per 64 positions on average but
 *   8.4 in the busiest lane of a warp, and ~250 instructions per run: two thirds of the issue slots run with a few lanes.
 *
 * Here a thread only appends the START of each of its runs to a list of the span's runs (one exclusive scan), and the list is
 * then coded run-parallel, 32 consecutive runs per warp step:
 *   A   value, stream, run end; __match_any_sync groups the lanes by stream; last run of every stream per 32-run block
 *   P1  per stream: last run of the stream before each block (so every run knows its predecessor in the stream)
 *   B   tokens of the run that are headed inside the span: distance token (sized from the predecessor's end; the stream's first
 *       run of the span stays deferred to k_layout), Q16 byte, one length token per 32 positions; byte offsets by a
 *       per-stream exclusive scan inside the warp
 *   P2  per stream: block offsets, SpanDir, one bump allocation per span
 *   C   the bytes
 * Spans the list cannot describe cheaply (a run covering a whole 64-position segment or reaching back beyond the halo, more
 * than RL_CAP runs) are queued for k_streams3 (`redo_list`).
 */
#pragma once
#include "rpq_streams3.cuh"

namespace rpq {

constexpr int RL_CAP = 1536;                 /* runs per span in the list */
constexpr int RL_BLOCKS = RL_CAP / 32;
constexpr u32 RL_NONE = 0xFFFFu;

/* dynamic shared memory of k_streams4 after the staged bytes (the staging tables of stage_quality_flat alias the run arrays) */
__host__ __device__ inline size_t streams4_smem(u32 nstreams) {
    size_t run_arrays = (size_t)RL_CAP * (4 * sizeof(unsigned short) + 1);
    const size_t stage_tables = 2 * (SQ_CAP + 1) * sizeof(u32);
    if (run_arrays < stage_tables) run_arrays = stage_tables;
    run_arrays = (run_arrays + 15) & ~(size_t)15;
    return (size_t)ST_SPAN + 2 * ST_HALO + 16 + run_arrays + 2 * (size_t)RL_BLOCKS * nstreams * sizeof(unsigned short);
}

__global__ void __launch_bounds__(S2_THREADS) k_streams4(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ u32 s_total[MAX_BINS + 2];
    __shared__ u32 s_base[MAX_BINS + 2];
    __shared__ u32 s_first[MAX_BINS + 2];
    __shared__ u32 s_warp[S2_THREADS / 32];
    __shared__ u64 s_slot;
    __shared__ u64 s_eq[S2_THREADS];                    /* every thread's "equals the previous position" mask: run ends without byte loops */
    __shared__ u32 s_tmp, s_redo, s_cross_p;
    const u32 span = blockIdx.x;
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    const u32 mode = job.mode;
    const u32 n = mode ? ck.seq_kept : ck.total_len;
    const u32 lo = (span - job.span_first[c]) * ST_SPAN;
    const u32 hi = lo + ST_SPAN < n ? lo + ST_SPAN : n;
    const u32 sm_lo = lo >= ST_HALO ? lo - ST_HALO : 0, sm_hi = hi + ST_HALO < n ? hi + ST_HALO : n;
    const u32 nstreams = job.nstreams;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u8* sm = dyn;
    u8* arrays = dyn + ST_SPAN + 2 * ST_HALO + 16;
    unsigned short* rs = reinterpret_cast<unsigned short*>(arrays);      /* [RL_CAP] run start, span relative */
    unsigned short* rend = rs + RL_CAP;                                   /* [RL_CAP] run end (exclusive, clipped to hi + 32), span relative */
    unsigned short* roff = rend + RL_CAP;                                 /* [RL_CAP] byte offset inside (block, stream) */
    unsigned short* rdm = roff + RL_CAP;                                  /* [RL_CAP] distance - 1 of the run's distance token, RL_NONE: none / deferred */
    u8* rcls = reinterpret_cast<u8*>(rdm + RL_CAP);                       /* [RL_CAP] stream index */
    size_t arr_bytes = (size_t)RL_CAP * (4 * sizeof(unsigned short) + 1);
    { const size_t st = 2 * (SQ_CAP + 1) * sizeof(u32); if (arr_bytes < st) arr_bytes = st; arr_bytes = (arr_bytes + 15) & ~(size_t)15; }
    unsigned short* t_last = reinterpret_cast<unsigned short*>(arrays + arr_bytes);     /* [RL_BLOCKS][nstreams] last run of the stream in the block -> P1: before the block */
    unsigned short* t_bytes = t_last + (size_t)RL_BLOCKS * nstreams;                     /* [RL_BLOCKS][nstreams] bytes of the stream in the block -> P2: offset of the block */

    s_lut[tid] = h.lut[tid];
    if (tid == 0) s_redo = 0;
    for (u32 k = tid; k < 8; k += S2_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO + 16)) sm[sm_hi - sm_lo + k] = mode == 0 ? h.major : (u8)0;
    if (mode == 0) {
        u32* s_off = reinterpret_cast<u32*>(arrays);
        if (!stage_quality_flat(b, ck, sm_lo, sm_hi, sm, job.span_read0[span], s_off, s_off + SQ_CAP + 1, &s_tmp))
            stage_quality_words(b, ck, sm_lo, sm_hi, sm, job.span_read0[span]);
    } else if (!stage_positions(b, h, ck, mode, sm_lo, sm_hi, sm)) {
        /* N positions, and every read that reaches into the span is plain bases: an empty span */
        if (tid == 0) {
            SpanDir d; d.bytes = 0; d.slot_off = 0; d.firstpos = NONE32; d.lastpos = NONE32; d.dst = 0; d.first_tok = 0; d.first_len = 0; d.pad = 0;
            job.dir[(size_t)span * nstreams] = d;
            job.span_slot[span] = 0;
        }
        return;
    }
    __syncthreads();

    /* ---- masks of the thread's 64 positions (as in k_streams3) */
    const u32 s = lo + (u32)tid * S2_SEG;
    const u32 e = s + S2_SEG < hi ? s + S2_SEG : hi;
    u64 nm = 0, eq = 0;
    if (s < hi) {
        const u32* W = reinterpret_cast<const u32*>(sm + (s - sm_lo));
        const bool major_is_stream = mode == 0 && s_lut[h.major] != LUT_SKIP;
        const u32 mmmm = 0x01010101u * h.major;
        u32 prevw = s > 0 ? (u32)sm[s - 1 - sm_lo] << 24 : 0u;
        u32 nmw[2] = {0, 0}, eqw[2] = {0, 0};
#pragma unroll
        for (int j = 0; j < S2_SEG / 4; j += 2) {
            const u32 w0 = W[j], w1 = W[j + 1];
            const u32 e0 = eq_bytes(w0, __funnelshift_l(prevw, w0, 8)), e1 = eq_bytes(w1, __funnelshift_l(w0, w1, 8));
            u32 n0, n1;
            if (mode == 0) { n0 = major_is_stream ? 0x80808080u : eq_bytes(w0, mmmm) ^ 0x80808080u; n1 = major_is_stream ? 0x80808080u : eq_bytes(w1, mmmm) ^ 0x80808080u; }
            else { n0 = eq_bytes(w0, 0x4E4E4E4Eu); n1 = eq_bytes(w1, 0x4E4E4E4Eu); }
            const u32 eb = (((e0 >> 4) | e1) * 0x00204081u) >> 24, nb = (((n0 >> 4) | n1) * 0x00204081u) >> 24;
            eqw[j >> 3] |= eb << (8 * ((j >> 1) & 3)); nmw[j >> 3] |= nb << (8 * ((j >> 1) & 3));
            prevw = w1;
        }
        if (s == 0) eqw[0] &= ~1u;
        nm = (u64)nmw[0] | ((u64)nmw[1] << 32); eq = (u64)eqw[0] | ((u64)eqw[1] << 32);
        const u32 valid = e - s;
        if (valid < 64u) nm &= (1ull << valid) - 1ull;
    }

    s_eq[tid] = eq;
    /* ---- the list of runs: every thread appends the starts inside its segment; the run that crosses into the span is entry 0 */
    const u64 starts = nm & ~eq;
    const u32 my = (u32)__popcll(starts);
    u32 wtot; const u32 ex = warp_excl_scan(my, lane, wtot);
    if (lane == 0) s_warp[warp] = wtot;
    bool redo = (nm & eq) == ~0ull;                         /* a run covers the whole segment: long-run machinery of k_streams3 */
    u32 has_cross = 0;
    if (warp == 0) {
        /* the run that crosses into the span: its start, 32 positions of the halo per step */
        const bool cross = __shfl_sync(0xffffffffu, (u32)(nm & eq & 1ull), 0) != 0;
        if (cross) {
            const u8 v0 = sm[lo - sm_lo];
            u32 p0 = lo;
            for (;;) {
                const bool ok = p0 >= sm_lo + 1u + (u32)lane;
                const bool same = ok && sm[p0 - 1u - (u32)lane - sm_lo] == v0;
                const u32 m = __ballot_sync(0xffffffffu, same);
                const u32 take = m == 0xffffffffu ? 32u : (u32)(__ffs((int)~m) - 1);
                p0 -= take;
                if (take < 32u) break;
            }
            if (tid == 0) {
                if (p0 == sm_lo && sm_lo > 0) redo = true;   /* starts before the halo */
                s_cross_p = p0; has_cross = 1;
            }
        }
    }
    if (redo) atomicOr(&s_redo, 1u);
    if (tid == 0) s_tmp = has_cross;
    __syncthreads();
    has_cross = s_tmp;
    u32 base = has_cross + ex, total = has_cross;
#pragma unroll
    for (int w = 0; w < S2_THREADS / 32; w++) { const u32 t = s_warp[w]; if (w < warp) base += t; total += t; }
    const u32 n_runs = total;
    if (s_redo || n_runs > (u32)RL_CAP) {
        if (tid == 0) {
            const bool dense = !s_redo;                              /* too many runs for the list: the dense coder's */
            if (mode == 0 && job.dense_list && (dense || job.list_takes_redo)) {
                if (dense) atomicAdd(job.dense_count, 1u);
                const u32 at = atomicAdd(job.list_count, 1u); job.dense_list[at] = span;
            } else { const u32 at = atomicAdd(job.redo_count, 1u); job.redo_list[at] = span; }                                         /* long runs: k_streams3 */
        }
        return;
    }
    {
        u32 m = (u32)starts, k = base;
        const u32 rel = s - lo;
        while (m) { const int i = __ffs((int)m) - 1; m &= m - 1u; rs[k++] = (unsigned short)(rel + (u32)i); }
        m = (u32)(starts >> 32);
        while (m) { const int i = __ffs((int)m) - 1; m &= m - 1u; rs[k++] = (unsigned short)(rel + 32u + (u32)i); }
    }
    const u32 nblocks = (n_runs + 31u) >> 5;
    for (u32 k = tid; k < nblocks * nstreams; k += S2_THREADS) { t_last[k] = (unsigned short)RL_NONE; t_bytes[k] = 0; }
    for (u32 k = tid; k < nstreams; k += S2_THREADS) s_first[k] = NONE32;
    __syncthreads();

    const u32 exc_stream = nstreams - 1;
    const u32 scan_lim = hi + 32u < n ? hi + 32u : n;
    const u32 rounds = (n_runs + S2_THREADS - 1) / S2_THREADS;
    /* ---- A */
    for (u32 m = 0; m < rounds; m++) {
        const u32 k = m * S2_THREADS + (u32)tid;
        const bool valid = k < n_runs;
        u32 cls = 0xFFu;
        if (valid) {
            const u32 x0 = (has_cross && k == 0) ? lo : lo + rs[k];
            const u8 v = sm[x0 - sm_lo];
            const u8 l = mode == 0 ? s_lut[v] : (u8)0;
            cls = l == LUT_EXC ? exc_stream : (u32)l;
            /* the run ends at the first position after x0 whose "equals the previous" bit is clear: from the segment masks (no run
             * covers a whole segment here), by bytes only for the up to 32 positions past the span that have no mask */
            const u32 r = x0 - lo, seg = r >> 6;
            u64 z = (~s_eq[seg] >> (r & 63u)) >> 1;
            u32 y;
            if (z) y = x0 + 1u + (u32)(__ffsll((long long)z) - 1);
            else {
                y = lo + 64u * (seg + 1u);
                if (y < hi) y += (u32)(__ffsll((long long)~s_eq[seg + 1u]) - 1);
            }
            if (y >= hi) { y = y < hi ? y : (x0 + 1u > hi ? x0 + 1u : hi); while (y < scan_lim && sm[y - sm_lo] == v) y++; }
            if (y > scan_lim) y = scan_lim;
            rend[k] = (unsigned short)(y - lo);
            rcls[k] = (u8)cls;
        }
        const u32 peers = __match_any_sync(0xffffffffu, cls);
        if (valid && (peers >> lane) <= 1u) t_last[(k >> 5) * nstreams + cls] = (unsigned short)k;       /* highest lane of its group */
    }
    __syncthreads();
    /* ---- P1: last run of the stream BEFORE each block (a warp per stream, a lane per block: "latest entry that is not NONE",
     * exclusive; one thread per stream walking the blocks was 48 dependent shared-memory round trips with the CTA waiting) */
    for (u32 st = warp; st < nstreams; st += S2_THREADS / 32) {
        u32 carry = RL_NONE;
        for (u32 j0 = 0; j0 < nblocks; j0 += 32) {
            const u32 j = j0 + (u32)lane;
            const u32 t = j < nblocks ? (u32)t_last[j * nstreams + st] : RL_NONE;
            u32 incl = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d && incl == RL_NONE) incl = u; }
            u32 excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = RL_NONE;
            if (excl == RL_NONE) excl = carry;
            if (j < nblocks) t_last[j * nstreams + st] = (unsigned short)excl;
            const u32 last = __shfl_sync(0xffffffffu, incl, 31);
            if (last != RL_NONE) carry = last;
        }
        if (lane == 0) s_total[st] = carry;                  /* the stream's last run of the span (for lastpos) */
    }
    __syncthreads();
    /* ---- B */
    for (u32 m = 0; m < rounds; m++) {
        const u32 k = m * S2_THREADS + (u32)tid;
        const bool valid = k < n_runs;
        const u32 cls = valid ? (u32)rcls[k] : 0xFFu;
        const u32 peers = __match_any_sync(0xffffffffu, cls);
        u32 bytes = 0;
        if (valid) {
            const bool crossing = has_cross && k == 0;
            const u32 p = crossing ? s_cross_p : lo + rs[k];
            const u32 r_end = lo + rend[k];
            const u32 stop = r_end < hi ? r_end : hi;
            if (cls == exc_stream && mode == 0) {
                const u32 a = p > lo ? p : lo;
                bytes = 5u * (stop - a);                      /* one record per position of the run inside the span */
            } else {
                u32 dm = RL_NONE;
                if (!crossing) {
                    const u32 lower = peers & ((1u << lane) - 1u);
                    u32 kp = lower ? (k - (u32)lane + (u32)(31 - __clz((int)lower))) : (u32)t_last[(k >> 5) * nstreams + cls];
                    if (kp != RL_NONE) dm = p - (lo + rend[kp]);           /* p - previous position of the stream - 1 */
                    else if (p == 0) dm = 0;
                    else s_first[cls] = p;                                 /* first of the stream in the span: sized by k_layout */
                    if (dm != RL_NONE) bytes += distance_len(dm);
                }
                rdm[k] = (unsigned short)dm;
                if (p == 0 && r_end > 1u && 1u < hi) bytes += 1;           /* Q16 */
                u32 head = p + (p == 0 ? 2u : 1u);
                if (head < lo) head += ((lo - head + 31u) / 32u) * 32u;
                if (head < stop) bytes += (stop - head + 31u) / 32u;
            }
        }
        /* byte offset inside (block, stream): the bytes of the lower lanes of the same stream.  Runs of up to 7 bytes (nearly all:
         * a distance token and a few length tokens) by population counts over the bit planes of the byte counts; else an exclusive
         * scan per stream present in the warp */
        u32 myoff = 0;
        u32 todo = 0;
        if (__any_sync(0xffffffffu, bytes >= 8u)) todo = __ballot_sync(0xffffffffu, valid);
        else {
            const u32 lower = peers & ((1u << lane) - 1u);
            const u32 b1 = __ballot_sync(0xffffffffu, bytes & 1u), b2 = __ballot_sync(0xffffffffu, bytes & 2u), b4 = __ballot_sync(0xffffffffu, bytes & 4u);
            myoff = (u32)__popc(lower & b1) + 2u * (u32)__popc(lower & b2) + 4u * (u32)__popc(lower & b4);
            if (valid && (peers >> lane) <= 1u) t_bytes[(k >> 5) * nstreams + cls] = (unsigned short)(myoff + bytes);
        }
        while (todo) {
            const int leader = __ffs((int)todo) - 1;
            const u32 c0 = __shfl_sync(0xffffffffu, cls, leader);
            const bool in = valid && cls == c0;
            const u32 members = __ballot_sync(0xffffffffu, in);
            u32 tot; const u32 exs = warp_excl_scan(in ? bytes : 0u, lane, tot);
            if (in) { myoff = exs; if ((members >> lane) <= 1u) t_bytes[(k >> 5) * nstreams + c0] = (unsigned short)tot; }
            todo &= ~members;
        }
        if (valid) roff[k] = (unsigned short)myoff;
    }
    __syncthreads();
    /* ---- P2: block offsets per stream (a warp per stream, a lane per block), directory, slot */
    for (u32 st = warp; st < nstreams; st += S2_THREADS / 32) {
        u32 acc = 0;
        for (u32 j0 = 0; j0 < nblocks; j0 += 32) {
            const u32 j = j0 + (u32)lane;
            const u32 t = j < nblocks ? (u32)t_bytes[j * nstreams + st] : 0u;
            u32 tot; const u32 exs = warp_excl_scan(t, lane, tot);
            if (j < nblocks) t_bytes[j * nstreams + st] = (unsigned short)(acc + exs);
            acc += tot;
        }
        if (lane == 0) {
            const u32 kl = s_total[st];
            u32 lastpos = NONE32;
            if (kl != RL_NONE && !(st == exc_stream && mode == 0)) { const u32 r_end = lo + rend[kl]; lastpos = (r_end < hi ? r_end : hi) - 1u; }
            s_total[st] = acc;
            SpanDir d; d.bytes = acc; d.slot_off = 0; d.firstpos = s_first[st]; d.lastpos = lastpos;
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
        s_tmp = acc;
    }
    __syncthreads();
    if (s_slot == ~0ull || s_tmp == 0) return;
    for (u32 st = tid; st < nstreams; st += S2_THREADS) job.dir[(size_t)span * nstreams + st].slot_off = s_base[st];
    /* ---- C */
    u8* slot = job.slots + s_slot;
    for (u32 k = tid; k < n_runs; k += S2_THREADS) {
        const u32 cls = rcls[k];
        const bool crossing = has_cross && k == 0;
        const u32 p = crossing ? s_cross_p : lo + rs[k];
        const u32 r_end = lo + rend[k];
        const u32 stop = r_end < hi ? r_end : hi;
        u8* o = slot + s_base[cls] + t_bytes[(k >> 5) * nstreams + cls] + roff[k];
        if (cls == exc_stream && mode == 0) {
            const u8 v = sm[(p > lo ? p : lo) - sm_lo];
            for (u32 q = p > lo ? p : lo; q < stop; q++) { o[0] = v; o[1] = (u8)q; o[2] = (u8)(q >> 8); o[3] = (u8)(q >> 16); o[4] = (u8)(q >> 24); o += 5; }
            continue;
        }
        const u32 dm = rdm[k];
        if (dm != RL_NONE) {
            if (dm < 128u) *o++ = (u8)dm;
            else { o[0] = (u8)(0x80u | (dm >> 8)); o[1] = (u8)dm; o += 2; }
        }
        if (p == 0 && r_end > 1u && 1u < hi) *o++ = 0x00;
        u32 head = p + (p == 0 ? 2u : 1u);
        if (head < lo) head += ((lo - head + 31u) / 32u) * 32u;
        for (; head < stop; head += 32u) { const u32 len = r_end - head < 32u ? r_end - head : 32u; *o++ = (u8)(0xC0u | (len - 1u)); }
    }
}

}  // namespace rpq
