/*
 * rpq_streams6.cuh - k_streams6: the position-stream coder (reference src/rfqcodec.cpp:625-765) for DENSE spans: quality columns
 * with ~40 values (BGI-SEQ, older Illumina), where nearly every position starts a run and emits one token byte into one of ~40
 * streams.  It replaces k_streams5, which kept k_streams4's run bookkeeping in position space: per-position arrays (112 KB, one
 * CTA of 1024 threads per SM) and a [block][stream] table swept by five phases - 480 thread instructions per position, 1.3 % of
 * the HBM roofline (profiles/README.md r01_v11).
 *
 * Here a token is owned by the POSITION that heads it, not by its run, and a warp walks its own 2048 positions of the span 32 at a
 * time with everything it needs in registers and a [warp][stream] table:
 *   - "equals the previous position" as one ballot per step (and the next step's, for run lengths up to 32 ahead);
 *   - the start of the run a position lies in = highest run start at or below its lane, else carried from the step before;
 *   - token heads: a run start (distance token, sized from the previous occurrence of the value: the highest lower lane of its
 *     __match_any_sync group, else the table), position 1 of a run that starts at position 0 (Q16), every 32nd position of a
 *     run after that (length token), every position of a value that is not in the header (5-byte exception record);
 *   - byte offsets inside (warp, stream): population counts of the size bits over the lower lanes of the group.
 * Two passes (count, write) with a per-stream pass over the eight warps in between, which sizes the first distance token of
 * every warp but the stream's first one in the span (that one is left to k_layout, as before).  Same contract as k_streams3 / 4
 * (SpanDir, slots).  25 KB of shared memory per CTA of 256 threads.  A span whose first run starts before the staged halo goes to
 * k_streams3 (redo list).  Quality streams only (mode 0).
 */
#pragma once
#include "rpq_streams4.cuh"

namespace rpq {

constexpr int S6_THREADS = 256;
constexpr int S6_WARPS = S6_THREADS / 32;
constexpr u32 S6_SUB = ST_SPAN / S6_WARPS;              /* positions per warp */
constexpr u32 S6_NS = MAX_BINS + 2;

__host__ __device__ inline size_t streams6_smem() {
    return (size_t)ST_SPAN + 2 * ST_HALO + 16 + 2 * (SQ_CAP + 1) * sizeof(u32) + 16;
}

struct S6Tables {
    u32 last[S6_WARPS][S6_NS];      /* last position of the stream's value in the warp's positions so far */
    u32 cnt[S6_WARPS][S6_NS];       /* bytes of the stream in the warp's positions so far (without the warp's deferred first token) */
    u32 first[S6_WARPS][S6_NS];     /* position of the warp's first distance token of the stream, if nothing in the warp precedes it */
    u32 wbase[S6_WARPS][S6_NS];     /* where the warp's bytes of the stream start inside the span's bytes of the stream */
    u32 fdm[S6_WARPS][S6_NS];       /* distance - 1 of that first token once the warps before are known (NONE32: left to k_layout) */
};

/* one pass over the warp's positions.  WRITE: bytes go to slot + s_base[stream] + wbase + offset; else they are counted. */
template <bool WRITE>
__device__ __forceinline__ void s6_walk(const u8* sm, u32 sm_lo, u32 n, u32 wlo, u32 whi, u32 carry_p0, const u8* s_lut, u32 exc_stream,
                                        S6Tables& T, int warp, int lane, u8* slot, const u32* s_base, u32& run_starts) {
    const u32 lt = (1u << lane) - 1u;
    auto byte_at = [&](u32 p) -> u32 { return sm[p - sm_lo]; };
    auto eq_at = [&](u32 p) -> bool { return p > 0 && p < n && byte_at(p) == byte_at(p - 1); };
    u32 eqm = __ballot_sync(0xffffffffu, eq_at(wlo + (u32)lane));
    for (u32 base = wlo; base < whi; base += 32) {
        const u32 p = base + (u32)lane;
        const bool live = p < whi;
        const u32 eqm_next = __ballot_sync(0xffffffffu, eq_at(p + 32u));
        const u32 v = live ? byte_at(p) : 0u;
        const u8 l = live ? s_lut[v] : LUT_SKIP;
        const u32 cls = l == LUT_EXC ? exc_stream : (u32)l;                /* LUT_SKIP (0xFF): the major quality, no stream */
        const u32 peers = __match_any_sync(0xffffffffu, live ? cls : 0x100u + (u32)lane);
        const u32 lower = peers & lt;
        /* the run this position lies in starts at p0 */
        const u32 starts = ~eqm;
        const u32 ls = starts & (lt | (1u << lane));
        const u32 p0 = ls ? base + (u32)(31 - __clz((int)ls)) : carry_p0;
        u32 size = 0, tok = 0;
        bool deferred = false;
        if (live && l != LUT_SKIP) {
            if (l == LUT_EXC) { size = 5; }
            else if (p == p0) {
                /* distance token (src/rfqcodec.cpp:648-676): from the previous occurrence of the value */
                u32 prev = lower ? base + (u32)(31 - __clz((int)lower)) : T.last[warp][cls];
                if (p == 0) { tok = 0; size = 1; }
                else if (prev != NONE32) {
                    const u32 dm = p - prev - 1u;
                    if (dm < 128u) { tok = dm; size = 1; }
                    else if (dm < (1u << 14)) { tok = (0x80u | (dm >> 8)) | ((dm & 0xFFu) << 8); size = 2; }
                    else { tok = (0xE0u | (dm >> 24)) | (((dm >> 16) & 0xFFu) << 8) | (((dm >> 8) & 0xFFu) << 16) | ((dm & 0xFFu) << 24); size = 4; }
                } else {
                    deferred = true;                                        /* nothing of the stream before it in this warp */
                    if (WRITE) {
                        const u32 dm = T.fdm[warp][cls];
                        if (dm != NONE32) {                                 /* sized between the passes; it leads the warp's bytes of the stream */
                            u8* o = slot + s_base[cls] + T.wbase[warp][cls];
                            if (dm < 128u) o[0] = (u8)dm;
                            else if (dm < (1u << 14)) { o[0] = (u8)(0x80u | (dm >> 8)); o[1] = (u8)dm; }
                            else { o[0] = (u8)(0xE0u | (dm >> 24)); o[1] = (u8)(dm >> 16); o[2] = (u8)(dm >> 8); o[3] = (u8)dm; }
                        }
                    } else T.first[warp][cls] = p;
                }
            } else {
                const u32 off = p - p0, s0 = p0 == 0 ? 2u : 1u;
                if (p0 == 0 && off == 1u) { tok = 0; size = 1; }            /* Q16: position 1 of a run that starts at position 0 */
                else if (off >= s0 && ((off - s0) & 31u) == 0u) {
                    /* length token: the positions of the run from here, at most 32 */
                    const u64 win = (((u64)eqm_next << 32) | eqm) >> (lane + 1);
                    const u32 follow = (u32)(__ffsll((long long)~win) - 1);
                    const u32 len = follow >= 31u ? 32u : follow + 1u;
                    tok = 0xC0u | (len - 1u); size = 1;
                }
            }
        }
        if (!WRITE) run_starts += (u32)__popc(__ballot_sync(0xffffffffu, live && l != LUT_SKIP && p == p0));
        /* offsets inside (warp, stream): the bytes of the lower lanes of the group */
        const u32 b1 = __ballot_sync(0xffffffffu, size & 1u), b2 = __ballot_sync(0xffffffffu, size & 2u), b4 = __ballot_sync(0xffffffffu, size & 4u);
        const bool streamed = live && l != LUT_SKIP;
        u32 at = 0;
        if (streamed) at = T.cnt[warp][cls];
        __syncwarp();
        if (streamed) {
            const u32 myoff = (u32)__popc(lower & b1) + 2u * (u32)__popc(lower & b2) + 4u * (u32)__popc(lower & b4);
            if (WRITE && size) {
                u8* o = slot + s_base[cls] + T.wbase[warp][cls] + at + myoff;
                if (size == 5u) { o[0] = (u8)v; o[1] = (u8)p; o[2] = (u8)(p >> 8); o[3] = (u8)(p >> 16); o[4] = (u8)(p >> 24); }
                else { o[0] = (u8)tok; if (size >= 2u) o[1] = (u8)(tok >> 8); if (size == 4u) { o[2] = (u8)(tok >> 16); o[3] = (u8)(tok >> 24); } }
            }
            if ((peers >> lane) <= 1u) {                                    /* the highest lane of the group: the value's last position so far */
                T.cnt[warp][cls] = at + myoff + size;
                T.last[warp][cls] = p;
            }
        }
        __syncwarp();
        (void)deferred;
        if (starts) carry_p0 = base + (u32)(31 - __clz((int)starts));
        eqm = eqm_next;
    }
}

__global__ void __launch_bounds__(S6_THREADS) k_streams6(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk, const u32* __restrict__ list) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ S6Tables T;
    __shared__ u32 s_total[S6_NS], s_base[S6_NS];
    __shared__ u64 s_slot;
    __shared__ u32 s_tmp, s_redo, s_runs, s_bytes;
    const u32 span = list ? list[blockIdx.x] : blockIdx.x;
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    const u32 n = ck.total_len;
    const u32 lo = (span - job.span_first[c]) * ST_SPAN;
    const u32 hi = lo + ST_SPAN < n ? lo + ST_SPAN : n;
    const u32 sm_lo = lo >= ST_HALO ? lo - ST_HALO : 0, sm_hi = hi + ST_HALO < n ? hi + ST_HALO : n;
    const u32 nstreams = job.nstreams;
    const u32 exc_stream = nstreams - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u8* sm = dyn;
    s_lut[tid] = h.lut[tid];
    if (tid == 0) { s_redo = 0; s_runs = 0; }
    for (u32 k = tid; k < 16; k += S6_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO + 16)) sm[sm_hi - sm_lo + k] = h.major;
    {
        u32* s_off = reinterpret_cast<u32*>(dyn + ST_SPAN + 2 * ST_HALO + 16);
        if (!stage_quality_flat(b, ck, sm_lo, sm_hi, sm, job.span_read0[span], s_off, s_off + SQ_CAP + 1, &s_tmp))
            stage_quality_words(b, ck, sm_lo, sm_hi, sm, job.span_read0[span]);
    }
    for (u32 k = tid; k < S6_WARPS * S6_NS; k += S6_THREADS) { (&T.last[0][0])[k] = NONE32; (&T.cnt[0][0])[k] = 0; (&T.first[0][0])[k] = NONE32; (&T.fdm[0][0])[k] = NONE32; (&T.wbase[0][0])[k] = 0; }
    __syncthreads();

    /* ---- this warp's positions, and the start of the run that reaches into them */
    const u32 wlo = lo + (u32)warp * S6_SUB < hi ? lo + (u32)warp * S6_SUB : hi;
    const u32 whi = wlo + S6_SUB < hi ? wlo + S6_SUB : hi;
    u32 carry_p0 = wlo;
    if (wlo < whi && wlo > 0) {
        /* walk back from wlo - 1 while the bytes stay equal: 32 positions per step */
        const u32 v0 = sm[wlo - sm_lo];
        u32 q = wlo;                                                       /* the run is known to hold [q, wlo] */
        bool open = sm[wlo - 1 - sm_lo] == v0;
        while (open && q > sm_lo) {
            const u32 cand = q - 1u - (u32)lane;                            /* lanes look at q-1, q-2, ... */
            const bool same = q >= 1u + (u32)lane && cand >= sm_lo && sm[cand - sm_lo] == v0;
            const u32 m = __ballot_sync(0xffffffffu, same);
            const u32 run = (u32)(__ffs((int)~m) - 1);                      /* leading lanes that are equal (32 if all: ffs(0) - 1 wraps to ~0) */
            const u32 take = m == 0xffffffffu ? 32u : run;
            q -= take;
            if (take < 32u) open = false;
        }
        carry_p0 = q;
        if (q == sm_lo && sm_lo > 0 && lane == 0) atomicOr(&s_redo, 1u);              /* the run may start before the staged halo: k_streams3's */
    }
    __syncthreads();
    if (s_redo) {
        if (tid == 0) { const u32 at = atomicAdd(job.redo_count, 1u); job.redo_list[at] = span; }
        return;
    }
    /* ---- pass 1: count */
    u32 run_starts = 0;
    s6_walk<false>(sm, sm_lo, n, wlo, whi, carry_p0, s_lut, exc_stream, T, warp, lane, nullptr, nullptr, run_starts);
    if (lane == 0 && run_starts) atomicAdd(&s_runs, run_starts);
    __syncthreads();
    if (!list && tid == 0 && s_runs <= (u32)RL_CAP) atomicAdd(job.dense_count, 1u);     /* a span k_streams4 would have coded itself */
    /* ---- per stream, over the warps: the first distance token of every warp but the first with anything, the warps' places, the
     * directory */
    for (u32 st = tid; st < nstreams; st += S6_THREADS) {
        u32 running = NONE32, base = 0, span_first = NONE32;
        for (int w = 0; w < S6_WARPS; w++) {
            const u32 f = T.first[w][st];
            u32 fsize = 0;
            if (f != NONE32) {
                if (running != NONE32) { const u32 dm = f - running - 1u; T.fdm[w][st] = dm; fsize = distance_len(dm); }
                else span_first = f;                                        /* the stream's first distance token of the span: k_layout sizes it */
            }
            T.wbase[w][st] = base;
            T.first[w][st] = fsize;                                         /* from here on: the bytes the warp's own first token takes */
            base += fsize + T.cnt[w][st];
            if (T.last[w][st] != NONE32) running = T.last[w][st];
        }
        s_total[st] = base;
        SpanDir d; d.bytes = base; d.slot_off = 0; d.firstpos = st == exc_stream ? NONE32 : span_first; d.lastpos = st == exc_stream ? NONE32 : running;
        d.dst = 0; d.first_tok = 0; d.first_len = 0; d.pad = 0;
        job.dir[(size_t)span * nstreams + st] = d;
    }
    __syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (u32 st = 0; st < nstreams; st++) { s_base[st] = acc; acc += s_total[st]; }
        const u64 at = atomicAdd(job.slot_cursor, (u64)acc);
        job.span_slot[span] = at;
        if (at + acc > job.slot_cap) { atomicOr(job.overflow, 1u); s_slot = ~0ull; } else s_slot = at;
        s_bytes = acc;
    }
    /* the tables start over for the second pass; a warp's own first token lies in front of its other bytes */
    for (u32 k = tid; k < S6_WARPS * S6_NS; k += S6_THREADS) { (&T.last[0][0])[k] = NONE32; (&T.cnt[0][0])[k] = (&T.first[0][0])[k]; }
    __syncthreads();
    if (s_slot == ~0ull || s_bytes == 0) return;
    for (u32 st = tid; st < nstreams; st += S6_THREADS) job.dir[(size_t)span * nstreams + st].slot_off = s_base[st];
    /* ---- pass 2: the bytes */
    s6_walk<true>(sm, sm_lo, n, wlo, whi, carry_p0, s_lut, exc_stream, T, warp, lane, job.slots + s_slot, s_base, run_starts);
}

}  // namespace rpq
