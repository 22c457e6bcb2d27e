/*
 * rpq_streams5.cuh - k_streams5: the position-stream coder (reference src/rfqcodec.cpp:625-765) for DENSE spans: quality columns
 * with ~40 values (BGI-SEQ, older Illumina), where nearly every position starts a run.  k_streams4's list of runs (1536 per
 * 16 K-position span) overflows there, and k_streams3, which took such spans, keeps per-thread per-stream state in shared memory:
 * 180 KB for 39 streams, one CTA of 8 warps per SM, one long dependent chain per thread (15 ms of a 19 ms encode at 0.94 GB,
 * profiles/README.md r01_v10).
 *
 * This is k_streams4's algorithm in POSITION space: a run is addressed by the span-relative position of its start (no list, no
 * capacity), blocks are 32 consecutive positions, and a CTA of 1024 threads (one per SM: 112 KB of per-position arrays) sweeps the
 * span's 16 K positions in 16 rounds per phase:
 *   A   value, stream, run end (from the "equals the previous position" masks); __match_any_sync groups a warp's 32 positions by
 *       stream; last run start of every stream per block
 *   P1  per stream (a warp each): last run start of the stream BEFORE each block
 *   B   token bytes of every run start (distance token sized from the predecessor's end, Q16 byte, one length token per 32
 *       positions headed inside the span); offsets inside (block, stream) from the lower peers
 *   P2  per stream: block offsets, SpanDir, one bump allocation per span
 *   C   the bytes
 * Same contract as k_streams3 / k_streams4 (SpanDir, slots, first distance token of a stream deferred to k_layout).  Spans it
 * cannot describe (a run covering a whole 64-position segment, a crossing run that starts before the halo, a stream of more than
 * 65535 bytes) go to k_streams3 through the redo list.  Quality streams only (mode 0).
 */
#pragma once
#include "rpq_streams4.cuh"

namespace rpq {

constexpr int S5_THREADS = 1024;
constexpr u32 S5_BLOCKS = ST_SPAN / 32;
constexpr u32 S5_ROUNDS = ST_SPAN / S5_THREADS;

/* one [block][stream] table holds first the predecessors, then the byte counts; with room for two (up to ~44 streams) the second
 * one is cleared once instead of block by block between a warp's reads and writes */
__host__ __device__ inline bool streams5_two_tables(u32 nstreams) {
    return (size_t)ST_SPAN + 2 * ST_HALO + 16 + (size_t)ST_SPAN * (3 * sizeof(unsigned short) + 1) + 2 * (size_t)S5_BLOCKS * nstreams * sizeof(unsigned short) <= 216u * 1024u;
}
__host__ __device__ inline size_t streams5_smem(u32 nstreams) {
    return (size_t)ST_SPAN + 2 * ST_HALO + 16 + (size_t)ST_SPAN * (3 * sizeof(unsigned short) + 1) +
           (streams5_two_tables(nstreams) ? 2u : 1u) * (size_t)S5_BLOCKS * nstreams * sizeof(unsigned short);
}

__global__ void __launch_bounds__(S5_THREADS) k_streams5(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk, const u32* __restrict__ list) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ u32 s_total[MAX_BINS + 2];
    __shared__ u32 s_base[MAX_BINS + 2];
    __shared__ u32 s_first[MAX_BINS + 2];
    __shared__ u64 s_start[ST_SPAN / 64], s_eq[ST_SPAN / 64];
    __shared__ u64 s_slot;
    __shared__ u32 s_tmp, s_redo, s_cross_p, s_cross, s_runs;
    const u32 span = list ? list[blockIdx.x] : blockIdx.x;
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    const u32 n = ck.total_len;
    const u32 lo = (span - job.span_first[c]) * ST_SPAN;
    const u32 hi = lo + ST_SPAN < n ? lo + ST_SPAN : n;
    const u32 sm_lo = lo >= ST_HALO ? lo - ST_HALO : 0, sm_hi = hi + ST_HALO < n ? hi + ST_HALO : n;
    const u32 nstreams = job.nstreams;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u8* sm = dyn;
    u8* arrays = dyn + ST_SPAN + 2 * ST_HALO + 16;
    unsigned short* rend = reinterpret_cast<unsigned short*>(arrays);     /* [ST_SPAN] run end (exclusive, clipped to hi + 32), span relative, by run start */
    unsigned short* roff = rend + ST_SPAN;                                /* [ST_SPAN] byte offset inside (block, stream) */
    unsigned short* rdm = roff + ST_SPAN;                                 /* [ST_SPAN] distance - 1 of the run's distance token, RL_NONE: none / deferred */
    u8* rcls = reinterpret_cast<u8*>(rdm + ST_SPAN);                      /* [ST_SPAN] stream index */
    unsigned short* tab = reinterpret_cast<unsigned short*>(rcls + ST_SPAN);   /* [S5_BLOCKS][nstreams]: A/P1 last run start of the stream in / before the block */
    const bool two = streams5_two_tables(nstreams);
    unsigned short* tbytes = two ? tab + (size_t)S5_BLOCKS * nstreams : tab;     /* B/P2 bytes of the stream in the block / before it (aliases tab if there is no room) */
    if (tid < 256) s_lut[tid] = h.lut[tid];
    if (tid == 0) { s_redo = 0; s_cross = 0; s_runs = 0; }
    for (u32 k = tid; k < 8; k += S5_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO + 16)) sm[sm_hi - sm_lo + k] = h.major;
    {
        u32* s_off = reinterpret_cast<u32*>(arrays);
        if (!stage_quality_flat(b, ck, sm_lo, sm_hi, sm, job.span_read0[span], s_off, s_off + SQ_CAP + 1, &s_tmp))
            stage_quality_words(b, ck, sm_lo, sm_hi, sm, job.span_read0[span]);
    }
    __syncthreads();

    /* ---- masks of 64 positions per thread (the first 256 threads), as in k_streams3 / k_streams4 */
    if (tid < ST_SPAN / 64) {
        const u32 s = lo + (u32)tid * 64u;
        const u32 e = s + 64u < hi ? s + 64u : hi;
        u64 nm = 0, eq = 0;
        if (s < hi) {
            const u32* W = reinterpret_cast<const u32*>(sm + (s - sm_lo));
            const bool major_is_stream = s_lut[h.major] != LUT_SKIP;
            const u32 mmmm = 0x01010101u * h.major;
            u32 prevw = s > 0 ? (u32)sm[s - 1 - sm_lo] << 24 : 0u;
            u32 nmw[2] = {0, 0}, eqw[2] = {0, 0};
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const u32 w0 = W[j], w1 = W[j + 1];
                const u32 e0 = eq_bytes(w0, __funnelshift_l(prevw, w0, 8)), e1 = eq_bytes(w1, __funnelshift_l(w0, w1, 8));
                const u32 n0 = major_is_stream ? 0x80808080u : eq_bytes(w0, mmmm) ^ 0x80808080u, n1 = major_is_stream ? 0x80808080u : eq_bytes(w1, mmmm) ^ 0x80808080u;
                const u32 eb = (((e0 >> 4) | e1) * 0x00204081u) >> 24, nb = (((n0 >> 4) | n1) * 0x00204081u) >> 24;
                eqw[j >> 3] |= eb << (8 * ((j >> 1) & 3)); nmw[j >> 3] |= nb << (8 * ((j >> 1) & 3));
                prevw = w1;
            }
            if (s == 0) eqw[0] &= ~1u;
            nm = (u64)nmw[0] | ((u64)nmw[1] << 32); eq = (u64)eqw[0] | ((u64)eqw[1] << 32);
            const u32 valid = e - s;
            if (valid < 64u) nm &= (1ull << valid) - 1ull;
        }
        s_start[tid] = nm & ~eq;
        s_eq[tid] = eq;
        if (!list) atomicAdd(&s_runs, (u32)__popcll(nm & ~eq));    /* coding every span of the batch: tell the host how many were not dense */
        bool redo = (nm & eq) == ~0ull;                     /* a run covers the whole segment */
        if (tid == 0 && (nm & eq & 1ull)) {                  /* a run crosses into the span: where it starts */
            const u8 v0 = sm[lo - sm_lo];
            u32 p0 = lo;
            while (p0 > sm_lo && sm[p0 - 1 - sm_lo] == v0) p0--;
            if (p0 == sm_lo && sm_lo > 0) redo = true;
            s_cross_p = p0; s_cross = 1;
        }
        if (redo) atomicOr(&s_redo, 1u);
    }
    for (u32 k = tid; k < S5_BLOCKS * nstreams; k += S5_THREADS) { tab[k] = (unsigned short)RL_NONE; if (two) tbytes[k] = 0; }
    for (u32 k = tid; k < nstreams; k += S5_THREADS) s_first[k] = NONE32;
    __syncthreads();
    if (!list && tid == 0 && s_runs <= (u32)RL_CAP) atomicAdd(job.dense_count, 1u);     /* a span k_streams4 would have coded itself */
    if (s_redo) {
        if (tid == 0) { const u32 at = atomicAdd(job.redo_count, 1u); job.redo_list[at] = span; }
        return;
    }
    const u32 has_cross = s_cross;
    const u32 exc_stream = nstreams - 1;
    const u32 scan_lim = hi + 32u < n ? hi + 32u : n;
    auto is_start = [&](u32 k) -> bool { return lo + k < hi && ((((s_start[k >> 6] >> (k & 63u)) & 1ull) != 0) || (k == 0 && has_cross)); };

    /* ---- A */
    for (u32 m = 0; m < S5_ROUNDS; m++) {
        const u32 k = m * S5_THREADS + (u32)tid;
        const bool valid = is_start(k);
        u32 cls = 0xFFu;
        if (valid) {
            const u32 x0 = lo + k;
            const u8 v = sm[x0 - sm_lo];
            const u8 l = s_lut[v];
            cls = l == LUT_EXC ? exc_stream : (u32)l;
            /* the run ends at the first position after x0 whose "equals the previous" bit is clear (no run covers a whole segment) */
            const u32 seg = k >> 6;
            const u64 z = (~s_eq[seg] >> (k & 63u)) >> 1;
            u32 y;
            if (z) y = x0 + 1u + (u32)(__ffsll((long long)z) - 1);
            else {
                y = lo + 64u * (seg + 1u);
                if (y < hi) y += (u32)(__ffsll((long long)~s_eq[seg + 1u]) - 1);
            }
            if (y >= hi) { y = x0 + 1u > hi ? x0 + 1u : hi; while (y < scan_lim && sm[y - sm_lo] == v) y++; }
            if (y > scan_lim) y = scan_lim;
            rend[k] = (unsigned short)(y - lo);
        }
        rcls[k] = (u8)cls;                                   /* 0xFF: no run starts here (what B and C test) */
        const u32 peers = __match_any_sync(0xffffffffu, cls);
        if (valid && (peers >> lane) <= 1u) tab[(k >> 5) * nstreams + cls] = (unsigned short)k;       /* highest lane of its group */
    }
    __syncthreads();
    /* ---- P1: last run start of the stream BEFORE each block; the stream's last run start of the span */
    for (u32 st = warp; st < nstreams; st += S5_THREADS / 32) {
        u32 running = RL_NONE;
        for (u32 j0 = 0; j0 < S5_BLOCKS; j0 += 32) {
            const u32 j = j0 + (u32)lane;
            const u32 t = tab[j * nstreams + st];
            u32 incl = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d && incl == RL_NONE) incl = up; }
            u32 before = __shfl_up_sync(0xffffffffu, incl, 1);          /* last start in the blocks j0 .. j-1 */
            if (lane == 0 || before == RL_NONE) before = running;        /* none there: the last one before this group of 32 blocks */
            tab[j * nstreams + st] = (unsigned short)before;
            const u32 last = __shfl_sync(0xffffffffu, incl, 31);
            if (last != RL_NONE) running = last;
        }
        if (lane == 0) s_total[st] = running;
    }
    __syncthreads();
    /* ---- B */
    for (u32 m = 0; m < S5_ROUNDS; m++) {
        const u32 k = m * S5_THREADS + (u32)tid;
        const u32 cls = rcls[k];
        const bool valid = cls != 0xFFu;
        const u32 peers = __match_any_sync(0xffffffffu, cls);
        const u32 lower = peers & ((1u << lane) - 1u);
        u32 bytes = 0;
        if (valid) {
            const bool crossing = has_cross && k == 0;
            const u32 p = crossing ? s_cross_p : lo + k;
            const u32 r_end = lo + rend[k];
            const u32 stop = r_end < hi ? r_end : hi;
            if (cls == exc_stream) {
                const u32 a = p > lo ? p : lo;
                bytes = 5u * (stop - a);                      /* one record per position of the run inside the span */
            } else {
                u32 dm = RL_NONE;
                if (!crossing) {
                    const u32 kp = lower ? (k - (u32)lane + (u32)(31 - __clz((int)lower))) : (u32)tab[(k >> 5) * nstreams + cls];
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
        if (!two) {
            __syncwarp();                                    /* every lane has read its predecessor from the table: the block's cells now count bytes */
            for (u32 st = lane; st < nstreams; st += 32) tab[(k >> 5) * nstreams + st] = 0;
            __syncwarp();
        }
        /* bytes of the lower lanes of the same stream: bit plane by bit plane, one ballot and one population count each; planes
         * nobody uses are skipped (a run's tokens are a few bytes; the records of an exception run - below 128 positions, or the
         * span would not be here - at most 635) */
        u32 myoff = 0;
#pragma unroll
        for (int bit = 0; bit < 10; bit++) {
            const u32 plane = __ballot_sync(0xffffffffu, (bytes >> bit) & 1u);
            if (plane) myoff += (u32)__popc(plane & lower) << bit;
        }
        if (valid) {
            roff[k] = (unsigned short)myoff;
            if ((peers >> lane) <= 1u) tbytes[(k >> 5) * nstreams + cls] = (unsigned short)(myoff + bytes);
        }
    }
    __syncthreads();
    /* ---- P2: block offsets per stream (a warp per stream, a lane per block), directory */
    for (u32 st = warp; st < nstreams; st += S5_THREADS / 32) {
        u32 acc = 0;
        for (u32 j0 = 0; j0 < S5_BLOCKS; j0 += 32) {
            const u32 j = j0 + (u32)lane;
            const u32 t = (u32)tbytes[j * nstreams + st];
            u32 tot; const u32 exs = warp_excl_scan(t, lane, tot);
            tbytes[j * nstreams + st] = (unsigned short)(acc + exs);
            acc += tot;
        }
        if (lane == 0) {
            const u32 kl = s_total[st];
            u32 lastpos = NONE32;
            if (kl != RL_NONE && st != exc_stream) { const u32 r_end = lo + rend[kl]; lastpos = (r_end < hi ? r_end : hi) - 1u; }
            s_total[st] = acc;
            if (acc > 0xFFFFu) atomicOr(&s_redo, 1u);         /* block offsets are 16 bits: such a span is k_streams3's */
            SpanDir d; d.bytes = acc; d.slot_off = 0; d.firstpos = st == exc_stream ? NONE32 : s_first[st]; d.lastpos = lastpos;
            d.dst = 0; d.first_tok = 0; d.first_len = 0; d.pad = 0;
            job.dir[(size_t)span * nstreams + st] = d;
        }
    }
    __syncthreads();
    if (s_redo) {
        if (tid == 0) { const u32 at = atomicAdd(job.redo_count, 1u); job.redo_list[at] = span; }
        return;
    }
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
    for (u32 st = tid; st < nstreams; st += S5_THREADS) job.dir[(size_t)span * nstreams + st].slot_off = s_base[st];
    /* ---- C */
    u8* slot = job.slots + s_slot;
    for (u32 m = 0; m < S5_ROUNDS; m++) {
        const u32 k = m * S5_THREADS + (u32)tid;
        const u32 cls = rcls[k];
        if (cls == 0xFFu) continue;
        const bool crossing = has_cross && k == 0;
        const u32 p = crossing ? s_cross_p : lo + k;
        const u32 r_end = lo + rend[k];
        const u32 stop = r_end < hi ? r_end : hi;
        u8* o = slot + s_base[cls] + tbytes[(k >> 5) * nstreams + cls] + roff[k];
        if (cls == exc_stream) {
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
