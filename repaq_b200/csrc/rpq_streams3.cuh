/*
 * rpq_streams3.cuh - k_streams3: the position-stream coder (reference src/rfqcodec.cpp:625-765 and :420-426), third
 * generation.  Same contract as k_streams2 (S2Tables / SpanDir / slots) but the work is organised around RUNS:
 *
 *   a thread owns 64 consecutive positions and reduces them to two 64-bit masks with SIMD byte compares:
 *       NM  position does not hold the major quality (mode 1: position holds 'N')
 *       EQ  position holds the same byte as the position before it
 *   a run of a stream value is then "lowest set NM bit + the EQ bits that follow"; every lane of a warp handles its k-th
 *   run in the same iteration (convergent), and both passes (count, write) replay the masks from registers.
 *   Staging of the quality bytes is word based (funnel shift / byte reverse), not byte based.
 *
 * v2 classified byte by byte inside a divergent loop: 1.72 G warp instructions for 180 M positions at 6.7 active threads
 * per instruction (profiles/r01_v2_ncu_full_k_streams2.csv).
 */
#pragma once
#include "rpq_streams2.cuh"

namespace rpq {

__device__ __forceinline__ u32 pack4(u32 m) { m &= 0x01010101u; return (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xFu; }
__device__ __forceinline__ int eq_run(u64 eq, int k) { if (k >= 64) return 0; const u64 t = ~(eq >> k); return t ? (__ffsll((long long)t) - 1) : (64 - k); }

/* quality bytes of the chunk's positions [lo, hi) into sm[pos - lo]; lo is a multiple of 4; word based.
 * `first_rel` = first read that can reach into the window (k_span_reads).  A warp takes every nwarps-th read; its lanes fetch
 * the offsets / lengths / line starts of 32 of them in parallel, then the warp copies those reads one after the other. */
__device__ inline void stage_quality_words(const EncBatchDev& b, const ChunkDev& ck, u32 lo, u32 hi, u8* sm, u32 first_rel) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (hi <= lo) return;
    u32* smw = reinterpret_cast<u32*>(sm);
    /* warp w takes reads first_rel + w + nwarps*k (k = lane within a round of 32): neighbouring reads go to different warps */
    for (u32 base = first_rel + (u32)warp; base < ck.count; base += 32u * nwarps) {
        const u32 rel_l = base + (u32)lane * nwarps;
        u32 off_l = 0xFFFFFFFFu, rl_l = 0, qw_l = 0;
        if (rel_l < ck.count) { const u32 i = ck.first + rel_l; off_l = b.qualoff[i]; rl_l = b.rlen[i]; qw_l = b.loc[i].w; }
        if (__shfl_sync(0xffffffffu, off_l, 0) >= hi) break;
        for (int k = 0; k < 32; k++) {
            const u32 off = __shfl_sync(0xffffffffu, off_l, k);
            if (off >= hi) break;                                         /* also ends at the padding lanes (off = ~0) */
            const u32 rl = __shfl_sync(0xffffffffu, rl_l, k), qstart = __shfl_sync(0xffffffffu, qw_l, k);
            if (off + rl <= lo) continue;
            const u32 rel = base + (u32)k * nwarps;
            const u32 f = (b.is_pe && b.two_files) ? (rel & 1u) : 0u;      /* chunks start at an even read: file = parity */
            const u8* q = b.t[f].text + qstart;
            const bool rev = ck.interleaved && (rel & 1u);
            const u32 d0 = (off > lo ? off : lo) - lo, d1 = (off + rl < hi ? off + rl : hi) - lo;      /* shared range [d0, d1) */
            const u32 w0 = (d0 + 3u) & ~3u, w1 = d1 & ~3u;
            auto src = [&](u32 x) -> u8 { const u32 j = x + lo - off; return rev ? q[rl - 1 - j] : q[j]; };
            if (w0 >= w1) { for (u32 x = d0 + lane; x < d1; x += 32) sm[x] = src(x); continue; }
            if ((u32)lane < w0 - d0) sm[d0 + lane] = src(d0 + lane);
            if ((u32)lane < d1 - w1) sm[w1 + lane] = src(w1 + lane);
            for (u32 x = w0 + 4u * lane; x < w1; x += 128u) {
                const u32 j = x + lo - off;                       /* source index of the word's first byte */
                const u8* g = rev ? q + (rl - 4u - j) : q + j;
                const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
                const u32* al = reinterpret_cast<const u32*>(ga & ~(uintptr_t)3);
                const u32 sh = (u32)(ga & 3u) * 8u;
                u32 w = al[0];
                if (sh) w = __funnelshift_r(w, al[1], sh);
                if (rev) w = __byte_perm(w, 0, 0x0123);
                smw[x >> 2] = w;
            }
        }
    }
}

/*
 * Flat staging: the reads reaching into [lo, hi) go into a small table (offset of the first quality, start of the quality line);
 * a thread then fills 16 bytes at a time: one binary search per 16 bytes, per word an unaligned 4-byte fetch (two aligned loads and a
 * funnel shift, byte-reversed on the reverse strand).  Every lane works on every step, whatever the read lengths are
 * (stage_quality_words gives a read to a warp: 150-base reads keep 38 of 64 lanes busy and pay ~60 instructions of per-read set-up).
 * Returns false when the table is too small; the caller then uses stage_quality_words.
 */
constexpr int SQ_CAP = 383;                        /* reads per staged window the table can hold */
__device__ inline bool stage_quality_flat(const EncBatchDev& b, const ChunkDev& ck, u32 lo, u32 hi, u8* sm, u32 first_rel, u32* s_off, u32* s_q, u32* s_n) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (hi <= lo) return true;
    if (tid == 0) *s_n = 0;
    __syncthreads();
    for (u32 k = tid; k <= (u32)SQ_CAP; k += nthreads) {
        const u32 rel = first_rel + k;
        u32 off = ck.total_len, qs = 0;
        if (rel < ck.count) { const u32 i = ck.first + rel; off = b.qualoff[i]; qs = b.loc[i].w; }
        s_off[k] = off;
        if (k < (u32)SQ_CAP) s_q[k] = qs;
        if (rel < ck.count && off < hi) atomicMax(s_n, k + 1u);
    }
    __syncthreads();
    const u32 n = *s_n;                                  /* reads 0..n-1 start below hi; s_off[n] >= hi is the end of the last one */
    if (n > (u32)SQ_CAP) return false;
    u32* smw = reinterpret_cast<u32*>(sm);
    const bool pe_files = b.is_pe && b.two_files;
    const bool il = ck.interleaved != 0;
    /* pass 1: the groups of 16 positions that lie inside one read.  The read of a group's first position is guessed from the mean read length
     * (exact when all reads are equally long) and corrected by stepping. */
    const u32 ngroups = (hi - lo + 15u) >> 4;
    const u32 off0 = s_off[0];
    const float per_pos = n ? (float)n / (float)(s_off[n] - off0) : 0.f;
    for (u32 g = tid; g < ngroups; g += nthreads) {
        const u32 pos0 = lo + 16u * g;
        u32 r = (u32)((float)(pos0 - off0) * per_pos);
        if (r >= n) r = n - 1u;
        while (s_off[r] > pos0) r--;
        u32 nxt = s_off[r + 1];
        while (pos0 >= nxt) { r++; nxt = s_off[r + 1]; }
        u32 off = s_off[r];
        const u8* q = b.t[pe_files ? ((first_rel + r) & 1u) : 0u].text + s_q[r];
        bool rev = il && ((first_rel + r) & 1u);
        if (pos0 + 16u <= nxt && pos0 + 16u <= hi) {
            /* the whole group lies inside one read (eight or nine of ten groups of 150-base reads): five aligned words, four funnel
             * shifts, one 128-bit store; the reverse strand reads the 16 bytes that end where the group's first position lies */
            const u32 j = pos0 - off;
            const uintptr_t ga = reinterpret_cast<uintptr_t>(rev ? q + ((nxt - off) - 16u - j) : q + j);
            const u32* al = reinterpret_cast<const u32*>(ga & ~(uintptr_t)3);
            const u32 sh = (u32)(ga & 3u) * 8u;
            const u32 a0 = al[0], a1 = al[1], a2 = al[2], a3 = al[3], a4 = al[4];
            const u32 w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh), w3 = __funnelshift_r(a3, a4, sh);
            uint4 v;
            if (rev) v = make_uint4(__byte_perm(w3, 0, 0x0123), __byte_perm(w2, 0, 0x0123), __byte_perm(w1, 0, 0x0123), __byte_perm(w0, 0, 0x0123));
            else v = make_uint4(w0, w1, w2, w3);
            *reinterpret_cast<uint4*>(smw + ((pos0 - lo) >> 2)) = v;
            continue;
        }
    }
    /* pass 2: the groups that hold the end of a read (bytes of two or more reads), and the partial group at hi: a thread per read
     * end, byte by byte.  (Taking them inside pass 1, word by word, cost as much as the whole of pass 1: one lane in ten had such a
     * group, so every warp walked the slow path with two or three lanes.) */
    for (u32 r = tid; r <= n; r += nthreads) {
        const u32 bnd = r < n ? s_off[r + 1] : hi;                 /* end of read r; r == n: the end of the window */
        if (r < n && bnd >= hi) continue;                          /* beyond the window (the window's end is r == n's) */
        if (bnd <= lo) continue;
        const u32 gpos = lo + ((bnd - lo) & ~15u);
        if (gpos == bnd) continue;                                 /* the read ends where a group ends */
        u32 r2 = r < n ? r : n - 1u;
        for (u32 pp = gpos; pp < gpos + 16u && pp < hi; pp++) {
            while (r2 > 0 && s_off[r2] > pp) r2--;
            while (pp >= s_off[r2 + 1]) r2++;
            const u32 off = s_off[r2], rl = s_off[r2 + 1] - off, j = pp - off;
            const u32 rel = first_rel + r2;
            const u8* q = b.t[pe_files ? (rel & 1u) : 0u].text + s_q[r2];
            sm[pp - lo] = (il && (rel & 1u)) ? q[rl - 1u - j] : q[j];
        }
    }
    return true;
}

struct RunCtx {
    const u8* sm; u32 sm_lo, n, lo, s, e;
    const u8* lut; u32 mode, nstreams; int tid;
};

/* replay the runs of one 64-position segment.  WRITE: emit token bytes into the slot.  Otherwise count them and keep up to
 * seven token bytes per (stream, thread) in `tok` (byte 7 = how many), so that the write pass is a plain copy; returns false
 * when some stream of the thread did not fit (the thread then replays its runs with WRITE). */
template <bool WRITE>
__device__ inline bool s3_runs(const RunCtx& R, u64 nm, const u64 eq, const S2Tables& T, u64* tok, u8* slot, const EncBatchDev& b, const ChunkDev& ck) {
    auto at = [&](u32 p) -> u8 { return p >= R.sm_lo ? R.sm[p - R.sm_lo] : stream_byte_slow(b, ck, R.mode, p); };
    const u32 exc_stream = R.nstreams - 1;
    bool fits = true;
    while (nm) {
        const int i = __ffsll((long long)nm) - 1;
        const u32 p = R.s + (u32)i;
        const u8 v = R.sm[p - R.sm_lo];
        int lin = 1 + eq_run(eq, i + 1);                       /* positions of the run inside the 64-bit window */
        if ((u32)i + (u32)lin > R.e - R.s) lin = (int)(R.e - R.s) - i;
        nm &= ~((lin >= 64 ? ~0ull : ((1ull << lin) - 1ull)) << i);
        const u8 cls = R.mode == 0 ? R.lut[v] : (u8)0;
        if (cls == LUT_EXC) {
            /* every position of the run inside the segment is an exception record {q, u32 LE pos} */
            const u32 idx = exc_stream * S2_THREADS + R.tid;
            u32 off = T.cnt[idx];
            for (u32 q = p; q < p + (u32)lin; q++) {
                if (WRITE) { u8* o = slot + off; o[0] = v; o[1] = (u8)q; o[2] = (u8)(q >> 8); o[3] = (u8)(q >> 16); o[4] = (u8)(q >> 24); }
                off += 5;
            }
            T.cnt[idx] = off;
            fits = false;
            continue;
        }
        if (cls == LUT_SKIP) continue;                          /* only when the major quality is not a stream: never set in NM */
        const u32 idx = (u32)cls * S2_THREADS + R.tid;
        /* run end, as far as the tokens headed in this segment need it: at most e + 32 */
        u32 r_end = p + (u32)lin;
        if (r_end == R.e) { const u32 lim = R.e + 32u < R.n ? R.e + 32u : R.n; while (r_end < lim && at(r_end) == v) r_end++; }
        const bool crossing = i == 0 && (eq & 1ull);
        u32 off = T.cnt[idx];
        u64 tk = 0; u32 tn = 0;
        if (!WRITE) { const u64 w = tok[idx]; tn = (u32)(w >> 56); tk = w & 0x00FFFFFFFFFFFFFFull; }
        auto put = [&](u32 byte) {
            if (WRITE) slot[off] = (u8)byte;
            else if (tn < 7u) { tk |= (u64)(byte & 0xFFu) << (8u * tn); tn++; }
            else fits = false;
            off++;
        };
        u32 head;
        if (crossing) {
            u32 p0 = p - 1; while (p0 > 0 && at(p0 - 1) == v) p0--;
            const u32 s0 = p0 == 0 ? 2u : 1u;
            head = p0 + s0;
            if (head < p) head += ((p - head + 31u) / 32u) * 32u;
        } else {
            /* distance token at the run start */
            const u32 lastrel = T.last[idx];
            u32 dm1 = 0; bool emit = true;
            if (lastrel != S2_NONE) dm1 = p - (R.lo + lastrel) - 1u;
            else if (p == 0) dm1 = 0;
            else {
                const u32 f = T.first[idx];
                if (WRITE && (f & S2_RESOLVED)) dm1 = T.fdist[idx];
                else { emit = false; if (!WRITE) T.first[idx] = (u16)(p - R.lo); }
            }
            if (emit) {
                if (dm1 < 128u) put(dm1);
                else if (dm1 < (1u << 14)) { put(0x80u | (dm1 >> 8)); put(dm1); }
                else { put(0xE0u | (dm1 >> 24)); put(dm1 >> 16); put(dm1 >> 8); put(dm1); }
            }
            if (p == 0 && r_end > 1 && 1u < R.e) put(0x00);       /* Q16 */
            head = p + (p == 0 ? 2u : 1u);
        }
        const u32 stop = r_end < R.e ? r_end : R.e;
        for (; head < stop; head += 32u) {
            const u32 len = r_end - head < 32u ? r_end - head : 32u;
            put(0xC0u | (len - 1u));
        }
        T.cnt[idx] = off;
        T.last[idx] = (u16)(stop - 1u - R.lo);
        if (!WRITE) tok[idx] = tk | ((u64)tn << 56);
    }
    return fits;
}

__global__ void __launch_bounds__(S2_THREADS) k_streams3(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk, u32 from_redo_list) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ u32 s_total[MAX_BINS + 2];
    __shared__ u32 s_base[MAX_BINS + 2];
    __shared__ u64 s_slot;
    __shared__ u32 s_bytes;
    const u32 span = from_redo_list ? job.redo_list[blockIdx.x] : blockIdx.x;     /* the spans k_streams4 passed on, or all of them */
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
    S2Tables T;
    u64* tok = reinterpret_cast<u64*>(dyn + ST_SPAN + 2 * ST_HALO);        /* [nstreams][256] recorded token bytes */
    T.cnt = reinterpret_cast<u32*>(tok + nstreams * S2_THREADS);
    T.first = reinterpret_cast<u16*>(T.cnt + nstreams * S2_THREADS);
    T.last = T.first + nstreams * S2_THREADS;
    T.fdist = T.last + nstreams * S2_THREADS;

    s_lut[tid] = h.lut[tid];
    for (u32 k = tid; k < nstreams * S2_THREADS; k += S2_THREADS) { tok[k] = 0; T.cnt[k] = 0; T.first[k] = (u16)S2_NONE; T.last[k] = (u16)S2_NONE; T.fdist[k] = 0; }
    for (u32 k = tid; k < 8; k += S2_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO)) sm[sm_hi - sm_lo + k] = mode == 0 ? h.major : (u8)0;
    if (mode == 0) {
        u32* s_off = reinterpret_cast<u32*>(T.fdist + nstreams * S2_THREADS);      /* [SQ_CAP + 1] */
        if (!stage_quality_flat(b, ck, sm_lo, sm_hi, sm, job.span_read0[span], s_off, s_off + SQ_CAP + 1, &s_bytes))
            stage_quality_words(b, ck, sm_lo, sm_hi, sm, job.span_read0[span]);
    } else stage_positions(b, h, ck, mode, sm_lo, sm_hi, sm);
    __syncthreads();

    /* ---- masks of the thread's 64 positions */
    const u32 s = lo + (u32)tid * S2_SEG;
    const u32 e = s + S2_SEG < hi ? s + S2_SEG : hi;
    u64 nm = 0, eq = 0;
    if (s < hi) {
        const u32* W = reinterpret_cast<const u32*>(sm + (s - sm_lo));
        const bool major_is_stream = mode == 0 && s_lut[h.major] != LUT_SKIP;
        const u32 mmmm = 0x01010101u * h.major;
        /* eight positions per multiply: SIMD byte compares leave a flag in bit 7 of every byte, pack8f gathers the flags of two
         * words into one byte in position order */
        u32 prevw = s > 0 ? (u32)sm[s - 1 - sm_lo] << 24 : 0u;           /* the byte before the segment in the top byte */
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
        if (s == 0) eqw[0] &= ~1u;                                       /* position 0 has no predecessor */
        nm = (u64)nmw[0] | ((u64)nmw[1] << 32); eq = (u64)eqw[0] | ((u64)eqw[1] << 32);
        const u32 valid = e - s;
        if (valid < 64u) nm &= (1ull << valid) - 1ull;
    }
    RunCtx R; R.sm = sm; R.sm_lo = sm_lo; R.n = n; R.lo = lo; R.s = s; R.e = e; R.lut = s_lut; R.mode = mode; R.nstreams = nstreams; R.tid = tid;
    bool fits = true;
    if (nm) fits = s3_runs<false>(R, nm, eq, T, tok, nullptr, b, ck);
    __syncthreads();

    /* per stream (a warp each): resolve first tokens against earlier segments of the span, exclusive scan of the byte counts */
    for (u32 st = warp; st < nstreams; st += S2_THREADS / 32) {
        const u32 base = st * S2_THREADS + lane * 8;
        u32 lastv[8]; u32 lane_last = S2_NONE;
#pragma unroll
        for (int k = 0; k < 8; k++) { lastv[k] = T.last[base + k]; if (lastv[k] != S2_NONE) lane_last = lastv[k]; }
        u32 incl = lane_last;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d && incl == S2_NONE) incl = t; }
        u32 prev = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) prev = S2_NONE;
        u32 cntv[8]; u32 lane_sum = 0; u32 span_first = S2_NONE;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 cv = T.cnt[base + k];
            const u32 f = T.first[base + k];
            if (f != S2_NONE) {
                if (prev != S2_NONE) { const u32 dm1 = f - prev - 1u; cv += distance_len(dm1); T.first[base + k] = (u16)(f | S2_RESOLVED); T.fdist[base + k] = (u16)dm1; }
                else span_first = f;
            }
            if (lastv[k] != S2_NONE) prev = lastv[k];
            cntv[k] = cv; lane_sum += cv;
        }
        u32 tot; const u32 ex = warp_excl_scan(lane_sum, lane, tot);
        u32 run = ex;
#pragma unroll
        for (int k = 0; k < 8; k++) { T.cnt[base + k] = run; run += cntv[k]; }
        const u32 sf = warp_min(span_first);
        const u32 sl = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            s_total[st] = tot;
            SpanDir d; d.bytes = tot; d.slot_off = 0; d.firstpos = sf == S2_NONE ? NONE32 : lo + sf; d.lastpos = sl == S2_NONE ? NONE32 : lo + sl;
            d.dst = 0; d.first_tok = 0; d.first_len = 0; d.pad = 0;
            job.dir[(size_t)span * nstreams + st] = d;
        }
    }
    __syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (u32 st = 0; st < nstreams; st++) { s_base[st] = acc; acc += s_total[st]; }
        s_bytes = acc;
        const u64 at = atomicAdd(job.slot_cursor, (u64)acc);
        job.span_slot[span] = at;
        if (at + acc > job.slot_cap) { atomicOr(job.overflow, 1u); s_slot = ~0ull; } else s_slot = at;
    }
    __syncthreads();
    if (s_slot == ~0ull) return;
    for (u32 st = tid; st < nstreams; st += S2_THREADS) job.dir[(size_t)span * nstreams + st].slot_off = s_base[st];
    for (u32 k = tid; k < nstreams * S2_THREADS; k += S2_THREADS) { T.cnt[k] += s_base[k / S2_THREADS]; T.last[k] = (u16)S2_NONE; }
    __syncthreads();
    if (!nm || !s_bytes) return;
    u8* slot = job.slots + s_slot;
    if (!fits) { s3_runs<true>(R, nm, eq, T, tok, slot, b, ck); return; }
    /* the write pass of a thread whose tokens were recorded: a first distance token sized by the scan, then the recorded bytes */
    for (u32 st = 0; st < nstreams; st++) {
        const u32 idx = st * S2_THREADS + (u32)tid;
        const u64 w = tok[idx];
        const u32 tn = (u32)(w >> 56);
        u8* o = slot + T.cnt[idx];
        if (T.first[idx] & S2_RESOLVED) {
            const u32 dm1 = T.fdist[idx];
            if (dm1 < 128u) { o[0] = (u8)dm1; o += 1; }
            else { o[0] = (u8)(0x80u | (dm1 >> 8)); o[1] = (u8)dm1; o += 2; }
        }
        for (u32 k = 0; k < tn; k++) o[k] = (u8)(w >> (8u * k));
    }
}

}  // namespace rpq
