/*
 * rpq_streams2.cuh - k_streams2: the position-stream coder (reference src/rfqcodec.cpp:625-765 and :420-426), second
 * generation.  Same contract as k_streams (SpanDir / slots, consumed by k_layout and k_gather) but ONE walk serves all
 * streams: a thread owns 64 consecutive positions of the span, skips words that hold only the major quality with a
 * single compare, and classifies the remaining bytes through the header's bin LUT.  Per-(stream, thread) byte counts,
 * first/last positions live in shared memory; a warp per stream turns them into offsets and resolves, inside the span,
 * every first distance token whose predecessor lies in an earlier segment.
 *
 * v1 walked every segment once per stream (bins x passes over shared memory) and was instruction bound
 * (profiles/r01_v1_ncu_full_k_streams.csv: 1.45 G warp instructions for 180 M positions).
 */
#pragma once
#include "rpq_encode.cuh"

namespace rpq {

constexpr int S2_THREADS = 256;
constexpr int S2_SEG = ST_SPAN / S2_THREADS;     /* 64 positions per thread */
constexpr u32 S2_NONE = 0x7FFFu;                 /* 15-bit "no position" */
constexpr u32 S2_RESOLVED = 0x8000u;             /* first token sized inside the span: pass 2 emits it */

struct S2Tables {
    u32* cnt;        /* [nstreams][256] bytes per (stream, thread); after the scan: write offset inside the slot */
    u16* first;      /* [nstreams][256] span-relative position of the thread's deferred first distance token | flags */
    u16* last;       /* [nstreams][256] span-relative last position of the stream value in the thread's segment */
    u16* fdist;      /* [nstreams][256] distance-1 of a first token resolved inside the span */
};

template <bool WRITE>
__device__ inline void s2_walk(const u8* sm, u32 sm_lo, u32 n, u32 lo, u32 s, u32 e, const u8* lut, u32 mode, u8 major, u32 nstreams,
                               const S2Tables& T, int tid, u8* slot, const EncBatchDev& b, const ChunkDev& ck) {
    /* sm holds positions [sm_lo, sm_lo + staged); every access below stays within [s - look-back, e + 32) which is staged
     * except for look-back beyond the left halo (stream_byte_slow) */
    auto at = [&](u32 p) -> u8 { return p >= sm_lo ? sm[p - sm_lo] : stream_byte_slow(b, ck, mode, p); };
    const u32 mmmm = 0x01010101u * major;
    const u32 exc_stream = nstreams - 1;            /* quality mode only */
    u32 p0 = 0; bool have_p0 = false;
    for (u32 w = s; w < e; w += 4) {
        const u32 word = *reinterpret_cast<const u32*>(sm + (w - sm_lo));
        if (mode == 0 ? (word == mmmm) : (__vcmpeq4(word, 0x4E4E4E4Eu) == 0)) continue;
#pragma unroll 1
        for (u32 k = 0; k < 4; k++) {
            const u32 p = w + k;
            if (p >= e) break;
            const u8 v = (u8)(word >> (8 * k));
            const u8 cls = mode == 0 ? lut[v] : (v == 'N' ? (u8)0 : LUT_SKIP);
            if (cls == LUT_SKIP) continue;
            if (cls == LUT_EXC) {
                const u32 idx = exc_stream * S2_THREADS + tid;
                if (WRITE) { u8* o = slot + T.cnt[idx]; o[0] = v; o[1] = (u8)p; o[2] = (u8)(p >> 8); o[3] = (u8)(p >> 16); o[4] = (u8)(p >> 24); }
                T.cnt[idx] += 5;
                continue;
            }
            const u32 idx = (u32)cls * S2_THREADS + tid;
            u32 bytes = 0; u8 tok[4];
            const bool cont = p > 0 && at(p - 1) == v;
            if (cont) {
                if (!have_p0) { p0 = p - 1; while (p0 > 0 && at(p0 - 1) == v) p0--; have_p0 = true; }
                const u32 rel = p - p0;
                if (p0 == 0 && rel == 1) { tok[0] = 0x00; bytes = 1; }                         /* Q16: second distance token */
                else {
                    const u32 s0 = p0 == 0 ? 2u : 1u;
                    if (rel >= s0 && ((rel - s0) & 31u) == 0) {
                        u32 len = 1;
                        while (len < 32 && p + len < n && at(p + len) == v) len++;
                        tok[0] = (u8)(0xC0u | (len - 1)); bytes = 1;
                    }
                }
            } else {
                p0 = p; have_p0 = true;
                const u32 lastrel = T.last[idx];
                u32 dm1 = 0; bool emit = true;
                if (lastrel != S2_NONE) dm1 = p - (lo + lastrel) - 1u;
                else if (p == 0) dm1 = 0;
                else {
                    const u32 f = T.first[idx];
                    if (WRITE && (f & S2_RESOLVED)) dm1 = T.fdist[idx];
                    else { emit = false; if (!WRITE) T.first[idx] = (u16)(p - lo); }
                }
                if (emit) {
                    if (dm1 < 128u) { tok[0] = (u8)dm1; bytes = 1; }
                    else if (dm1 < (1u << 14)) { tok[0] = (u8)(0x80u | (dm1 >> 8)); tok[1] = (u8)dm1; bytes = 2; }
                    else { tok[0] = (u8)(0xE0u | (dm1 >> 24)); tok[1] = (u8)(dm1 >> 16); tok[2] = (u8)(dm1 >> 8); tok[3] = (u8)dm1; bytes = 4; }
                }
            }
            T.last[idx] = (u16)(p - lo);
            if (bytes) {
                if (WRITE) { u8* o = slot + T.cnt[idx]; for (u32 q = 0; q < bytes; q++) o[q] = tok[q]; }
                T.cnt[idx] += bytes;
            }
        }
    }
}

__global__ void __launch_bounds__(S2_THREADS) k_streams2(EncBatchDev b, HeaderDev h, StreamJob job, const u32* __restrict__ span_chunk) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u8 s_lut[256];
    __shared__ u32 s_total[MAX_BINS + 2];
    __shared__ u32 s_base[MAX_BINS + 2];
    __shared__ u64 s_slot;
    __shared__ u32 s_bytes;
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
    S2Tables T;
    T.cnt = reinterpret_cast<u32*>(dyn + ST_SPAN + 2 * ST_HALO);
    T.first = reinterpret_cast<u16*>(T.cnt + nstreams * S2_THREADS);
    T.last = T.first + nstreams * S2_THREADS;
    T.fdist = T.last + nstreams * S2_THREADS;

    s_lut[tid] = h.lut[tid];
    for (u32 k = tid; k < nstreams * S2_THREADS; k += S2_THREADS) { T.cnt[k] = 0; T.first[k] = (u16)S2_NONE; T.last[k] = (u16)S2_NONE; T.fdist[k] = 0; }
    /* the walk reads whole words: pad the tail of the staging area */
    for (u32 k = tid; k < 8; k += S2_THREADS) if (sm_hi - sm_lo + k < (u32)(ST_SPAN + 2 * ST_HALO)) sm[sm_hi - sm_lo + k] = mode == 0 ? h.major : (u8)0;
    stage_positions(b, h, ck, mode, sm_lo, sm_hi, sm);
    __syncthreads();

    const u32 s = lo + (u32)tid * S2_SEG;
    const u32 e = s + S2_SEG < hi ? s + S2_SEG : hi;
    if (s < hi) s2_walk<false>(sm, sm_lo, n, lo, s, e, s_lut, mode, h.major, nstreams, T, tid, nullptr, b, ck);
    __syncthreads();

    /* per stream (a warp each): resolve first tokens against earlier segments of the span, exclusive scan of the byte counts */
    for (u32 st = warp; st < nstreams; st += S2_THREADS / 32) {
        const u32 base = st * S2_THREADS + lane * 8;
        /* inclusive max of `last` over the lane's 8 entries, then across lanes (NONE = no position; positions ascend with t) */
        u32 lastv[8]; u32 lane_last = S2_NONE;
#pragma unroll
        for (int k = 0; k < 8; k++) { lastv[k] = T.last[base + k]; if (lastv[k] != S2_NONE) lane_last = lastv[k]; }
        /* exclusive "latest non-NONE" scan across lanes */
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
        const u32 sf = warp_min(span_first);                    /* at most one lane has it */
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
    /* offsets become absolute inside the slot; `last` is rebuilt by pass 2 */
    for (u32 k = tid; k < nstreams * S2_THREADS; k += S2_THREADS) { T.cnt[k] += s_base[k / S2_THREADS]; T.last[k] = (u16)S2_NONE; }
    __syncthreads();
    if (s < hi && s_bytes) s2_walk<true>(sm, sm_lo, n, lo, s, e, s_lut, mode, h.major, nstreams, T, tid, job.slots + s_slot, b, ck);
}

}  // namespace rpq
