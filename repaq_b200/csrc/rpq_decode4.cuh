/*
 * rpq_decode4.cuh - the decode path without a quality plane in HBM.
 *
 * Generation 3 decoded every position stream into a 1-byte-per-base plane (k_fill + k_dec_streams: 1.5 GB written, 2.9 GB of
 * sector read-modify-write traffic, 1.4 GB read again by the formatter for a 3.4 GB FASTQ pair) because a stream can only be
 * parsed from its first byte.  Here a stream is parsed twice instead, and only its own bytes ever move:
 *
 *   k_dec_qindex   a warp per (chunk, stream) walks the tokens 128 stream bytes per step (decodeSingleQualByCol, reference
 *                  src/rfqcodec.cpp:957-1007, without storing anything) and leaves one 8-byte CHECKPOINT per step: where the
 *                  first token of the step starts and the position the stream has reached.  0.06 bytes per stream byte.
 *   k_dec_tiledir  per formatter tile (G reads of ONE chunk) and stream: which steps of the stream hold the tile's positions (two
 *                  bisections over the checkpoints), which exception records are the tile's.
 *   k_dec_planes   a small CTA per tile: the tile's qualities pre-filled with the major quality (src/rfqcodec.cpp:1089), every
 *                  (stream, step) of the directory decoded from its checkpoint by whichever warp takes it, the exception records
 *                  (:1034-1043) and the N positions (:856-858, a bitmap) applied; the finished tile goes to its slot in HBM
 *                  (19 KB per 128 reads of 150 bases: written and read once).
 *   k_dec_format4  the formatter: the tile's slot and its piece of the 2-bit column arrive by TMA while the name lines are written;
 *                  two threads per read emit the records into shared memory, one TMA bulk store per output stream takes them out.
 */
#pragma once
#include "rpq_decode2.cuh"

namespace rpq {

constexpr u32 QX_STEP = 128;                        /* stream bytes per warp step = per checkpoint */
constexpr int QX_WARPS = 4;                         /* k_dec_qindex: (chunk, stream) pairs per CTA */

/* checkpoints of chunk `c`: its streams' entries lie back to back from here (a stream of n bytes owns n / 128 + 2 entries; all
 * streams of a chunk together never need more than bytes / 128 + 2 * n_streams, which is what the body offset leaves room for) */
__device__ __forceinline__ u64 qx_chunk_base(const DecChunk& ck, u32 c, u32 n_streams) { return ck.in_off / QX_STEP + 2ull * n_streams * c; }
inline size_t qx_entries(u64 body_len, u32 n_chunks, u32 n_streams) { return (size_t)(body_len / QX_STEP + 2ull * n_streams * n_chunks + 64); }

/* one position stream of a chunk: its bytes, its length, the value it stands for, its first checkpoint */
struct QStream { const u8* p; u32 len; u32 ck; };

/*
 * The quality column of a chunk (reference src/rfqcodec.cpp:1009-1033): u32 LE lengths of the nb streams, the streams, the
 * exception records.  Lengths are clamped to the column as the column is walked, so that the streams of a damaged chunk are
 * disjoint pieces of the column.  Warp-cooperative (nb <= 64: two streams per lane); every lane gets the stream `want`
 * (nb: the exception records; nb + 1: the N positions).
 */
__device__ inline QStream qx_stream(const DecBatchDev& b, const HeaderDev& h, const DecChunk& ck, u32 want, int lane) {
    const u8* in = b.body + ck.in_off;
    const u8* qcol = in + ck.off_qual;
    const u32 nb = h.nb;
    QStream s; s.p = nullptr; s.len = 0; s.ck = 0;
    unsigned long long off0 = 0, off1 = 0, tot = 0;
    u32 l0 = 0, l1 = 0, e0 = 0, e1 = 0, etot = 0;
    const bool table_ok = !(h.flags & RPQ_DONT_ENCODE_QUAL) && 4ull * nb <= ck.qual_size;
    if (table_ok) {
        const u32 k0 = (u32)lane, k1 = (u32)lane + 32u;
        const u32 v0 = k0 < nb ? ld32(qcol + 4 * k0) : 0u, v1 = k1 < nb ? ld32(qcol + 4 * k1) : 0u;
        unsigned long long i0 = v0, i1 = v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t0 = __shfl_up_sync(0xffffffffu, i0, d), t1 = __shfl_up_sync(0xffffffffu, i1, d);
            if (lane >= d) { i0 += t0; i1 += t1; }
        }
        const unsigned long long tot0 = __shfl_sync(0xffffffffu, i0, 31);
        off0 = 4ull * nb + i0 - v0; off1 = 4ull * nb + tot0 + i1 - v1;
        tot = 4ull * nb + tot0 + __shfl_sync(0xffffffffu, i1, 31);
        auto clamp = [&](unsigned long long off, u32 v) -> u32 { return off >= ck.qual_size ? 0u : (off + v > ck.qual_size ? (u32)(ck.qual_size - off) : v); };
        l0 = k0 < nb ? clamp(off0, v0) : 0u; l1 = k1 < nb ? clamp(off1, v1) : 0u;
        /* checkpoint entries before each stream */
        u32 c0 = k0 < nb ? l0 / QX_STEP + 2u : 0u, c1 = k1 < nb ? l1 / QX_STEP + 2u : 0u;
        u32 j0 = c0, j1 = c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t0 = __shfl_up_sync(0xffffffffu, j0, d), t1 = __shfl_up_sync(0xffffffffu, j1, d); if (lane >= d) { j0 += t0; j1 += t1; } }
        const u32 ct0 = __shfl_sync(0xffffffffu, j0, 31);
        e0 = j0 - c0; e1 = ct0 + j1 - c1; etot = ct0 + __shfl_sync(0xffffffffu, j1, 31);
    }
    if (want < nb) {
        const int src = (int)(want & 31u);
        const unsigned long long o = __shfl_sync(0xffffffffu, want < 32u ? off0 : off1, src);
        s.len = __shfl_sync(0xffffffffu, want < 32u ? l0 : l1, src);
        s.ck = __shfl_sync(0xffffffffu, want < 32u ? e0 : e1, src);
        s.p = qcol + o;
    } else if (want == nb) {
        if (table_ok && tot < ck.qual_size) { s.p = qcol + tot; s.len = (u32)(ck.qual_size - tot); }
        s.ck = etot;
    } else {
        s.p = in + ck.off_npos; s.len = (h.flags & RPQ_ENCODE_N_POS) ? ck.npos_size : 0u;
        s.ck = etot;
    }
    return s;
}

/* a position stream as a warp reads it: aligned words, 4 bytes per lane and step */
struct QCursor {
    const u32* A; const u32* Aend; u32 sh; u32 slen;
    __device__ __forceinline__ u32 ldw(u32 k) const { const u32* w = A + k; return w < Aend ? *w : 0u; }
};
__device__ __forceinline__ QCursor qx_cursor(const DecBatchDev& b, const QStream& s) {
    QCursor q;
    const uintptr_t sa = reinterpret_cast<uintptr_t>(s.p);
    q.A = reinterpret_cast<const u32*>(sa & ~(uintptr_t)3);
    q.sh = 8u * (u32)(sa & 3u);
    q.Aend = reinterpret_cast<const u32*>((reinterpret_cast<uintptr_t>(b.body + b.body_len) + 3u) & ~(uintptr_t)3);
    q.slen = s.len;
    return q;
}

/*
 * One step of the token walk (the body of k_dec_streams' loop): the 128 stream bytes from `base`, four per lane.  Token length
 * depends on the first byte only (0xxxxxxx 1, 10xxxxxx 2, 110xxxxx 1, 111xxxxx 4).  Where a lane's tokens start depends on how
 * many payload bytes spill in from the lane before (0..3): every lane tabulates its exit spill for the four possible entries;
 * most tables are constant, so the chain resolves in a round or two.  Each lane then decodes its (at most four) tokens; one warp
 * scan of the per-lane advances places them.  emit(first, end1) is called for every token of the lane: it covers the positions
 * [first, end1).  `skip` (payload bytes at the start of the step that belong to the previous token) and `next` (1 + the last
 * position reached) are carried from step to step; cura / nexta are the lane's words of this step and of the next one.
 */
template <class Emit>
__device__ __forceinline__ void qx_step(const QCursor& S, u32 base, u32 cura, u32 nexta, u32& skip, u32& next, int lane, Emit&& emit) {
    u32 a1 = __shfl_down_sync(0xffffffffu, cura, 1), a2 = __shfl_down_sync(0xffffffffu, cura, 2);
    const u32 n0 = __shfl_sync(0xffffffffu, nexta, 0), n1 = __shfl_sync(0xffffffffu, nexta, 1);
    if (lane == 31) { a1 = n0; a2 = n1; } else if (lane == 30) a2 = n0;
    u64 B = (u64)__funnelshift_r(cura, a1, S.sh) | ((u64)__funnelshift_r(a1, a2, S.sh) << 32);
    const u32 p0 = base + 4u * (u32)lane;
    const u32 left = p0 < S.slen ? S.slen - p0 : 0u;       /* stream bytes from this lane's first byte on */
    if (left < 8u) B = left ? B & ((1ull << (8u * left)) - 1ull) : 0ull;
    const u32 nv = left < 4u ? left : 4u;
    const u32 cur = (u32)B;
    u32 adv[4], run[4]; u32 lane_adv = 0;
    /* a step of one-byte tokens only (0xxxxxxx distances, 110xxxxx lengths: the rule in the dense streams of columns with ~40
     * values) with nothing spilling in needs no token chain: every byte heads a token */
    const u32 multi = cur & (~(cur << 1) | (cur << 2)) & 0x80808080u;       /* bit 7 of the bytes 10xxxxxx and 111xxxxx */
    if (skip == 0u && __ballot_sync(0xffffffffu, multi != 0u) == 0u) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 b0 = (cur >> (8 * k)) & 0xFFu;
            const bool len_tok = (b0 & 0x80u) != 0;
            adv[k] = (u32)k < nv ? (b0 & (len_tok ? 0x1Fu : 0x7Fu)) + 1u : 0u;
            run[k] = len_tok ? adv[k] : 0u;
            lane_adv += adv[k];
        }
    } else {
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
        u32 next_head = r_in;
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
    }
    u32 tot; const u32 ex = warp_excl_scan(lane_adv, lane, tot);
    u32 acc = next + ex;                                            /* 1 + the position before this lane's first token */
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (adv[k]) {
            const u32 end1 = acc + adv[k];                          /* 1 + the position of the token's last element */
            acc = end1;
            emit(run[k] ? end1 - run[k] : end1 - 1u, end1);
        }
    }
    next += tot;
}

/* stream indices by decreasing length in chunk `c0`: see k_dec_stream_order */

/*
 * k_dec_qindex: grid (chunks / QX_WARPS, streams); blockIdx.y picks the stream through `order` (longest first).  Stream nb is the
 * list of exception records: nothing to index, but the formatter looks its range up by bisection, which needs the records in
 * increasing position (as the reference writes them, src/rfqcodec.cpp:750-758): a list that is not raises ERRBIT_RFQ and the
 * host decodes the batch through the plane.
 */
__global__ void __launch_bounds__(32 * QX_WARPS) k_dec_qindex(DecBatchDev b, HeaderDev h, u32 n_qstreams, u32 n_streams, const u32* __restrict__ order, uint2* __restrict__ ckpt) {
    const u32 c = blockIdx.x * QX_WARPS + (threadIdx.x >> 5);
    if (c >= b.n_chunks) return;
    const u32 task = order[blockIdx.y];
    const u32 st = task < n_qstreams ? task : (u32)h.nb + 1u;        /* quality streams 0..nb-1, exceptions nb; then the N positions */
    const int lane = threadIdx.x & 31;
    const DecChunk& ck = b.chunks[c];
    const QStream s = qx_stream(b, h, ck, st, lane);
    if (st == h.nb) {
        const u32 n = s.len / 5u;
        u32 bad = 0;
        for (u32 k = lane; k + 1 < n; k += 32) { if (ld32(s.p + 5ull * k + 1) > ld32(s.p + 5ull * k + 6)) bad = 1; }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(b.err, ERRBIT_RFQ);
        return;
    }
    if (!s.len) return;
    const QCursor S = qx_cursor(b, s);
    uint2* out = ckpt + qx_chunk_base(ck, c, n_streams) + s.ck;
    u32 next = 0, skip = 0;
    u32 cura = S.ldw((u32)lane);
    for (u32 base = 0; base < S.slen; base += QX_STEP) {
        const u32 nexta = base + QX_STEP < S.slen + 8 ? S.ldw((base >> 2) + 32u + (u32)lane) : 0u;
        if (lane == 0) out[base / QX_STEP] = make_uint2(skip, next);
        qx_step(S, base, cura, nexta, skip, next, lane, [](u32, u32) {});
        cura = nexta;
    }
}

struct Fmt4Cfg {
    u32 reads_per_cta;     /* G; blockDim.x = 2 G */
    u32 plane_cap;         /* shared bytes of the quality tile (incl. 16 bytes of alignment slack) */
    u32 out_cap;           /* shared bytes per output stream staging (incl. 16 bytes of alignment slack) */
    u32 nbits_words;       /* shared words of the N bitmap (0 unless the header has ENCODE_N_POS) */
    u32 n_streams;         /* checkpoint layout: streams per chunk (quality streams + exceptions + N positions) */
    u32 seq_cap;           /* k_dec_format4: shared bytes of the tile's piece of the 2-bit column (incl. alignment slack) */
};

/*
 * The tile directory: for every formatter tile (G reads of one chunk) and stream, which steps of the stream hold tokens for the
 * tile's positions.  Every step has its checkpoint, so the formatter decodes the steps of all its streams independently and in
 * parallel - no search and no token chain inside the formatter.  For the exception records: the records of the tile.
 */
struct TileDir {
    u32 off;            /* the stream's first byte, from the chunk's first byte (exceptions: the tile's first record) */
    u32 len;            /* stream bytes (exceptions: records of the tile) */
    u32 ck;             /* checkpoint of the first step, from the chunk's first checkpoint */
    u32 step0, n_steps; /* steps [step0, step0 + n_steps) */
    u32 pad[3];
};
/* tiles of chunk c lie at tile index read_base / G + c (+ tile): disjoint, since a chunk of n reads has at most n / G + 1 tiles */
__device__ __forceinline__ u64 tile_index(const DecChunk& ck, u32 c, u32 G) { return (u64)(ck.read_base / G) + c; }
inline size_t tile_dir_entries(u32 n_reads, u32 n_chunks, u32 G, u32 n_streams) { return ((size_t)n_reads / G + n_chunks + 2) * n_streams; }

/* grid (chunks / QX_WARPS, streams): a warp per (chunk, stream), a lane per tile */
__global__ void __launch_bounds__(32 * QX_WARPS) k_dec_tiledir(DecBatchDev b, HeaderDev h, u32 G, u32 n_qstreams, u32 n_streams, const uint2* __restrict__ ckpt, TileDir* __restrict__ dir) {
    const u32 c = blockIdx.x * QX_WARPS + (threadIdx.x >> 5);
    if (c >= b.n_chunks) return;
    const u32 task = blockIdx.y;
    const u32 st = task < n_qstreams ? task : (u32)h.nb + 1u;
    const int lane = threadIdx.x & 31;
    const DecChunk& ck = b.chunks[c];
    const QStream s = qx_stream(b, h, ck, st, lane);
    const uint2* e = ckpt + qx_chunk_base(ck, c, n_streams) + s.ck;
    const u32 n_ck = (s.len + QX_STEP - 1) / QX_STEP;
    const u32 n_tiles = (ck.reads + G - 1) / G;
    const bool compact = st == (u32)h.nb + 1u;             /* the N positions count compacted bases */
    const u32* offs = compact ? b.seqoff : b.qualoff;
    const u32 total = compact ? ck.seq_kept : ck.total_len;
    TileDir* out = dir + tile_index(ck, c, G) * n_streams + task;
    const u32 base_off = (u32)(s.p ? (u64)(s.p - (b.body + ck.in_off)) : 0u);
    for (u32 t = (u32)lane; t < n_tiles; t += 32) {
        const u32 r0 = t * G, r1 = r0 + G < ck.reads ? r0 + G : ck.reads;
        const u32 lo = offs[ck.read_base + r0];
        u32 hi = r1 < ck.reads ? offs[ck.read_base + r1] : total;
        if (hi > ck.total_len) hi = ck.total_len;           /* positions >= the chunk's length are ignored (Q20) */
        TileDir d; d.off = base_off; d.len = s.len; d.ck = s.ck; d.step0 = 0; d.n_steps = 0; d.pad[0] = d.pad[1] = d.pad[2] = 0;
        if (st == h.nb) {
            /* exception records {q, u32 LE pos} in increasing position (checked by k_dec_qindex): those with lo <= pos < hi */
            const u32 n = s.len / 5u;
            auto below = [&](u32 lim) { u32 a = 0, z = n; while (a < z) { const u32 m = (a + z) >> 1; if (ld32(s.p + 5ull * m + 1) < lim) a = m + 1; else z = m; } return a; };
            const u32 f = n ? below(lo) : 0u, g = n ? below(hi) : 0u;
            d.off = base_off + 5u * f; d.len = g > f ? g - f : 0u;
        } else if (n_ck && hi > lo) {
            /* the last step whose entry position is <= lo (the tokens before it lie below lo) ... */
            u32 a = 0, z = n_ck;
            while (z - a > 1u) { const u32 m = (a + z) >> 1; if (e[m].y <= lo) a = m; else z = m; }
            /* ... to the last step whose entry position is < hi */
            u32 a2 = a, z2 = n_ck;
            while (z2 - a2 > 1u) { const u32 m = (a2 + z2) >> 1; if (e[m].y < hi) a2 = m; else z2 = m; }
            d.step0 = a; d.n_steps = a2 - a + 1u; d.ck = s.ck + a;
        }
        out[(size_t)t * n_streams] = d;
    }
}

/* one step (128 stream bytes) of a position stream into the tile: the positions [lo, hi) as bytes `q` at tile[pos - org], or
 * (BITS) as bits of the N bitmap - one warp */
template <bool BITS>
__device__ __forceinline__ void qx_decode_step(const DecBatchDev& b, const u8* stream, u32 slen, u32 step, uint2 entry, u32 lo, u32 hi, u8 q, u8* tile, u32* bits, u32 org, int lane) {
    QStream s; s.p = stream; s.len = slen; s.ck = 0;
    const QCursor S = qx_cursor(b, s);
    u32 skip = entry.x, next = entry.y;
    const u32 base = step * QX_STEP;
    const u32 cura = S.ldw((base >> 2) + (u32)lane);
    const u32 nexta = base + QX_STEP < S.slen + 8 ? S.ldw((base >> 2) + 32u + (u32)lane) : 0u;
    qx_step(S, base, cura, nexta, skip, next, lane, [&](u32 first, u32 end1) {
        const u32 x = first > lo ? first : lo, y = end1 < hi ? end1 : hi;
        if (BITS) { for (u32 p = x; p < y; p++) atomicOr(&bits[(p - org) >> 5], 1u << ((p - org) & 31u)); }
        else if (x < y) {                                              /* mostly one position (a distance token) */
            u8* o = tile + (x - org);
            o[0] = q;
            for (u32 k = 1; k < y - x; k++) o[k] = q;
        }
    });
}

/* the shared-memory image of a tile's qualities and N positions, as k_dec_planes leaves it in HBM (one slot per tile) and
 * k_dec_format4 stages it: plane_cap bytes of qualities (position p of the chunk at byte p - (P0 & ~15)), then the N bitmap */
__device__ __forceinline__ u32 slot_bytes(const Fmt4Cfg& cfg) { return cfg.plane_cap + 4u * cfg.nbits_words; }

/*
 * k_dec_planes: grid (tiles per chunk, chunks), 128 threads.  A tile's qualities: allQual(seqLen, majorQual()) (reference
 * src/rfqcodec.cpp:1089), then every (stream, step) the tile directory lists as a task of its own - the steps of all streams are
 * decoded independently and in parallel from their checkpoints (decodeSingleQualByCol :957-1007, exceptions :1034-1043, N
 * positions :856-858) - and the finished tile goes to its slot with 16-byte stores.  Small CTAs (a tile is 19 KB for 128 reads of
 * 150 bases): eleven per SM, the loads of a step are hidden behind the other CTAs' steps.
 */
constexpr int PL_THREADS = 128;
__global__ void __launch_bounds__(PL_THREADS) k_dec_planes(DecBatchDev b, HeaderDev h, Fmt4Cfg cfg, u32 chunk_first, const uint2* __restrict__ ckpt, const TileDir* __restrict__ dir, u8* __restrict__ planes) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u32 s_task;
    __shared__ TileDir s_dir[MAX_BINS + 3];
    __shared__ u32 s_pre[MAX_BINS + 4];                /* tasks before each stream */
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 G = cfg.reads_per_cta;
    const u32 c = chunk_first + blockIdx.y;
    const DecChunk& ck = b.chunks[c];
    const u32 r0 = blockIdx.x * G;
    if (r0 >= ck.reads) return;
    const u32 n_here = ck.reads - r0 < G ? ck.reads - r0 : G;
    u8* s_plane = dyn;
    u32* s_nbits = reinterpret_cast<u32*>(dyn + cfg.plane_cap);
    const u8* in = b.body + ck.in_off;
    const bool raw_qual = (h.flags & RPQ_DONT_ENCODE_QUAL) != 0;
    const bool npos_mode = (h.flags & RPQ_ENCODE_N_POS) != 0;
    /* ---- the tile's ranges: quality positions [P0, P1) and compacted bases [C0, C1) of the chunk */
    const u32 i_first = ck.read_base + r0, i_last = i_first + n_here - 1u;
    const u32 P0 = b.qualoff[i_first], P1 = b.qualoff[i_last] + b.rlen[i_last];
    const u32 C0 = b.seqoff[i_first], C1 = r0 + n_here < ck.reads ? b.seqoff[i_last + 1u] : ck.seq_kept;
    const u32 porg = P0 & ~15u, corg = C0 & ~31u;
    const u32 plane_bytes = ((P1 - porg + 15u) >> 4) << 4;
    const u32 n_q = raw_qual ? 0u : (u32)h.nb + 1u;                    /* quality streams + exception records */
    const u32 n_tasks = n_q + (npos_mode ? 1u : 0u);
    if (tid == 0) s_task = 0;
    {
        const TileDir* src = dir + (tile_index(ck, c, G) + blockIdx.x) * cfg.n_streams;
        for (u32 k = tid; k < n_tasks * (u32)(sizeof(TileDir) / 4); k += blockDim.x) reinterpret_cast<u32*>(s_dir)[k] = reinterpret_cast<const u32*>(src)[k];
    }
    /* allQual(seqLen, majorQual()) (src/rfqcodec.cpp:1089) for this tile; DONT_ENCODE_QUAL: the column itself (:903-908) */
    if (!raw_qual) {
        const u32 m4 = 0x01010101u * h.major;
        const uint4 fill = make_uint4(m4, m4, m4, m4);
        for (u32 k = tid; k < plane_bytes / 16u; k += blockDim.x) reinterpret_cast<uint4*>(s_plane)[k] = fill;
    } else {
        const u8* qcol = in + ck.off_qual;
        const u32 have = ck.qual_size < ck.total_len ? ck.qual_size : ck.total_len;       /* positions the column holds; the rest keeps the major quality */
        for (u32 p = porg + tid; p < porg + plane_bytes; p += blockDim.x) s_plane[p - porg] = p < have ? qcol[p] : h.major;
    }
    for (u32 k = tid; k < cfg.nbits_words; k += blockDim.x) s_nbits[k] = 0;
    __syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (u32 t = 0; t < n_tasks; t++) { s_pre[t] = acc; acc += (n_q && t == n_q - 1u) ? (s_dir[t].len + 31u) / 32u : s_dir[t].n_steps; }
        s_pre[n_tasks] = acc;
    }
    __syncthreads();
    auto stream_tasks = [&]() {
        const u32 total = s_pre[n_tasks];
        const u8* cbase = b.body + ck.in_off;
        const uint2* cck = ckpt + qx_chunk_base(ck, c, cfg.n_streams);
        u32 t = 0;                                                     /* the stream of task k: the last one with s_pre[t] <= k; tasks come in increasing order */
        for (;;) {
            u32 k = 0;
            if (lane == 0) k = atomicAdd(&s_task, 1u);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= total) break;
            while (t + 1u < n_tasks && s_pre[t + 1u] <= k) t++;
            const TileDir& d = s_dir[t];
            const u32 j = k - s_pre[t];
            const u32 st = t < n_q ? t : (u32)h.nb + 1u;
            if (st == h.nb) {
                /* 32 exception records per task (src/rfqcodec.cpp:1034-1043) */
                const u32 rec = 32u * j + (u32)lane;
                if (rec < d.len) { const u8* p = cbase + d.off + 5ull * rec; const u32 pos = ld32(p + 1); if (pos >= P0 && pos < P1 && pos < ck.total_len) s_plane[pos - porg] = p[0]; }
            } else if (st < h.nb) qx_decode_step<false>(b, cbase + d.off, d.len, d.step0 + j, cck[d.ck + j], P0, P1 < ck.total_len ? P1 : ck.total_len, h.normal_bins[st], s_plane, nullptr, porg, lane);
            else qx_decode_step<true>(b, cbase + d.off, d.len, d.step0 + j, cck[d.ck + j], C0, C1 < ck.total_len ? C1 : ck.total_len, 0, nullptr, s_nbits, corg, lane);
        }
    };
    stream_tasks();
    __syncthreads();
    /* ---- the tile to its slot */
    {
        uint4* dst = reinterpret_cast<uint4*>(planes + (size_t)(tile_index(ck, c, G) + blockIdx.x) * slot_bytes(cfg));
        const uint4* src = reinterpret_cast<const uint4*>(dyn);
        const u32 nq16 = plane_bytes / 16u, nb16 = cfg.nbits_words / 4u, boff = cfg.plane_cap / 16u;
        for (u32 k = tid; k < nq16; k += blockDim.x) dst[k] = src[k];
        for (u32 k = tid; k < nb16; k += blockDim.x) dst[boff + k] = src[boff + k];
    }
}

/*
 * k_dec_format4: grid (tiles per chunk, chunks of the window); CTA = G reads of one chunk, two threads per read in different
 * warps.  Shared: quality tile | N bitmap (the tile's slot, one TMA bulk copy, awaited only where the qualities are first needed:
 * the name lines are formatted while it is in flight) | record staging per output stream.
 */
__global__ void __launch_bounds__(256) k_dec_format4(DecBatchDev b, HeaderDev h, Fmt4Cfg cfg, u32 chunk_first, const u8* __restrict__ planes) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u32 s_lut_fwd[256], s_lut_rc[256];
#ifndef RPQ_EMU
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar);
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 G = cfg.reads_per_cta;
    const u32 c = chunk_first + blockIdx.y;
    const DecChunk& ck = b.chunks[c];
    const u32 r0 = blockIdx.x * G;
    if (r0 >= ck.reads) return;
    const u32 n_here = ck.reads - r0 < G ? ck.reads - r0 : G;
    const int rt = tid < (int)G ? tid : tid - (int)G, half = tid < (int)G ? 0 : 1;
    u8* s_plane = dyn;
    u32* s_nbits = reinterpret_cast<u32*>(dyn + cfg.plane_cap);
    u8* s_out[2] = {dyn + cfg.plane_cap + 4u * cfg.nbits_words, dyn + cfg.plane_cap + 4u * cfg.nbits_words + cfg.out_cap};
    const u32 nstreams = b.split_pairs ? 2u : 1u;
    u8* s_seq = dyn + cfg.plane_cap + 4u * cfg.nbits_words + nstreams * cfg.out_cap;
    const u8* in = b.body + ck.in_off;
    const u32 fl = ck.flags;
    const bool raw_qual = (h.flags & RPQ_DONT_ENCODE_QUAL) != 0;
    const bool npos_mode = (h.flags & RPQ_ENCODE_N_POS) != 0;

    /* ---- the tile's ranges: quality positions [P0, P1) and compacted bases [C0, C1) of the chunk */
    const u32 i_first = ck.read_base + r0, i_last = i_first + n_here - 1u;
    const u32 P0 = b.qualoff[i_first], P1 = b.qualoff[i_last] + b.rlen[i_last];
    const u32 C0 = b.seqoff[i_first], C1 = r0 + n_here < ck.reads ? b.seqoff[i_last + 1u] : ck.seq_kept;
    const u32 porg = P0 & ~15u, corg = C0 & ~31u;
    const u32 plane_bytes = ((P1 - porg + 15u) >> 4) << 4;

    for (u32 v = tid; v < 256; v += blockDim.x) {
        u32 f = 0, r = 0;
        for (u32 k = 0; k < 4; k++) {
            const u32 code = (v >> (2 * k)) & 3u;
            const u32 ch = code == 0 ? 'G' : code == 1 ? 'A' : code == 2 ? 'T' : 'C';
            const u32 cc = code == 0 ? 'C' : code == 1 ? 'T' : code == 2 ? 'A' : 'G';
            f |= ch << (8 * k);
            r |= cc << (8 * (3 - k));
        }
        s_lut_fwd[v] = f; s_lut_rc[v] = r;
    }
#ifndef RPQ_EMU
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();                                           /* the tables and the barrier object are there for everybody: the only CTA-wide wait before the records are done */

    const bool active = rt < (int)n_here;
    const u32 i = i_first + rt;
    u32 r = 0, rl = 0, stream = 0, olen = 0;
    u64 oabs = 0;
    u32 qrel = 0;
    /* everything the thread will need from the per-read tables and the chunk's columns is requested here, in one go: the loads are
     * independent, their latencies overlap each other and the two barriers that follow */
    u32 pre_so = 0, pre_x = 0, pre_y = 0, pre_prev_rl = 0, pre_ls = 0, pre_l1 = 0, pre_lane = 0, pre_tile = 0;
    int pre_ov = 0;
    if (active) {
        r = r0 + rt; rl = b.rlen[i]; olen = b.olen[i];
        stream = b.split_pairs ? (r & 1u) : 0u;
        oabs = ck.out_off[stream] + b.outoff[i];
        qrel = b.qualoff[i];
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const u32 xy = il ? r >> 1 : r;
        pre_ls = (fl & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
        if (half == 1) {
            pre_l1 = (fl & (RPQ_NAME1_SAME | RPQ_NAME1_LEN_SAME)) ? in[ck.off_n1len] : in[ck.off_n1len + r];
            if (h.flags & RPQ_HAS_LANE) pre_lane = (fl & RPQ_LANE_SAME) ? in[ck.off_lane] : in[ck.off_lane + xy];
            if (h.flags & RPQ_HAS_TILE) { const u32 k = (fl & RPQ_TILE_SAME) ? 0u : xy; pre_tile = (u32)in[ck.off_tile + 2 * k] | ((u32)in[ck.off_tile + 2 * k + 1] << 8); }
            if (h.flags & RPQ_HAS_X) pre_x = b.xs[ck.read_base + xy];
            if (h.flags & RPQ_HAS_Y) pre_y = b.ys[ck.read_base + xy];
        } else {
            pre_so = b.seqoff[i];
            if (il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP) && (r & 1u)) { pre_ov = (int)(signed char)in[ck.off_ov + (r >> 1)] - (int)h.overlap_shift; pre_prev_rl = b.rlen[i - 1]; }
        }
    }
    /* where the tile's records start and end in each output stream: from the tile's first and last reads of the stream (the same
     * few table entries for every thread: no exchange through shared memory, no barrier behind the table loads) */
    u64 t_start[2] = {~0ull, ~0ull}, t_end[2] = {0, 0};
    for (u32 sidx = 0; sidx < nstreams; sidx++) {
        if (sidx >= n_here) continue;                              /* a tile of one read has one stream */
        const u32 rt_first = sidx;                                 /* r0 is even: the first read of stream s is read s of the tile */
        const u32 rt_last = nstreams == 1u ? n_here - 1u : (((n_here - 1u) & 1u) == sidx ? n_here - 1u : n_here - 2u);
        t_start[sidx] = ck.out_off[sidx] + b.outoff[i_first + rt_first];
        t_end[sidx] = ck.out_off[sidx] + b.outoff[i_first + rt_last] + b.olen[i_first + rt_last];
    }

    /* ---- the tile's slot and its piece of the 2-bit column: two TMA bulk copies counted on one mbarrier.  The piece: the bytes
     * that hold the compact bases [C0, C1) of the tile's reads (an overlapped mate reads its partner's bases: pairs never straddle
     * tiles) plus the 8 bytes a 16-base load may reach past them, from a 16-byte aligned address, never past the end of the body;
     * whatever a (damaged) chunk makes a read look for outside the piece comes from the body itself. */
    const u32 slot_n = slot_bytes(cfg);
    const u8* slot = planes + (size_t)(tile_index(ck, c, G) + blockIdx.x) * slot_n;
    const u8* seqcol = in + ck.off_seq;
    const u8* sq_src; u32 sq_n = 0;
    {
        const uintptr_t want0 = reinterpret_cast<uintptr_t>(seqcol + (C0 >> 2)), want1 = reinterpret_cast<uintptr_t>(seqcol + ((C1 + 3u) >> 2)) + 8u;
        const uintptr_t end_body = reinterpret_cast<uintptr_t>(b.body + b.body_len) & ~(uintptr_t)15;
        const uintptr_t a0 = want0 & ~(uintptr_t)15;
        uintptr_t a1 = (want1 + 15u) & ~(uintptr_t)15;
        if (a1 > end_body) a1 = end_body;
        sq_src = reinterpret_cast<const u8*>(a0);
        if (a1 > a0 && C1 > C0) { sq_n = (u32)(a1 - a0); if (sq_n > cfg.seq_cap) sq_n = cfg.seq_cap & ~15u; }
    }
#ifdef RPQ_EMU
    {
        const uint4* src = reinterpret_cast<const uint4*>(slot);
        uint4* dst = reinterpret_cast<uint4*>(s_plane);
        for (u32 k = tid; k < plane_bytes / 16u; k += blockDim.x) dst[k] = src[k];
        for (u32 k = tid; k < cfg.nbits_words / 4u; k += blockDim.x) dst[cfg.plane_cap / 16u + k] = src[cfg.plane_cap / 16u + k];
        for (u32 k = tid; k < sq_n; k += blockDim.x) s_seq[k] = sq_src[k];
    }
    __syncthreads();
    auto plane_ready = [&]() {};
#else
    if (tid == 0) {
        const u32 nbytes = cfg.nbits_words ? slot_n : plane_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(nbytes + sq_n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((u32)__cvta_generic_to_shared(s_plane)), "l"(slot), "r"(nbytes), "r"(mbar) : "memory");
        if (sq_n)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((u32)__cvta_generic_to_shared(s_seq)), "l"(sq_src), "r"(sq_n), "r"(mbar) : "memory");
    }
    auto plane_ready = [&]() {
        u32 done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    };
#endif

    u8* o = nullptr; u32 ls = 0, name_end = 0;
    if (active) {
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const bool odd = (r & 1u) != 0;
        const u32 xy = il ? r >> 1 : r;
        o = s_out[stream] + (u32)(oabs - ((stream ? t_start[1] : t_start[0]) & ~15ull));
        /* ---- strand length first: it fixes where every part of the record lies */
        ls = pre_ls;
        name_end = olen - (2u * rl + ls + 3u);          /* bytes of the name line including its line break */
        if (half == 1) {
            u32 w_at = 0;
            /* ---- name (reference src/rfqcodec.cpp:1157-1231) */
            const u32 l1 = pre_l1;
            const u8* n1 = in + ck.off_n1 + ((fl & RPQ_NAME1_SAME) ? 0u : b.n1off[i]);
            for (u32 k = 0; k < l1; k++) o[w_at + k] = n1[k];
            w_at += l1;
            if (h.flags & RPQ_HAS_LANE) { o[w_at++] = ':'; w_at += put_dec(o + w_at, pre_lane); }
            if (h.flags & RPQ_HAS_TILE) { o[w_at++] = ':'; w_at += put_dec(o + w_at, pre_tile); }
            if (h.flags & RPQ_HAS_X) { o[w_at++] = ':'; w_at += put_dec(o + w_at, pre_x); }
            if (h.flags & RPQ_HAS_Y) { o[w_at++] = ':'; w_at += put_dec(o + w_at, pre_y); }
            if (h.flags & RPQ_HAS_NAME2) {
                const u32 l2 = (fl & (RPQ_NAME2_SAME | RPQ_NAME2_LEN_SAME)) ? in[ck.off_n2len] : in[ck.off_n2len + r];
                const u8* n2 = in + ck.off_n2 + ((fl & RPQ_NAME2_SAME) ? 0u : b.n2off[i]);
                for (u32 k = 0; k < l2; k++) o[w_at + k] = n2[k];
                if ((fl & RPQ_NAME2_SAME) && il && odd && h.name2_diff_char != 0 && h.name2_diff_pos < l2) o[w_at + h.name2_diff_pos] = h.name2_diff_char;
                w_at += l2;
            }
            o[w_at++] = '\n';
            const u8* sp = in + ck.off_strand + ((fl & RPQ_STRAND_SAME) ? 0u : b.soff[i]);
            u8* o_str = o + name_end + rl + 1;
            for (u32 k = 0; k < ls; k++) o_str[k] = sp[k];
            o_str[ls] = '\n';
            o_str[ls + 1 + rl] = '\n';
        } else o[name_end + rl] = '\n';
    }
    plane_ready();                                                 /* the quality tile and the N bitmap are there */

    if (active) {
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const bool ov_on = il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP);
        const bool odd = (r & 1u) != 0;
        const u8* q = s_plane + (qrel - porg);
        u8* o_seq = o + name_end;
        u8* o_qual = o_seq + rl + 1 + ls + 1;
        /* ---- sequence + quality, four positions per step */
        const u8* seqb = seqcol;
        const long long sq_off = sq_src - seqb;                        /* column byte the staged piece starts at */
        auto code_byte = [&](long long bi) -> u32 { const long long d = bi - sq_off; return (d >= 0 && d < (long long)sq_n) ? s_seq[d] : seqb[bi]; };
        const long long so = half == 0 ? pre_so : 0;
        int ov = 0; u32 prev_rl = 0;
        if (ov_on && odd) { ov = pre_ov; prev_rl = pre_prev_rl; }
        const bool rc = il && odd;
        const u8 nq = (u8)h.n_base_qual;
        const u32 nq4 = 0x01010101u * nq;
        const long long unpacked = ck.seq_size * 4u < ck.total_len ? ck.seq_size * 4u : ck.total_len;
        /* the N bitmap covers the compact positions [corg, cend) */
        const long long cend = (long long)corg + 32ll * cfg.nbits_words;
        auto nbit = [&](long long ci) -> u32 { return (ci >= (long long)corg && ci < cend && (u64)ci < ck.total_len) ? (s_nbits[(u32)(ci - corg) >> 5] >> ((u32)(ci - corg) & 31u)) & 1u : 0u; };
        /* compact index of output position jo: piece A for jo < bnd, piece B after; ci = c + sgn * jo */
        long long cA, cB; u32 bnd = rl; const int sgn = rc ? -1 : 1;
        if (!rc) { cA = so; cB = so; }
        else if (ov >= 0) { cA = so - ov + (long long)rl - 1; cB = cA; }
        else { const u32 a = (u32)(-ov); cA = so - (long long)prev_rl + a - 1; cB = so + (long long)rl - 1; bnd = a < rl ? a : rl; }
        auto slow_base = [&](u32 jo) -> u8 {
            const long long ci = (jo < bnd ? cA : cB) + (long long)sgn * jo;
            u8 base = 'N';
            if (ci >= 0 && ci < unpacked) { const u32 code = (code_byte(ci >> 2) >> (2 * (ci & 3))) & 3u; base = code == 0 ? 'G' : code == 1 ? 'A' : code == 2 ? 'T' : 'C'; }
            if (npos_mode) { if (nbit(ci)) base = 'N'; }
            else if (q[rc ? rl - 1 - jo : jo] == nq) base = 'N';
            return rc ? complement_base(base) : base;
        };
        /* Both lines are produced destination-first: up to three bytes until the output is word aligned, then whole aligned
         * 32-bit words, then up to three bytes.  The source phase (plane byte offset, 2-bit phase) is constant along a line. */
        const int d = rc ? -1 : 1;
        if (half == 1) {
            /* quality line: a byte-shifted (reverse strand: byte-reversed) copy out of the tile, one PRMT per word */
            u32 need = (4u - (u32)(reinterpret_cast<uintptr_t>(o_qual) & 3u)) & 3u; if (need > rl) need = rl;
            for (u32 jo = 0; jo < need; jo++) o_qual[jo] = q[rc ? rl - 1 - jo : jo];
            const u32 nw = (rl - need) >> 2;
            u32* dw = reinterpret_cast<u32*>(o_qual + need);
            if (nw) {
                const uintptr_t p0 = reinterpret_cast<uintptr_t>(rc ? q + (rl - 4u - need) : q + need);
                const u32* w = reinterpret_cast<const u32*>(p0 & ~(uintptr_t)3);
                const u32 sel = (rc ? 0x0123u : 0x3210u) + 0x1111u * (u32)(p0 & 3u);
#pragma unroll 4
                for (u32 m = 0; m < nw; m++) { dw[m] = __byte_perm(w[0], w[1], sel); w += d; }
            }
            for (u32 jo = need + 4u * nw; jo < rl; jo++) o_qual[jo] = q[rc ? rl - 1 - jo : jo];
        } else {
            u32 need = (4u - (u32)(reinterpret_cast<uintptr_t>(o_seq) & 3u)) & 3u; if (need > rl) need = rl;
            for (u32 jo = 0; jo < need; jo++) o_seq[jo] = slow_base(jo);
            const u32 nw = (rl - need) >> 2;
            u32* dw = reinterpret_cast<u32*>(o_seq + need);
            /* compact positions a word may load codes for: inside the unpacked range, and not closer than 8 bytes to the end of the body */
            long long lim = unpacked;
            {
                const long long avail = (long long)b.body_len - (long long)(ck.in_off + ck.off_seq);
                const long long safe = avail > 8 ? 4 * (avail - 8) : 0;
                if (lim > safe) lim = safe;
            }
            /* words [m0, m1) of the piece jo in [jlo, jhi) with compact base cb whose four positions can take the fast path */
            /* ... and inside the staged piece of the column, whose loads of 8 bytes must stay inside it too */
            long long lim_lo = 4 * sq_off;
            if (lim_lo < 0) lim_lo = 0;
            { const long long st_hi = 4 * (sq_off + (long long)sq_n) - 32; if (lim > st_hi) lim = st_hi; }
            auto range = [&](long long cb, long long jlo, long long jhi, u32& m0, u32& m1) {
                long long a, z;                                    /* allowed first positions jo of a word: a <= jo <= z */
                if (!rc) { a = lim_lo - cb; z = lim - cb - 4; } else { a = cb - lim + 1; z = cb - 3 - lim_lo; }
                if (npos_mode) {                                   /* and inside the N bitmap */
                    if (!rc) { if (a < (long long)corg - cb) a = (long long)corg - cb; if (z > cend - cb - 4) z = cend - cb - 4; }
                    else { if (a < cb - cend + 1) a = cb - cend + 1; if (z > cb - 3 - (long long)corg) z = cb - 3 - (long long)corg; }
                }
                if (a < jlo) a = jlo;
                if (z > jhi - 4) z = jhi - 4;
                long long lo_m = a <= (long long)need ? 0 : (a - need + 3) >> 2;
                long long hi_m = z < (long long)need ? 0 : ((z - need) >> 2) + 1;
                if (hi_m > (long long)nw) hi_m = nw;
                if (lo_m > hi_m) lo_m = hi_m;
                m0 = (u32)lo_m; m1 = (u32)hi_m;
            };
            u32 a0, a1, b0, b1;
            if (cA == cB) { range(cA, 0, rl, a0, a1); b0 = b1 = a1; }
            else { range(cA, 0, bnd, a0, a1); range(cB, bnd, rl, b0, b1); if (b0 < a1) b0 = a1; if (b1 < b0) b1 = b0; }
            auto slow_word = [&](u32 m) {
                const u32 jo = need + 4u * m;
                dw[m] = (u32)slow_base(jo) | ((u32)slow_base(jo + 1) << 8) | ((u32)slow_base(jo + 2) << 16) | ((u32)slow_base(jo + 3) << 24);
            };
            const u32* lut = rc ? s_lut_rc : s_lut_fwd;
            /* sixteen bases per 32-bit load of the 2-bit column, four per table lookup; 'N' from a SIMD compare of the quality word */
            auto fast_run = [&](u32 m0, u32 m1, long long cb) {
                if (m0 >= m1) return;
                const u32 jo0 = need + 4u * m0;
                const uintptr_t p0 = reinterpret_cast<uintptr_t>(rc ? q + (rl - 4u - jo0) : q + jo0);
                const u32* qw = reinterpret_cast<const u32*>(p0 & ~(uintptr_t)3);
                const u32 qsel = (rc ? 0x0123u : 0x3210u) + 0x1111u * (u32)(p0 & 3u);
                const long long ci0 = rc ? cb - (long long)jo0 - 15 : cb + (long long)jo0;     /* lowest compact index of the first 16 */
                const uintptr_t ca = reinterpret_cast<uintptr_t>(s_seq + ((ci0 >> 2) - sq_off));     /* shared memory; same 16-byte phase as the body */
                const u32* cw = reinterpret_cast<const u32*>(ca & ~(uintptr_t)3);
                const u32 sh = 8u * (u32)(ca & 3u) + 2u * (u32)(ci0 & 3);
                for (u32 m = m0; m < m1; m += 4) {
                    const u32 codes = __funnelshift_r(cw[0], cw[1], sh);
                    cw += d;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        if (m + t < m1) {
                            const u32 code8 = (codes >> (rc ? 24 - 8 * t : 8 * t)) & 0xFFu;
                            u32 bw = lut[code8];
                            if (npos_mode) {
                                const u32 jo = need + 4u * (m + t);
                                const long long ci_lo = rc ? cb - (long long)jo - 3 : cb + (long long)jo;
                                const u32 rel = (u32)(ci_lo - (long long)corg);
                                const u32 wi = rel >> 5, bp = rel & 31u;
                                u32 m4 = (s_nbits[wi] >> bp) & 0xFu;
                                if (bp > 28u && wi + 1u < cfg.nbits_words) m4 |= (s_nbits[wi + 1] << (32u - bp)) & 0xFu;
                                if (rc) m4 = ((m4 & 1u) << 3) | ((m4 & 2u) << 1) | ((m4 & 4u) >> 1) | ((m4 & 8u) >> 3);
                                const u32 mask = ((m4 | (m4 << 7) | (m4 << 14) | (m4 << 21)) & 0x01010101u) * 0xFFu;
                                bw = (bw & ~mask) | (0x4E4E4E4Eu & mask);
                            } else {
                                const u32 qv = __byte_perm(qw[0], qw[1], qsel);
                                qw += d;
                                const u32 fl7 = eq_bytes(qv, nq4);
                                if (fl7) { const u32 mask = (fl7 >> 7) * 0xFFu; bw = (bw & ~mask) | (0x4E4E4E4Eu & mask); }
                            }
                            dw[m + t] = bw;
                        }
                    }
                }
            };
            u32 m = 0;
            for (; m < a0; m++) slow_word(m);
            fast_run(a0, a1, cA); if (m < a1) m = a1;
            for (; m < b0; m++) slow_word(m);
            fast_run(b0, b1, cB); if (m < b1) m = b1;
            for (; m < nw; m++) slow_word(m);
            for (u32 jo = need + 4u * nw; jo < rl; jo++) o_seq[jo] = slow_base(jo);
        }
    }
    /* ---- the records leave shared memory: the 16-byte aligned body of each output stream as ONE TMA bulk store (shared ->
     * global, issued by one thread; `UBLKCP` in SASS), the few bytes before and after it with plain stores */
#ifndef RPQ_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        /* this thread's shared stores, visible to the bulk copy engine */
#endif
    __syncthreads();
    for (u32 s = 0; s < nstreams; s++) {
        const u64 a = s ? t_start[1] : t_start[0], e = s ? t_end[1] : t_end[0];
        if (a == ~0ull || e <= a) continue;
        const u64 base = a & ~15ull;
        u8* g = b.out[s];
        const u8* sm = s_out[s];
        const u64 v0 = (a + 15) & ~15ull, v1 = e & ~15ull;
        if (v0 >= v1) { for (u64 p = a + tid; p < e; p += blockDim.x) g[p] = sm[p - base]; continue; }
        for (u64 p = a + tid; p < v0; p += blockDim.x) g[p] = sm[p - base];
#ifdef RPQ_EMU
        const u32 nvec = (u32)((v1 - v0) >> 4);
        uint4* gd = reinterpret_cast<uint4*>(g + v0);
        const uint4* sd = reinterpret_cast<const uint4*>(sm + (v0 - base));
        for (u32 k = tid; k < nvec; k += blockDim.x) gd[k] = sd[k];
#else
        if (tid == (int)(32u * s)) {                                    /* one thread per stream, in different warps */
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(g + v0), "r"((u32)__cvta_generic_to_shared(sm + (v0 - base))), "r"((u32)(v1 - v0)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#endif
        for (u64 p = v1 + tid; p < e; p += blockDim.x) g[p] = sm[p - base];
    }
#ifndef RPQ_EMU
    if (tid == 0 || tid == 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    /* shared memory must outlive the reads of the copy */
#endif
    (void)warp;
}

}  // namespace rpq
