/*
 * rpq_encode.cuh - the encode kernels: everything RfqCodec::encodeChunk (reference src/rfqcodec.cpp:163-586) does
 * for a chunk, for all chunks of a batch at once, writing the serialised RfqChunk (src/rfqchunk.cpp:230-312) directly.
 *
 *   k_meta0        FastqMeta::parse of every chunk's first read                (src/fastqmeta.cpp:22-80, rfqcodec.cpp:181-192)
 *   k_meta         one warp per read / pair: tokenise, compare with read 0 (:225-234), the PE consistency test
 *                  (:233-270, Q10), reverse complement + overlap of the pair (:372-385, :1391-1438)
 *   k_chunk_finish chunk flags (:437-448), interleave decision, per-read scans of kept bases / qualities / name parts
 *   k_coords       X and Y varint-delta streams (:1262-1330)
 *   k_streams      position streams of the quality column (:625-765) and of 'N' in the compacted bases (:420-426)
 *   k_layout       stream placement, column sizes, mSize (src/rfqchunk.cpp:141-159, Q2)
 *   k_emit / k_head / k_gather  2-bit pack (:588-604, Q14) and all columns into the final byte layout
 */
#pragma once
#include "rpq_common.cuh"

namespace rpq {

/* ================================================================== FastqMeta::parse on a warp ==== */

/* atoi as glibc does it: (int)strtol(s, 0, 10) - isspace skip, sign, digits, saturating at LONG_MIN/MAX */
/* (not inlined: the rare path of four call sites per kernel, a fifth of k_meta3's code otherwise - its warps wait for instructions) */
__device__ __noinline__ int atoi_like(const u8* s, int n) {
    int i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= '\t' && s[i] <= '\r'))) i++;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; i++; }
    u64 acc = 0; bool sat = false;
    const u64 lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) {
        const u32 d = (u32)(s[i] - '0');
        if (!sat) {
            if (acc > (lim - d) / 10) { sat = true; acc = lim; }
            else acc = acc * 10 + d;
        }
    }
    const long long v = neg ? (long long)(0ull - acc) : (long long)acc;
    return (int)v;
}

/*
 * Closed form of the reference's positional state machine (src/fastqmeta.cpp:22-80).  Let c1<c2<... be the colon
 * positions and S the first space.  The scan stops at the 7th colon if it precedes every space, else at S.
 *   stop at c7      : lane=(c3,c4) tile=(c4,c5) x=(c5,c6) y=(c6,c7)                  name1=[0,c3) name2=[c7,len)
 *   stop at S, n=6  : lane=(c3,c4) tile=(c4,c5) x=(c5,c6) y=(c6,S)                   name1=[0,c3) name2=[S,len)
 *   stop at S, n=5  : lane=(c3,c4) tile=(c5,S)  (the space overwrites tile)          name1=[0,c3)
 *   stop at S, n=4  : lane=(c4,S)  (the space overwrites lane and moves the start)   name1=[0,c4)
 *   otherwise       : no lane/tile/x/y, whole name is name1
 * where (a,b) = atoi of name[a+1, b) and n = colons before S.  `name` is in shared memory; all lanes return the same.
 */
__device__ inline ReadMeta warp_tokenise(const u8* name, int len, int lane) {
    int ncol = 0, c3 = -1, c4 = -1, c5 = -1, c6 = -1, c7 = -1, S = -1;
    bool done = false;
    for (int r = 0; r * 32 < len && !done; r++) {
        const int p = r * 32 + lane;
        const u8 c = p < len ? name[p] : 0;
        u32 cm = __ballot_sync(0xffffffffu, c == ':');
        const u32 sm = __ballot_sync(0xffffffffu, c == ' ');
        int sp = -1;
        if (sm) { sp = __ffs((int)sm) - 1; cm &= (1u << sp) - 1u; }
        while (cm && ncol < 7) {
            const int q = __ffs((int)cm) - 1;
            cm &= cm - 1;
            ncol++;
            const int pos = r * 32 + q;
            if (ncol == 3) c3 = pos; else if (ncol == 4) c4 = pos; else if (ncol == 5) c5 = pos; else if (ncol == 6) c6 = pos; else if (ncol == 7) c7 = pos;
        }
        if (ncol == 7) done = true;
        else if (sm) { S = r * 32 + sp; done = true; }
    }
    ReadMeta m;
    m.x = 0; m.y = 0; m.tile = 0; m.lane = 0; m.has = 0;
    m.name_len = (u8)len; m.strand_len = 0;
    m.name1_len = (u8)len; m.name2_off = (u8)len;
    int stop = -1, n1 = -1;
    int fa[4] = {-1, -1, -1, -1}, fb[4] = {0, 0, 0, 0};     /* fields: lane, tile, x, y as (from, to) */
    if (ncol == 7) { stop = c7; n1 = c3; fa[0] = c3; fb[0] = c4; fa[1] = c4; fb[1] = c5; fa[2] = c5; fb[2] = c6; fa[3] = c6; fb[3] = c7; }
    else if (S >= 0 && ncol == 6) { stop = S; n1 = c3; fa[0] = c3; fb[0] = c4; fa[1] = c4; fb[1] = c5; fa[2] = c5; fb[2] = c6; fa[3] = c6; fb[3] = S; }
    else if (S >= 0 && ncol == 5) { stop = S; n1 = c3; fa[0] = c3; fb[0] = c4; fa[1] = c5; fb[1] = S; }
    else if (S >= 0 && ncol == 4) { stop = S; n1 = c4; fa[0] = c4; fb[0] = S; }
    if (stop > 0) {           /* coordsEndAt > 0 (coordsStartAt > 0 always holds once a 4th colon was seen) */
        int v = 0;
        if (lane < 4 && fa[lane] >= 0) v = atoi_like(name + fa[lane] + 1, fb[lane] - fa[lane] - 1);
        const int vl = __shfl_sync(0xffffffffu, v, 0), vt = __shfl_sync(0xffffffffu, v, 1);
        const int vx = __shfl_sync(0xffffffffu, v, 2), vy = __shfl_sync(0xffffffffu, v, 3);
        m.lane = (u8)vl; m.tile = (u16)vt; m.x = (u32)vx; m.y = (u32)vy;
        m.has = 1; m.name1_len = (u8)n1; m.name2_off = (u8)stop;
    }
    return m;
}

/* copy a line of the FASTQ image into shared memory (bytes; coalesced across the warp) */
__device__ __forceinline__ void warp_load_line(const u8* __restrict__ text, u32 start, int len, u8* dst, int lane) {
    for (int k = lane; k < len; k += 32) dst[k] = text[start + k];
}

constexpr int META_WARPS = 8;
constexpr int META_SEQ_SMEM = 512;        /* reads up to this length are staged in shared memory for the overlap search */

__global__ void __launch_bounds__(32 * META_WARPS) k_meta0(EncBatchDev b) {
    __shared__ u8 s_name[META_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 c = blockIdx.x * META_WARPS + w;
    if (c >= b.n_chunks) return;
    const u32 i = b.chunk_first[c];
    u32 f, rec; read_locus(b, i, f, rec);
    const TextDev& t = b.t[f];
    const u32 s = line_start(t, 4 * rec);
    const int len = (int)(line_end(t, 4 * rec) - s);
    warp_load_line(t.text, s, len, s_name[w], lane);
    __syncwarp();
    ReadMeta m = warp_tokenise(s_name[w], len, lane);
    m.strand_len = (u8)(line_end(t, 4 * rec + 2) - line_start(t, 4 * rec + 2));
    if (lane == 0) b.meta0[c] = m;
}

/* ------------------------------------------------------------------ overlap (src/rfqcodec.cpp:1391-1438) ---- */
/*
 * Forward: smallest o in [12, minlen] with r1[len1-o, len1) == r2[0, o): the 12-byte head of r2 is searched in r1 from
 * the right (largest start first = smallest o); a lane tests one start, hits are verified by the whole warp.
 * Backward: same with the roles swapped; reported as -o.  A1/A2 are byte accessors (shared memory or the text).
 */
template <class A1, class A2>
__device__ inline int warp_overlap_dir(const A1& a, int la, const A2& p, int lp, int lane) {
    /* find smallest o>=12, o<=min(la,lp) with a[la-o+i] == p[i] for i<o */
    const int minlen = la < lp ? la : lp;
    if (minlen < 12) return 0;
    const u32 w0 = (u32)p(0) | ((u32)p(1) << 8) | ((u32)p(2) << 16) | ((u32)p(3) << 24);
    for (int o0 = 12; o0 <= minlen; o0 += 32) {
        const int o = o0 + lane;
        bool hit = false;
        if (o <= minlen) {
            const int s = la - o;
            const u32 w = (u32)a(s) | ((u32)a(s + 1) << 8) | ((u32)a(s + 2) << 16) | ((u32)a(s + 3) << 24);
            hit = w == w0;
        }
        u32 m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            const int oc = o0 + l, s = la - oc;
            bool ok = true;
            for (int i = lane; i < oc; i += 32) if (a(s + i) != p(i)) ok = false;
            if (__all_sync(0xffffffffu, ok)) return oc;
        }
    }
    return 0;
}

struct SmemBytes { const u8* p; __device__ __forceinline__ u8 operator()(int i) const { return p[i]; } };
struct TextFwd { const u8* p; __device__ __forceinline__ u8 operator()(int i) const { return p[i]; } };
struct TextRc { const u8* p; int len; __device__ __forceinline__ u8 operator()(int i) const { return complement_base(p[len - 1 - i]); } };

template <class A1, class A2>
__device__ inline int warp_overlap(const A1& r1, int len1, const A2& r2, int len2, int lane) {
    int o = warp_overlap_dir(r1, len1, r2, len2, lane);
    if (o) return o;
    o = warp_overlap_dir(r2, len2, r1, len1, lane);
    return -o;
}

/* ------------------------------------------------------------------ k_meta ---- */
constexpr u32 AB_READ_LEN = 1u << 0, AB_N1LEN = 1u << 1, AB_N2LEN = 1u << 2, AB_SLEN = 1u << 3, AB_LANE = 1u << 4, AB_TILE = 1u << 5,
              AB_N1 = 1u << 6, AB_STRAND = 1u << 8;        /* same bit positions as the chunk flags (src/rfqchunk.h:25-41) */

__device__ __forceinline__ bool warp_bytes_equal(const u8* a, const u8* bb, int n, int lane) {
    bool ok = true;
    for (int k = lane; k < n; k += 32) if (a[k] != bb[k]) ok = false;
    return __all_sync(0xffffffffu, ok) != 0;
}

/* one warp per unit: a read (SE) or a pair (PE) */
__global__ void __launch_bounds__(32 * META_WARPS) k_meta(EncBatchDev b, HeaderDev h, u32 n_units, int seq_in_smem) {
    __shared__ u8 s_name[META_WARPS][2][256];
    RPQ_DYN_SMEM(dyn);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 u = blockIdx.x * META_WARPS + w;
    if (u >= n_units) return;
    const u32 per = b.is_pe ? 2u : 1u;
    const u32 i0 = u * per;
    const u32 c = chunk_of_read(b, i0);
    const ChunkDev& ck = b.chunks[c];
    const u32 first = b.chunk_first[c];
    const ReadMeta m0 = b.meta0[c];
    u32 f0, rec0; read_locus(b, first, f0, rec0);
    const TextDev& t0 = b.t[f0];
    const u8* name0 = t0.text + line_start(t0, 4 * rec0);
    const u8* strand0 = t0.text + line_start(t0, 4 * rec0 + 2);
    const u32 rlen0 = b.rlen[first];
    const int n2len0 = (int)m0.name_len - (int)m0.name2_off;
    (void)ck;

    u32 clear = 0;                 /* and_bits to clear for this chunk */
    ReadMeta mm[2];
    bool eq0[2] = {true, true};    /* name2 == read0.name2 */
    for (u32 k = 0; k < per; k++) {
        const u32 i = i0 + k;
        u32 f, rec; read_locus(b, i, f, rec);
        const TextDev& t = b.t[f];
        const u32 ns = line_start(t, 4 * rec);
        const int nlen = (int)(line_end(t, 4 * rec) - ns);
        u8* nm = s_name[w][k];
        warp_load_line(t.text, ns, nlen < 256 ? nlen : 255, nm, lane);
        __syncwarp();
        ReadMeta m = warp_tokenise(nm, nlen < 256 ? nlen : 255, lane);
        const u32 ss = line_start(t, 4 * rec + 2);
        const int slen = (int)(line_end(t, 4 * rec + 2) - ss);
        m.strand_len = (u8)slen;
        mm[k] = m;
        if (lane == 0) b.meta[i] = m;
        /* compare with the chunk's first read: src/rfqcodec.cpp:225-232 */
        if (b.rlen[i] != rlen0) clear |= AB_READ_LEN;
        if (m.name1_len != m0.name1_len) clear |= AB_N1LEN;
        const int n2len = (int)m.name_len - (int)m.name2_off;
        if (n2len != n2len0) clear |= AB_N2LEN;
        if (m.strand_len != m0.strand_len) clear |= AB_SLEN;
        if (m.lane != m0.lane) clear |= AB_LANE;
        if (m.tile != m0.tile) clear |= AB_TILE;
        if (m.name1_len != m0.name1_len || !warp_bytes_equal(nm, name0, m.name1_len, lane)) clear |= AB_N1;
        {
            bool ok = m.strand_len == m0.strand_len;
            if (ok) { const u8* sp = t.text + ss; bool e = true; for (int q = lane; q < slen; q += 32) if (sp[q] != strand0[q]) e = false; ok = __all_sync(0xffffffffu, e) != 0; }
            if (!ok) clear |= AB_STRAND;
        }
        eq0[k] = (n2len == n2len0) && warp_bytes_equal(nm + m.name2_off, name0 + m0.name2_off, n2len, lane);
    }
    if (lane == 0 && clear) atomicAnd(&b.chunks[c].and_bits, ~clear);

    const u32 rel = i0 - first;         /* chunk-relative index of the unit's first read (even for PE) */
    if (!b.is_pe) {
        if (!eq0[0] && lane == 0) {
            if (rel & 1u) atomicMax(&b.chunks[c].last_odd_neq, rel + 1); else atomicOr(&b.chunks[c].even_neq, 1u);
        }
        return;
    }
    /* ---- paired end: Q10 bookkeeping (src/rfqcodec.cpp:233-270) */
    if (lane == 0) {
        if (!eq0[0]) atomicOr(&b.chunks[c].even_neq, 1u);
        if (!eq0[1]) atomicMax(&b.chunks[c].last_odd_neq, rel + 2);          /* 1 + (rel+1) */
    }
    if (!h.support_interleaved) return;                                      /* canBePeInterleaved false from the start */
    {
        /* A: R1.name2 with the header's diff char substituted == R2.name2 */
        const int l1 = (int)mm[0].name_len - (int)mm[0].name2_off, l2 = (int)mm[1].name_len - (int)mm[1].name2_off;
        bool okA = l1 == l2;
        if (okA) {
            const u8* a = s_name[w][0] + mm[0].name2_off; const u8* q = s_name[w][1] + mm[1].name2_off;
            bool e = true;
            for (int k = lane; k < l1; k += 32) { u8 ch = a[k]; if (h.name2_diff_char != 0 && k == (int)h.name2_diff_pos) ch = h.name2_diff_char; if (ch != q[k]) e = false; }
            okA = __all_sync(0xffffffffu, e) != 0;
        }
        const bool okB = mm[0].lane == mm[1].lane && mm[0].tile == mm[1].tile && mm[0].x == mm[1].x && mm[0].y == mm[1].y;
        if (lane == 0) {
            if (!okA) atomicMin(&b.chunks[c].fA, rel + 1);
            if (!okB) atomicMin(&b.chunks[c].fB, rel + 1);
        }
    }
    if (!(h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) { if (lane == 0) b.ov[u] = 0; return; }
    /* ---- overlap of R1 with revcomp(R2); used only if the chunk stays interleaved */
    u32 f1, r1, f2, r2; read_locus(b, i0, f1, r1); read_locus(b, i0 + 1, f2, r2);
    const u8* q1 = b.t[f1].text + line_start(b.t[f1], 4 * r1 + 1);
    const u8* q2 = b.t[f2].text + line_start(b.t[f2], 4 * r2 + 1);
    const int len1 = (int)b.rlen[i0], len2 = (int)b.rlen[i0 + 1];
    int o;
    if (seq_in_smem) {
        u8* s1 = dyn + (size_t)w * 2 * META_SEQ_SMEM; u8* s2 = s1 + META_SEQ_SMEM;
        for (int k = lane; k < len1; k += 32) s1[k] = q1[k];
        for (int k = lane; k < len2; k += 32) s2[k] = complement_base(q2[len2 - 1 - k]);
        __syncwarp();
        SmemBytes a{s1}, p{s2};
        o = warp_overlap(a, len1, p, len2, lane);
    } else {
        TextFwd a{q1}; TextRc p{q2, len2};
        o = warp_overlap(a, len1, p, len2, lane);
    }
    /* shift so that it fits a signed byte, else drop (src/rfqcodec.cpp:378-383) */
    if (o + (int)h.overlap_shift > 127) o = 0;
    if (o + (int)h.overlap_shift < -127) o = 0;
    if (lane == 0) b.ov[u] = (short)o;
}

__global__ void k_init_chunks(EncBatchDev b) {
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.n_chunks) return;
    ChunkDev z; memset(&z, 0, sizeof z);
    z.and_bits = 0xFFFFFFFFu; z.fA = NONE32; z.fB = NONE32;
    b.chunks[c] = z;
}

/* ================================================================== k_chunk_finish ==== */
constexpr int FIN_THREADS = 256;

struct Scan5 { u32 a, b, c, d, e; };
__device__ __forceinline__ Scan5 operator+(const Scan5& x, const Scan5& y) { return Scan5{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d, x.e + y.e}; }

/* kept bases of read i (chunk-relative index rel) once the interleave decision is known: src/rfqcodec.cpp:388-403 */
__device__ __forceinline__ u32 kept_bases(const EncBatchDev& b, const HeaderDev& h, bool interleaved, u32 i, u32 rel) {
    const u32 rl = b.rlen[i];
    if (interleaved && (rel & 1u) && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) {
        const int o = b.ov[i >> 1];
        return rl - (u32)(o < 0 ? -o : o);
    }
    return rl;
}

__global__ void __launch_bounds__(FIN_THREADS) k_chunk_finish(EncBatchDev b, HeaderDev h) {
    __shared__ Scan5 s_warp[FIN_THREADS / 32];
    __shared__ Scan5 s_carry;
    const u32 c = blockIdx.x;
    ChunkDev& ck = b.chunks[c];
    const u32 first = b.chunk_first[c], count = b.chunk_first[c + 1] - first;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    /* Q10: T = min(first name2 failure, first coordinate failure + 1); the chunk stays interleaved iff there is none */
    const bool pe_start = b.is_pe && h.support_interleaved;
    u32 T = 0;
    if (pe_start) { const u32 fa = ck.fA, fb = ck.fB; T = fa < (fb == NONE32 ? NONE32 : fb + 1) ? fa : (fb == NONE32 ? NONE32 : fb + 1); }
    const bool interleaved = pe_start && T == NONE32;
    /* name2Same: every even read, and every odd read with index >= T, equals read0.name2 */
    const bool odd_bad = ck.last_odd_neq != 0 && (ck.last_odd_neq - 1) >= T;
    const bool n2_same = !ck.even_neq && !odd_bad;

    if (tid == 0) s_carry = Scan5{0, 0, 0, 0, 0};
    __syncthreads();
    for (u32 base = 0; base < count; base += FIN_THREADS) {
        const u32 rel = base + tid;
        Scan5 v{0, 0, 0, 0, 0};
        if (rel < count) {
            const u32 i = first + rel;
            const ReadMeta m = b.meta[i];
            v.a = kept_bases(b, h, interleaved, i, rel);
            v.b = b.rlen[i];
            v.c = m.name1_len; v.d = (u32)m.name_len - (u32)m.name2_off; v.e = m.strand_len;
        }
        Scan5 inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Scan5 t;
            t.a = __shfl_up_sync(0xffffffffu, inc.a, d); t.b = __shfl_up_sync(0xffffffffu, inc.b, d); t.c = __shfl_up_sync(0xffffffffu, inc.c, d);
            t.d = __shfl_up_sync(0xffffffffu, inc.d, d); t.e = __shfl_up_sync(0xffffffffu, inc.e, d);
            if (lane >= d) inc = inc + t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        Scan5 pre = s_carry;
        for (int q = 0; q < warp; q++) pre = pre + s_warp[q];
        if (rel < count) {
            const u32 i = first + rel;
            b.seqoff[i] = pre.a + inc.a - v.a;
            b.qualoff[i] = pre.b + inc.b - v.b;
            b.n1off[i] = pre.c + inc.c - v.c;
            b.n2off[i] = pre.d + inc.d - v.d;
            b.soff[i] = pre.e + inc.e - v.e;
        }
        __syncthreads();
        if (tid == 0) { Scan5 t = s_carry; for (int q = 0; q < FIN_THREADS / 32; q++) t = t + s_warp[q]; s_carry = t; }
        __syncthreads();
    }
    if (tid != 0) return;
    const Scan5 tot = s_carry;
    const ReadMeta m0 = b.meta0[c];
    u32 flags = ck.and_bits & (AB_READ_LEN | AB_N1LEN | AB_N2LEN | AB_SLEN | AB_LANE | AB_TILE | AB_N1 | AB_STRAND);
    if (n2_same) flags |= RPQ_NAME2_SAME;
    if (interleaved) flags |= RPQ_PE_INTERLEAVED;
    const u32 s = count, xy = interleaved ? s / 2 : s;
    ck.first = first; ck.count = count;
    ck.flags = flags; ck.interleaved = interleaved ? 1u : 0u; ck.xy_num = xy;
    ck.seq_kept = tot.a; ck.total_len = tot.b; ck.tot_n1 = tot.c; ck.tot_n2 = tot.d; ck.tot_strand = tot.e;
    const u32 rlb = h.read_length_bytes;
    ck.readlen_size = (flags & RPQ_READ_LEN_SAME) ? rlb : rlb * s;
    ck.n1len_size = (flags & RPQ_NAME1_LEN_SAME) ? 1 : s;
    ck.n2len_size = (flags & RPQ_NAME2_LEN_SAME) ? 1 : s;
    ck.slen_size = (flags & RPQ_STRAND_LEN_SAME) ? 1 : s;
    ck.lane_size = (flags & RPQ_LANE_SAME) ? 1 : xy;
    ck.tile_size = 2 * ((flags & RPQ_TILE_SAME) ? 1 : xy);
    ck.n1_size = (flags & RPQ_NAME1_SAME) ? m0.name1_len : tot.c;
    ck.n2_size = (flags & RPQ_NAME2_SAME) ? (u32)m0.name_len - (u32)m0.name2_off : tot.d;
    ck.strand_size = (flags & RPQ_STRAND_SAME) ? m0.strand_len : tot.e;
    ck.seq_size = (tot.a + 3) / 4;
    ck.ov_size = (interleaved && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) ? s / 2 : 0;
    /* where the chunk's last record ends in the text(s): for Q13 and for the caller's resume offsets */
    {
        const u32 last = first + count - 1;
        u32 f, rec; read_locus(b, last, f, rec);
        const u32 e_last = caller_break_first(b.t[f], 4 * rec + 3);    /* first break character after the last quality line */
        if (b.is_pe && b.two_files) {
            u32 f1, rec1; read_locus(b, last - 1, f1, rec1);
            ck.r1_end = caller_break_first(b.t[0], 4 * rec1 + 3); ck.r2_end = e_last;
        } else { ck.r1_end = e_last; ck.r2_end = e_last; }
    }
}

/* ================================================================== k_coords: src/rfqcodec.cpp:1262-1330 ==== */
constexpr int CO_THREADS = 256;

/* grid (n_chunks, 2): y=0 X column, y=1 Y column.  out = tmp + 3*first (3 bytes per read is the reference's own bound) */
__global__ void __launch_bounds__(CO_THREADS) k_coords(EncBatchDev b, HeaderDev h, u8* tmp_x, u8* tmp_y) {
    __shared__ u32 s_wtot[CO_THREADS / 32];
    __shared__ u32 s_wmax[CO_THREADS / 32];
    __shared__ u32 s_carry_off, s_carry_nonrep;      /* bytes so far; 1 + index of the last non-repeat element so far */
    const u32 c = blockIdx.x; const int col = blockIdx.y;
    if (!(h.flags & (col ? RPQ_HAS_Y : RPQ_HAS_X))) return;
    ChunkDev& ck = b.chunks[c];
    const u32 first = ck.first, n = ck.xy_num, step = ck.interleaved ? 2u : 1u;
    u8* out = (col ? tmp_y : tmp_x) + 3ull * first;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry_off = 0; s_carry_nonrep = 0; }
    __syncthreads();
    bool err = false;
    for (u32 base = 0; base < n; base += CO_THREADS) {
        const u32 k = base + tid;
        u32 v = 0, prev = 1000, nxt = 0; bool valid = k < n, has_next = false;
        if (valid) {
            const ReadMeta m = b.meta[first + k * step]; v = col ? m.y : m.x;
            if (k > 0) { const ReadMeta p = b.meta[first + (k - 1) * step]; prev = col ? p.y : p.x; }
            if (k + 1 < n) { const ReadMeta q = b.meta[first + (k + 1) * step]; nxt = col ? q.y : q.x; has_next = true; }
        }
        const bool rep = valid && v == prev;
        /* j = index inside the run of repeats = k - (1 + last non-repeat index) */
        u32 mark = (valid && !rep) ? k + 1 : 0;
        u32 inc = mark;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = t > inc ? t : inc; }
        if (lane == 31) s_wmax[warp] = inc;
        __syncthreads();
        u32 nonrep = s_carry_nonrep;
        for (int q = 0; q < warp; q++) nonrep = s_wmax[q] > nonrep ? s_wmax[q] : nonrep;
        nonrep = inc > nonrep ? inc : nonrep;            /* 1 + last non-repeat index <= k */
        u32 sz = 0; u8 t0 = 0, t1 = 0, t2 = 0;
        if (valid) {
            if (rep) {
                const u32 j = k - nonrep;                 /* 0-based position in the repeat run */
                const bool run_ends = !has_next || nxt != v;
                if ((j & 31u) == 31u || run_ends) { sz = 1; t0 = (u8)(0xC0 | (j & 31u)); }
            } else {
                const int diff = (int)(v - prev);
                if (diff > 0 && diff <= 64) { sz = 1; t0 = (u8)(0x80 | (diff - 1)); }
                else if (v <= 32767u) { sz = 2; t0 = (u8)(v >> 8); t1 = (u8)v; }
                else if (v < (1u << 21)) { sz = 3; t0 = (u8)(0xE0 | (v >> 16)); t1 = (u8)(v >> 8); t2 = (u8)v; }
                else err = true;
            }
        }
        u32 wt; const u32 ex = warp_excl_scan(sz, lane, wt);
        if (lane == 0) s_wtot[warp] = wt;
        __syncthreads();
        u32 off = s_carry_off;
        for (int q = 0; q < warp; q++) off += s_wtot[q];
        off += ex;
        if (sz >= 1) out[off] = t0;
        if (sz >= 2) out[off + 1] = t1;
        if (sz >= 3) out[off + 2] = t2;
        __syncthreads();
        if (tid == 0) {
            u32 o = s_carry_off, nr = s_carry_nonrep;
            for (int q = 0; q < CO_THREADS / 32; q++) { o += s_wtot[q]; nr = s_wmax[q] > nr ? s_wmax[q] : nr; }
            s_carry_off = o; s_carry_nonrep = nr;
        }
        __syncthreads();
    }
    if (err) atomicOr(b.err, ERRBIT_COORD);
    if (tid == 0) { if (col) ck.y_size = s_carry_off; else ck.x_size = s_carry_off; }
}

/* ================================================================== k_streams: position streams ==== */
/*
 * encodeSingleQualByCol (src/rfqcodec.cpp:625-710) in closed form.  For every maximal run [p, p+L) of a stream value:
 *   p      -> distance token of d = p - (previous position of the value, -1 if none): d<=128: 1 byte d-1;
 *             d<=16384: 2 bytes 0x80|hi,lo; else 4 bytes 0xE0|b3,b2,b1,b0 of d-1
 *   p == 0 -> position 1, if it belongs to the run, is a second distance token (0x00)            (the `cur > 1` guard, Q16)
 *   the rest, from s0 = (p==0 ? 2 : 1), in groups of <= 32: one byte 0xC0|(len-1) per group.
 * A token is owned by its head position.  A CTA takes a SPAN of consecutive positions of one chunk, stages them in
 * shared memory (for the quality column: gathered from the reads' quality lines, reversed for R2 of an interleaved
 * chunk; for N positions: the kept, possibly reverse-complemented bases) and codes them (k_streams4: run-parallel from bit
 * masks; k_streams7: dense columns, a thread per 64 positions; k_streams3 / k_streams2: the earlier generations, kept for the
 * spans and inputs those cannot take).  Only the very first distance token of a stream inside a span can depend on data left
 * of the span; it is left to k_layout (`firstpos`).
 */
constexpr int ST_SPAN = 16384;       /* positions per CTA */
constexpr int ST_HALO = 64;          /* staged on both sides of the span (look-back for run starts, look-ahead for run ends) */

struct SpanDir {                     /* per (span, stream): what the span produced */
    u32 bytes;                       /* token bytes written to the slot for this stream (without the deferred token) */
    u32 slot_off;                    /* where they start inside the span's slot */
    u32 firstpos;                    /* position of the deferred first distance token, NONE32 if none */
    u32 lastpos;                     /* last position of the value inside the span, NONE32 if none */
    u32 dst;                         /* k_layout: destination offset of this piece inside the column */
    u32 first_tok;                   /* k_layout: deferred token, bytes little-end first, */
    u32 first_len;                   /*           and its length (0..4) */
    u32 pad;
};

struct StreamJob {                   /* one per kind (quality / N positions) */
    const u32* span_first;           /* [n_chunks+1] first span of each chunk */
    SpanDir* dir;                    /* [(total spans) * (nstreams)] */
    u8* slots;                       /* token bytes of all spans, bump-allocated (worst case 5 bytes per position) */
    u64 slot_cap;
    u64* slot_cursor;                /* bump pointer */
    u64* span_slot;                  /* [spans] where each span's bytes start in `slots` */
    u32* overflow;                   /* set when slot_cap was too small: the host grows it and repeats the batch */
    const u32* n_spans;              /* actual number of spans (the grid is an upper bound) */
    u32* span_read0;                 /* [spans] chunk-relative index of the first read whose positions reach into the span's staging window */
    u32* redo_count;                 /* k_streams4: spans left to k_streams3 ... */
    u32* redo_list;                  /* ... and which */
    u32* dense_count;                /* k_streams4: quality spans with more runs than its list holds, left to k_streams7 ... */
    u32* dense_list;                 /* ... and which (NULL: they go to the redo list) */
    u32* list_count;                 /* entries of dense_list */
    u32 list_takes_redo;             /* the coder behind dense_list also takes the spans with long runs (k_streams7) */
    u32 nstreams;                    /* nb + 1 (exceptions) for quality; 1 for N positions */
    u32 mode;                        /* 0 quality, 1 N positions */
};

/* chunk position -> byte, slow path (binary search over the chunk's reads); only for look-back beyond the halo */
__device__ inline u8 stream_byte_slow(const EncBatchDev& b, const ChunkDev& ck, u32 mode, u32 pos) {
    const u32* offs = mode ? b.seqoff : b.qualoff;
    u32 lo = 0, hi = ck.count;          /* largest rel with offs[first+rel] <= pos (reads with zero kept bases share offsets: take the last) */
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (offs[ck.first + mid] <= pos) lo = mid; else hi = mid; }
    const u32 i = ck.first + lo, j = pos - offs[i];
    u32 f, rec; read_locus(b, i, f, rec);
    const TextDev& t = b.t[f];
    const bool rev = ck.interleaved && (lo & 1u);
    const u32 rl = b.rlen[i];
    if (mode == 0) { const u8* q = t.text + line_start(t, 4 * rec + 3); return rev ? q[rl - 1 - j] : q[j]; }
    const u8* s = t.text + line_start(t, 4 * rec + 1);
    if (!rev) return s[j];
    const int o = b.ov[i >> 1];
    const u32 jj = o > 0 ? j + (u32)o : j;             /* forward overlap drops the first o bases of revcomp(R2) */
    return complement_base(s[rl - 1 - jj]);
}

/* stage positions [lo, hi) of the chunk's concatenation into sm[pos - lo] */
/* returns false when nothing but cleared bytes was staged (N positions of a window whose reads are all plain bases): the caller may
 * skip the span.  Called by every thread of the CTA. */
__device__ inline bool stage_positions(const EncBatchDev& b, const HeaderDev& h, const ChunkDev& ck, u32 mode, u32 lo, u32 hi, u8* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (hi <= lo) return true;
    const u32* offs = mode ? b.seqoff : b.qualoff;
    u32 a = 0, z = ck.count;            /* first read that can contain `lo`: largest rel with offs <= lo */
    while (z - a > 1) { const u32 mid = (a + z) >> 1; if (offs[ck.first + mid] <= lo) a = mid; else z = mid; }
    /* N positions: a read of plain bases only has none (k_meta3 has looked at every character), so the window is cleared and
     * only the other reads - a few per thousand - are staged */
    const bool sparse = mode != 0 && b.unclean != nullptr;
    __shared__ int s_staged;
    if (sparse) {
        if (threadIdx.x == 0) s_staged = 0;
        const u32 n16 = (hi - lo + 15u) >> 4;
        for (u32 k = threadIdx.x; k < n16; k += blockDim.x) reinterpret_cast<uint4*>(sm)[k] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
    }
    int staged = 0;
    for (u32 rel = a + warp; rel < ck.count; rel += nwarps) {
        const u32 i = ck.first + rel;
        const u32 off = offs[i];
        if (off >= hi) break;
        if (sparse && !b.unclean[i]) continue;
        const u32 rl = b.rlen[i];
        const u32 n = mode ? kept_bases(b, h, ck.interleaved != 0, i, rel) : rl;
        if (off + n <= lo) continue;
        staged = 1;
        u32 f, rec; read_locus(b, i, f, rec);
        const TextDev& t = b.t[f];
        const bool rev = ck.interleaved && (rel & 1u);
        const u32 j0 = lo > off ? lo - off : 0, j1 = (off + n < hi ? off + n : hi) - off;
        if (mode == 0) {
            const u8* q = t.text + b.loc[i].w;
            if (rev) for (u32 j = j0 + lane; j < j1; j += 32) sm[off + j - lo] = q[rl - 1 - j];
            else for (u32 j = j0 + lane; j < j1; j += 32) sm[off + j - lo] = q[j];
        } else {
            const u8* s = t.text + b.loc[i].y;
            if (rev) { const int o = b.ov[i >> 1]; const u32 sh = o > 0 ? (u32)o : 0u; for (u32 j = j0 + lane; j < j1; j += 32) sm[off + j - lo] = complement_base(s[rl - 1 - (j + sh)]); }
            else for (u32 j = j0 + lane; j < j1; j += 32) sm[off + j - lo] = s[j];
        }
    }
    if (!sparse) return true;
    if (staged && lane == 0) s_staged = 1;
    __syncthreads();
    return s_staged != 0;
}

struct TokSink {            /* count-only or writing */
    u8* out; u32 n;
    __device__ __forceinline__ void put(u8 v) { if (out) out[n] = v; n++; }
};

__device__ __forceinline__ void put_distance(TokSink& s, u32 dm1) {           /* dm1 = distance - 1 */
    if (dm1 < 128u) s.put((u8)dm1);
    else if (dm1 < (1u << 14)) { s.put((u8)(0x80u | (dm1 >> 8))); s.put((u8)dm1); }
    else { s.put((u8)(0xE0u | (dm1 >> 24))); s.put((u8)(dm1 >> 16)); s.put((u8)(dm1 >> 8)); s.put((u8)dm1); }
}
__device__ __forceinline__ u32 distance_len(u32 dm1) { return dm1 < 128u ? 1u : dm1 < (1u << 14) ? 2u : 4u; }

constexpr int LAY_THREADS = 128;

__device__ inline void layout_stream_set(const StreamJob& job, u32 c, u32 st, u32 table_bytes, u32* stream_len) {
    const u32 s0 = job.span_first[c], s1 = job.span_first[c + 1];
    u32 len = 0, prev_last = NONE32; bool have = false;
    for (u32 sp = s0; sp < s1; sp++) {
        SpanDir& d = job.dir[(size_t)sp * job.nstreams + st];
        u32 flen = 0, ftok = 0;
        if (d.firstpos != NONE32) {
            const u32 dm1 = have ? d.firstpos - prev_last - 1u : d.firstpos;      /* last = -1: d-1 = firstpos */
            TokSink t{reinterpret_cast<u8*>(&ftok), 0};
            put_distance(t, dm1);
            flen = t.n;
        }
        d.first_len = flen; d.first_tok = ftok;
        d.dst = len;                         /* relative to the stream start; the caller adds the stream base */
        len += flen + d.bytes;
        if (d.lastpos != NONE32) { prev_last = d.lastpos; have = true; }
    }
    (void)table_bytes;
    *stream_len = len;
}

__global__ void __launch_bounds__(LAY_THREADS) k_layout(EncBatchDev b, HeaderDev h, StreamJob qjob, StreamJob njob, int have_q, int have_n) {
    __shared__ u32 s_len[MAX_BINS + 2];
    const u32 c = blockIdx.x;
    ChunkDev& ck = b.chunks[c];
    const int tid = threadIdx.x;
    u32 qual_size = 0, npos_size = 0;
    if (h.flags & RPQ_DONT_ENCODE_QUAL) qual_size = ck.total_len;
    else if (have_q) {
        const u32 ns = qjob.nstreams;           /* nb streams + exceptions */
        for (u32 st = tid; st < ns; st += LAY_THREADS) layout_stream_set(qjob, c, st, 0, &s_len[st]);
        __syncthreads();
        if (tid == 0) {
            u32 base = 4u * (ns - 1);            /* the u32 length table comes first (src/rfqcodec.cpp:729-733) */
            for (u32 st = 0; st < ns; st++) { const u32 l = s_len[st]; s_len[st] = base; base += l; }
            s_len[ns] = base;
        }
        __syncthreads();
        const u32 s0 = qjob.span_first[c], s1 = qjob.span_first[c + 1];
        for (u32 k = tid; k < (s1 - s0) * ns; k += LAY_THREADS) { const u32 sp = s0 + k / ns, st = k % ns; qjob.dir[(size_t)sp * ns + st].dst += s_len[st]; }
        qual_size = s_len[ns];
        __syncthreads();
    }
    if (have_n) {
        if (tid == 0) layout_stream_set(njob, c, 0, 0, &s_len[MAX_BINS + 1]);
        __syncthreads();
        npos_size = s_len[MAX_BINS + 1];
    }
    if (tid != 0) return;
    ck.qual_size = qual_size; ck.npos_size = npos_size;
    if (!(ck.flags & RPQ_NAME1_SAME) || ((h.flags & RPQ_HAS_NAME2) && !(ck.flags & RPQ_NAME2_SAME)) || !(ck.flags & RPQ_STRAND_SAME))
        atomicOr(b.err, INFOBIT_NEED_NAMES);
    /* column offsets in RfqChunk::write order (src/rfqchunk.cpp:230-312) */
    u32 o = 18u + ((h.flags & RPQ_ENCODE_N_POS) ? 4u : 0u);
    ck.off_readlen = o; o += ck.readlen_size;
    ck.off_n1len = o; o += ck.n1len_size;
    ck.off_n2len = o; if (h.flags & RPQ_HAS_NAME2) o += ck.n2len_size;
    ck.off_slen = o; o += ck.slen_size;
    ck.off_lane = o; if (h.flags & RPQ_HAS_LANE) o += ck.lane_size;
    ck.off_tile = o; if (h.flags & RPQ_HAS_TILE) o += ck.tile_size;
    ck.off_x = o; if (h.flags & RPQ_HAS_X) o += 4u + ck.x_size;
    ck.off_y = o; if (h.flags & RPQ_HAS_Y) o += 4u + ck.y_size;
    ck.off_n1 = o; o += ck.n1_size;
    ck.off_n2 = o; if (h.flags & RPQ_HAS_NAME2) o += ck.n2_size;
    ck.off_strand = o; o += ck.strand_size;
    ck.off_seq = o; o += ck.seq_size;
    ck.off_qual = o; o += qual_size;
    ck.off_ov = o; o += ck.ov_size;
    ck.off_npos = o; if (h.flags & RPQ_ENCODE_N_POS) o += npos_size;
    ck.bytes = o;
    /* mSize as the reference computes it (Q2): tile bytes land in mLaneBufSize, lane bytes and mTileBufSize are never
     * counted, name2 columns are counted even when the header has no NAME2 */
    u32 ms = 18u + ck.readlen_size + ck.n1len_size + ck.n2len_size + ck.slen_size + ck.tile_size + ck.n1_size + ck.n2_size + ck.strand_size
           + ck.seq_size + qual_size + ck.ov_size;
    if (h.flags & RPQ_ENCODE_N_POS) ms += 4u + npos_size;
    if (h.flags & RPQ_HAS_X) ms += 4u + ck.x_size;
    if (h.flags & RPQ_HAS_Y) ms += 4u + ck.y_size;
    ck.msize = ms;
}

/* exclusive scan of chunk sizes -> out_offset; one CTA (n_chunks is a few thousand) */
__global__ void __launch_bounds__(256) k_chunk_offsets(EncBatchDev b, u64* total) {
    __shared__ u64 s_w[8];
    __shared__ u64 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u32 base = 0; base < b.n_chunks; base += 256) {
        const u32 c = base + tid;
        const u64 v = c < b.n_chunks ? b.chunks[c].bytes : 0;
        u64 inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        u64 pre = s_carry;
        for (int q = 0; q < warp; q++) pre += s_w[q];
        if (c < b.n_chunks) b.chunks[c].out_offset = pre + inc - v;
        __syncthreads();
        if (tid == 0) { u64 t = s_carry; for (int q = 0; q < 8; q++) t += s_w[q]; s_carry = t; }
        __syncthreads();
    }
    if (tid == 0) { total[0] = s_carry; total[1] = *b.err; }     /* the error bits ride along: one read-back for the host */
}

/* spans of every chunk for one stream kind; one CTA.  span_chunk has room for the host's upper bound. */
__global__ void __launch_bounds__(256) k_span_plan(EncBatchDev b, u32 mode, u32* span_first, u32* span_chunk, u32 cap, u32* n_spans) {
    __shared__ u32 s_w[8];
    __shared__ u32 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u32 base = 0; base < b.n_chunks; base += 256) {
        const u32 c = base + tid;
        u32 v = 0;
        if (c < b.n_chunks) { const u32 n = mode ? b.chunks[c].seq_kept : b.chunks[c].total_len; v = (n + ST_SPAN - 1) / ST_SPAN; }
        u32 wt; const u32 ex = warp_excl_scan(v, lane, wt);
        if (lane == 0) s_w[warp] = wt;
        __syncthreads();
        u32 pre = s_carry;
        for (int q = 0; q < warp; q++) pre += s_w[q];
        if (c < b.n_chunks) {
            const u32 f = pre + ex;
            span_first[c] = f;
            for (u32 k = 0; k < v && f + k < cap; k++) span_chunk[f + k] = c;
        }
        __syncthreads();
        if (tid == 0) { u32 t = s_carry; for (int q = 0; q < 8; q++) t += s_w[q]; s_carry = t; }
        __syncthreads();
    }
    if (tid == 0) { span_first[b.n_chunks] = s_carry; *n_spans = s_carry < cap ? s_carry : cap; }
}

/* first read of every span's staging window [lo - ST_HALO, ...): one thread per span, off the critical path of k_streams3 */
__global__ void k_span_reads(EncBatchDev b, StreamJob job, const u32* __restrict__ span_chunk) {
    const u32 span = blockIdx.x * blockDim.x + threadIdx.x;
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    const u32 lo = (span - job.span_first[c]) * ST_SPAN;
    const u32 sm_lo = lo >= ST_HALO ? lo - ST_HALO : 0;
    const u32* offs = job.mode ? b.seqoff : b.qualoff;
    u32 a = 0, z = ck.count;
    while (z - a > 1) { const u32 mid = (a + z) >> 1; if (offs[ck.first + mid] <= sm_lo) a = mid; else z = mid; }
    job.span_read0[span] = a;
}

/* ================================================================== emit ==== */
__device__ __forceinline__ void put_u32le(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }
__device__ __forceinline__ void put_u16le(u8* p, u16 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }

/* chunk header, single-value columns, X/Y streams: one warp per chunk */
__global__ void __launch_bounds__(128) k_head(EncBatchDev b, HeaderDev h, u8* out, const u8* tmp_x, const u8* tmp_y, rpq_encode_in params) {
    const int lane = threadIdx.x & 31;
    const u32 c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= b.n_chunks) return;
    ChunkDev& ck = b.chunks[c];
    u8* o = out + ck.out_offset;
    const ReadMeta m0 = b.meta0[c];
    u32 f0, rec0; read_locus(b, ck.first, f0, rec0);
    const TextDev& t0 = b.t[f0];
    if (lane == 0) {
        /* Q13: NO_LINE_BREAK bits are OR-ed in after mSize was computed */
        u32 flags = ck.flags;
        /* the post-loop flush (src/repaq.cpp:594-616,710-754): a last chunk that did not reach chunk_bases.  The reader has by then
         * tried the record after it: what counts is how far it got (k_cut_ends) */
        const bool flush = params.final && c == b.n_chunks - 1 && ck.total_len < params.chunk_bases;
        u64 e1 = ck.r1_end, e2 = b.two_files ? ck.r2_end : ck.r1_end;
        if (flush) { if (b.reach[0] > e1) e1 = b.reach[0]; const u32 r2 = b.two_files ? b.reach[1] : b.reach[0]; if (r2 > e2) e2 = r2; }
        if (e1 >= params.nobreak_from[0]) flags |= RPQ_NO_LINE_BREAK_AT_END;
        if (b.is_pe && e2 >= params.nobreak_from[b.two_files ? 1 : 0]) flags |= RPQ_NO_LINE_BREAK_AT_END_R2;
        /* tail_flags belong to the post-loop flush only */
        if (flush) flags |= params.tail_flags;
        ck.flags = flags;
        put_u32le(o, ck.msize); put_u32le(o + 4, ck.count); put_u16le(o + 8, (u16)flags);
        put_u32le(o + 10, ck.seq_size); put_u32le(o + 14, ck.qual_size);
        if (h.flags & RPQ_ENCODE_N_POS) put_u32le(o + 18, ck.npos_size);
        const u32 fl = ck.flags;
        if (fl & RPQ_READ_LEN_SAME) { const u32 rl = b.rlen[ck.first]; if (h.read_length_bytes == 1) o[ck.off_readlen] = (u8)rl; else put_u16le(o + ck.off_readlen, (u16)rl); }
        if (fl & RPQ_NAME1_LEN_SAME) o[ck.off_n1len] = m0.name1_len;
        if ((h.flags & RPQ_HAS_NAME2) && (fl & RPQ_NAME2_LEN_SAME)) o[ck.off_n2len] = (u8)(m0.name_len - m0.name2_off);
        if (fl & RPQ_STRAND_LEN_SAME) o[ck.off_slen] = m0.strand_len;
        if ((h.flags & RPQ_HAS_LANE) && (fl & RPQ_LANE_SAME)) o[ck.off_lane] = m0.lane;
        if ((h.flags & RPQ_HAS_TILE) && (fl & RPQ_TILE_SAME)) put_u16le(o + ck.off_tile, m0.tile);
        if (h.flags & RPQ_HAS_X) put_u32le(o + ck.off_x, ck.x_size);
        if (h.flags & RPQ_HAS_Y) put_u32le(o + ck.off_y, ck.y_size);
    }
    __syncwarp();
    const u32 fl = ck.flags;
    const u8* name0 = t0.text + line_start(t0, 4 * rec0);
    if (fl & RPQ_NAME1_SAME) for (u32 k = lane; k < m0.name1_len; k += 32) o[ck.off_n1 + k] = name0[k];
    if ((h.flags & RPQ_HAS_NAME2) && (fl & RPQ_NAME2_SAME)) { const u32 l = (u32)m0.name_len - m0.name2_off; for (u32 k = lane; k < l; k += 32) o[ck.off_n2 + k] = name0[m0.name2_off + k]; }
    if (fl & RPQ_STRAND_SAME) { const u8* s0 = t0.text + line_start(t0, 4 * rec0 + 2); for (u32 k = lane; k < m0.strand_len; k += 32) o[ck.off_strand + k] = s0[k]; }
    if (h.flags & RPQ_HAS_X) { const u8* src = tmp_x + 3ull * ck.first; for (u32 k = lane; k < ck.x_size; k += 32) o[ck.off_x + 4 + k] = src[k]; }
    if (h.flags & RPQ_HAS_Y) { const u8* src = tmp_y + 3ull * ck.first; for (u32 k = lane; k < ck.y_size; k += 32) o[ck.off_y + 4 + k] = src[k]; }
}

/*
 * One warp per read: its entries in the per-read columns, its name parts, its kept bases 2-bit packed (Q14: the bit
 * stream is continuous across the reads of the chunk; a byte is written by the read that owns its first base and the
 * owner pulls the up to 3 remaining bases from the following reads), its overlap byte.
 */
constexpr int EMIT_WARPS = 8;

__device__ inline u8 kept_base_at(const EncBatchDev& b, const HeaderDev& h, const ChunkDev& ck, u32 rel, u32 j) {
    /* j-th kept base of read rel (chunk-relative) */
    const u32 i = ck.first + rel;
    u32 f, rec; read_locus(b, i, f, rec);
    const TextDev& t = b.t[f];
    const u8* s = t.text + line_start(t, 4 * rec + 1);
    if (ck.interleaved && (rel & 1u)) {
        const u32 rl = b.rlen[i];
        const int o = (h.flags & RPQ_ENCODE_PE_BY_OVERLAP) ? (int)b.ov[i >> 1] : 0;
        const u32 jj = o > 0 ? j + (u32)o : j;
        return complement_base(s[rl - 1 - jj]);
    }
    return s[j];
}

__global__ void __launch_bounds__(32 * EMIT_WARPS) k_emit(EncBatchDev b, HeaderDev h, u8* out) {
    const int lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * EMIT_WARPS + (threadIdx.x >> 5);
    if (i >= b.n_reads) return;
    const u32 c = chunk_of_read(b, i);
    const ChunkDev& ck = b.chunks[c];
    const u32 rel = i - ck.first;
    u8* o = out + ck.out_offset;
    const u32 fl = ck.flags;
    const ReadMeta m = b.meta[i];
    u32 f, rec; read_locus(b, i, f, rec);
    const TextDev& t = b.t[f];
    const u32 rl = b.rlen[i];
    if (lane == 0) {
        if (!(fl & RPQ_READ_LEN_SAME)) { if (h.read_length_bytes == 1) o[ck.off_readlen + rel] = (u8)rl; else put_u16le(o + ck.off_readlen + 2 * rel, (u16)rl); }
        if (!(fl & RPQ_NAME1_LEN_SAME)) o[ck.off_n1len + rel] = m.name1_len;
        if ((h.flags & RPQ_HAS_NAME2) && !(fl & RPQ_NAME2_LEN_SAME)) o[ck.off_n2len + rel] = (u8)(m.name_len - m.name2_off);
        if (!(fl & RPQ_STRAND_LEN_SAME)) o[ck.off_slen + rel] = m.strand_len;
        const bool xy_owner = !ck.interleaved || !(rel & 1u);
        const u32 xy = ck.interleaved ? rel >> 1 : rel;
        if (xy_owner) {
            if ((h.flags & RPQ_HAS_LANE) && !(fl & RPQ_LANE_SAME)) o[ck.off_lane + xy] = m.lane;
            if ((h.flags & RPQ_HAS_TILE) && !(fl & RPQ_TILE_SAME)) put_u16le(o + ck.off_tile + 2 * xy, m.tile);
        }
        if (ck.ov_size && (rel & 1u)) o[ck.off_ov + (rel >> 1)] = (u8)(signed char)((int)b.ov[i >> 1] + (int)h.overlap_shift);
    }
    const u8* name = t.text + line_start(t, 4 * rec);
    if (!(fl & RPQ_NAME1_SAME)) { u8* d = o + ck.off_n1 + b.n1off[i]; for (u32 k = lane; k < m.name1_len; k += 32) d[k] = name[k]; }
    if ((h.flags & RPQ_HAS_NAME2) && !(fl & RPQ_NAME2_SAME)) { u8* d = o + ck.off_n2 + b.n2off[i]; const u32 l = (u32)m.name_len - m.name2_off; for (u32 k = lane; k < l; k += 32) d[k] = name[m.name2_off + k]; }
    if (!(fl & RPQ_STRAND_SAME)) { const u8* s = t.text + line_start(t, 4 * rec + 2); u8* d = o + ck.off_strand + b.soff[i]; for (u32 k = lane; k < m.strand_len; k += 32) d[k] = s[k]; }

    /* 2-bit pack */
    const u32 so = b.seqoff[i];
    const u32 kept = kept_bases(b, h, ck.interleaved != 0, i, rel);
    if (kept == 0) return;
    const u32 b0 = (so + 3u) >> 2, b1 = (so + kept - 1u) >> 2;       /* bytes whose first base (4B) lies in [so, so+kept) */
    const u8* own_seq = t.text + line_start(t, 4 * rec + 1);
    const bool own_rev = ck.interleaved && (rel & 1u);
    u32 own_shift = 0;
    if (own_rev && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) { const int ovv = b.ov[i >> 1]; own_shift = ovv > 0 ? (u32)ovv : 0u; }
    for (u32 B = b0 + lane; B <= b1; B += 32) {
        u32 v = 0;
#pragma unroll
        for (u32 k = 0; k < 4; k++) {
            const u32 pos = 4 * B + k;
            if (pos >= ck.seq_kept) break;
            u8 ch;
            if (pos < so + kept) { const u32 j = pos - so; ch = own_rev ? complement_base(own_seq[rl - 1 - (j + own_shift)]) : own_seq[j]; }
            else {
                /* spill into the following read(s) that still have kept bases */
                u32 r2 = rel + 1;
                while (r2 < ck.count && b.seqoff[ck.first + r2] + kept_bases(b, h, ck.interleaved != 0, ck.first + r2, r2) <= pos) r2++;
                ch = kept_base_at(b, h, ck, r2, pos - b.seqoff[ck.first + r2]);
            }
            v |= base_code(ch) << (2 * k);
        }
        o[ck.off_seq + B] = (u8)v;
    }
}

/* the first bases of a chunk belong to byte 0 only if read 0 owns position 0 - always true; but a chunk whose first reads
 * have no kept bases cannot happen (R1 is always stored whole).  Nothing else to do for byte ownership. */

/* raw quality copy for DONT_ENCODE_QUAL (src/rfqcodec.cpp:612-615): one warp per read */
__global__ void __launch_bounds__(32 * EMIT_WARPS) k_raw_qual(EncBatchDev b, u8* out) {
    const int lane = threadIdx.x & 31;
    const u32 i = blockIdx.x * EMIT_WARPS + (threadIdx.x >> 5);
    if (i >= b.n_reads) return;
    const u32 c = chunk_of_read(b, i);
    const ChunkDev& ck = b.chunks[c];
    const u32 rel = i - ck.first;
    u32 f, rec; read_locus(b, i, f, rec);
    const TextDev& t = b.t[f];
    const u8* q = t.text + line_start(t, 4 * rec + 3);
    const u32 rl = b.rlen[i];
    u8* d = out + ck.out_offset + ck.off_qual + b.qualoff[i];
    if (ck.interleaved && (rel & 1u)) for (u32 k = lane; k < rl; k += 32) d[k] = q[rl - 1 - k];
    else for (u32 k = lane; k < rl; k += 32) d[k] = q[k];
}

/* pieces of the span slots -> final position; one small CTA (64 threads) per span: a span's pieces are a few hundred bytes
 * behind a chain of four dependent loads, so the kernel lives on how many spans are in flight (32 CTAs per SM), not on threads
 * per span; also the stream length table */
__global__ void __launch_bounds__(256) k_gather(EncBatchDev b, StreamJob job, const u32* __restrict__ span_chunk, u8* out, int is_npos) {
    const u32 span = blockIdx.x;
    if (span >= *job.n_spans) return;
    const u32 c = span_chunk[span];
    const ChunkDev& ck = b.chunks[c];
    u8* col = out + ck.out_offset + (is_npos ? ck.off_npos : ck.off_qual);
    const u8* slot = job.slots + job.span_slot[span];
    const u32 ns = job.nstreams;
    /* the span's directory first (all entries in flight at once), then a warp per stream: forty streams of a dense column were
     * forty dependent directory loads, one after the other, with a few hundred bytes of copying between them */
    __shared__ SpanDir s_dir[MAX_BINS + 2];
    {
        const uint4* g = reinterpret_cast<const uint4*>(job.dir + (size_t)span * ns);
        uint4* sd = reinterpret_cast<uint4*>(s_dir);
        for (u32 k = threadIdx.x; k < 2u * ns; k += blockDim.x) sd[k] = g[k];
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (u32 st = warp; st < ns; st += nwarps) {
        const SpanDir d = s_dir[st];
        if (lane < d.first_len) col[d.dst + lane] = (u8)(d.first_tok >> (8 * lane));
        u8* dst = col + d.dst + d.first_len;
        const u8* src = slot + d.slot_off;
        for (u32 k = lane; k < d.bytes; k += 32u) dst[k] = src[k];
    }
    /* stream length table (quality only): written by the chunk's first span */
    if (!is_npos && span == job.span_first[c]) {
        const u32 s1 = job.span_first[c + 1];
        for (u32 st = threadIdx.x; st + 1 < ns; st += blockDim.x) {
            /* length = start of the next stream - start of this one; starts are the dst of the chunk's first span */
            const u32 a = job.dir[(size_t)span * ns + st].dst;
            const u32 z = job.dir[(size_t)span * ns + st + 1].dst;
            put_u32le(col + 4 * st, z - a);
        }
        (void)s1;
    }
}

}  // namespace rpq
