/*
 * rpq_meta3.cuh - k_meta3: per-read metadata, third generation.  Same staging and outputs as k_meta2 (rpq_meta2.cuh) but the
 * CTA re-maps its threads per phase so that twice as many threads share the same shared-memory footprint, and the overlap
 * search does three instructions per candidate:
 *
 *   stage   one TMA bulk copy (cp.async.bulk, mbarrier byte counting) per record: name + sequence + strand lines -> its slot
 *   A       thread per READ : FastqMeta::parse (reference src/fastqmeta.cpp:22-80), comparisons with the chunk's first read
 *                             (src/rfqcodec.cpp:225-234)
 *   A2      thread per PAIR : the PE consistency test (src/rfqcodec.cpp:233-270, Q10)
 *   B       thread per READ : 2-bit packing as stored (src/rfqcodec.cpp:593-604); an odd read of a file whose header supports
 *                             interleaving is packed as revcomp (src/read.cpp:77-115), which is what both the overlap search
 *                             and the stored stream need; "clean" = only A/C/G/T(/acgt for R2) characters
 *   C       thread per (PAIR, DIRECTION): RfqCodec::overlap (src/rfqcodec.cpp:1391-1438) on packed words, 16 candidate shifts
 *                             per pair of words; a packed match of two clean reads is exact, otherwise the bytes are compared
 *   D       thread per PAIR : forward beats backward, +-127 clamp (src/rfqcodec.cpp:378-383); packed reads to HBM, coalesced
 *
 * v2 ran all phases in one thread per pair: 18.75 % occupancy, 28 % issue utilisation (profiles/r01_v2_ncu_full_k_meta2.csv).
 */
#pragma once
#include "rpq_meta2.cuh"

namespace rpq {

/* ---- SIMD-in-a-register helpers for the name path ---- */
/* 0x80 in every byte of x that is zero (exact, no borrow between bytes) */
__device__ __forceinline__ u32 zero_bytes(u32 x) { return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu); }
/* the four 0x80 flags of a word as 4 adjacent bits (the partial products of the multiply never overlap) */
__device__ __forceinline__ u32 flags4(u32 t) { return (((t >> 7) * 0x00204081u) >> 21) & 0xFu; }
/* four 2-bit codes held in the low 2 bits of each byte -> one byte, by one multiply (products land 6 bits apart: no carries) */
__device__ __forceinline__ u32 squeeze4m(u32 c) { return (c * 0x01041040u) >> 24; }
/* value of four decimal digits held as 0..9 in bytes 0 (most significant) .. 3 */
__device__ __forceinline__ u32 swar_dec4(u32 t) { const u32 p = (t * 10u + (t >> 8)) & 0x00FF00FFu; return (p * 100u + (p >> 16)) & 0xFFFFu; }

/* atoi of name[a+1, b2): fields of 1..8 plain digits (every real coordinate) in two SIMD steps, anything else (empty, sign,
 * blanks, letters, 9+ digits with glibc's saturation) through atoi_like */
__device__ __forceinline__ int field_value(const u32* words, u32 off, int a, int b2) {
    const int n = b2 - a - 1;
    if (n >= 1 && n <= 8) {
        const u32 o = off + (u32)a + 1u;
        const int nh = n > 4 ? n - 4 : 0, nl = n - nh;                 /* leading 0..4 and trailing 1..4 characters */
        const u32 wl = ld4(words, o + (u32)nh) ^ 0x30303030u, ml = 0xFFFFFFFFu >> (8 * (4 - nl));
        u32 bad = ((((wl & 0x7f7f7f7fu) + 0x76767676u) | wl) & 0x80808080u) & ml;      /* a byte that is not 0..9 */
        u32 v = swar_dec4((wl & ml) << (8 * (4 - nl)));
        if (nh) {
            const u32 wh = ld4(words, o) ^ 0x30303030u, mh = 0xFFFFFFFFu >> (8 * (4 - nh));
            bad |= ((((wh & 0x7f7f7f7fu) + 0x76767676u) | wh) & 0x80808080u) & mh;
            v += swar_dec4((wh & mh) << (8 * (4 - nh))) * 10000u;
        }
        if (!bad) return (int)v;
    }
    return atoi_like(reinterpret_cast<const u8*>(words) + off + a + 1, n);
}

/* FastqMeta::parse (reference src/fastqmeta.cpp:22-80) on a name staged in shared memory at byte offset `off` of `words`:
 * same closed form as thread_tokenise (rpq_meta2.cuh), but ':' and ' ' are located 16 bytes at a time as bit masks.  Names
 * whose scan does not end within 64 bytes take thread_tokenise. */
__device__ inline ReadMeta thread_tokenise3(const u32* words, u32 off, int len) {
    unsigned long long cm = 0, sm = 0;
    bool known = false;
    for (int base = 0; base < len && base < 64 && !known; base += 16) {
        u32 c16 = 0, s16 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 w = ld4(words, off + (u32)(base + 4 * k));
            c16 |= flags4(zero_bytes(w ^ 0x3A3A3A3Au)) << (4 * k);
            s16 |= flags4(zero_bytes(w ^ 0x20202020u)) << (4 * k);
        }
        const int left = len - base;
        if (left < 16) { const u32 m = (1u << left) - 1u; c16 &= m; s16 &= m; }
        cm |= (unsigned long long)c16 << base; sm |= (unsigned long long)s16 << base;
        known = s16 != 0 || __popcll(cm) >= 7;
    }
    if (!known && len > 64) return thread_tokenise(reinterpret_cast<const u8*>(words) + off, len);
    int S = -1;
    if (sm) { S = __ffsll((long long)sm) - 1; cm &= (1ull << S) - 1ull; }       /* only colons before the first space count */
    const int ncol = __popcll(cm);
    unsigned long long t = cm;
    t &= t - 1; t &= t - 1;                                                       /* colons 1 and 2 */
    const int c3 = __ffsll((long long)t) - 1; t &= t - 1;
    const int c4 = __ffsll((long long)t) - 1; t &= t - 1;
    const int c5 = __ffsll((long long)t) - 1; t &= t - 1;
    const int c6 = __ffsll((long long)t) - 1; t &= t - 1;
    const int c7 = __ffsll((long long)t) - 1;
    ReadMeta m;
    m.x = 0; m.y = 0; m.tile = 0; m.lane = 0; m.has = 0; m.name_len = (u8)len; m.strand_len = 0; m.name1_len = (u8)len; m.name2_off = (u8)len;
    auto fld = [&](int a, int b2) { return field_value(words, off, a, b2); };
    int stop = -1;
    if (ncol >= 7) { stop = c7; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c4, c5); m.x = (u32)fld(c5, c6); m.y = (u32)fld(c6, c7); }
    else if (S >= 0 && ncol == 6) { stop = S; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c4, c5); m.x = (u32)fld(c5, c6); m.y = (u32)fld(c6, S); }
    else if (S >= 0 && ncol == 5) { stop = S; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c5, S); }
    else if (S >= 0 && ncol == 4) { stop = S; m.name1_len = (u8)c4; m.lane = (u8)fld(c4, S); }
    if (stop > 0) { m.has = 1; m.name2_off = (u8)stop; }
    else { m.name1_len = (u8)len; m.name2_off = (u8)len; m.lane = 0; m.tile = 0; m.x = 0; m.y = 0; }
    return m;
}

/* n bytes at byte offset `off` of the shared word array against n bytes of global text at g, four at a time (aligned global
 * words, funnel-shifted); reads up to 7 bytes past g + n (the text buffers are padded) */
__device__ __forceinline__ bool equal_shared_global(const u32* words, u32 off, const u8* g, int n) {
    const u32* gw = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)3);
    const u32 gsh = (u32)(reinterpret_cast<uintptr_t>(g) & 3u) * 8u;
    u32 diff = 0;
    int k = 0;
    u32 lo = n > 0 ? gw[0] : 0u;
    for (; k + 4 <= n; k += 4) { const u32 hi = gw[(k >> 2) + 1]; diff |= ld4(words, off + (u32)k) ^ __funnelshift_r(lo, hi, gsh); lo = hi; }
    if (k < n) diff |= (ld4(words, off + (u32)k) ^ __funnelshift_r(lo, gw[(k >> 2) + 1], gsh)) & ((1u << (8 * (n - k))) - 1u);
    return diff == 0;
}

/* The read as it will be stored, 16 bases per word.  One code path for both strands (mates sit in neighbouring lanes: two
 * functions would run one after the other).  rev: the read is packed as its reverse complement; lower case counts as a
 * plain base there (src/read.cpp:92-113).  Whole words of 16 bases take four consecutive shared words each (one new load per
 * four bases, carried over), the codes are checked against the characters they stand for once per word, and only a word
 * that holds something else (N, lower case in a forward read, ...) is recoded exactly. */
__device__ inline bool pack_read_c(const u32* words, u32 off, int len, bool rev, u32* dst, int pkw) {
    bool clean = true;
    const u32 sel = rev ? 0x0123u : 0x3210u;
    const u32 lower = rev ? 0x20202020u : 0u, flip = rev ? 0x03030303u : 0u;
    const u32 table = rev ? 0x63746167u : 0x43544147u;          /* code G0 A1 T2 C3 -> the character it came from */
    const int full = len >> 4;
    /* group k (bases 4k..4k+3 of the result) = source bytes [off + 4k, +4) forward, [off + len - 4 - 4k, +4) reversed */
    const u32 edge = rev ? off + (u32)len : off;
    const u32 sh = (edge & 3u) * 8u;
    const int step = rev ? -1 : 1;
    int ni = (int)(edge >> 2) + step;                          /* next shared word to fetch: upwards forward, downwards reversed */
    u32 carry = words[edge >> 2];
    for (int j = 0; j < full; j++) {
        u32 wq[4], bad = 0, acc = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const u32 x = words[ni];
            ni += step;
            const u32 raw = __funnelshift_r(rev ? x : carry, rev ? carry : x, sh);      /* selects, not branches */
            carry = x;
            const u32 w = __byte_perm(raw, 0, sel);
            const u32 l = w | lower;
            const u32 c = ((l ^ (l >> 1)) & 0x02020202u) | ((~l >> 2) & 0x01010101u);
            const u32 t = c | (c >> 4);
            bad |= __byte_perm(table, 0u, __byte_perm(t, 0u, 0x4420)) ^ l;
            acc |= squeeze4m(c ^ flip) << (8 * g);
            wq[g] = w;
        }
        if (bad) {                                              /* exact: every byte that is not a plain base is code 0 */
            clean = false;
            acc = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) acc |= squeeze4m(rev ? codes_rc(wq[g]) : codes_fwd(wq[g])) << (8 * g);
        }
        dst[j] = acc;
    }
    for (int j = full; j < pkw; j++) {                          /* the last, partial word and the zero padding */
        u32 acc = 0;
        const int base = j * 16;
        if (base < len) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int p = base + 4 * g;
                if (p >= len) break;
                const int left = len - p;
                u32 w;
                if (left >= 4 || !rev) w = __byte_perm(ld4(words, off + (u32)(rev ? len - 4 - p : p)), 0, sel);
                else { w = 0; for (int q = 0; q < left; q++) w |= (u32)reinterpret_cast<const u8*>(words)[off + (u32)(len - 1 - p - q)] << (8 * q); }
                const u32 vm = left >= 4 ? 0xFFFFFFFFu : ((1u << (8 * left)) - 1u);
                const u32 l = w | lower;
                u32 c = ((l ^ (l >> 1)) & 0x02020202u) | ((~l >> 2) & 0x01010101u);
                const u32 t = c | (c >> 4);
                const u32 expect = __byte_perm(table, 0u, __byte_perm(t, 0u, 0x4420));
                if ((expect ^ l) & vm) { clean = false; c = rev ? codes_rc(w) : codes_fwd(w); }
                else c ^= flip;
                acc |= squeeze4m(c & vm) << (8 * g);
            }
        }
        dst[j] = acc;
    }
    return clean;
}

/* do the packed sequences agree on o bases: a from base s, p from base 0 */
__device__ __forceinline__ bool packed_equal(const u32* a, int s, const u32* p, int o, int pkw) {
    for (int k = 0; k < o; k += 16) {
        const int s2 = s + k;
        const u32 w2 = (u32)s2 >> 4, sh2 = ((u32)s2 & 15u) * 2u;
        const u32 win = __funnelshift_r(a[w2], (int)(w2 + 1) < pkw ? a[w2 + 1] : 0u, sh2);
        const int rem = o - k;
        const u32 m = rem >= 16 ? 0xFFFFFFFFu : ((1u << (2 * rem)) - 1u);
        if (((win ^ p[k >> 4]) & m) != 0) return false;
    }
    return true;
}

/* smallest o in [12, min(la, lp)] with a[la-o+i] == p[i] (i < o); `exact(o)` confirms a packed match when a read is not clean */
template <class V>
__device__ inline int overlap_search(const u32* a, int la, const u32* p, int lp, int pkw, bool packed_is_exact, const V& exact) {
    const int minlen = la < lp ? la : lp;
    if (minlen < 12) return 0;
    const u32 pat = p[0];
    auto confirm = [&](int o) -> bool { return packed_equal(a, la - o, p, o, pkw) && (packed_is_exact || exact(o)); };
    /* o = 12..15: masked window */
    for (int o = 12; o <= 15 && o <= minlen; o++) {
        const int s = la - o;
        const u32 wi = (u32)s >> 4, sh = ((u32)s & 15u) * 2u;
        const u32 win = __funnelshift_r(a[wi], (int)(wi + 1) < pkw ? a[wi + 1] : 0u, sh);
        if ((((win ^ pat) & ((1u << (2 * o)) - 1u)) == 0) && confirm(o)) return o;
    }
    if (minlen < 16) return 0;
    /* o >= 16: window start s from la-16 down to la-minlen, a full 16-base compare per shift */
    const int s_hi = la - 16, s_lo = la - minlen;
    const u32 pat8 = (pat & 0xFFFFu) * 0x00010001u;              /* the first 8 bases of the pattern in both halves */
    for (int wi = s_hi >> 4; wi >= (s_lo >> 4); wi--) {
        const u32 lo = a[wi], hi = wi + 1 < pkw ? a[wi + 1] : 0u;
        /* filter: the two halves of the window at shift sh are the first 8 bases of the candidates sh and sh + 8; a zero
         * half of (window ^ pat8) exists iff (x - 0x00010001) & ~x & 0x80008000 is not 0 (exact as an any-test) */
        u32 any = 0;
#pragma unroll
        for (int sh = 0; sh < 8; sh++) { const u32 x = __funnelshift_r(lo, hi, 2 * sh) ^ pat8; any |= (x - 0x00010001u) & ~x; }
        if (!(any & 0x80008000u)) continue;
        u32 hits = 0;
#pragma unroll
        for (int sh = 0; sh < 16; sh++) hits |= (u32)(__funnelshift_r(lo, hi, 2 * sh) == pat) << sh;
        /* restrict to [s_lo, s_hi] */
        const int base = wi << 4;
        if (base + 15 > s_hi) hits &= (2u << (s_hi - base)) - 1u;
        if (base < s_lo) hits &= ~((1u << (s_lo - base)) - 1u);
        while (hits) {
            const int sh = 31 - __clz((int)hits);              /* largest start first = smallest o first */
            hits &= ~(1u << sh);
            const int o = la - (base + sh);
            if (confirm(o)) return o;
        }
    }
    return 0;
}

__global__ void __launch_bounds__(256) k_meta3(EncBatchDev b, HeaderDev h, u32 n_units, Meta2Cfg cfg) {
    RPQ_DYN_SMEM(dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const u32 P = cfg.units_per_cta, per = b.is_pe ? 2u : 1u;
    const u32 u0 = blockIdx.x * P;
    const u32 n_here = n_units - u0 < P ? n_units - u0 : P;
    const u32 n_reads_here = n_here * per;
    u32* slots = reinterpret_cast<u32*>(dyn);                         /* [P*per][slot_words] */
    u32* pkS = slots + (size_t)P * per * cfg.slot_words;              /* [P*per][pkw] */
    ReadMeta* s_meta = reinterpret_cast<ReadMeta*>(pkS + (size_t)P * per * cfg.pkw);   /* [P*per] */
    u32* s_seq = reinterpret_cast<u32*>(s_meta + (size_t)P * per);    /* [P*per] byte offset of the sequence inside the slot | rlen << 16 */
    u8* s_flag = reinterpret_cast<u8*>(s_seq + (size_t)P * per);      /* [P*per] bit0 name2 == read0.name2, bit1 clean */
    short* s_ov = reinterpret_cast<short*>(s_flag + (((size_t)P * per + 3) & ~(size_t)3));   /* [P*2] per direction */
    const bool rc_odd = b.is_pe && h.support_interleaved;

    /* ---- stage: one TMA bulk copy (cp.async.bulk global -> shared, 16-byte granules) per record head, all in flight at once,
     * completion counted in bytes on one mbarrier.  The source's 16-byte phase is kept inside the slot. */
#ifdef RPQ_EMU
    for (u32 r = warp; r < n_reads_here; r += nwarps) {
        const u32 i = u0 * per + r;
        const uint4 lc = b.loc[i];
        u32 f, rec; read_locus(b, i, f, rec);
        const u8* src = b.t[f].text + (lc.x & ~15u);
        const u32 nb = (((lc.x & 15u) + (lc.w - lc.x) + 15u) >> 4) << 4;
        u8* dst = reinterpret_cast<u8*>(slots + (size_t)r * cfg.slot_words);
        for (u32 k = lane; k < nb && k < cfg.slot_words * 4u; k += 32) dst[k] = src[k];
    }
    (void)nwarps;
    __syncthreads();
#else
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mbar), "r"(n_reads_here) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((u32)tid < n_reads_here) {
        const u32 r = tid, i = u0 * per + r;
        const uint4 lc = b.loc[i];
        u32 f, rec; read_locus(b, i, f, rec);
        const u8* src = b.t[f].text + (lc.x & ~15u);
        u32 nb = (((lc.x & 15u) + (lc.w - lc.x) + 15u) >> 4) << 4;
        if (nb > cfg.slot_words * 4u) nb = cfg.slot_words * 4u;
        const u32 dst = (u32)__cvta_generic_to_shared(slots + (size_t)r * cfg.slot_words);
        asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(nb), "r"(mbar) : "memory");
    }
    {
        u32 done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar) : "memory");
        }
    }
    (void)lane; (void)warp; (void)nwarps;
#endif

    /* ---- A: thread per read */
    if ((u32)tid < n_reads_here) {
        const u32 r = tid, i = u0 * per + r;
        const u32 c = chunk_of_read(b, i);
        const u32 first = b.chunk_first[c];
        const ReadMeta m0 = b.meta0[c];
        const uint4 lc0 = b.loc[first];
        u32 f0, rec0; read_locus(b, first, f0, rec0);
        const u8* name0 = b.t[f0].text + lc0.x;
        const u8* strand0 = b.t[f0].text + lc0.z;
        const u32 rlen0 = lc0.z - lc0.y - 1u - b.t[f0].crlf;
        const int n2len0 = (int)m0.name_len - (int)m0.name2_off;
        const uint4 lc = b.loc[i];
        u32 f, rec; read_locus(b, i, f, rec);
        const u32 crlf = b.t[f].crlf;
        const u32* words = slots + (size_t)r * cfg.slot_words;
        const int nlen = (int)(lc.y - lc.x - 1u - crlf);
        const int rlen = (int)(lc.z - lc.y - 1u - crlf);
        const int slen = (int)(lc.w - lc.z - 1u - crlf);
        const u32 noff = lc.x & 15u;                                  /* the name's byte offset inside the slot */
        ReadMeta m = thread_tokenise3(words, noff, nlen < 256 ? nlen : 255);
        m.strand_len = (u8)slen;
        b.meta[i] = m;
        s_meta[r] = m;
        s_seq[r] = ((lc.x & 15u) + (lc.y - lc.x)) | ((u32)rlen << 16);
        u32 clear = 0;
        if ((u32)rlen != rlen0) clear |= AB_READ_LEN;
        if (m.name1_len != m0.name1_len) clear |= AB_N1LEN;
        const int n2len = (int)m.name_len - (int)m.name2_off;
        if (n2len != n2len0) clear |= AB_N2LEN;
        if (m.strand_len != m0.strand_len) clear |= AB_SLEN;
        if (m.lane != m0.lane) clear |= AB_LANE;
        if (m.tile != m0.tile) clear |= AB_TILE;
        if (m.name1_len != m0.name1_len || !equal_shared_global(words, noff, name0, m.name1_len)) clear |= AB_N1;
        if (m.strand_len != m0.strand_len || !equal_shared_global(words, noff + (lc.z - lc.x), strand0, slen)) clear |= AB_STRAND;
        const bool eq0 = (n2len == n2len0) && equal_shared_global(words, noff + m.name2_off, name0 + m0.name2_off, n2len);
        ChunkDev& ck = b.chunks[c];
        if (clear && (*(volatile u32*)&ck.and_bits & clear)) atomicAnd(&ck.and_bits, ~clear);
        const u32 rel = i - first;
        if (!eq0) { if (rel & 1u) atomicMax(&ck.last_odd_neq, rel + 1); else if (!*(volatile u32*)&ck.even_neq) atomicOr(&ck.even_neq, 1u); }
        /* ---- B: the read as it will be stored */
        const u32 so = s_seq[r] & 0xFFFFu;
        const bool clean = pack_read_c(words, so, rlen, rc_odd && (r & 1u), pkS + (size_t)r * cfg.pkw, (int)cfg.pkw);
        s_flag[r] = (u8)((eq0 ? 1u : 0u) | (clean ? 2u : 0u));
        if (b.unclean) b.unclean[i] = clean ? 0 : 1;                  /* the N-position coder only stages these reads */
    }
    __syncthreads();
    if (rc_odd) {
        /* ---- A2: thread per pair (Q10) */
        if ((u32)tid < n_here) {
            const u32 u = u0 + tid, i0 = u * 2;
            const u32 c = chunk_of_read(b, i0);
            ChunkDev& ck = b.chunks[c];
            const u32 rel = i0 - b.chunk_first[c];
            const ReadMeta ma = s_meta[2 * tid], mb = s_meta[2 * tid + 1];
            const u32* w1 = slots + (size_t)(2 * tid) * cfg.slot_words;
            const u32* w2 = slots + (size_t)(2 * tid + 1) * cfg.slot_words;
            const u32 o1 = (b.loc[i0].x & 15u) + ma.name2_off, o2 = (b.loc[i0 + 1].x & 15u) + mb.name2_off;
            const int l1 = (int)ma.name_len - (int)ma.name2_off, l2 = (int)mb.name_len - (int)mb.name2_off;
            bool okA = l1 == l2;
            if (okA) {                                             /* four bytes at a time; R1's byte at name2_diff_pos reads as name2_diff_char */
                const int dp = h.name2_diff_char != 0 ? (int)h.name2_diff_pos : -1;
                u32 diff = 0;
                for (int q = 0; q < l1; q += 4) {
                    u32 a = ld4(w1, o1 + (u32)q);
                    if ((u32)(dp - q) < 4u) { const u32 sh = 8u * (u32)(dp - q); a = (a & ~(0xFFu << sh)) | ((u32)h.name2_diff_char << sh); }
                    u32 x = a ^ ld4(w2, o2 + (u32)q);
                    if (l1 - q < 4) x &= (1u << (8 * (l1 - q))) - 1u;
                    diff |= x;
                }
                okA = diff == 0;
            }
            const bool okB = ma.lane == mb.lane && ma.tile == mb.tile && ma.x == mb.x && ma.y == mb.y;
            if (!okA) atomicMin(&ck.fA, rel + 1);
            if (!okB) atomicMin(&ck.fB, rel + 1);
        }
        /* ---- C: thread per (pair, direction) */
        if ((h.flags & RPQ_ENCODE_PE_BY_OVERLAP) && (u32)tid < 2 * n_here) {
            const u32 pair = tid >> 1, dir = tid & 1u;
            const u32* A = pkS + (size_t)(2 * pair) * cfg.pkw;          /* r1 as stored */
            const u32* R = pkS + (size_t)(2 * pair + 1) * cfg.pkw;      /* revcomp(r2) */
            const u32 sa = s_seq[2 * pair], sb = s_seq[2 * pair + 1];
            const int len1 = (int)(sa >> 16), len2 = (int)(sb >> 16);
            const u8* s1 = reinterpret_cast<const u8*>(slots + (size_t)(2 * pair) * cfg.slot_words) + (sa & 0xFFFFu);
            const u8* s2 = reinterpret_cast<const u8*>(slots + (size_t)(2 * pair + 1) * cfg.slot_words) + (sb & 0xFFFFu);
            const bool exact_packed = (s_flag[2 * pair] & 2u) && (s_flag[2 * pair + 1] & 2u);
            int o;
            if (dir == 0) {
                auto vf = [&](int oo) { for (int q = 0; q < oo; q++) if (s1[len1 - oo + q] != complement_base(s2[len2 - 1 - q])) return false; return true; };
                o = overlap_search(A, len1, R, len2, (int)cfg.pkw, exact_packed, vf);
            } else {
                auto vb = [&](int oo) { for (int q = 0; q < oo; q++) if (complement_base(s2[oo - 1 - q]) != s1[q]) return false; return true; };
                o = overlap_search(R, len2, A, len1, (int)cfg.pkw, exact_packed, vb);
            }
            s_ov[tid] = (short)o;
        }
        __syncthreads();
        if ((u32)tid < n_here) {
            int o = 0;
            if (h.flags & RPQ_ENCODE_PE_BY_OVERLAP) {
                o = s_ov[2 * tid] ? (int)s_ov[2 * tid] : -(int)s_ov[2 * tid + 1];
                if (o + (int)h.overlap_shift > 127) o = 0;
                if (o + (int)h.overlap_shift < -127) o = 0;
            }
            b.ov[u0 + tid] = (short)o;
        }
    }
    /* ---- D: packed reads to global memory, coalesced */
    {
        const u32 nw = n_reads_here * cfg.pkw;
        u32* g = b.pk + (size_t)u0 * per * cfg.pkw;
        for (u32 k = tid; k < nw; k += blockDim.x) g[k] = pkS[k];
    }
}

}  // namespace rpq
