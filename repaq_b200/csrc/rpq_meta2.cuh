/*
 * rpq_meta2.cuh - k_meta2: per-read metadata, second generation: ONE THREAD per read / pair on text staged in
 * shared memory (v1 k_meta used a warp per pair with byte loops and needed 2400 warp instructions per pair,
 * profiles/r01_v1_ncu_full_k_meta.csv).
 *
 * A CTA takes P consecutive units (reads, or pairs).  Stage: the "head" of every record (name, sequence and strand
 * lines, i.e. everything before the quality line) is copied word by word into a fixed-size slot, keeping the source's
 * byte phase so that aligned global words land on aligned shared words.  Work, per thread:
 *   FastqMeta::parse of both names                         (reference src/fastqmeta.cpp:22-80)
 *   the comparisons with the chunk's first read            (src/rfqcodec.cpp:225-234) and the PE test (:233-270, Q10)
 *   2-bit packing of the reads as they will be stored      (src/rfqcodec.cpp:593-604; G=0 A=1 T=2 C=3, anything else 0)
 *     and of revcomp(R2)                                   (src/read.cpp:77-115)
 *   RfqCodec::overlap                                      (src/rfqcodec.cpp:1391-1438): a 16-base packed window of one
 *     read is slid over the other (funnel shift per candidate); hits are verified on the bytes.
 * The packed reads go to global memory for k_emit2, so the sequence lines are read from HBM exactly once.
 */
#pragma once
#include "rpq_encode.cuh"

namespace rpq {

struct Meta2Cfg {
    u32 units_per_cta;   /* = blockDim.x */
    u32 slot_words;      /* shared words per record head (odd: conflict-free stride) */
    u32 pkw;             /* packed words per read (odd) */
};

/* ---- thread-level FastqMeta::parse on bytes in shared memory: same closed form as warp_tokenise */
__device__ inline ReadMeta thread_tokenise(const u8* name, int len) {
    int ncol = 0, c3 = -1, c4 = -1, c5 = -1, c6 = -1, c7 = -1, S = -1;
    for (int i = 0; i < len; i++) {
        const u8 c = name[i];
        if (c == ':') {
            ncol++;
            if (ncol == 3) c3 = i; else if (ncol == 4) c4 = i; else if (ncol == 5) c5 = i; else if (ncol == 6) c6 = i;
            else if (ncol == 7) { c7 = i; break; }
        } else if (c == ' ') { S = i; break; }
    }
    ReadMeta m;
    m.x = 0; m.y = 0; m.tile = 0; m.lane = 0; m.has = 0; m.name_len = (u8)len; m.strand_len = 0; m.name1_len = (u8)len; m.name2_off = (u8)len;
    auto fld = [&](int a, int b2) { return atoi_like(name + a + 1, b2 - a - 1); };
    int stop = -1;
    if (ncol == 7) { stop = c7; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c4, c5); m.x = (u32)fld(c5, c6); m.y = (u32)fld(c6, c7); }
    else if (S >= 0 && ncol == 6) { stop = S; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c4, c5); m.x = (u32)fld(c5, c6); m.y = (u32)fld(c6, S); }
    else if (S >= 0 && ncol == 5) { stop = S; m.name1_len = (u8)c3; m.lane = (u8)fld(c3, c4); m.tile = (u16)fld(c5, S); }
    else if (S >= 0 && ncol == 4) { stop = S; m.name1_len = (u8)c4; m.lane = (u8)fld(c4, S); }
    if (stop > 0) { m.has = 1; m.name2_off = (u8)stop; }
    else { m.name1_len = (u8)len; m.name2_off = (u8)len; m.lane = 0; m.tile = 0; m.x = 0; m.y = 0; }
    return m;
}

__device__ __forceinline__ bool bytes_equal(const u8* a, const u8* g, int n) {
    for (int k = 0; k < n; k++) if (a[k] != g[k]) return false;
    return true;
}

/* 4 bytes at an arbitrary byte offset of a word-aligned shared array */
__device__ __forceinline__ u32 ld4(const u32* words, u32 byteoff) {
    const u32 w = byteoff >> 2, sh = (byteoff & 3u) * 8u;
    return __funnelshift_r(words[w], words[w + 1], sh);
}
/* four 2-bit codes held in the low 2 bits of each byte -> one byte */
__device__ __forceinline__ u32 squeeze4(u32 c) { const u32 c2 = c | (c >> 6); return (c2 & 0xFu) | ((c2 >> 12) & 0xF0u); }
/* src/rfqcodec.cpp:593-599, exact: A=1 T=2 C=3, everything else (G, N, lower case, ...) 0 */
__device__ __forceinline__ u32 codes_fwd(u32 w) {
    return (__vcmpeq4(w, 0x41414141u) & 0x01010101u) | (__vcmpeq4(w, 0x54545454u) & 0x02020202u) | (__vcmpeq4(w, 0x43434343u) & 0x03030303u);
}
/* code of complement_base(c) (src/read.cpp:92-113): A/a->T=2  T/t->A=1  G/g->C=3  C/c->G=0  else N=0 */
__device__ __forceinline__ u32 codes_rc(u32 w) {
    const u32 l = w | 0x20202020u;
    return (__vcmpeq4(l, 0x61616161u) & 0x02020202u) | (__vcmpeq4(l, 0x74747474u) & 0x01010101u) | (__vcmpeq4(l, 0x67676767u) & 0x03030303u);
}

/* pack `len` bases starting at byte offset `off` of `words` into dst[0 .. (len+15)/16); unused high bits are 0 */
__device__ inline void pack_forward(const u32* words, u32 off, int len, u32* dst, int pkw) {
    for (int j = 0; j < pkw; j++) {
        u32 acc = 0;
        const int base = j * 16;
        if (base < len) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int p = base + 4 * g;
                if (p >= len) break;
                u32 c = codes_fwd(ld4(words, off + (u32)p));
                const int left = len - p;
                if (left < 4) c &= (1u << (8 * left)) - 1u;
                acc |= squeeze4(c) << (8 * g);
            }
        }
        dst[j] = acc;
    }
}
/* pack revcomp: base k of the result is complement(seq[len-1-k]) */
__device__ inline void pack_revcomp(const u32* words, u32 off, int len, u32* dst, int pkw) {
    for (int j = 0; j < pkw; j++) {
        u32 acc = 0;
        const int base = j * 16;
        if (base < len) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int k = base + 4 * g;                 /* result positions k..k+3 <- source positions len-1-k .. len-4-k */
                if (k >= len) break;
                const int left = len - k;
                u32 w;
                if (left >= 4) w = __byte_perm(ld4(words, off + (u32)(len - 4 - k)), 0, 0x0123);
                else { w = 0; for (int q = 0; q < left; q++) w |= (u32)reinterpret_cast<const u8*>(words)[off + (u32)(len - 1 - k - q)] << (8 * q); }
                u32 c = codes_rc(w);
                if (left < 4) c &= (1u << (8 * left)) - 1u;
                acc |= squeeze4(c) << (8 * g);
            }
        }
        dst[j] = acc;
    }
}

/*
 * smallest o in [12, min(la, lp)] with a[la-o+i] == p[i] for i < o, where a / p are packed (16 bases per word, base j
 * at bits 2(j&15)) and the bytes are checked through `verify(o)` on a packed match.  0 if none.
 */
template <class V>
__device__ inline int overlap_dir_packed(const u32* a, int la, const u32* p, int lp, int pkw, const V& verify) {
    const int minlen = la < lp ? la : lp;
    if (minlen < 12) return 0;
    const u32 pat = p[0];
    for (int o = 12; o <= minlen; o++) {
        const int s = la - o;                                   /* window start in a */
        const u32 wi = (u32)s >> 4, sh = ((u32)s & 15u) * 2u;
        const u32 lo = a[wi], hi = (int)(wi + 1) < pkw ? a[wi + 1] : 0u;
        const u32 win = __funnelshift_r(lo, hi, sh);
        const u32 mask = o >= 16 ? 0xFFFFFFFFu : ((1u << (2 * o)) - 1u);
        if (((win ^ pat) & mask) != 0) continue;
        /* the first min(o,16) packed bases agree: compare the rest packed, then the bytes */
        bool ok = true;
        for (int k = 16; k < o && ok; k += 16) {
            const int s2 = s + k;
            const u32 w2 = (u32)s2 >> 4, sh2 = ((u32)s2 & 15u) * 2u;
            const u32 win2 = __funnelshift_r(a[w2], (int)(w2 + 1) < pkw ? a[w2 + 1] : 0u, sh2);
            const int rem = o - k;
            const u32 m2 = rem >= 16 ? 0xFFFFFFFFu : ((1u << (2 * rem)) - 1u);
            if (((win2 ^ p[k >> 4]) & m2) != 0) ok = false;
        }
        if (ok && verify(o)) return o;
    }
    return 0;
}

}  // namespace rpq

namespace rpq {

/*
 * k_emit2: one THREAD per read.  Per-read column entries, and the read's share of the continuous 2-bit stream (Q14,
 * reference src/rfqcodec.cpp:590-604) assembled from the packed words k_meta2 left in global memory: an output word
 * (16 bases) is owned by the read that holds its first base; the owner pulls missing bases from the following reads.
 */
__device__ __forceinline__ u32 packed_window(const u32* src, u32 j0, u32 pkw) {
    const u32 wi = j0 >> 4, sh = (j0 & 15u) * 2u;
    const u32 lo = wi < pkw ? src[wi] : 0u, hi = wi + 1 < pkw ? src[wi + 1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}

__global__ void __launch_bounds__(256) k_emit2(EncBatchDev b, HeaderDev h, u8* out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_reads) return;
#ifndef RPQ_EMU
    {   /* the read's packed words (two or three sectors) are on their way while the chunk and per-read tables are fetched */
        const char* row = reinterpret_cast<const char*>(b.pk + (size_t)i * b.pkw);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 4 * b.pkw - 4));
    }
#endif
    const u32 c = chunk_of_read(b, i);
    const ChunkDev& ck = b.chunks[c];
    const u32 rel = i - ck.first;
    u8* o = out + ck.out_offset;
    const u32 fl = ck.flags;
    const ReadMeta m = b.meta[i];
    const u32 rl = b.rlen[i];
    if (!(fl & RPQ_READ_LEN_SAME)) { if (h.read_length_bytes == 1) o[ck.off_readlen + rel] = (u8)rl; else put_u16le(o + ck.off_readlen + 2 * rel, (u16)rl); }
    if (!(fl & RPQ_NAME1_LEN_SAME)) o[ck.off_n1len + rel] = m.name1_len;
    if ((h.flags & RPQ_HAS_NAME2) && !(fl & RPQ_NAME2_LEN_SAME)) o[ck.off_n2len + rel] = (u8)(m.name_len - m.name2_off);
    if (!(fl & RPQ_STRAND_LEN_SAME)) o[ck.off_slen + rel] = m.strand_len;
    const bool il = ck.interleaved != 0;
    if (!il || !(rel & 1u)) {
        const u32 xy = il ? rel >> 1 : rel;
        if ((h.flags & RPQ_HAS_LANE) && !(fl & RPQ_LANE_SAME)) o[ck.off_lane + xy] = m.lane;
        if ((h.flags & RPQ_HAS_TILE) && !(fl & RPQ_TILE_SAME)) put_u16le(o + ck.off_tile + 2 * xy, m.tile);
    }
    if (ck.ov_size && (rel & 1u)) o[ck.off_ov + (rel >> 1)] = (u8)(signed char)((int)b.ov[i >> 1] + (int)h.overlap_shift);

    const u32 so = b.seqoff[i];
    const u32 kept = kept_bases(b, h, il, i, rel);
    if (kept == 0) return;
    const u32 pkw = b.pkw;
    /* 16 kept bases of read ii (chunk-relative rr) from kept-base index j0.  In a file whose header supports interleaving the
     * odd reads are packed as revcomp (k_meta3); a chunk that lost the interleaved form (Q10) needs them forward: from the text. */
    const bool rc_file = b.is_pe && h.support_interleaved;
    auto window = [&](u32 ii, u32 rr, u32 j0) -> u32 {
        if (il && (rr & 1u)) {
            const int ov = (h.flags & RPQ_ENCODE_PE_BY_OVERLAP) ? (int)b.ov[ii >> 1] : 0;
            return packed_window(b.pk + (size_t)ii * pkw, j0 + (ov > 0 ? (u32)ov : 0u), pkw);
        }
        if (rc_file && (rr & 1u)) {
            u32 f, rec; read_locus(b, ii, f, rec);
            const u8* sq = b.t[f].text + b.loc[ii].y;
            const u32 rl2 = b.rlen[ii];
            u32 w = 0;
            for (u32 k = 0; k < 16u && j0 + k < rl2; k++) w |= base_code(sq[j0 + k]) << (2 * k);
            return w;
        }
        return packed_window(b.pk + (size_t)ii * pkw, j0, pkw);
    };
    u8* col = o + ck.off_seq;
    /* The column starts wherever the chunk layout puts it; output words are therefore counted from the 4-byte boundary at or
     * below it: word V holds the column bytes [4V - al, 4V - al + 4), i.e. the stream's bases [16V - 4 al, 16V - 4 al + 16), and
     * leaves with one aligned 32-bit store.  Only word 0 of a misaligned column and the column's last word are partial. */
    const u32 al = (u32)(reinterpret_cast<uintptr_t>(col) & 3u);
    const u32 sh_so = so + 4u * al;                                                /* the read's first base, counted from word 0 */
    u8* colw = col - al;
    const u32 col_end = ck.seq_size + al;                                          /* end of the column, counted from colw */
    const u32 w0 = (sh_so + 15u) >> 4, w1 = (sh_so + kept - 1u) >> 4;
    auto store_word = [&](u32 V, u32 word) {
        const u32 bo = 4u * V;
        if ((V > 0u || al == 0u) && bo + 4u <= col_end) { *reinterpret_cast<u32*>(colw + bo) = word; return; }
        const u32 from = V == 0u ? al : 0u;
        const u32 nbytes = col_end - bo < 4u ? col_end - bo : 4u;
        for (u32 k = from; k < nbytes; k++) colw[bo + k] = (u8)(word >> (8 * k));
    };
    /* word V built from this read's kept bases and, where they end early, the reads that follow */
    auto boundary_word = [&](u32 V) {
        const u32 lead = V == 0u ? 4u * al : 0u;                                    /* base slots of word 0 below the column */
        const u32 p0 = 16u * V + lead - 4u * al;                                    /* stream position of the first real base */
        const u32 room = 16u - lead;
        u32 have = so + kept - p0; if (have > room) have = room;                    /* own bases in this word */
        u32 word = window(i, rel, p0 - so);
        if (have < 16u) {
            word &= (1u << (2 * have)) - 1u;
            /* pull from the following reads that still have kept bases */
            u32 r2 = rel + 1, filled = have;
            while (filled < room && r2 < ck.count && p0 + filled < ck.seq_kept) {
                const u32 i2 = ck.first + r2;
                const u32 k2 = kept_bases(b, h, il, i2, r2);
                if (k2) {
                    u32 take = room - filled; if (take > k2) take = k2;
                    u32 w2 = window(i2, r2, 0u);
                    if (take < 16u) w2 &= (1u << (2 * take)) - 1u;
                    word |= w2 << (2 * filled);
                    filled += take;
                }
                r2++;
            }
        }
        store_word(V, lead ? word << (2 * lead) : word);
    };
    if (so == 0u && al != 0u) boundary_word(0u);                                   /* word 0 of a misaligned column: nobody's first base is at its start */
    u32 W = w0;
    /* words that lie entirely inside this read: a funnel shift of two neighbouring packed words, the loads a few words ahead */
    if (w1 > w0 && !(rc_file && (rel & 1u) && !il)) {
        const int ov = (il && (rel & 1u) && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP)) ? (int)b.ov[i >> 1] : 0;
        const u32 a = 16u * w0 - sh_so + (ov > 0 ? (u32)ov : 0u);
        const u32* src = b.pk + (size_t)i * pkw;
        const u32 sh = (a & 15u) * 2u;
        u32 k = a >> 4;
        u32 lo = k < pkw ? src[k] : 0u;
#pragma unroll 4
        for (; W < w1; W++) {
            k++;
            const u32 hi = k < pkw ? src[k] : 0u;
            *reinterpret_cast<u32*>(colw + 4u * W) = __funnelshift_r(lo, hi, sh);  /* W >= 1 and 4W + 4 <= col_end here */
            lo = hi;
        }
    }
    for (; W <= w1; W++) boundary_word(W);
}

/* name1 / name2 / strand arenas for chunks whose parts are not all the same: EIGHT LANES per read (text bytes, coalesced in
 * 8-byte pieces).  Names are a few dozen bytes: a whole warp per read spent ~60 instructions of per-read set-up on one
 * iteration of copying (0.85 ms for 4 M BGI-shape reads, whose names are all different). */
constexpr int EN_LANES = 8;
__global__ void __launch_bounds__(256) k_emit_names(EncBatchDev b, HeaderDev h, u8* out) {
    const u32 l8 = threadIdx.x & (EN_LANES - 1);
    const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) / EN_LANES;
    if (i >= b.n_reads) return;
    const u32 c = chunk_of_read(b, i);
    const ChunkDev& ck = b.chunks[c];
    const u32 fl = ck.flags;
    const bool need1 = !(fl & RPQ_NAME1_SAME), need2 = (h.flags & RPQ_HAS_NAME2) && !(fl & RPQ_NAME2_SAME), need3 = !(fl & RPQ_STRAND_SAME);
    if (!need1 && !need2 && !need3) return;
    u8* o = out + ck.out_offset;
    const ReadMeta m = b.meta[i];
    const uint4 lc = b.loc[i];
    u32 f, rec; read_locus(b, i, f, rec);
    const u8* text = b.t[f].text;
    const u8* name = text + lc.x;
    if (need1) { u8* d = o + ck.off_n1 + b.n1off[i]; for (u32 k = l8; k < m.name1_len; k += EN_LANES) d[k] = name[k]; }
    if (need2) { u8* d = o + ck.off_n2 + b.n2off[i]; const u32 l = (u32)m.name_len - m.name2_off; for (u32 k = l8; k < l; k += EN_LANES) d[k] = name[m.name2_off + k]; }
    if (need3) { const u8* s = text + lc.z; u8* d = o + ck.off_strand + b.soff[i]; for (u32 k = l8; k < m.strand_len; k += EN_LANES) d[k] = s[k]; }
}

}  // namespace rpq
