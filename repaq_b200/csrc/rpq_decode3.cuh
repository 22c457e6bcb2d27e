/*
 * rpq_decode3.cuh - k_dec_format3: the record formatter, third generation.  Same frame as k_dec_format2 (a thread per read,
 * quality plane and records staged in shared memory, 16-byte stores to HBM) but the sequence and quality lines are
 * produced FOUR positions per step:
 *   - four quality bytes come from one funnel-shifted (and, for the reverse strand, byte-reversed) shared-memory word;
 *   - four bases come from one byte of the 2-bit column through a 256-entry lookup (forward, or reverse-complement:
 *     reference src/rfqcodec.cpp:833-853 and src/read.cpp:77-115 fused);
 *   - 'N' restoration is a SIMD compare of the quality word with the N quality (:1093-1100) or four bits of the N bitmap;
 *   - both lines are written through a small byte sink that emits aligned 32-bit shared stores.
 * v2 did all of this per base (~9 k thread instructions per read, profiles/r01_v2_ncu_full_k_dec_format2.csv).
 */
#pragma once
#include "rpq_decode2.cuh"

namespace rpq {

/* sequential byte sink into shared memory: after the first (unaligned) bytes every put4 is one aligned 32-bit store */
struct Sink {
    u8* dst; u32 res, nres, need;
    __device__ __forceinline__ void init(u8* d) { dst = d; res = 0; nres = 0; need = (4u - (u32)(reinterpret_cast<uintptr_t>(d) & 3u)) & 3u; }
    __device__ __forceinline__ void put4(u32 w) {
        if (need) {
            for (u32 k = 0; k < need; k++) dst[k] = (u8)(w >> (8u * k));
            dst += need; res = w >> (8u * need); nres = 4u - need; need = 0;
            return;
        }
        if (nres == 0) { *reinterpret_cast<u32*>(dst) = w; dst += 4; return; }
        *reinterpret_cast<u32*>(dst) = res | (w << (8u * nres));
        dst += 4;
        res = w >> (8u * (4u - nres));
    }
    __device__ __forceinline__ void put1(u8 c) {
        if (need) { *dst++ = c; need--; return; }
        res |= (u32)c << (8u * nres); nres++;
        if (nres == 4u) { *reinterpret_cast<u32*>(dst) = res; dst += 4; res = 0; nres = 0; }
    }
    __device__ __forceinline__ void flush() { for (u32 k = 0; k < nres; k++) dst[k] = (u8)(res >> (8u * k)); dst += nres; nres = 0; res = 0; }
};

/* 4 bytes at any byte address of shared memory */
__device__ __forceinline__ u32 lds4(const u8* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32* w = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    const u32 sh = (u32)(a & 3u) * 8u;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}

__global__ void __launch_bounds__(256) k_dec_format3(DecBatchDev b, HeaderDev h, Fmt2Cfg cfg, u32 read_first, u32 read_end) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u64 s_start[2], s_end[2];
    __shared__ u64 s_q0, s_q1;
    __shared__ u32 s_lut_fwd[256], s_lut_rc[256];
    const int tid = threadIdx.x;
    const u32 G = cfg.reads_per_cta;
    /* two threads per read, in different warps (no divergence): threads [0,G) write name + sequence line, [G,2G) strand + quality line */
    const int rt = tid < (int)G ? tid : tid - (int)G, half = tid < (int)G ? 0 : 1;
    const u32 i0 = read_first + blockIdx.x * G;          /* the launch covers reads [read_first, read_end) */
    const u32 n_here = read_end - i0 < G ? read_end - i0 : G;
    u8* s_plane = dyn;
    u8* s_out[2] = {dyn + cfg.plane_cap, dyn + cfg.plane_cap + cfg.out_cap};
    const u32 nstreams = b.split_pairs ? 2u : 1u;

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
    if (tid < 2) { s_start[tid] = ~0ull; s_end[tid] = 0; }
    __syncthreads();
    const bool active = rt < (int)n_here;
    const u32 i = i0 + rt;
    u32 c = 0, r = 0, rl = 0, stream = 0, olen = 0;
    u64 oabs = 0, qabs = 0;
    if (active) {
        c = b.read_chunk[i];
        const DecChunk& ck = b.chunks[c];
        r = i - ck.read_base; rl = b.rlen[i]; olen = b.olen[i];
        stream = b.split_pairs ? (r & 1u) : 0u;
        oabs = ck.out_off[stream] + b.outoff[i];
        qabs = ck.plane_off + b.qualoff[i];
        if (half == 0) {
            if ((u32)rt < nstreams) s_start[stream] = oabs;
            if ((u32)rt + nstreams >= n_here) s_end[stream] = oabs + olen;
            if (rt == 0) s_q0 = qabs;
            if ((u32)rt == n_here - 1) s_q1 = qabs + rl;
        }
    }
    __syncthreads();
    const u64 q0 = s_q0, q1 = s_q1, qa = q0 & ~15ull;
    {
        const u32 nvec = (u32)((q1 - qa + 15) >> 4);
        const uint4* src = reinterpret_cast<const uint4*>(b.plane + qa);
        uint4* dst = reinterpret_cast<uint4*>(s_plane);
        for (u32 k = tid; k < nvec; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    if (active) {
        const DecChunk& ck = b.chunks[c];
        const u8* in = b.body + ck.in_off;
        const u32 fl = ck.flags;
        const bool il = (fl & RPQ_PE_INTERLEAVED) != 0;
        const bool ov_on = il && (h.flags & RPQ_ENCODE_PE_BY_OVERLAP);
        const bool odd = (r & 1u) != 0;
        const u32 xy = il ? r >> 1 : r;
        u8* o = s_out[stream] + (u32)(oabs - (s_start[stream] & ~15ull));
        const u8* q = s_plane + (u32)(qabs - qa);
        /* ---- strand length first: it fixes where every part of the record lies */
        const u32 ls = (fl & (RPQ_STRAND_SAME | RPQ_STRAND_LEN_SAME)) ? in[ck.off_slen] : in[ck.off_slen + r];
        const u32 name_end = olen - (2u * rl + ls + 3u);          /* bytes of the name line including its line break */
        if (half == 0) {
        u32 w_at = 0;
        /* ---- name (reference src/rfqcodec.cpp:1157-1231) */
        const u32 l1 = (fl & (RPQ_NAME1_SAME | RPQ_NAME1_LEN_SAME)) ? in[ck.off_n1len] : in[ck.off_n1len + r];
        const u8* n1 = in + ck.off_n1 + ((fl & RPQ_NAME1_SAME) ? 0u : b.n1off[i]);
        for (u32 k = 0; k < l1; k++) o[w_at + k] = n1[k];
        w_at += l1;
        if (h.flags & RPQ_HAS_LANE) { o[w_at++] = ':'; w_at += put_dec(o + w_at, (fl & RPQ_LANE_SAME) ? in[ck.off_lane] : in[ck.off_lane + xy]); }
        if (h.flags & RPQ_HAS_TILE) { const u32 k = (fl & RPQ_TILE_SAME) ? 0u : xy; o[w_at++] = ':'; w_at += put_dec(o + w_at, (u32)in[ck.off_tile + 2 * k] | ((u32)in[ck.off_tile + 2 * k + 1] << 8)); }
        if (h.flags & RPQ_HAS_X) { o[w_at++] = ':'; w_at += put_dec(o + w_at, b.xs[ck.read_base + xy]); }
        if (h.flags & RPQ_HAS_Y) { o[w_at++] = ':'; w_at += put_dec(o + w_at, b.ys[ck.read_base + xy]); }
        if (h.flags & RPQ_HAS_NAME2) {
            const u32 l2 = (fl & (RPQ_NAME2_SAME | RPQ_NAME2_LEN_SAME)) ? in[ck.off_n2len] : in[ck.off_n2len + r];
            const u8* n2 = in + ck.off_n2 + ((fl & RPQ_NAME2_SAME) ? 0u : b.n2off[i]);
            for (u32 k = 0; k < l2; k++) o[w_at + k] = n2[k];
            if ((fl & RPQ_NAME2_SAME) && il && odd && h.name2_diff_char != 0 && h.name2_diff_pos < l2) o[w_at + h.name2_diff_pos] = h.name2_diff_char;
            w_at += l2;
        }
        o[w_at++] = '\n';
        }
        const u8* sp = in + ck.off_strand + ((fl & RPQ_STRAND_SAME) ? 0u : b.soff[i]);
        u8* o_seq = o + name_end;
        u8* o_str = o_seq + rl + 1;
        u8* o_qual = o_str + ls + 1;
        if (half == 0) o_seq[rl] = '\n';
        else {
            for (u32 k = 0; k < ls; k++) o_str[k] = sp[k];
            o_str[ls] = '\n';
            o_qual[rl] = '\n';
        }

        /* ---- sequence + quality, four positions per step */
        const u8* seqb = in + ck.off_seq;
        const u32* nmap = b.nmap + ck.nmap_off;
        const long long so = b.seqoff[i];
        int ov = 0; u32 prev_rl = 0;
        if (ov_on && odd) { ov = (int)(signed char)in[ck.off_ov + (r >> 1)] - (int)h.overlap_shift; prev_rl = b.rlen[i - 1]; }
        const bool rc = il && odd;
        const bool npos_mode = (h.flags & RPQ_ENCODE_N_POS) != 0;
        const u8 nq = (u8)h.n_base_qual;
        const u32 nq4 = 0x01010101u * nq;
        const long long unpacked = ck.seq_size * 4u < ck.total_len ? ck.seq_size * 4u : ck.total_len;
        /* compact index of output position jo: piece A for jo < bnd, piece B after; ci = c + sgn * jo */
        long long cA, cB; u32 bnd = rl; const int sgn = rc ? -1 : 1;
        if (!rc) { cA = so; cB = so; }
        else if (ov >= 0) { cA = so - ov + (long long)rl - 1; cB = cA; }
        else { const u32 a = (u32)(-ov); cA = so - (long long)prev_rl + a - 1; cB = so + (long long)rl - 1; bnd = a < rl ? a : rl; }
        auto slow_base = [&](u32 jo) -> u8 {
            const long long ci = (jo < bnd ? cA : cB) + (long long)sgn * jo;
            u8 base = 'N';
            if (ci >= 0 && ci < unpacked) { const u32 code = (seqb[ci >> 2] >> (2 * (ci & 3))) & 3u; base = code == 0 ? 'G' : code == 1 ? 'A' : code == 2 ? 'T' : 'C'; }
            if (npos_mode) { if (ci >= 0 && (u64)ci < ck.total_len && ((nmap[ci >> 5] >> (ci & 31)) & 1u)) base = 'N'; }
            else if (q[rc ? rl - 1 - jo : jo] == nq) base = 'N';
            return rc ? complement_base(base) : base;
        };
        const u32 ngroups = rl >> 2;
        if (half == 1) {
            /* quality line (reversed for the reverse strand) */
            Sink qs; qs.init(o_qual);
            for (u32 g = 0; g < ngroups; g++) {
                const u32 jo = 4u * g;
                qs.put4(rc ? __byte_perm(lds4(q + (rl - 4u - jo)), 0, 0x0123) : lds4(q + jo));
            }
            for (u32 jo = ngroups * 4u; jo < rl; jo++) qs.put1(q[rc ? rl - 1 - jo : jo]);
            qs.flush();
        } else {
            Sink ss; ss.init(o_seq);
            for (u32 g = 0; g < ngroups; g++) {
                const u32 jo = 4u * g;
                const bool piece_a = jo + 3u < bnd, piece_b = jo >= bnd;
                const long long cbase = piece_a ? cA : cB;
                const long long ci_lo = rc ? cbase - (long long)jo - 3 : cbase + (long long)jo;
                u32 bw;
                if ((piece_a || piece_b) && ci_lo >= 0 && ci_lo + 3 < unpacked) {
                    const u32 k = (u32)(ci_lo >> 2), ph = 2u * (u32)(ci_lo & 3);
                    const u32 two = (u32)seqb[k] | (ph ? (u32)seqb[k + 1] << 8 : 0u);
                    const u32 code8 = (two >> ph) & 0xFFu;
                    bw = rc ? s_lut_rc[code8] : s_lut_fwd[code8];
                    u32 mask;
                    if (npos_mode) {
                        const u32 wi = (u32)(ci_lo >> 5), bp = (u32)(ci_lo & 31);
                        u32 m4 = (nmap[wi] >> bp) & 0xFu;
                        if (bp > 28u) m4 |= (nmap[wi + 1] << (32u - bp)) & 0xFu;
                        if (rc) m4 = ((m4 & 1u) << 3) | ((m4 & 2u) << 1) | ((m4 & 4u) >> 1) | ((m4 & 8u) >> 3);
                        mask = ((m4 | (m4 << 7) | (m4 << 14) | (m4 << 21)) & 0x01010101u) * 0xFFu;
                    } else {
                        const u32 qw = rc ? __byte_perm(lds4(q + (rl - 4u - jo)), 0, 0x0123) : lds4(q + jo);
                        mask = __vcmpeq4(qw, nq4);
                    }
                    bw = (bw & ~mask) | (0x4E4E4E4Eu & mask);
                } else {
                    bw = (u32)slow_base(jo) | ((u32)slow_base(jo + 1) << 8) | ((u32)slow_base(jo + 2) << 16) | ((u32)slow_base(jo + 3) << 24);
                }
                ss.put4(bw);
            }
            for (u32 jo = ngroups * 4u; jo < rl; jo++) ss.put1(slow_base(jo));
            ss.flush();
        }
    }
    __syncthreads();
    for (u32 s = 0; s < nstreams; s++) {
        const u64 a = s_start[s], e = s_end[s];
        if (a == ~0ull || e <= a) continue;
        const u64 base = a & ~15ull;
        u8* g = b.out[s];
        const u8* sm = s_out[s];
        const u64 v0 = (a + 15) & ~15ull, v1 = e & ~15ull;
        if (v0 >= v1) { for (u64 p = a + tid; p < e; p += blockDim.x) g[p] = sm[p - base]; continue; }
        for (u64 p = a + tid; p < v0; p += blockDim.x) g[p] = sm[p - base];
        const u32 nvec = (u32)((v1 - v0) >> 4);
        uint4* gd = reinterpret_cast<uint4*>(g + v0);
        const uint4* sd = reinterpret_cast<const uint4*>(sm + (v0 - base));
        for (u32 k = tid; k < nvec; k += blockDim.x) gd[k] = sd[k];
        for (u64 p = v1 + tid; p < e; p += blockDim.x) g[p] = sm[p - base];
    }
}

}  // namespace rpq
