/*
 * rpq_decode3.cuh - k_dec_format3: the record formatter, third generation.  Same frame as k_dec_format2 (a thread per read,
 * quality plane and records staged in shared memory, 16-byte stores to HBM) but the sequence and quality lines are
 * produced FOUR positions per step:
 *   - four quality bytes come from one funnel-shifted (and, for the reverse strand, byte-reversed) shared-memory word;
 *   - four bases come from one byte of the 2-bit column through a 256-entry lookup (forward, or reverse-complement:
 *     reference src/rfqcodec.cpp:833-853 and src/read.cpp:77-115 fused);
 *   - 'N' restoration is a SIMD compare of the quality word with the N quality (:1093-1100) or four bits of the N bitmap;
 *   - both lines are written destination-first: whole aligned 32-bit shared stores, the source phase hoisted out of the loop
 *     (one PRMT per quality word, one funnel shift per sixteen bases).
 * v2 did all of this per base (~9 k thread instructions per read, profiles/r01_v2_ncu_full_k_dec_format2.csv).
 */
#pragma once
#include "rpq_decode2.cuh"

namespace rpq {

__global__ void __launch_bounds__(256) k_dec_format3(DecBatchDev b, HeaderDev h, Fmt2Cfg cfg, u32 read_first, u32 read_end) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u64 s_start[2], s_end[2];
    __shared__ u64 s_q0, s_q1;
    __shared__ u32 s_lut_fwd[256], s_lut_rc[256];
    const int tid = threadIdx.x;
    const u32 G = cfg.reads_per_cta;
    /* two threads per read, in different warps (no divergence): threads [0,G) write the sequence line, [G,2G) the name, strand and
     * quality lines (about the same number of instructions each) */
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
    /* the CTA's slice of the quality plane: one TMA bulk copy, awaited only where a thread first needs the plane (the name lines
     * are formatted while it is in flight) */
    const u32 plane_bytes = (u32)((q1 - qa + 15) >> 4) << 4;
#ifdef RPQ_EMU
    {
        const uint4* src = reinterpret_cast<const uint4*>(b.plane + qa);
        uint4* dst = reinterpret_cast<uint4*>(s_plane);
        for (u32 k = tid; k < plane_bytes / 16u; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    auto plane_ready = [&]() {};
#else
    __shared__ __align__(8) unsigned long long s_mbar;
    const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (plane_bytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(plane_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((u32)__cvta_generic_to_shared(s_plane)), "l"(b.plane + qa), "r"(plane_bytes), "r"(mbar) : "memory");
        } else asm volatile("mbarrier.arrive.shared.b64 _, [%0];" ::"r"(mbar) : "memory");
    }
    __syncthreads();                                           /* the barrier object is initialised for everybody */
    auto plane_ready = [&]() {
        u32 done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
    };
#endif

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
        if (half == 1) {
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

        plane_ready();
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
        /* Both lines are produced destination-first: up to three bytes until the output is word aligned, then whole aligned
         * 32-bit words, then up to three bytes.  The source phase (plane byte offset, 2-bit phase) is constant along a line. */
        const int d = rc ? -1 : 1;
        if (half == 1) {
            /* quality line: a byte-shifted (reverse strand: byte-reversed) copy out of the staged plane, one PRMT per word */
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
            auto range = [&](long long cb, long long jlo, long long jhi, u32& m0, u32& m1) {
                long long a, z;                                    /* allowed first positions jo of a word: a <= jo <= z */
                if (!rc) { a = -cb; z = lim - cb - 4; } else { a = cb - lim + 1; z = cb - 3; }
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
                const uintptr_t ca = reinterpret_cast<uintptr_t>(seqb) + (uintptr_t)(ci0 >> 2);
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
                                const u32 wi = (u32)(ci_lo >> 5), bp = (u32)(ci_lo & 31);
                                u32 m4 = (nmap[wi] >> bp) & 0xFu;
                                if (bp > 28u) m4 |= (nmap[wi + 1] << (32u - bp)) & 0xFu;
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
        const u64 a = s_start[s], e = s_end[s];
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
}

}  // namespace rpq
