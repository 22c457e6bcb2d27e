/*
 * rpq_index.cuh - FASTQ line index, record lengths and the chunk cut.
 *
 * Replaces, for a whole FASTQ image resident in HBM, the byte loop of FastqReader::getLine (reference
 * src/fastqreader.cpp:94-156, the '\n' scan at :100-105) and the chunk-cut loop of Repaq::compress / compressPE
 * (src/repaq.cpp:546-553, 656-663: append record, total += bases, flush when total >= chunkSize).
 *
 * Supported line ends: all "\n" or all "\r\n" (the reference additionally accepts lone '\r' and silently swallows
 * an empty line after a break; those inputs are rejected here with RPQ_ERR_FASTQ).
 */
#pragma once
#include "rpq_common.cuh"

namespace rpq {

constexpr int IDX_THREADS = 256;
constexpr int IDX_CHUNKS = 4;                                /* 16-byte pieces per lane and round */
constexpr int IDX_ROUNDS = 4;                                /* rounds per CTA */
constexpr int IDX_WARP_BYTES = 32 * 16 * IDX_CHUNKS;         /* 2 KiB per warp and round */
constexpr int IDX_ROUND_BYTES = (IDX_THREADS / 32) * IDX_WARP_BYTES;   /* 16 KiB */
constexpr int IDX_TILE = IDX_ROUNDS * IDX_ROUND_BYTES;       /* 64 KiB per CTA */

struct IndexCounters {
    u32 ticket;     /* dynamic tile id */
    u32 n_nl;       /* '\n' bytes */
    u32 n_cr;       /* '\r' bytes */
    u32 n_crlf;     /* '\n' preceded by '\r' */
    /* written by k_index_finish */
    u32 n_lines;
    u32 crlf;
    u32 bad_eol;    /* mixed or lone '\r' line ends */
    u32 pad;
};

constexpr u64 TS_AGG = 1ull << 62, TS_PREFIX = 2ull << 62, TS_MASK = 3ull << 62;

__device__ __forceinline__ u32 nl_mask16(uint4 v, u8 c) {
    const u32 cc = 0x01010101u * c;
    u32 m0 = __vcmpeq4(v.x, cc), m1 = __vcmpeq4(v.y, cc), m2 = __vcmpeq4(v.z, cc), m3 = __vcmpeq4(v.w, cc);
    /* one bit per byte: take bit 0 of every byte lane and pack */
    auto pack = [](u32 m) -> u32 { m &= 0x01010101u; return (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xFu; };
    return pack(m0) | (pack(m1) << 4) | (pack(m2) << 8) | (pack(m3) << 12);
}

/*
 * One pass over the text: positions of every '\n', in order.  A CTA takes a 64 KiB tile (4 rounds of 16 KiB, uint4 loads),
 * keeps the newline masks in registers, publishes its count and gets its rank base from a chained scan over tiles with a
 * warp-wide decoupled look-back (32 predecessors per probe), then writes the positions.
 */
__global__ void __launch_bounds__(IDX_THREADS) k_index_lines(const u8* __restrict__ text, u64 len, u32* __restrict__ nl, u32 nl_cap,
                                                            u64* tile_state, IndexCounters* ctr) {
    __shared__ u32 s_tile;
    __shared__ u32 s_tot[IDX_ROUNDS][IDX_THREADS / 32];
    __shared__ u32 s_prefix;
    __shared__ u32 s_cr, s_crlf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_tile = atomicAdd(&ctr->ticket, 1u); s_cr = 0; s_crlf = 0; }
    __syncthreads();
    const u32 tile = s_tile;
    const u64 tbase = (u64)tile * IDX_TILE;

    u32 nlm[IDX_ROUNDS][IDX_CHUNKS / 2];       /* two 16-bit masks per register */
    u32 exc[IDX_ROUNDS][IDX_CHUNKS / 2];       /* two 16-bit in-warp exclusive offsets per register */
    u32 ncr = 0, ncrlf = 0;
#pragma unroll
    for (int r = 0; r < IDX_ROUNDS; r++) {
        const u64 wbase = tbase + (u64)r * IDX_ROUND_BYTES + (u64)warp * IDX_WARP_BYTES;
        u32 wtot = 0;
#pragma unroll
        for (int k = 0; k < IDX_CHUNKS; k++) {
            const u64 p = wbase + (u64)k * 512 + (u64)lane * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            u32 valid = 0xFFFFu;
            if (p + 16 <= len) v = *reinterpret_cast<const uint4*>(text + p);
            else if (p < len) {
                u32 w[4] = {0, 0, 0, 0};
                const int n = (int)(len - p);
                for (int i = 0; i < n; i++) w[i >> 2] |= (u32)text[p + i] << (8 * (i & 3));
                v = make_uint4(w[0], w[1], w[2], w[3]);
                valid = (1u << n) - 1u;
            } else valid = 0;
            const u32 m = nl_mask16(v, '\n') & valid;
            const u32 c = nl_mask16(v, '\r') & valid;
            u32 pair = m & (c << 1);
            if ((m & 1u) && p > 0 && text[p - 1] == '\r') pair |= 1u;
            ncr += (u32)__popc(c); ncrlf += (u32)__popc(pair);
            u32 t; const u32 ex = wtot + warp_excl_scan((u32)__popc(m), lane, t); wtot += t;
            if (k & 1) { nlm[r][k >> 1] |= m << 16; exc[r][k >> 1] |= ex << 16; } else { nlm[r][k >> 1] = m; exc[r][k >> 1] = ex; }
        }
        if (lane == 0) s_tot[r][warp] = wtot;
    }
    ncr = warp_sum(ncr); ncrlf = warp_sum(ncrlf);
    if (lane == 0) { if (ncr) atomicAdd(&s_cr, ncr); if (ncrlf) atomicAdd(&s_crlf, ncrlf); }
    __syncthreads();
    u32 btot = 0;
#pragma unroll
    for (int r = 0; r < IDX_ROUNDS; r++)
#pragma unroll
        for (int w = 0; w < IDX_THREADS / 32; w++) btot += s_tot[r][w];

    if (warp == 0) {
        volatile u64* st = tile_state;
        u32 prefix = 0;
        if (tile > 0) {
            if (lane == 0) { st[tile] = TS_AGG | btot; __threadfence(); }
            int j = (int)tile - 1;                       /* newest predecessor not yet accounted for */
            for (;;) {
                const int idx = j - lane;
                u64 s = 2ull << 62;                       /* before tile 0: an empty prefix (TS_PREFIX | 0) */
                if (idx >= 0) s = st[idx];
                const u32 unset = __ballot_sync(0xffffffffu, (s & TS_MASK) == 0);
                const u32 pre = __ballot_sync(0xffffffffu, (s & TS_MASK) == TS_PREFIX);
                /* usable lanes: those before the first unset one, up to and including the first prefix */
                const int first_unset = unset ? __ffs((int)unset) - 1 : 32;
                const int first_pre = pre ? __ffs((int)pre) - 1 : 32;
                const int upto = first_pre < first_unset ? first_pre + 1 : first_unset;   /* lanes [0, upto) are summed */
                u32 v = lane < upto ? (u32)s : 0u;
                prefix += warp_sum(v);
                if (first_pre < first_unset) break;
                j -= upto;
                if (upto == 0) RPQ_SPIN_HINT();
            }
        }
        if (lane == 0) {
            __threadfence();
            st[tile] = TS_PREFIX | (u64)(prefix + btot);
            s_prefix = prefix;
            if (s_cr) atomicAdd(&ctr->n_cr, s_cr);
            if (s_crlf) atomicAdd(&ctr->n_crlf, s_crlf);
            if ((u64)(tile + 1) * IDX_TILE >= len) ctr->n_nl = prefix + btot;     /* the last tile knows the total */
        }
    }
    __syncthreads();
    u32 base = s_prefix;
#pragma unroll
    for (int r = 0; r < IDX_ROUNDS; r++) {
        u32 wpre = 0;
#pragma unroll
        for (int w = 0; w < IDX_THREADS / 32; w++) { const u32 t = s_tot[r][w]; if (w < warp) wpre += t; }
        const u64 wbase = tbase + (u64)r * IDX_ROUND_BYTES + (u64)warp * IDX_WARP_BYTES;
#pragma unroll
        for (int k = 0; k < IDX_CHUNKS; k++) {
            u32 m = (nlm[r][k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
            u32 o = base + wpre + ((exc[r][k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
            const u32 p = (u32)(wbase + (u64)k * 512 + (u64)lane * 16);
            while (m) {
                const int bb = __ffs((int)m) - 1;
                m &= m - 1;
                if (o < nl_cap) nl[o] = p + (u32)bb;
                o++;
            }
        }
#pragma unroll
        for (int w = 0; w < IDX_THREADS / 32; w++) base += s_tot[r][w];
    }
}

/* line-end mode, the virtual final newline, line count */
__global__ void k_index_finish(const u8* text, u64 len, u32* nl, u32 nl_cap, IndexCounters* ctr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const u32 n = ctr->n_nl;
    u32 crlf = 0, bad = 0;
    if (ctr->n_cr) { crlf = 1; if (ctr->n_cr != ctr->n_crlf || ctr->n_crlf != n) bad = 1; }
    u32 lines = n;
    if (len > 0 && text[len - 1] != '\n') {
        if (n < nl_cap) nl[n] = (u32)len + crlf;       /* so that line_end() == len */
        lines = n + 1;
    }
    ctr->n_lines = lines; ctr->crlf = crlf; ctr->bad_eol = bad;
}

/* per-unit (read, or pair) statistics */
struct UnitStats {
    u32 first_empty;      /* first unit with an empty line: input ends there (src/fastqreader.cpp:180-181,190-191) */
    u32 first_qual_len;   /* first unit whose quality length != sequence length */
    u32 first_name_len;   /* first unit with a name/strand line > 255 bytes */
    u32 first_read_len;   /* first unit with a read > 65535 bases */
    u32 min_bases, max_bases;
    u32 max_read;         /* longest single read */
    u32 max_head;         /* most bytes from the start of a name line to the start of the quality line */
    u32 n_chunks;         /* written by k_cut */
    u32 units_in_chunks;  /* units covered by the emitted chunks */
};

__global__ void k_unit_lengths(EncBatchDev b, u32 n_units, u32* __restrict__ rlen, u32* __restrict__ unit_bases, UnitStats* st) {
    const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const u32 per = b.is_pe ? 2u : 1u;
    u32 bases = 0;
    bool empty = false, badq = false, badn = false, badl = false;
    u32 longest = 0, head = 0;
    for (u32 k = 0; k < per; k++) {
        const u32 i = u * per + k;
        u32 f, rec; read_locus(b, i, f, rec);
        const TextDev& t = b.t[f];
        u32 ln[4];
#pragma unroll
        for (u32 j = 0; j < 4; j++) { const u32 L = 4 * rec + j; ln[j] = line_end(t, L) - line_start(t, L); }
        if (!ln[0] || !ln[1] || !ln[2] || !ln[3]) empty = true;
        if (ln[1] != ln[3]) badq = true;
        if (ln[0] > 255 || ln[2] > 255) badn = true;
        if (ln[1] > 65535) badl = true;
        rlen[i] = ln[1];
        {
            uint4 lc;
            lc.x = line_start(t, 4 * rec); lc.y = line_start(t, 4 * rec + 1); lc.z = line_start(t, 4 * rec + 2); lc.w = line_start(t, 4 * rec + 3);
            b.loc[i] = lc;
            head = lc.w - lc.x > head ? lc.w - lc.x : head;
        }
        longest = ln[1] > longest ? ln[1] : longest;
        bases += ln[1];
    }
    unit_bases[u] = bases;
    if (empty) atomicMin(&st->first_empty, u);
    if (badq) atomicMin(&st->first_qual_len, u);
    if (badn) atomicMin(&st->first_name_len, u);
    if (badl) atomicMin(&st->first_read_len, u);
    atomicMin(&st->min_bases, bases);
    atomicMax(&st->max_bases, bases);
    atomicMax(&st->max_read, longest);
    atomicMax(&st->max_head, head);
}

/*
 * The greedy cut (Q19): chunk k+1 starts after the first unit at which the running base count since the chunk
 * start reaches chunk_bases.  `prefix` is the INCLUSIVE prefix sum of unit_bases.  One CTA.
 *   uniform unit size: arithmetic.   otherwise: a warp walks the chain, 32-ary search per chunk.
 */
__global__ void k_cut(const u64* __restrict__ prefix, u32 n_units, u32 chunk_bases, u32 uniform_bases, int final, u32 per,
                      u32* __restrict__ chunk_first, u32 cap, UnitStats* st) {
    if (blockIdx.x != 0) return;
    if (uniform_bases) {
        const u32 upc = (chunk_bases + uniform_bases - 1) / uniform_bases;      /* units per chunk */
        const u32 full = n_units / upc, rem = n_units % upc;
        const u32 n = full + ((final && rem) ? 1u : 0u);
        for (u32 c = threadIdx.x; c <= full && c < cap; c += blockDim.x) chunk_first[c] = c * upc * per;
        if (threadIdx.x == 0) {
            if (final && rem && n < cap) chunk_first[n] = n_units * per;
            st->n_chunks = n;
            st->units_in_chunks = (final && rem) ? n_units : full * upc;
        }
        return;
    }
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    u32 cur = 0, n = 0;
    if (lane == 0 && cap) chunk_first[0] = 0;
    while (cur < n_units) {
        const u64 base = cur ? prefix[cur - 1] : 0ull;
        const u64 target = base + chunk_bases;
        if (prefix[n_units - 1] < target) break;           /* not enough bases left for a full chunk */
        /* smallest j in [cur, n_units) with prefix[j] >= target: 32-ary narrowing */
        u32 lo = cur, hi = n_units - 1;                    /* invariant: answer in [lo, hi] */
        while (hi > lo) {
            const u32 span = hi - lo;                      /* probe points lo + span*(lane+1)/33 */
            const u32 pt = lo + (u32)(((u64)span * (u32)(lane + 1)) / 33u);
            const bool ge = prefix[pt] >= target;
            const u32 m = __ballot_sync(0xffffffffu, ge);
            if (m == 0) { lo = __shfl_sync(0xffffffffu, pt, 31) + 1; }
            else {
                const int fl = __ffs((int)m) - 1;
                const u32 new_hi = __shfl_sync(0xffffffffu, pt, fl);
                const u32 below = fl ? __shfl_sync(0xffffffffu, pt, fl - 1) + 1 : lo;
                hi = new_hi; lo = below;
            }
        }
        cur = lo + 1;
        n++;
        if (lane == 0 && n < cap) chunk_first[n] = cur * per;
    }
    u32 covered = cur;
    if (final && cur < n_units) { n++; covered = n_units; if (lane == 0 && n < cap) chunk_first[n] = n_units * per; }
    if (lane == 0) { st->n_chunks = n; st->units_in_chunks = covered; }
}

}  // namespace rpq
