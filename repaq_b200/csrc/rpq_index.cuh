/*
 * rpq_index.cuh - FASTQ line index, record lengths and the chunk cut.
 *
 * Replaces, for a whole FASTQ image resident in HBM, the byte loop of FastqReader::getLine (reference
 * src/fastqreader.cpp:94-156, the '\n' scan at :100-105) and the chunk-cut loop of Repaq::compress / compressPE
 * (src/repaq.cpp:546-553, 656-663: append record, total += bases, flush when total >= chunkSize).
 *
 * Line breaks are the reference reader's: '\n', '\r', "\r\n", and a '\n' directly after any break (a swallowed blank
 * line), with its 1 MiB buffer-edge exception (k_index_lines).
 */
#pragma once
#include "rpq_common.cuh"

namespace rpq {

constexpr int IDX_THREADS = 512;
constexpr int IDX_ROW = 128;                                 /* bytes per thread: 8 pieces of 16 bytes */
constexpr int IDX_PIECES = IDX_ROW / 16;
constexpr int IDX_WARP_BYTES = 32 * IDX_ROW;                 /* 4 KiB per warp: one TMA bulk copy */
constexpr int IDX_TILE = IDX_THREADS * IDX_ROW;              /* 64 KiB per CTA */
constexpr int IDX_SMEM = IDX_TILE;                            /* the tile */

struct IndexCounters {
    u32 overflow;   /* a tile holds more line breaks than its region of the tile-local index: the host indexes again with full regions */
    u32 n_nl;       /* line breaks */
    u32 n_w2;       /* line breaks of two bytes (the second one a swallowed '\n') */
    u32 pad0;
    /* written by k_index_finish */
    u32 n_lines;
    u32 crlf;       /* 1: every line break has two bytes */
    u32 irregular;  /* breaks of one AND of two bytes: the text goes through k_canon_* first */
    u32 pad;
};

constexpr u64 RD_BUF = 1ull << 20;                           /* FQ_BUF_SIZE of the reference's reader (src/fastqreader.cpp:5) */

/* exact per-byte equality, SIMD in a register: bit 7 of every byte of the result is set iff that byte of v equals c.
 * ((v ^ cccc) & 0x7f7f7f7f) + 0x7f7f7f7f carries into bit 7 iff the low seven bits differ; bit 7 itself must be clear in v
 * (c < 0x80).  Three instructions per word, no carries between bytes. */
__device__ __forceinline__ u32 eq_lowascii(u32 v, u32 cccc) {
    const u32 t = ((v ^ cccc) & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | v) & 0x80808080u;
}
/* bit 7 of every byte CLEAR iff that byte equals c (the complement of eq_lowascii before masking): lets four words be
 * AND-ed together for an "any byte equals c" test */
__device__ __forceinline__ u32 ne_lowascii(u32 v, u32 cccc) {
    return (((v ^ cccc) & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v;
}
/* flags at bit 7 of the bytes of two words -> 8 consecutive bits in text order (lo word first) at bits 24..31.
 * (lo >> 4 | hi) * (1 + 2^7 + 2^14 + 2^21): every partial product lands on a distinct bit, so there are no carries. */
__device__ __forceinline__ u32 pack8(u32 lo, u32 hi) { return ((lo >> 4) | hi) * 0x00204081u; }
__device__ __forceinline__ u32 mask16(u32 m0, u32 m1, u32 m2, u32 m3) {
    return (pack8(m0, m1) >> 24) | ((pack8(m2, m3) >> 16) & 0xFF00u);
}
__device__ __forceinline__ bool is_brk(u8 c) { return c == '\n' || c == '\r'; }

/* may the '\n' at file offset q be taken as the second byte of a line break?  Not if it is the first or the last byte of one
 * of the reader's 1 MiB buffers: `if (end < mBufDataLen - 1 && mBuf[end] == '\n')` (src/fastqreader.cpp:113-116) */
__device__ __forceinline__ bool rd_may_swallow(u64 q) { const u64 r = q & (RD_BUF - 1); return r != 0 && r != RD_BUF - 1; }

/*
 * One pass over the text: the line breaks as FastqReader::getLine sees them (reference src/fastqreader.cpp:94-156), in order.
 * A line ends at the first '\r' or '\n'; a '\n' that directly follows the byte that ended a line is taken as part of that
 * break (so "\r\n" is one break - but so is "\n\n": a single blank line is invisible to the reference), unless it is the first
 * or last byte of a 1 MiB reader buffer.  With T = "byte is \r or \n", A(q) = '\n' at q & T(q-1) & may_swallow(q), a byte is
 * swallowed iff A(q) & !A(q-1) (exact for the first three bytes of a run of break characters; a longer run contains an empty
 * line, where the input ends anyway), and a line ends at E = T & !swallowed.  nl[] receives, per line, the offset of the LAST
 * byte of its break, so that the next line starts one byte later.
 * A CTA per 64 KiB tile, brought into shared memory by one TMA bulk copy per warp (4 KiB); a thread owns 128 contiguous bytes
 * (its pieces read in a lane-rotated order, so that the 128-bit shared loads of a quarter warp fall into eight different bank
 * groups), turns them into exact 128-bit masks with SIMD-in-register compares, and counts.  Rows without '\r' and without two
 * break characters in a row (every row of a plain file) never leave the '\n' mask.
 *
 * No CTA waits for another one.  The rank of a line among all lines of the text needs the counts of all earlier tiles; the
 * single-pass version got them from a chained scan with decoupled look-back and spent half of every CTA's life there: with ~450
 * tiles in flight and 1.5 us of work per tile a tile always finds its predecessors still looking back themselves, and sums
 * hundreds of aggregates in dependent rounds (r02: 46 % of the stall samples at the barrier behind the look-back, 2.8 TB/s).  Here a
 * tile writes its line ends into its own region of a tile-local index (lcap entries per tile) and its count into a table;
 * k_index_scan turns the counts into ranks (one CTA, 52 K tiles of a 3.4 GB text) and k_index_compact moves the entries to their
 * ranks (4 bytes per line in and out: 3 % of the text's bytes).
 */
__global__ void __launch_bounds__(IDX_THREADS, 3) k_index_lines(const u8* __restrict__ text, u64 len, u64 file_off, u32* __restrict__ nl_local, u32 lcap,
                                                               u32* __restrict__ tile_count, IndexCounters* ctr) {
    RPQ_DYN_SMEM(dyn);
    __shared__ u32 s_wtot[IDX_THREADS / 32];
    __shared__ u32 s_w2;
#ifndef RPQ_EMU
    __shared__ __align__(8) unsigned long long s_mbar[IDX_THREADS / 32];
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_w2 = 0;
#ifndef RPQ_EMU
    if (tid < IDX_THREADS / 32) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((u32)__cvta_generic_to_shared(&s_mbar[tid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();
    const u32 tile = blockIdx.x;
    const u64 tbase = (u64)tile * IDX_TILE;
    u8* row = dyn + (size_t)tid * IDX_ROW;

    /* ---- stage */
    const u64 wbase = tbase + (u64)warp * IDX_WARP_BYTES;
    const u32 wbytes = wbase >= len ? 0u : (len - wbase >= IDX_WARP_BYTES ? (u32)IDX_WARP_BYTES : (u32)(((len - wbase) + 15) & ~15ull));
#ifdef RPQ_EMU
    for (u32 k = lane; k < wbytes; k += 32) { const u64 p = wbase + k; dyn[(size_t)warp * IDX_WARP_BYTES + k] = p < len ? text[p] : 0; }
    __syncwarp();
#else
    if (wbytes) {
        const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar[warp]);
        if (lane == 0) {
            const u32 dst = (u32)__cvta_generic_to_shared(dyn + (size_t)warp * IDX_WARP_BYTES);
            asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(mbar), "r"(wbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(text + wbase), "r"(wbytes), "r"(mbar) : "memory");
        }
        u32 done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar) : "memory");
    }
#endif
    const u64 p0 = tbase + (u64)tid * IDX_ROW;                /* first byte of this thread's row */
    const int lim = p0 >= len ? 0 : (len - p0 >= IDX_ROW ? IDX_ROW : (int)(len - p0));
    if (lim < IDX_ROW) for (int i = lim; i < IDX_ROW; i++) row[i] = 0;     /* the end of the text: zeros match nothing */

    /* ---- masks of the row: '\n' and (rarely) '\r' */
    u64 nlo = 0, nhi = 0, clo = 0, chi = 0;
#pragma unroll
    for (int k = 0; k < IDX_PIECES; k++) {
        const int kk = (k + lane) & (IDX_PIECES - 1);
        const uint4 v = *reinterpret_cast<const uint4*>(row + 16 * kk);
        const u32 m = mask16(eq_lowascii(v.x, 0x0A0A0A0Au), eq_lowascii(v.y, 0x0A0A0A0Au), eq_lowascii(v.z, 0x0A0A0A0Au), eq_lowascii(v.w, 0x0A0A0A0Au));
        const u32 nocr = ne_lowascii(v.x, 0x0D0D0D0Du) & ne_lowascii(v.y, 0x0D0D0D0Du) & ne_lowascii(v.z, 0x0D0D0D0Du) & ne_lowascii(v.w, 0x0D0D0D0Du);
        const u64 placed = (u64)m << (16 * (kk & 3));
        if (kk & 4) nhi |= placed; else nlo |= placed;
        if (~nocr & 0x80808080u) {
            const u32 c = mask16(eq_lowascii(v.x, 0x0D0D0D0Du), eq_lowascii(v.y, 0x0D0D0D0Du), eq_lowascii(v.z, 0x0D0D0D0Du), eq_lowascii(v.w, 0x0D0D0D0Du));
            const u64 pc = (u64)c << (16 * (kk & 3));
            if (kk & 4) chi |= pc; else clo |= pc;
        }
    }
    /* ---- line ends E and, per line end, whether the byte after it belongs to the break (swn) */
    auto byte_at = [&](long long rel) -> u8 {                /* text[p0 + rel], rel in [-2, 128]: this warp's shared copy, or the text */
        const long long q = (long long)p0 + rel;
        if (q < 0 || (u64)q >= len) return 0;
        const long long in_warp = q - (long long)wbase;
        return (in_warp >= 0 && in_warp < IDX_WARP_BYTES) ? dyn[(size_t)warp * IDX_WARP_BYTES + in_warp] : text[q];
    };
    u64 elo = nlo | clo, ehi = nhi | chi, swn_lo = 0, swn_hi = 0;
    {
        const u64 tlo = elo, thi = ehi;
        u64 alo = nlo & (tlo << 1), ahi = nhi & ((thi << 1) | (tlo >> 63));
        if ((nlo & 1ull) && is_brk(byte_at(-1))) alo |= 1ull;
        if ((alo | ahi) != 0) {                                /* a '\n' right after a break character */
            /* not at the first or the last byte of a reader buffer: at most one buffer edge near the row */
            const u64 abs0 = file_off + p0;
            const u64 edge = (abs0 + IDX_ROW) & ~(RD_BUF - 1);   /* the largest multiple of 1 MiB <= abs0 + 128 */
            if (edge >= abs0 && edge > 0) {
                const u32 r = (u32)(edge - abs0);                /* 0..128: clear positions r and r - 1 */
                if (r < 64u) alo &= ~(1ull << r); else if (r < 128u) ahi &= ~(1ull << (r - 64u));
                if (r >= 1u && r - 1u < 64u) alo &= ~(1ull << (r - 1u)); else if (r >= 65u) ahi &= ~(1ull << (r - 65u));
            }
            /* A at the byte before the row */
            const u8 b1 = byte_at(-1);
            const u64 a_prev = (p0 >= 1 && b1 == '\n' && is_brk(byte_at(-2)) && rd_may_swallow(file_off + p0 - 1)) ? 1ull : 0ull;
            const u64 slo = alo & ~((alo << 1) | a_prev), shi = ahi & ~((ahi << 1) | (alo >> 63));     /* swallowed */
            elo = tlo & ~slo; ehi = thi & ~shi;
            swn_lo = (slo >> 1) | (shi << 63); swn_hi = shi >> 1;
        }
        if (ehi >> 63) {                                       /* a line ends at the row's last byte: is the next byte swallowed? */
            const u64 q = p0 + IDX_ROW;
            if (q < len && byte_at(IDX_ROW) == '\n' && rd_may_swallow(file_off + q) && !(ahi >> 63)) swn_hi |= 1ull << 63;
        }
    }
    const u32 cnt = (u32)(__popcll(elo) + __popcll(ehi));
    const u32 w2 = (u32)(__popcll(swn_lo & elo) + __popcll(swn_hi & ehi));
    u32 wtot;
    const u32 ex_in_warp = warp_excl_scan(cnt, lane, wtot);
    const u32 w2w = warp_sum(w2);
    if (lane == 0) { s_wtot[warp] = wtot; if (w2w) atomicAdd(&s_w2, w2w); }
    __syncthreads();
    /* line ends in the warps before this one, and in the tile: every warp scans the 16 warp totals in its lanes */
    u32 btot;
    const u32 wscan = warp_excl_scan(lane < IDX_THREADS / 32 ? s_wtot[lane] : 0u, lane, btot);
    const u32 wpre = __shfl_sync(0xffffffffu, wscan, warp);

    if (tid == 0) {
        tile_count[tile] = btot;
        if (btot > lcap) atomicOr(&ctr->overflow, 1u);
        if (s_w2) atomicAdd(&ctr->n_w2, s_w2);
    }

    /* ---- positions, in text order: the last byte of every break, into the tile's region */
    if (cnt) {
        u32 o = wpre + ex_in_warp;
        u32* out = nl_local + (size_t)tile * lcap;
        const u32 base = (u32)p0;
        /* 32 positions at a time: a line end every ~90 bytes, one to three per row */
#pragma unroll
        for (int q = 0; q < 4; q++) {
            u32 m = q == 0 ? (u32)elo : q == 1 ? (u32)(elo >> 32) : q == 2 ? (u32)ehi : (u32)(ehi >> 32);
            const u32 sw = q == 0 ? (u32)swn_lo : q == 1 ? (u32)(swn_lo >> 32) : q == 2 ? (u32)swn_hi : (u32)(swn_hi >> 32);
            while (m) { const int bb = __ffs((int)m) - 1; m &= m - 1u; if (o < lcap) out[o] = base + 32u * (u32)q + (u32)bb + ((sw >> bb) & 1u); o++; }
        }
    }
}

/* ranks of the tiles' first lines: exclusive scan of the tile counts, one CTA (a thread takes a contiguous piece) */
constexpr int IDX_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(IDX_SCAN_THREADS) k_index_scan(const u32* __restrict__ tile_count, u32 tiles, u32* __restrict__ tile_first, IndexCounters* ctr) {
    __shared__ u32 s_w[IDX_SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 per = (tiles + IDX_SCAN_THREADS - 1) / IDX_SCAN_THREADS;
    const u32 a = (u32)tid * per, z = a + per < tiles ? a + per : tiles;
    u32 sum = 0;
    for (u32 k = a; k < z; k++) sum += tile_count[k];
    u32 wtot; const u32 ex = warp_excl_scan(sum, lane, wtot);
    if (lane == 0) s_w[warp] = wtot;
    __syncthreads();
    u32 run = ex, total = 0;
#pragma unroll
    for (int w = 0; w < IDX_SCAN_THREADS / 32; w++) { const u32 t = s_w[w]; if (w < warp) run += t; total += t; }
    for (u32 k = a; k < z; k++) { tile_first[k] = run; run += tile_count[k]; }
    if (tid == 0) { tile_first[tiles] = total; ctr->n_nl = total; }
}

/* the tile-local entries to their ranks: a CTA per tile */
__global__ void __launch_bounds__(256) k_index_compact(const u32* __restrict__ nl_local, u32 lcap, const u32* __restrict__ tile_count, const u32* __restrict__ tile_first,
                                                      u32* __restrict__ nl, u32 nl_cap) {
    const u32 tile = blockIdx.x;
    const u32 n = tile_count[tile] < lcap ? tile_count[tile] : lcap, first = tile_first[tile];
    const u32* src = nl_local + (size_t)tile * lcap;
    for (u32 k = threadIdx.x; k < n; k += blockDim.x) if (first + k < nl_cap) nl[first + k] = src[k];
}

/* is the break that ends a line at text offset e (its last byte) two bytes long?  Only a swallowed '\n' can follow the byte that
 * ended the line, and a line holds no break characters itself. */
__device__ __forceinline__ bool brk_is_two(const u8* text, u32 line_start, u32 e) { return text[e] == '\n' && e > line_start && is_brk(text[e - 1]); }

/* line-end mode, the virtual final newline, line count.  eof: the text ends where the input ends, so an unterminated last line
 * is a line (FastqReader::getLine returns it, src/fastqreader.cpp:107-120); otherwise the text is a window of a longer input:
 * the unterminated tail belongs to a record of the next window, and so does a line whose one-byte break is the window's last
 * byte (whether a '\n' after it belongs to the break is for the next window to see). */
__global__ void k_index_finish(const u8* text, u64 len, u32* nl, u32 nl_cap, IndexCounters* ctr, int eof) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u32 n = ctr->n_nl, w2 = ctr->n_w2;
    if (n && n <= nl_cap && nl[n - 1] == (u32)(len - 1)) {
        const u32 start = n > 1 ? nl[n - 2] + 1u : 0u;
        const bool two = brk_is_two(text, start, (u32)(len - 1));
        if (!eof && !two) n--;
        else if (eof && two && w2 == 1 && text[len - 2] == '\n' && n < nl_cap) {
            /* a text of one-byte breaks that ends "\n\n": the reference's reader never takes the LAST byte of its input as the
             * second byte of a break (end < mBufDataLen - 1), it reads an empty last line - which changes nothing, but keeps the
             * text regular */
            nl[n - 1] = (u32)(len - 2); nl[n] = (u32)(len - 1); n++; w2 = 0;
        }
    }
    const u32 crlf = (n && w2 == n) ? 1u : 0u;
    u32 lines = n;
    if (eof && len > 0 && !is_brk(text[len - 1])) {
        if (n < nl_cap) nl[n] = (u32)len + crlf;       /* so that line_end() == len */
        lines = n + 1;
    }
    ctr->n_nl = n; ctr->n_w2 = w2;
    ctr->n_lines = lines; ctr->crlf = crlf; ctr->irregular = (w2 != 0 && w2 != n) ? 1u : 0u;
}

/* ---- irregular texts (line breaks of one and of two bytes: blank lines the reader swallows, a "\r\n" on a reader-buffer edge,
 * mixed line ends): the lines are copied into a text of plain "\n" breaks, exactly the lines the reference's reader delivers,
 * and everything after the index works on that copy. */
__global__ void k_canon_lens(const u8* __restrict__ text, u64 len, const u32* __restrict__ nl, u32 n_lines, u32 n_brk, u32* __restrict__ out_len) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lines) return;
    const u32 start = j ? nl[j - 1] + 1u : 0u;
    u32 end;
    if (j < n_brk) { const u32 e = nl[j]; end = e - (brk_is_two(text, start, e) ? 1u : 0u); }
    else end = (u32)len;                               /* the unterminated last line */
    out_len[j] = end - start + 1u;                     /* content + '\n' */
}
__global__ void k_canon_copy(const u8* __restrict__ text, const u32* __restrict__ nl, const u32* __restrict__ lens, const u64* __restrict__ pre, u32 n_lines,
                             u8* __restrict__ out, u32* __restrict__ nl_out) {
    const int lane = threadIdx.x & 31;
    const u32 j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_lines) return;
    const u32 start = j ? nl[j - 1] + 1u : 0u;
    const u32 n = lens[j] - 1u;
    const u64 dst = pre[j] - lens[j];
    for (u32 k = lane; k < n; k += 32) out[dst + k] = text[start + k];
    if (lane == 0) { out[dst + n] = '\n'; nl_out[j] = (u32)(dst + n); }
}

/* per-unit (read, or pair) statistics */
struct UnitStats {
    u32 first_empty;      /* first unit with an empty line: input ends there (src/fastqreader.cpp:180-181,190-191) */
    u32 first_qual_len;   /* first unit whose quality line is shorter than its sequence */
    u32 first_name_len;   /* first unit with a name/strand line > 255 bytes */
    u32 first_read_len;   /* first unit with a read > 65535 bases */
    u32 min_bases, max_bases;
    u32 max_read;         /* longest single read */
    u32 max_head;         /* most bytes from the start of a name line to the start of the quality line */
    u32 n_chunks;         /* written by k_cut */
    u32 units_in_chunks;  /* units covered by the emitted chunks */
    u32 end[2];           /* written by k_cut_ends: offset of the line break that ends the last covered record, per file */
    u32 reach[2];         /* written by k_cut_ends (final batches): how far the reference's reader has read in each file when its loop ends */
};

__global__ void k_unit_lengths(EncBatchDev b, u32 n_units, u32* __restrict__ rlen, u32* __restrict__ unit_bases, UnitStats* st) {
    const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = u < n_units;
    const u32 per = b.is_pe ? 2u : 1u;
    u32 bases = 0;
    bool empty = false, badq = false, badn = false, badl = false;
    u32 longest = 0, head = 0;
    if (live) {
        for (u32 k = 0; k < per; k++) {
            const u32 i = u * per + k;
            u32 f, rec; read_locus(b, i, f, rec);
            const TextDev& t = b.t[f];
            uint4 lc;
            lc.x = line_start(t, 4 * rec); lc.y = line_start(t, 4 * rec + 1); lc.z = line_start(t, 4 * rec + 2); lc.w = line_start(t, 4 * rec + 3);
            const u32 e3 = line_end(t, 4 * rec + 3);
            /* a line ends where the next one starts, minus its break */
            const u32 brk = 1u + t.crlf;
            const u32 ln0 = lc.y - brk - lc.x, ln1 = lc.z - brk - lc.y, ln2 = lc.w - brk - lc.z, ln3 = e3 - lc.w;
            if (!ln0 || !ln1 || !ln2 || !ln3) empty = true;
            if (ln3 < ln1) badq = true;                       /* a longer quality line is cut to the sequence's length, as the reference does (src/rfqcodec.cpp:332-407 copies seq.length() bytes); a shorter one makes it read past its string */
            if (ln0 > 255 || ln2 > 255) badn = true;
            if (ln1 > 65535) badl = true;
            rlen[i] = ln1;
            b.loc[i] = lc;
            head = lc.w - lc.x > head ? lc.w - lc.x : head;
            longest = ln1 > longest ? ln1 : longest;
            bases += ln1;
        }
        unit_bases[u] = bases;
    }
    /* one request per CTA and statistic: all CTAs update the same eight words, so even a filtering load per warp queues up at one
     * L2 slice (1.2 M requests per 3.4 GB batch).  Warp reduce, shared-memory atomics, then eight threads compare with the word in
     * memory and only issue a global atomic that would change it (on uniform reads only the first CTAs do). */
    __shared__ u32 s_stat[8];
    const u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 8) s_stat[threadIdx.x] = threadIdx.x < 5 ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    const u32 w_empty = __reduce_min_sync(FULL, empty ? u : 0xFFFFFFFFu), w_badq = __reduce_min_sync(FULL, badq ? u : 0xFFFFFFFFu);
    const u32 w_badn = __reduce_min_sync(FULL, badn ? u : 0xFFFFFFFFu), w_badl = __reduce_min_sync(FULL, badl ? u : 0xFFFFFFFFu);
    const u32 w_min = __reduce_min_sync(FULL, live ? bases : 0xFFFFFFFFu), w_max = __reduce_max_sync(FULL, bases);
    const u32 w_long = __reduce_max_sync(FULL, longest), w_head = __reduce_max_sync(FULL, head);
    if (lane == 0) {
        if (w_empty != 0xFFFFFFFFu) atomicMin(&s_stat[0], w_empty);
        if (w_badq != 0xFFFFFFFFu) atomicMin(&s_stat[1], w_badq);
        if (w_badn != 0xFFFFFFFFu) atomicMin(&s_stat[2], w_badn);
        if (w_badl != 0xFFFFFFFFu) atomicMin(&s_stat[3], w_badl);
        atomicMin(&s_stat[4], w_min);
        atomicMax(&s_stat[5], w_max);
        atomicMax(&s_stat[6], w_long);
        atomicMax(&s_stat[7], w_head);
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        /* UnitStats starts with first_empty, first_qual_len, first_name_len, first_read_len, min_bases, max_bases, max_read, max_head */
        u32* word = reinterpret_cast<u32*>(st) + threadIdx.x;
        const u32 v = s_stat[threadIdx.x], cur = *reinterpret_cast<volatile u32*>(word);
        if (threadIdx.x < 5) { if (v < cur) atomicMin(word, v); }
        else if (v > cur) atomicMax(word, v);
    }
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

/* where the text covered by the chunks ends (one launch behind k_cut, so that the host reads everything back at once) */
/*
 * Where the reference's reader stands when its loop ends (final batches only).  After the last good record it tries one more
 * (FastqReader::read, src/fastqreader.cpp:166-196): name, sequence and strand line are read whatever they hold, the quality line
 * only if none of the three is empty; a line that is not there is the end of the file.  What matters is the first break character
 * of the last line it read: the buffer holding it has been loaded, and loading the file's last, short buffer raises
 * NO_LINE_BREAK_AT_END when the file does not end in a line feed (:31-46) - which the post-loop flush then writes into its
 * chunk (Q13).  An input that ends at an empty line, at a ragged record or because the mate file is shorter can so flag a chunk
 * whose own last record ends before that buffer.
 */
__device__ inline u32 reader_attempt(const TextDev& t, u32 rec, u32 caller_len, bool& complete) {
    const u32 L = 4u * rec;
    complete = false;
    if (L + 2u >= t.n_lines) return caller_len;
    bool empty3 = false;
    for (u32 k = 0; k < 3u; k++) if (line_end(t, L + k) == line_start(t, L + k)) empty3 = true;
    if (empty3) return caller_break_first(t, L + 2u);
    if (L + 3u >= t.n_lines) return caller_len;
    complete = line_end(t, L + 3u) != line_start(t, L + 3u);
    return caller_break_first(t, L + 3u);
}

__global__ void k_cut_ends(UnitStats* st, EncBatchDev b, int two, int pe, u32 n_units, int final, u32 len0, u32 len1) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    st->end[0] = st->end[1] = 0;
    st->reach[0] = st->reach[1] = 0;
    if (final) {
        /* unit n_units is the one the reader fails on (an empty line, a ragged or missing record) */
        bool ok;
        if (two) { st->reach[0] = reader_attempt(b.t[0], n_units, len0, ok); st->reach[1] = reader_attempt(b.t[1], n_units, len1, ok); }
        else if (pe) { u32 r = reader_attempt(b.t[0], 2u * n_units, len0, ok); if (ok) r = reader_attempt(b.t[0], 2u * n_units + 1u, len0, ok); st->reach[0] = r; }
        else st->reach[0] = reader_attempt(b.t[0], n_units, len0, ok);
    }
    if (st->units_in_chunks == 0) return;
    const u32 last_unit = st->units_in_chunks - 1;
    const u32 rec0 = two ? last_unit : (pe ? 2 * last_unit + 1 : last_unit);
    st->end[0] = caller_break_last(b.t[0], 4 * rec0 + 3);
    if (two) st->end[1] = caller_break_last(b.t[1], 4 * last_unit + 3);
}

}  // namespace rpq
