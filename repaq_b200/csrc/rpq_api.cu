/*
 * rpq_api.cu - the C ABI of include/repaq_b200.h: context, device memory, and the launch sequences of the encode and
 * decode paths.  Everything between the input text and the serialised chunks happens on the GPU; the host only sizes
 * buffers from a handful of scalars read back between stages.
 */
#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "rpq_common.cuh"
#include "rpq_index.cuh"
#include "rpq_encode.cuh"
#include "rpq_streams2.cuh"
#include "rpq_streams3.cuh"
#include "rpq_streams4.cuh"
#include "rpq_streams7.cuh"
#include "rpq_meta2.cuh"
#include "rpq_meta3.cuh"
#include "rpq_decode.cuh"
#include "rpq_decode2.cuh"
#include "rpq_decode4.cuh"
#include "rpq_compare.cuh"
#include "rpq_host.h"

using namespace rpq;

namespace rpq {
/* Small device tables and scalars go to (mapped) pinned host memory with plain stores from an SM instead of a copy-engine
 * transfer: a copy-engine transfer queues behind whatever bulk copies other contexts have in flight in the same direction
 * (a decoder returning gigabytes of FASTQ beside this encoder), and the host waits for every one of these read-backs. */
__global__ void __launch_bounds__(256) k_fetch(const u32* __restrict__ src, u32* __restrict__ dst, u32 nwords) {
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < nwords; k += gridDim.x * blockDim.x) dst[k] = src[k];
}
}  // namespace rpq

namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct rpq_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    char err[768] = {0};
    rpq_header hdr;
    bool have_hdr = false;
    HeaderDev hd;
    /* grow-only device buffers */
    DevBuf loc, pk, pk_rc, text[2], nl[2], nl_local, tile_state, counters, rlen, unit_bases, prefix, scan_tmp, ustats, chunk_first, chunks, meta, meta0, ov,
        seqoff, qualoff, n1off, n2off, soff, errbits, tmpx, tmpy, span_first[2], span_chunk[2], dir[2], slots[2], span_slot[2], span_read0[2], redo_list[2], dense_list, unclean, misc, out,
        d_in, d_desc, d_tmp[8], d_tmp2, out2, d_slabs, d_ckpt, d_dir, canon[2], nl2[2], canon_len, canon_pre;
    int streams5 = 1;                      /* RPQ_DEBUG_STREAMS5=0: every span k_streams4 cannot code goes to k_streams3 (A/B, at most 46 streams); =2: k_streams7 codes every quality span (test coverage) */
    bool no_streams4 = false;              /* RPQ_DEBUG_NO_STREAMS4=1: k_streams3 codes every span (test coverage, A/B) */
    bool dense_hint = false;               /* most quality spans of the last batch were dense: the next one goes to k_streams7 directly (k_streams4 would stage
                                              and count every span only to hand it over) */
    u64 stats_dense_spans = 0;             /* spans k_streams4 handed to k_streams7 (dense ones and those with long runs) */
    u64 stats_redo_spans = 0;              /* spans k_streams4 handed to k_streams3 */
    u32 fmt_reads = 0;                     /* RPQ_DEBUG_FMT_READS=n: reads per formatter CTA (tuning experiments) */
    bool no_par_walk = false;              /* RPQ_DEBUG_NO_PAR_WALK=1: the chunk chain of a device-resident body is followed by one warp (A/B) */
    u64 par_walk_min = 0;                  /* RPQ_DEBUG_PAR_WALK_MIN=<bytes>: smallest body the parallel walk is used for (tests) */
    bool force_v1 = false;                 /* RPQ_DEBUG_FORCE_V1=1: take the long-read fallback kernels (test coverage) */
    u32 d2h_depth = 1;                     /* RPQ_D2H_DEPTH=n: windows of decoded FASTQ whose copies may be queued at a time (0: wait for each) */
    bool no_pipeline = false;              /* RPQ_NO_PIPELINE=1: host batches are never cut into pipelined windows */
    uint64_t pipe_window = 0;              /* RPQ_DEBUG_PIPE_WINDOW=<bytes>: window size of the pipelined host path (tests) */
    std::vector<rpq_ctx*> lanes;           /* sub-contexts (own stream + buffers) of the pipelined host path */
    cudaStream_t copy_stream = 0;          /* decode: device-to-host copies of finished windows */
    bool copy_stream_ok = false;
    cudaStream_t side_stream = 0;          /* decode: the stream checkpoints (k_dec_qindex) beside the table kernels */
    bool side_stream_ok = false;
    RtEvent side_ev, fork_ev;              /* work on the side stream done; the point of the main stream it may start from */
    RtEvent win_ev[16], cp_ev[16];
    void* pinned_small = nullptr;          /* 64 KiB scratch for scalar read-backs */
    DevBuf host_out, host_out2;            /* pinned host buffers for results */
    DevBuf pinned_tab;                     /* pinned host copy of the chunk table (k_fetch) */
    std::vector<rpq_chunk_info> infos;
    std::vector<ChunkDev> h_chunks;
    DecBatchDev last_dec;                  /* tables and outputs of the last rpq_decode (rpq_compare reads them) */
    rpq_stats stats;
    RtEvent ev[8];
    uint32_t launches = 0;
    double slot_factor = 0.35;
    /* optional per-kernel timing (rpq_set_profiling): an event pair around every launch */
    bool profiling = false;
    struct ProfRec { const char* name; RtEvent a, b; };
    std::vector<ProfRec> prof;
    std::string prof_text;             /* slot bytes per position, grown on overflow (worst case 5) */
};

namespace {

int fail(rpq_ctx* c, int code, const std::string& msg) { snprintf(c->err, sizeof c->err, "%s", msg.c_str()); return code; }

bool ensure(rpq_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return true;
    if (b.p) { rt_stream_sync(c->stream); if (c->side_stream_ok) rt_stream_sync(c->side_stream); rt_free_device(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    b.p = rt_malloc_device(want);
    if (!b.p) return false;
    b.cap = want;
    return true;
}
bool ensure_pinned(rpq_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return true;
    if (b.p) { rt_stream_sync(c->stream); rt_free_pinned(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 4096;
    b.p = rt_malloc_pinned(want);
    if (!b.p) return false;
    b.cap = want;
    return true;
}

inline void prof_begin(rpq_ctx* c, const char* name) {
    if (!c->profiling) return;
    rpq_ctx::ProfRec r; r.name = name; rt_event_create(&r.a); rt_event_create(&r.b);
    rt_event_record(&r.a, c->stream);
    c->prof.push_back(r);
}
inline void prof_end(rpq_ctx* c) { if (c->profiling) rt_event_record(&c->prof.back().b, c->stream); }

#define LAUNCH(ctx, kern, grid, block, smem, ...)                                           \
    do { if ((grid) > 0) { prof_begin(ctx, #kern); RPQ_LAUNCH(kern, grid, block, smem, (ctx)->stream, __VA_ARGS__); prof_end(ctx); (ctx)->launches++; } } while (0)

/* the context's second stream (created on first use): decode fills the quality plane on it, encode runs kernels there that do
 * not depend on what the main stream is busy with */
bool side_ready(rpq_ctx* c) {
    if (!c->side_stream_ok && !rt_stream_create(&c->side_stream)) { c->side_stream_ok = true; rt_event_create(&c->side_ev); rt_event_create(&c->fork_ev); }
    return c->side_stream_ok;
}

int check_launch(rpq_ctx* c, const char* where) {
    char buf[256];
    if (rt_last_error(buf, sizeof buf)) return fail(c, RPQ_ERR_CUDA, std::string(where) + ": " + buf);
    return RPQ_OK;
}

void build_header_dev(const rpq_header& h, HeaderDev& d) {
    memset(&d, 0, sizeof d);
    d.flags = h.flags; d.read_length_bytes = h.read_length_bytes; d.name2_diff_pos = h.name2_diff_pos; d.name2_diff_char = h.name2_diff_char;
    d.n_base_qual = h.n_base_qual; d.overlap_shift = h.overlap_shift; d.support_interleaved = h.support_interleaved;
    d.major = h.qual_buf[0];
    uint8_t nbuf[130];
    int nb = host_normal_bins(&h, nbuf);
    if (nb > MAX_BINS) nb = MAX_BINS;              /* only reachable with DONT_ENCODE_QUAL, where streams are not used */
    d.nb = (u8)nb;
    for (int i = 0; i < 256; i++) d.lut[i] = (i == d.major) ? LUT_SKIP : LUT_EXC;
    for (int i = nb - 1; i >= 0; i--) { d.normal_bins[i] = nbuf[i]; d.lut[nbuf[i]] = (u8)i; }
    /* the tables of the quality run-length coder (decode only, row a10): makeQualBitTable / computeNormalQualBits */
    for (int i = 0; i < h.qual_bins; i++) { const int bit = i > 0 ? 2 * i - 1 : 0; if (bit < 128) d.rle_b2q[bit] = h.qual_buf[i < 128 ? i : 127]; }
    { int mx = (int)h.qual_bins * 2 - 3; if (mx < 1) mx = 1; d.rle_nq_bits = (u8)(mx >= 64 ? 1 : mx >= 32 ? 2 : mx >= 16 ? 3 : mx >= 8 ? 4 : mx >= 4 ? 5 : mx >= 2 ? 6 : 7); }
}

}  // namespace

/* ================================================================== header (host) ==== */
extern "C" int rpq_make_header(const char* r1, uint64_t r1_len, const char* r2, uint64_t r2_len, int interleaved, uint32_t chunk_bases,
                               rpq_header* out, char* err, size_t err_cap) {
    if (!out) return RPQ_ERR_ARG;
    if (!r1 || !r1_len) { if (err && err_cap) snprintf(err, err_cap, "the input holds no FASTQ record"); return RPQ_NO_RECORDS; }
    return host_make_header(r1, r1_len, r2, r2_len, interleaved, chunk_bases, out, err, err_cap);
}
extern "C" size_t rpq_header_write(const rpq_header* h, uint8_t* out, size_t cap) { return host_header_write(h, out, cap); }
extern "C" int rpq_header_read(const uint8_t* in, size_t len, rpq_header* out, size_t* consumed, char* err, size_t err_cap) {
    return host_header_read(in, len, out, consumed, err, err_cap);
}

/* ================================================================== context ==== */
extern "C" int rpq_create(int device, rpq_ctx** out) {
    if (!out) return RPQ_ERR_ARG;
    *out = nullptr;
    if (rt_device_count() <= device || device < 0) return RPQ_ERR_CUDA;     /* no CPU fallback */
    if (rt_set_device(device)) return RPQ_ERR_CUDA;
    rpq_ctx* c = new rpq_ctx();
    c->device = device;
    if (rt_stream_create(&c->stream)) { delete c; return RPQ_ERR_CUDA; }
    c->pinned_small = rt_malloc_pinned(65536);
    if (!c->pinned_small) { rt_stream_destroy(c->stream); delete c; return RPQ_ERR_NOMEM; }      /* every scalar read-back goes through it */
    for (auto& e : c->ev) rt_event_create(&e);
    memset(&c->stats, 0, sizeof c->stats);
    memset(&c->hdr, 0, sizeof c->hdr);
    memset(&c->last_dec, 0, sizeof c->last_dec);
    { const char* e = getenv("RPQ_DEBUG_FORCE_V1"); c->force_v1 = e && e[0] == '1'; }
    { const char* e = getenv("RPQ_DEBUG_NO_PAR_WALK"); c->no_par_walk = e && e[0] == '1'; }
    { const char* e = getenv("RPQ_DEBUG_PAR_WALK_MIN"); c->par_walk_min = e ? strtoull(e, nullptr, 10) : (32ull << 20); }
    { const char* e = getenv("RPQ_DEBUG_STREAMS5"); if (e) c->streams5 = atoi(e); }
    { const char* e = getenv("RPQ_DEBUG_NO_STREAMS4"); c->no_streams4 = e && e[0] == '1'; }
    { const char* e = getenv("RPQ_DEBUG_FMT_READS"); c->fmt_reads = e ? (u32)atoi(e) : 0u; }
    { const char* e = getenv("RPQ_NO_PIPELINE"); c->no_pipeline = e && e[0] == '1'; }
    { const char* e = getenv("RPQ_D2H_DEPTH"); if (e) c->d2h_depth = (u32)atoi(e) > 8u ? 8u : (u32)atoi(e); }
    { const char* e = getenv("RPQ_DEBUG_PIPE_WINDOW"); c->pipe_window = e ? strtoull(e, nullptr, 10) : 0; }
#ifndef RPQ_EMU
    cudaFuncSetAttribute(k_index_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, IDX_SMEM);
    cudaFuncSetAttribute(k_streams2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_streams3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_streams4, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_streams7, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_meta3, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_dec_format4, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
    cudaFuncSetAttribute(k_dec_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
#endif
    *out = c;
    return RPQ_OK;
}

extern "C" void rpq_destroy(rpq_ctx* c) {
    if (!c) return;
    rt_set_device(c->device);
    rt_stream_sync(c->stream);
    for (rpq_ctx* l : c->lanes) rpq_destroy(l);
    c->lanes.clear();
    if (c->copy_stream_ok) { rt_stream_sync(c->copy_stream); rt_stream_destroy(c->copy_stream); for (auto& e : c->win_ev) rt_event_destroy(&e); for (auto& e : c->cp_ev) rt_event_destroy(&e); }
    if (c->side_stream_ok) { rt_stream_sync(c->side_stream); rt_stream_destroy(c->side_stream); rt_event_destroy(&c->side_ev); rt_event_destroy(&c->fork_ev); }
    DevBuf* all[] = {&c->loc, &c->pk, &c->pk_rc, &c->text[0], &c->text[1], &c->nl[0], &c->nl[1], &c->nl_local, &c->tile_state, &c->counters, &c->rlen, &c->unit_bases, &c->prefix, &c->scan_tmp,
                     &c->ustats, &c->chunk_first, &c->chunks, &c->meta, &c->meta0, &c->ov, &c->seqoff, &c->qualoff, &c->n1off, &c->n2off, &c->soff,
                     &c->errbits, &c->tmpx, &c->tmpy, &c->span_first[0], &c->span_first[1], &c->span_chunk[0], &c->span_chunk[1], &c->dir[0], &c->dir[1],
                     &c->slots[0], &c->slots[1], &c->span_slot[0], &c->span_slot[1], &c->span_read0[0], &c->span_read0[1], &c->redo_list[0], &c->redo_list[1], &c->dense_list, &c->unclean, &c->misc, &c->out, &c->d_in, &c->d_desc, &c->out2, &c->d_slabs, &c->d_ckpt, &c->d_dir, &c->canon[0], &c->canon[1], &c->nl2[0], &c->nl2[1], &c->canon_len, &c->canon_pre,
                     &c->d_tmp2, &c->d_tmp[0], &c->d_tmp[1], &c->d_tmp[2], &c->d_tmp[3], &c->d_tmp[4], &c->d_tmp[5], &c->d_tmp[6], &c->d_tmp[7]};
    for (DevBuf* b : all) rt_free_device(b->p);
    rt_free_pinned(c->pinned_small);
    rt_free_pinned(c->host_out.p); rt_free_pinned(c->host_out2.p); rt_free_pinned(c->pinned_tab.p);
    for (auto& e : c->ev) rt_event_destroy(&e);
    rt_stream_destroy(c->stream);
    delete c;
}

extern "C" const char* rpq_last_error(const rpq_ctx* c) { return c ? c->err : "no context (is a CUDA device present?)"; }
extern "C" void* rpq_stream(rpq_ctx* c) { return c ? (void*)(uintptr_t)c->stream : nullptr; }
extern "C" int rpq_get_stats(const rpq_ctx* c, rpq_stats* out) { if (!c || !out) return RPQ_ERR_ARG; *out = c->stats; return RPQ_OK; }

extern "C" int rpq_set_profiling(rpq_ctx* c, int on) {
    if (!c) return RPQ_ERR_ARG;
    rt_stream_sync(c->stream);
    for (auto& r : c->prof) { rt_event_destroy(&r.a); rt_event_destroy(&r.b); }
    c->prof.clear();
    c->profiling = on != 0;
    return RPQ_OK;
}

/* "kernel launches total_ms\n" per kernel, accumulated since rpq_set_profiling(ctx, 1) */
extern "C" const char* rpq_get_profile(rpq_ctx* c) {
    if (!c) return "";
    rt_stream_sync(c->stream);
    std::vector<std::string> names; std::vector<double> ms; std::vector<int> cnt;
    for (auto& r : c->prof) {
        size_t k = 0;
        for (; k < names.size(); k++) if (names[k] == r.name) break;
        if (k == names.size()) { names.push_back(r.name); ms.push_back(0); cnt.push_back(0); }
        ms[k] += rt_event_ms(r.a, r.b); cnt[k]++;
    }
    c->prof_text.clear();
    char line[256];
    for (size_t k = 0; k < names.size(); k++) { snprintf(line, sizeof line, "%s %d %.6f\n", names[k].c_str(), cnt[k], ms[k]); c->prof_text += line; }
    return c->prof_text.c_str();
}

extern "C" void* rpq_host_alloc(size_t bytes) { return rt_malloc_pinned(bytes); }
extern "C" void rpq_host_free(void* p) { rt_free_pinned(p); }

extern "C" int rpq_set_header(rpq_ctx* c, const rpq_header* h) {
    if (!c || !h) return RPQ_ERR_ARG;
    if (h->read_length_bytes != 1 && h->read_length_bytes != 2 && h->read_length_bytes != 4)
        return fail(c, RPQ_ERR_HEADER, "header incorrect: read length bytes should be 1/2/4");
    c->hdr = *h; c->have_hdr = true;
    build_header_dev(*h, c->hd);
    return RPQ_OK;
}

/* ================================================================== encode ==== */
namespace {

/* scalar read-back through the pinned scratch; synchronises the stream */
template <class T> int read_back(rpq_ctx* c, const void* dev, T* host, size_t count = 1) {
    static_assert(sizeof(T) % 4 == 0, "read_back moves 32-bit words");
    RPQ_LAUNCH(k_fetch, 1, 64, 0, c->stream, (const u32*)dev, (u32*)c->pinned_small, (u32)(sizeof(T) * count / 4));
    if (rt_stream_sync(c->stream)) return RPQ_ERR_CUDA;
    memcpy(host, c->pinned_small, sizeof(T) * count);
    return RPQ_OK;
}

/* bulk device -> pinned host on `stream` */
int push_to_host(rpq_ctx* c, void* host, const void* dev, size_t bytes, cudaStream_t stream) {
    (void)c;
    if (!bytes) return 0;
    return rt_memcpy_d2h(host, dev, bytes, stream);
}

/* queues the copy of a device table (a multiple of 4 bytes) into c->pinned_tab; valid after the stream is synchronised */
int fetch_table(rpq_ctx* c, const void* dev, size_t bytes) {
    if (!ensure_pinned(c, c->pinned_tab, bytes + 64)) return RPQ_ERR_NOMEM;
    const u32 nwords = (u32)(bytes / 4);
    if (nwords) RPQ_LAUNCH(k_fetch, (nwords + 1023) / 1024 < 64 ? (nwords + 1023) / 1024 : 64, 256, 0, c->stream, (const u32*)dev, c->pinned_tab.as<u32>(), nwords);
    return RPQ_OK;
}

int index_text(rpq_ctx* c, int f, const u8* d_text, u64 len, u64 file_off, IndexCounters* hc, bool eof) {
    const u32 tiles = (u32)((len + IDX_TILE - 1) / IDX_TILE);
    if (!ensure(c, c->tile_state, sizeof(u32) * 2 * ((size_t)tiles + 1)) || !ensure(c, c->counters, sizeof(IndexCounters) * 2)) return RPQ_ERR_NOMEM;
    u32* tile_count = c->tile_state.as<u32>();
    u32* tile_first = tile_count + tiles + 1;
    /* a tile's region of the tile-local index: lines of 16 bytes and more fit (IDX_TILE / 16 entries); a text with shorter lines
     * is indexed again with room for a line end at every byte */
    size_t cap = (size_t)(len / 16) + 4096;
    u32 lcap = IDX_TILE / 16;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!ensure(c, c->nl[f], sizeof(u32) * cap) || !ensure(c, c->nl_local, sizeof(u32) * (size_t)lcap * tiles + 64)) return RPQ_ERR_NOMEM;
        cap = c->nl[f].cap / sizeof(u32);
        IndexCounters* dc = c->counters.as<IndexCounters>() + f;
        rt_memset(dc, 0, sizeof(IndexCounters), c->stream);
        if (tiles) {
            LAUNCH(c, k_index_lines, tiles, IDX_THREADS, IDX_SMEM, d_text, len, file_off, c->nl_local.as<u32>(), lcap, tile_count, dc);
            LAUNCH(c, k_index_scan, 1, IDX_SCAN_THREADS, 0, (const u32*)tile_count, tiles, tile_first, dc);
            LAUNCH(c, k_index_compact, tiles, 256, 0, (const u32*)c->nl_local.as<u32>(), lcap, (const u32*)tile_count, (const u32*)tile_first, c->nl[f].as<u32>(), (u32)cap);
        }
        LAUNCH(c, k_index_finish, 1, 32, 0, d_text, len, c->nl[f].as<u32>(), (u32)cap, dc, eof ? 1 : 0);
        if (int rc = read_back(c, dc, hc)) return rc;
        if ((size_t)hc->n_nl + 2 <= cap && !hc->overflow) return RPQ_OK;
        cap = (size_t)hc->n_nl + 16;           /* pathological line density: index again with room for every line */
        lcap = IDX_TILE;
    }
    return RPQ_OK;
}

/* A text with line breaks of one and of two bytes (blank lines the reference's reader swallows, a "\r\n" on an edge of its
 * buffer, mixed line ends): the lines the reader delivers are copied into a text of plain '\n' breaks, with its own index;
 * *text / *len are replaced.  The original index stays in c->nl[f] (offsets reported to the caller come from it). */
int canon_text(rpq_ctx* c, int f, const u8** text, u64* len, IndexCounters* hc) {
    const u32 n = hc->n_lines;
    if (!n) return RPQ_OK;
    if (!ensure(c, c->canon_len, 4 * (size_t)n) || !ensure(c, c->canon_pre, 8 * (size_t)n) || !ensure(c, c->nl2[f], 4 * (size_t)n + 64) ||
        !ensure(c, c->scan_tmp, rt_scan_tmp_bytes(n)))
        return RPQ_ERR_NOMEM;
    LAUNCH(c, k_canon_lens, (n + 255) / 256, 256, 0, *text, *len, (const u32*)c->nl[f].as<u32>(), n, hc->n_nl, c->canon_len.as<u32>());
    rt_inclusive_sum_u32_u64(c->canon_len.as<u32>(), c->canon_pre.as<u64>(), n, c->scan_tmp.p, c->scan_tmp.cap, c->stream);
    u64 total = 0;
    if (int rc = read_back(c, c->canon_pre.as<u64>() + (n - 1), &total)) return rc;
    if (!ensure(c, c->canon[f], total + 64)) return RPQ_ERR_NOMEM;
    LAUNCH(c, k_canon_copy, (n + 7) / 8, 256, 0, *text, (const u32*)c->nl[f].as<u32>(), (const u32*)c->canon_len.as<u32>(), (const u64*)c->canon_pre.as<u64>(), n,
           c->canon[f].as<u8>(), c->nl2[f].as<u32>());
    *text = c->canon[f].as<u8>(); *len = total;
    hc->n_nl = n; hc->n_w2 = 0; hc->crlf = 0;
    return RPQ_OK;
}

}  // namespace

#include "rpq_api_encode.inc"
#include "rpq_api_decode.inc"
#include "rpq_api_compare.inc"
