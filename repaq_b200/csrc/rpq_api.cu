/*
 * rpq_api.cu - the C ABI of include/repaq_b200.h: context, device memory, and the launch sequences of the encode and
 * decode paths.  Everything between the input text and the serialised chunks happens on the GPU; the host only sizes
 * buffers from a handful of scalars read back between stages.
 */
#include <algorithm>
#include <string>
#include <vector>

#include "rpq_common.cuh"
#include "rpq_index.cuh"
#include "rpq_encode.cuh"
#include "rpq_streams2.cuh"
#include "rpq_streams3.cuh"
#include "rpq_meta2.cuh"
#include "rpq_meta3.cuh"
#include "rpq_decode.cuh"
#include "rpq_decode2.cuh"
#include "rpq_decode3.cuh"
#include "rpq_host.h"

using namespace rpq;

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
    DevBuf loc, pk, pk_rc, text[2], nl[2], tile_state, counters, rlen, unit_bases, prefix, scan_tmp, ustats, chunk_first, chunks, meta, meta0, ov,
        seqoff, qualoff, n1off, n2off, soff, errbits, tmpx, tmpy, span_first[2], span_chunk[2], dir[2], slots[2], span_slot[2], misc, out,
        d_in, d_desc, d_tmp[8], d_tmp2, out2;
    bool force_v1 = false;                 /* RPQ_DEBUG_FORCE_V1=1: take the long-read fallback kernels (test coverage) */
    void* pinned_small = nullptr;          /* 64 KiB scratch for scalar read-backs */
    DevBuf host_out, host_out2;            /* pinned host buffers for results */
    std::vector<rpq_chunk_info> infos;
    std::vector<ChunkDev> h_chunks;
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
    if (b.p) { rt_stream_sync(c->stream); rt_free_device(b.p); b.p = nullptr; b.cap = 0; }
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
}

}  // namespace

/* ================================================================== header (host) ==== */
extern "C" int rpq_make_header(const char* r1, uint64_t r1_len, const char* r2, uint64_t r2_len, int interleaved, uint32_t chunk_bases,
                               rpq_header* out, char* err, size_t err_cap) {
    if (!r1 || !out) return RPQ_ERR_ARG;
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
    for (auto& e : c->ev) rt_event_create(&e);
    memset(&c->stats, 0, sizeof c->stats);
    memset(&c->hdr, 0, sizeof c->hdr);
    { const char* e = getenv("RPQ_DEBUG_FORCE_V1"); c->force_v1 = e && e[0] == '1'; }
#ifndef RPQ_EMU
    cudaFuncSetAttribute(k_streams2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_streams3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_meta3, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_dec_format3, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
#endif
    *out = c;
    return RPQ_OK;
}

extern "C" void rpq_destroy(rpq_ctx* c) {
    if (!c) return;
    rt_set_device(c->device);
    rt_stream_sync(c->stream);
    DevBuf* all[] = {&c->loc, &c->pk, &c->pk_rc, &c->text[0], &c->text[1], &c->nl[0], &c->nl[1], &c->tile_state, &c->counters, &c->rlen, &c->unit_bases, &c->prefix, &c->scan_tmp,
                     &c->ustats, &c->chunk_first, &c->chunks, &c->meta, &c->meta0, &c->ov, &c->seqoff, &c->qualoff, &c->n1off, &c->n2off, &c->soff,
                     &c->errbits, &c->tmpx, &c->tmpy, &c->span_first[0], &c->span_first[1], &c->span_chunk[0], &c->span_chunk[1], &c->dir[0], &c->dir[1],
                     &c->slots[0], &c->slots[1], &c->span_slot[0], &c->span_slot[1], &c->misc, &c->out, &c->d_in, &c->d_desc, &c->out2,
                     &c->d_tmp2, &c->d_tmp[0], &c->d_tmp[1], &c->d_tmp[2], &c->d_tmp[3], &c->d_tmp[4], &c->d_tmp[5], &c->d_tmp[6], &c->d_tmp[7]};
    for (DevBuf* b : all) rt_free_device(b->p);
    rt_free_pinned(c->pinned_small);
    rt_free_pinned(c->host_out.p); rt_free_pinned(c->host_out2.p);
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

extern "C" int rpq_set_header(rpq_ctx* c, const rpq_header* h) {
    if (!c || !h) return RPQ_ERR_ARG;
    if (h->read_length_bytes != 1 && h->read_length_bytes != 2 && h->read_length_bytes != 4)
        return fail(c, RPQ_ERR_HEADER, "header incorrect: read length bytes should be 1/2/4");
    if (!(h->flags & RPQ_DONT_ENCODE_QUAL) && !(h->flags & RPQ_ENCODE_QUAL_BY_COL))
        return fail(c, RPQ_ERR_HEADER, "header selects the quality run-length coder, which ALGORITHM_VER 2 never produces");
    c->hdr = *h; c->have_hdr = true;
    build_header_dev(*h, c->hd);
    return RPQ_OK;
}

/* ================================================================== encode ==== */
namespace {

/* scalar read-back through the pinned scratch; synchronises the stream */
template <class T> int read_back(rpq_ctx* c, const void* dev, T* host, size_t count = 1) {
    if (rt_memcpy_d2h(c->pinned_small, dev, sizeof(T) * count, c->stream)) return RPQ_ERR_CUDA;
    if (rt_stream_sync(c->stream)) return RPQ_ERR_CUDA;
    memcpy(host, c->pinned_small, sizeof(T) * count);
    return RPQ_OK;
}

int index_text(rpq_ctx* c, int f, const u8* d_text, u64 len, IndexCounters* hc) {
    const u32 tiles = (u32)((len + IDX_TILE - 1) / IDX_TILE);
    if (!ensure(c, c->tile_state, sizeof(u64) * (tiles + 1)) || !ensure(c, c->counters, sizeof(IndexCounters) * 2)) return RPQ_ERR_NOMEM;
    size_t cap = (size_t)(len / 16) + 4096;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!ensure(c, c->nl[f], sizeof(u32) * cap)) return RPQ_ERR_NOMEM;
        cap = c->nl[f].cap / sizeof(u32);
        IndexCounters* dc = c->counters.as<IndexCounters>() + f;
        rt_memset(c->tile_state.p, 0, sizeof(u64) * (tiles + 1), c->stream);
        rt_memset(dc, 0, sizeof(IndexCounters), c->stream);
        LAUNCH(c, k_index_lines, tiles, IDX_THREADS, 0, d_text, len, c->nl[f].as<u32>(), (u32)cap, c->tile_state.as<u64>(), dc);
        LAUNCH(c, k_index_finish, 1, 32, 0, d_text, len, c->nl[f].as<u32>(), (u32)cap, dc);
        if (int rc = read_back(c, dc, hc)) return rc;
        if ((size_t)hc->n_nl + 1 <= cap) return RPQ_OK;
        cap = (size_t)hc->n_nl + 16;           /* pathological line density: index again with room for every line */
    }
    return RPQ_OK;
}

}  // namespace

extern "C" int rpq_encode(rpq_ctx* c, const rpq_encode_in* in, rpq_encode_out* out) {
    if (!c || !in || !out) return RPQ_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!c->have_hdr) return fail(c, RPQ_ERR_ARG, "rpq_set_header() must be called before rpq_encode()");
    if (in->r1_len == 0 || (in->r2 && in->r2_len == 0)) return RPQ_OK;          /* no records: no chunks (src/rfqcodec.cpp:165-166) */
    if (!in->r1 || in->r1_len >= (1ull << 32) || in->r2_len >= (1ull << 32)) return fail(c, RPQ_ERR_ARG, "FASTQ batch must be < 4 GiB per file");
    if (in->chunk_bases == 0) return fail(c, RPQ_ERR_ARG, "chunk_bases must be positive");
    rt_set_device(c->device);
    c->launches = 0; memset(&c->stats, 0, sizeof c->stats);
    c->infos.clear();
    const HeaderDev& hd = c->hd;
    const bool two = in->r2 != nullptr;
    const bool pe = two || in->interleaved;

    rt_event_record(&c->ev[0], c->stream);
    /* ---- input residency */
    const u8* d_text[2] = {nullptr, nullptr};
    const u64 lens[2] = {in->r1_len, two ? in->r2_len : 0};
    const char* srcs[2] = {in->r1, in->r2};
    for (int f = 0; f < (two ? 2 : 1); f++) {
        if (in->mem == RPQ_MEM_DEVICE) {
            if (((uintptr_t)srcs[f] & 15u) != 0) return fail(c, RPQ_ERR_ARG, "device FASTQ buffers must be 16-byte aligned");
            d_text[f] = (const u8*)srcs[f];
        } else {
            if (!ensure(c, c->text[f], lens[f] + 64)) return fail(c, RPQ_ERR_NOMEM, "out of device memory (text)");
            if (rt_memcpy_h2d(c->text[f].p, srcs[f], lens[f], c->stream)) return fail(c, RPQ_ERR_CUDA, "H2D copy failed");
            c->stats.h2d_bytes += lens[f];
            d_text[f] = c->text[f].as<u8>();
        }
    }
    rt_event_record(&c->ev[1], c->stream);

    /* ---- line index */
    IndexCounters ic[2]; memset(ic, 0, sizeof ic);
    for (int f = 0; f < (two ? 2 : 1); f++) {
        if (lens[f] == 0) continue;
        if (int rc = index_text(c, f, d_text[f], lens[f], &ic[f])) return rc == RPQ_ERR_NOMEM ? fail(c, rc, "out of device memory (line index)") : fail(c, rc, "line index failed");
        if (ic[f].bad_eol) return fail(c, RPQ_ERR_FASTQ, "unsupported line ends: expected all \"\\n\" or all \"\\r\\n\"");
    }
    EncBatchDev b; memset(&b, 0, sizeof b);
    for (int f = 0; f < 2; f++) { b.t[f].text = d_text[f]; b.t[f].len = lens[f]; b.t[f].nl = c->nl[f].as<u32>(); b.t[f].n_lines = ic[f].n_lines; b.t[f].crlf = ic[f].crlf; }
    b.is_pe = pe; b.two_files = two;
    const u32 per = pe ? 2u : 1u;
    u32 rec0 = ic[0].n_lines / 4, rec1 = ic[1].n_lines / 4;
    u32 n_units = two ? std::min(rec0, rec1) : (in->interleaved ? rec0 / 2 : rec0);
    if (n_units == 0) { rt_event_record(&c->ev[7], c->stream); rt_stream_sync(c->stream); return RPQ_OK; }

    /* ---- record lengths, greedy chunk cut */
    const size_t n_reads_max = (size_t)n_units * per;
    if (!ensure(c, c->rlen, 4 * n_reads_max) || !ensure(c, c->unit_bases, 4 * (size_t)n_units) || !ensure(c, c->prefix, 8 * (size_t)n_units) ||
        !ensure(c, c->ustats, sizeof(UnitStats)) || !ensure(c, c->scan_tmp, rt_scan_tmp_bytes(n_units)) || !ensure(c, c->errbits, 64))
        return fail(c, RPQ_ERR_NOMEM, "out of device memory (lengths)");
    if (!ensure(c, c->loc, 16 * n_reads_max)) return fail(c, RPQ_ERR_NOMEM, "out of device memory (record index)");
    b.rlen = c->rlen.as<u32>(); b.err = c->errbits.as<u32>(); b.loc = c->loc.as<uint4>();
    {
        UnitStats init; memset(&init, 0xFF, sizeof init); init.max_bases = 0; init.max_read = 0; init.max_head = 0; init.n_chunks = 0; init.units_in_chunks = 0;
        memcpy(c->pinned_small, &init, sizeof init);
        rt_memcpy_h2d(c->ustats.p, c->pinned_small, sizeof init, c->stream);
        rt_memset(c->errbits.p, 0, 64, c->stream);
    }
    LAUNCH(c, k_unit_lengths, (n_units + 255) / 256, 256, 0, b, n_units, c->rlen.as<u32>(), c->unit_bases.as<u32>(), c->ustats.as<UnitStats>());
    UnitStats us;
    if (int rc = read_back(c, c->ustats.p, &us)) return fail(c, rc, "CUDA failure in k_unit_lengths");
    if (us.first_empty < n_units) n_units = us.first_empty;       /* the reference stops at the first record with an empty line */
    if (n_units == 0) { rt_event_record(&c->ev[7], c->stream); rt_stream_sync(c->stream); return RPQ_OK; }
    const u32 uniform = (us.min_bases == us.max_bases) ? us.min_bases : 0;
    if (!uniform) rt_inclusive_sum_u32_u64(c->unit_bases.as<u32>(), c->prefix.as<u64>(), n_units, c->scan_tmp.p, c->scan_tmp.cap, c->stream);
    const size_t chunk_cap = (size_t)((lens[0] + lens[1]) / in->chunk_bases) + 4;
    if (!ensure(c, c->chunk_first, 4 * (chunk_cap + 1))) return fail(c, RPQ_ERR_NOMEM, "out of device memory (chunks)");
    LAUNCH(c, k_cut, 1, 256, 0, c->prefix.as<u64>(), n_units, in->chunk_bases, uniform, in->final, per, c->chunk_first.as<u32>(), (u32)chunk_cap + 1,
           c->ustats.as<UnitStats>());
    if (int rc = read_back(c, c->ustats.p, &us)) return fail(c, rc, "CUDA failure in k_cut");
    const u32 n_chunks = us.n_chunks;
    const u32 n_reads = us.units_in_chunks * per;
    /* only records that end up in a chunk are validated (a batch that is not final may end inside a record) */
    if (us.first_qual_len < us.units_in_chunks) return fail(c, RPQ_ERR_FASTQ, "quality and sequence lengths differ in record " + std::to_string(us.first_qual_len));
    if (us.first_name_len < us.units_in_chunks) return fail(c, RPQ_ERR_FASTQ, "name or strand line longer than 255 bytes in record " + std::to_string(us.first_name_len));
    if (us.first_read_len < us.units_in_chunks) return fail(c, RPQ_ERR_FASTQ, "read longer than 65535 bases in record " + std::to_string(us.first_read_len));
    if (n_chunks == 0) { rt_event_record(&c->ev[7], c->stream); rt_stream_sync(c->stream); return RPQ_OK; }
    if (n_chunks > chunk_cap) return fail(c, RPQ_ERR_FASTQ, "internal: chunk table overflow");
    b.n_reads = n_reads; b.n_chunks = n_chunks; b.chunk_first = c->chunk_first.as<u32>();
    b.uniform_reads_per_chunk = uniform ? ((in->chunk_bases + uniform - 1) / uniform) * per : 0;

    /* ---- per-read metadata, chunk flags, scans */
    if (!ensure(c, c->chunks, sizeof(ChunkDev) * n_chunks) || !ensure(c, c->meta, sizeof(ReadMeta) * (size_t)n_reads) ||
        !ensure(c, c->meta0, sizeof(ReadMeta) * n_chunks) || !ensure(c, c->ov, 2 * ((size_t)n_reads / 2 + 1)) ||
        !ensure(c, c->seqoff, 4 * (size_t)n_reads) || !ensure(c, c->qualoff, 4 * (size_t)n_reads) || !ensure(c, c->n1off, 4 * (size_t)n_reads) ||
        !ensure(c, c->n2off, 4 * (size_t)n_reads) || !ensure(c, c->soff, 4 * (size_t)n_reads) || !ensure(c, c->tmpx, 3 * (size_t)n_reads + 16) ||
        !ensure(c, c->tmpy, 3 * (size_t)n_reads + 16))
        return fail(c, RPQ_ERR_NOMEM, "out of device memory (read tables)");
    b.chunks = c->chunks.as<ChunkDev>(); b.meta = c->meta.as<ReadMeta>(); b.meta0 = c->meta0.as<ReadMeta>(); b.ov = c->ov.as<short>();
    b.seqoff = c->seqoff.as<u32>(); b.qualoff = c->qualoff.as<u32>(); b.n1off = c->n1off.as<u32>(); b.n2off = c->n2off.as<u32>(); b.soff = c->soff.as<u32>();
    LAUNCH(c, k_init_chunks, (n_chunks + 255) / 256, 256, 0, b);
    LAUNCH(c, k_meta0, (n_chunks + META_WARPS - 1) / META_WARPS, 32 * META_WARPS, 0, b);
    /* v2 path (thread per read on staged text) whenever a CTA's record heads fit in shared memory; very long reads
     * fall back to the warp-per-pair kernels that read the text directly */
    Meta2Cfg m2; memset(&m2, 0, sizeof m2);
    bool v2 = false;
    {
        const u32 units = us.units_in_chunks;
        m2.slot_words = ((us.max_head + 6u) / 4u + 1u) | 1u;
        m2.pkw = ((us.max_read + 15u) / 16u + 1u) | 1u;
        size_t smem2 = 0;
        for (u32 P : {128u, 64u, 32u}) {
            const size_t nr = (size_t)P * per;
            smem2 = 4 * nr * (m2.slot_words + m2.pkw) + 16 * nr + 4 * nr + ((nr + 3) & ~(size_t)3) + 4 * (size_t)P + 64;
            if (smem2 <= 160u * 1024u && !c->force_v1) { m2.units_per_cta = P; v2 = true; break; }
        }
        if (v2) {
            if (!ensure(c, c->pk, 4ull * n_reads * m2.pkw + 64)) return fail(c, RPQ_ERR_NOMEM, "out of device memory (packed reads)");
            b.pk = c->pk.as<u32>(); b.pk_rc = nullptr; b.pkw = m2.pkw;
            LAUNCH(c, k_meta3, (units + m2.units_per_cta - 1) / m2.units_per_cta, m2.units_per_cta * per, smem2, b, hd, units, m2);
        } else {
            const int use_smem = pe && us.max_read <= (u32)META_SEQ_SMEM;     /* longer reads: the overlap search reads the text directly */
            LAUNCH(c, k_meta, (units + META_WARPS - 1) / META_WARPS, 32 * META_WARPS, use_smem ? META_WARPS * 2 * META_SEQ_SMEM : 0, b, hd, units, use_smem);
        }
    }
    LAUNCH(c, k_chunk_finish, n_chunks, FIN_THREADS, 0, b, hd);
    if (hd.flags & (RPQ_HAS_X | RPQ_HAS_Y)) {
        dim3 g(n_chunks, 2);
        prof_begin(c, "k_coords");
        RPQ_LAUNCH(k_coords, g, CO_THREADS, 0, c->stream, b, hd, c->tmpx.as<u8>(), c->tmpy.as<u8>());
        prof_end(c);
        c->launches++;
    }

    /* ---- position streams: quality column and (header ENCODE_N_POS) the N positions */
    const bool have_q = !(hd.flags & RPQ_DONT_ENCODE_QUAL);
    const bool have_n = (hd.flags & RPQ_ENCODE_N_POS) != 0;
    u64 total_bases = 0;
    if (uniform) total_bases = (u64)uniform * us.units_in_chunks;
    else if (int rc = read_back(c, c->prefix.as<u64>() + (us.units_in_chunks - 1), &total_bases)) return fail(c, rc, "CUDA failure (prefix)");
    const u32 span_cap = (u32)(total_bases / ST_SPAN) + n_chunks + 1;
    StreamJob jobs[2]; memset(jobs, 0, sizeof jobs);
    if (!ensure(c, c->misc, 256)) return fail(c, RPQ_ERR_NOMEM, "out of device memory");
    for (;;) {
        bool again = false;
        rt_memset(c->misc.p, 0, 256, c->stream);
        for (int k = 0; k < 2; k++) {
            if (!(k == 0 ? have_q : have_n)) continue;
            StreamJob& j = jobs[k];
            j.mode = (u32)k; j.nstreams = k == 0 ? (u32)hd.nb + 1u : 1u;
            const u64 slot_cap = (u64)((double)total_bases * c->slot_factor) + (1u << 20);
            if (!ensure(c, c->span_first[k], 4 * ((size_t)n_chunks + 1)) || !ensure(c, c->span_chunk[k], 4 * (size_t)span_cap) ||
                !ensure(c, c->dir[k], sizeof(SpanDir) * (size_t)span_cap * j.nstreams) || !ensure(c, c->span_slot[k], 8 * (size_t)span_cap) ||
                !ensure(c, c->slots[k], slot_cap))
                return fail(c, RPQ_ERR_NOMEM, "out of device memory (stream slots)");
            j.span_first = c->span_first[k].as<u32>(); j.dir = c->dir[k].as<SpanDir>(); j.slots = c->slots[k].as<u8>(); j.slot_cap = c->slots[k].cap;
            j.slot_cursor = c->misc.as<u64>() + 2 * k; j.overflow = c->misc.as<u32>() + 16; j.span_slot = c->span_slot[k].as<u64>();
            j.n_spans = c->misc.as<u32>() + 20 + k;
            LAUNCH(c, k_span_plan, 1, 256, 0, b, (u32)k, c->span_first[k].as<u32>(), c->span_chunk[k].as<u32>(), span_cap, c->misc.as<u32>() + 20 + k);
            const size_t smem = ST_SPAN + 2 * ST_HALO + (size_t)j.nstreams * S2_THREADS * (sizeof(u32) + 3 * sizeof(u16));
            if (c->force_v1) LAUNCH(c, k_streams2, span_cap, S2_THREADS, smem, b, hd, j, c->span_chunk[k].as<u32>());
            else LAUNCH(c, k_streams3, span_cap, S2_THREADS, smem, b, hd, j, c->span_chunk[k].as<u32>());
        }
        u32 ovf = 0;
        if (have_q || have_n) { if (int rc = read_back(c, c->misc.as<u32>() + 16, &ovf)) return fail(c, rc, "CUDA failure in k_streams"); }
        if (ovf) { c->slot_factor = 5.1; again = true; }
        if (!again) break;
    }

    /* ---- layout, offsets */
    LAUNCH(c, k_layout, n_chunks, LAY_THREADS, 0, b, hd, jobs[0], jobs[1], (int)have_q, (int)have_n);
    LAUNCH(c, k_chunk_offsets, 1, 256, 0, b, c->misc.as<u64>() + 16);
    u64 total_out = 0;
    if (int rc = read_back(c, c->misc.as<u64>() + 16, &total_out)) return fail(c, rc, "CUDA failure in k_layout");
    u32 errbits = 0;
    if (int rc = read_back(c, c->errbits.p, &errbits)) return fail(c, rc, "CUDA failure (error bits)");
    if (errbits & ERRBIT_COORD) return fail(c, RPQ_ERR_COORD, "The X/Y coordinate cannot be larger than 2M");
    if (!ensure(c, c->out, total_out + 64)) return fail(c, RPQ_ERR_NOMEM, "out of device memory (output)");
    u8* d_out = c->out.as<u8>();

    /* ---- emit */
    LAUNCH(c, k_head, (n_chunks + 3) / 4, 128, 0, b, hd, d_out, c->tmpx.as<u8>(), c->tmpy.as<u8>(), *in);
    if (v2) {
        LAUNCH(c, k_emit2, (n_reads + 255) / 256, 256, 0, b, hd, d_out);
        if (errbits & INFOBIT_NEED_NAMES) LAUNCH(c, k_emit_names, (n_reads + 7) / 8, 256, 0, b, hd, d_out);
    } else {
        LAUNCH(c, k_emit, (n_reads + EMIT_WARPS - 1) / EMIT_WARPS, 32 * EMIT_WARPS, 0, b, hd, d_out);
    }
    if (have_q) LAUNCH(c, k_gather, span_cap, 256, 0, b, jobs[0], c->span_chunk[0].as<u32>(), d_out, 0);
    else LAUNCH(c, k_raw_qual, (n_reads + EMIT_WARPS - 1) / EMIT_WARPS, 32 * EMIT_WARPS, 0, b, d_out);
    if (have_n) LAUNCH(c, k_gather, span_cap, 256, 0, b, jobs[1], c->span_chunk[1].as<u32>(), d_out, 1);
    rt_event_record(&c->ev[2], c->stream);

    /* ---- results */
    c->h_chunks.resize(n_chunks);
    if (rt_memcpy_d2h(c->h_chunks.data(), c->chunks.p, sizeof(ChunkDev) * n_chunks, c->stream)) return fail(c, RPQ_ERR_CUDA, "D2H failed");
    const u8* result = d_out;
    if (in->out_mem == RPQ_MEM_HOST) {
        if (!ensure_pinned(c, c->host_out, total_out + 64)) return fail(c, RPQ_ERR_NOMEM, "out of pinned host memory");
        if (rt_memcpy_d2h(c->host_out.p, d_out, total_out, c->stream)) return fail(c, RPQ_ERR_CUDA, "D2H failed");
        c->stats.d2h_bytes += total_out;
        result = c->host_out.as<u8>();
    }
    rt_event_record(&c->ev[3], c->stream);
    if (rt_stream_sync(c->stream)) return fail(c, RPQ_ERR_CUDA, "CUDA failure at the end of rpq_encode");
    if (int rc = check_launch(c, "rpq_encode")) return rc;

    c->infos.resize(n_chunks);
    for (u32 k = 0; k < n_chunks; k++) {
        const ChunkDev& ck = c->h_chunks[k]; rpq_chunk_info& ci = c->infos[k];
        memset(&ci, 0, sizeof ci);
        ci.offset = ck.out_offset; ci.bytes = ck.bytes; ci.msize = ck.msize; ci.reads = ck.count; ci.flags = (uint16_t)ck.flags;
        ci.seq_size = ck.seq_size; ci.qual_size = ck.qual_size; ci.npos_size = ck.npos_size; ci.x_size = ck.x_size; ci.y_size = ck.y_size;
        ci.name1_size = ck.n1_size; ci.name2_size = ck.n2_size; ci.strand_size = ck.strand_size;
        ci.r1_end = ck.r1_end; ci.r2_end = ck.r2_end;
    }
    out->data = result; out->bytes = total_out; out->n_chunks = n_chunks; out->chunks = c->infos.data(); out->n_reads = n_reads;
    {
        /* consumed = just past the line break of the last record (clipped: the final line may have none) */
        const ChunkDev& last = c->h_chunks[n_chunks - 1];
        auto past = [&](u32 brk, int f) { u64 p = (u64)brk + 1 + ic[f].crlf; return p > lens[f] ? lens[f] : p; };
        out->r1_consumed = past(last.r1_end, 0);
        out->r2_consumed = two ? past(last.r2_end, 1) : 0;
    }
    c->stats.launches = c->launches;
    c->stats.ms_h2d = rt_event_ms(c->ev[0], c->ev[1]);
    c->stats.ms_kernels = rt_event_ms(c->ev[1], c->ev[2]);
    c->stats.ms_d2h = rt_event_ms(c->ev[2], c->ev[3]);
    c->stats.ms_total = rt_event_ms(c->ev[0], c->ev[3]);
    return RPQ_OK;
}

#include "rpq_api_decode.inc"
