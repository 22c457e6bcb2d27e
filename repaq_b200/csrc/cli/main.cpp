/*
 * repaq_b200_cli - host C++ driver with the reference's command line (reference src/main.cpp:31-49, src/options.cpp:36-111) for the
 * modes on the hot path - `repaq -c`, `repaq -d`, `repaq --compare`, `-v` / `-f` (src/repaq.cpp:36-759) - on top of the C ABI.
 *
 * File-level rules kept from the reference: header written first (src/repaq.cpp:554-557), an input without records leaves an EMPTY
 * output, Q13 NO_LINE_BREAK thresholds (src/fastqreader.cpp:31-46), trailing-newline trimming on decode (src/repaq.cpp:300-328,
 * 375-413), `.gz` FASTQ in and out through zlib (src/fastqreader.cpp:33,49-52, src/writer.cpp:40-43), `.rfq.xz` by re-running this
 * program through the `xz` executable (src/main.cpp:133-178), the compare report of reportCompareResult (src/repaq.cpp:235-259).
 *
 * Where it differs by design: files are STREAMED.  A reader thread per input fills page-locked windows (rpq_host_alloc) while the
 * GPU works on the previous one (64 MiB of text per file and window, 16 MiB of .rfq when decoding), results leave through a writer
 * thread; host memory is bounded whatever the file size (the reference's reader refills one 1 MiB buffer, src/fastqreader.cpp:5,31-46).
 * -v / -f: every batch is decoded again and checked against its input on the GPU (rpq_compare) after it has been written, the
 * first difference is reported on stderr in the words of completeCheckAndOutput (src/repaq.cpp:430-528); like the reference, the
 * output is written either way.  (The reference's -f checks every tenth chunk to save CPU time, src/repaq.cpp:575; here -f checks
 * every chunk, as -v does: a superset.)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "repaq_b200.h"

static std::vector<std::string> g_outputs;               /* files this run has created: removed again if it fails */
static void error_exit(const std::string& msg) {         /* src/util.h:246-249 */
    fprintf(stderr, "ERROR: %s\n", msg.c_str());
    for (const std::string& p : g_outputs) unlink(p.c_str());
    _exit(255);
}

static bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }
static bool is_fastq_name(const std::string& f) { return ends_with(f, ".fq") || ends_with(f, ".fastq") || ends_with(f, ".fq.gz") || ends_with(f, ".fastq.gz"); }   /* src/options.cpp:22-27 */
static bool is_rfq_name(const std::string& f) { return ends_with(f, ".rfq") || ends_with(f, ".rfq.xz"); }                                                        /* src/options.cpp:29-34 */

/* ---- byte sources and sinks: plain files, stdin / stdout, `.gz` through zlib (as FastqReader / Writer pick by file name) */
struct Source {
    FILE* f = nullptr; gzFile gz = nullptr; std::string path;
    void open(const std::string& p) {
        path = p;
        if (ends_with(p, ".gz")) { gz = gzopen(p.c_str(), "r"); if (!gz) error_exit("Failed to open file: " + p); gzbuffer(gz, 1 << 20); }
        else { f = p == "/dev/stdin" ? stdin : fopen(p.c_str(), "rb"); if (!f) error_exit("Failed to open file: " + p); }
    }
    /* up to n bytes; fewer only at the end of the input */
    size_t read(char* dst, size_t n) {
        size_t got = 0;
        while (got < n) {
            if (gz) {
                const int k = gzread(gz, dst + got, (unsigned)((n - got) < (1u << 30) ? (n - got) : (1u << 30)));
                if (k < 0) error_exit("Error to read gzip file");
                if (k == 0) break;
                got += (size_t)k;
            } else {
                const size_t k = fread(dst + got, 1, n - got, f);
                if (k == 0) break;
                got += k;
            }
        }
        return got;
    }
    void close() { if (gz) gzclose(gz); else if (f && f != stdin) fclose(f); gz = nullptr; f = nullptr; }
};

struct Sink {
    FILE* f = nullptr; gzFile gz = nullptr; std::string path;
    void open(const std::string& p) {
        path = p;
        if (ends_with(p, ".gz")) {                        /* src/writer.cpp:40-43: level 3, 1 MiB buffer */
            gz = gzopen(p.c_str(), "w"); if (!gz) error_exit("Failed to open file: " + p);
            gzsetparams(gz, 3, Z_DEFAULT_STRATEGY); gzbuffer(gz, 1024 * 1024);
            g_outputs.push_back(p);
        } else if (p == "/dev/stdout") f = stdout;
        else { f = fopen(p.c_str(), "wb"); if (!f) error_exit("Failed to open file: " + p); g_outputs.push_back(p); }
    }
    void write(const void* p, size_t n) {
        const char* s = (const char*)p;
        while (n) {
            const size_t k = n < (1u << 30) ? n : (1u << 30);
            if (gz) { if (gzwrite(gz, s, (unsigned)k) != (int)k) error_exit("Failed to write file: " + path); }
            else if (fwrite(s, 1, k, f) != k) error_exit("Failed to write file: " + path);
            s += k; n -= k;
        }
    }
    void close() { if (gz) { gzflush(gz, Z_FINISH); gzclose(gz); } else if (f) { if (f == stdout) fflush(f); else fclose(f); } gz = nullptr; f = nullptr; }
};

#include <chrono>
static void stamp(const char* what) {                    /* RPQ_CLI_TIMING=1: where the wall time of a run goes */
    static const bool on = getenv("RPQ_CLI_TIMING") != NULL;
    static const auto t0 = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[cli %8.1f ms] %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
}
static size_t env_size(const char* name, size_t dflt) { const char* e = getenv(name); if (!e) return dflt; const unsigned long long v = strtoull(e, NULL, 10); return v ? (size_t)v : dflt; }

/*
 * A reader thread that fills page-locked windows ahead of the consumer.  A window's data starts HEAD bytes into its buffer: the
 * consumer copies the text the previous call did not cover (an unfinished chunk, an unfinished record) in front of it, so that
 * every call sees one contiguous text without the bulk of the data ever being copied on the host.
 */
struct Window { char* base = nullptr; size_t n = 0; bool eof = false; };
class StreamReader {
public:
    size_t head = 0, cap = 0;                              /* bytes reserved in front of the data / data bytes per window */
    uint64_t total = 0;                                    /* bytes read so far (consumer side: of the windows handed out) */
    int last_byte = -1;
    void start(const std::string& path, size_t head_bytes, size_t cap_bytes, int n_windows) {
        src_.open(path); head = head_bytes; cap = cap_bytes;
        win_.resize(n_windows); state_.assign(n_windows, 0); want_.assign(n_windows, cap_bytes);
        th_ = std::thread([this] { run(); });             /* a window's buffer is page-locked when the reader first needs it: short inputs pay for one */
    }
    /* window k (in order); blocks until it has been read */
    Window* get(uint64_t k) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return state_[k % win_.size()] == 2 && seq_[k % win_.size()] == k; });
        return &win_[k % win_.size()];
    }
    /* the consumer is done with window k; the window that will reuse its buffer reads `want` bytes (<= cap) */
    void release(uint64_t k, size_t want) {
        std::lock_guard<std::mutex> lk(m_);
        state_[k % win_.size()] = 0; want_[k % win_.size()] = want < cap ? want : cap;
        cv_.notify_all();
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(m_); quit_ = true; cv_.notify_all(); }
        if (th_.joinable()) th_.join();
        for (Window& w : win_) rpq_host_free(w.base);
        win_.clear(); src_.close();
    }
private:
    void run() {
        for (uint64_t k = 0;; k++) {
            const size_t s = k % win_.size();
            size_t want;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || state_[s] == 0; });
                if (quit_) return;
                state_[s] = 1; want = want_[s];
            }
            Window& w = win_[s];
            if (!w.base) { w.base = (char*)rpq_host_alloc(head + cap + 64); if (!w.base) error_exit("out of page-locked host memory (no CUDA device?): repaq_b200 has no CPU fallback"); }
            w.n = src_.read(w.base + head, want);
            w.eof = w.n < want;
            {
                std::lock_guard<std::mutex> lk(m_);
                seq_[s] = k; state_[s] = 2;
                cv_.notify_all();
            }
            if (w.eof) return;
        }
    }
    Source src_; std::thread th_; std::mutex m_; std::condition_variable cv_;
    std::vector<Window> win_; std::vector<int> state_; std::vector<size_t> want_; uint64_t seq_[8] = {0}; bool quit_ = false;
};

/* results leave through a writer thread: the caller hands over a buffer it will not touch until done() says so */
class StreamWriter {
public:
    void start(const std::string& path) { sink_.open(path); th_ = std::thread([this] { run(); }); }
    void put(std::vector<char>&& owned) { std::lock_guard<std::mutex> lk(m_); q_.push_back(Job{std::move(owned), nullptr, 0, 0}); cv_.notify_all(); }
    /* borrowed memory: valid until wait(ticket) returns */
    uint64_t put_borrowed(const void* p, size_t n) { std::lock_guard<std::mutex> lk(m_); q_.push_back(Job{{}, p, n, ++issued_}); cv_.notify_all(); return issued_; }
    void wait(uint64_t ticket) { std::unique_lock<std::mutex> lk(m_); cv_.wait(lk, [&] { return done_ >= ticket; }); }
    void finish() {
        { std::lock_guard<std::mutex> lk(m_); quit_ = true; cv_.notify_all(); }
        if (th_.joinable()) th_.join();
        sink_.close();
    }
private:
    struct Job { std::vector<char> own; const void* p; size_t n; uint64_t ticket; };
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return quit_ || !q_.empty(); });
                if (q_.empty()) return;
                j = std::move(q_.front()); q_.pop_front();
            }
            if (j.p) sink_.write(j.p, j.n); else sink_.write(j.own.data(), j.own.size());
            if (j.ticket) { std::lock_guard<std::mutex> lk(m_); done_ = j.ticket; cv_.notify_all(); }
        }
    }
    Sink sink_; std::thread th_; std::mutex m_; std::condition_variable cv_; std::deque<Job> q_; uint64_t issued_ = 0, done_ = 0; bool quit_ = false;
};

/* Q13: the reader raises hasNoLineBreakAtEnd when it loads a SHORT 1 MiB buffer that does not end in '\n' (src/fastqreader.cpp:42-45) */
static void nobreak_rule(uint64_t n, int last_byte, uint64_t& from, bool& tail) {
    const uint64_t MiB = 1ull << 20;
    const bool nl = n && last_byte == '\n';
    if (n % MiB == 0) { from = nl ? UINT64_MAX : n; tail = true; }
    else { from = nl ? UINT64_MAX : (n / MiB) * MiB; tail = false; }
}

struct Opt { std::string in1, in2, out1, out2, rfq_compare, json_compare; bool compress = false, decompress = false, compare = false, interleaved = false, to_stdout = false, from_stdin = false, verify = false; int k = 1000; int device = 0; int threads = 1, level = 3; };

/* one input of the compress loop: its reader, the text of the current call (the uncovered rest of the previous window + the
 * current window) and where that text starts in the file */
struct Feed {
    StreamReader rd;
    uint64_t k = 0;                 /* current window */
    Window* cur = nullptr;
    char* text = nullptr; size_t len = 0;
    uint64_t file_pos = 0;          /* file offset of text[0] */
    bool eof = false;               /* cur is the last window */
    uint64_t size = 0; bool size_known = false;
    void first() { cur = rd.get(0); text = cur->base + rd.head; len = cur->n; eof = cur->eof; track(cur); }
    /* more windows in front of the current text until it holds `need` bytes or the input has ended */
    void at_least(size_t need) {
        while (len < need && !eof) {
            Window* nx = rd.get(k + 1);
            track(nx);
            if (len > rd.head) error_exit("the streaming window is too small for this chunk size (raise RPQ_CLI_FQ_HEAD)");
            char* dst = nx->base + rd.head - len;
            memmove(dst, text, len);
            rd.release(k, rd.cap);
            k++; cur = nx; text = dst; len += nx->n; eof = nx->eof;
        }
    }
    void track(Window* w) { if (w->n) rd.last_byte = (unsigned char)w->base[rd.head + w->n - 1]; rd.total += w->n; if (w->eof) { size = rd.total; size_known = true; } }
};

static int do_compress(const Opt& o) {
    const bool two = !o.in2.empty();
    const uint32_t chunk_bases = (uint32_t)(o.k < 100 ? 100 : o.k) * 1000u;        /* src/main.cpp:69 */
    /* a window holds at least six chunks of text (a chunk of b bases is ~2.5 b bytes of FASTQ per file); the room in front of it
     * takes what a call leaves uncovered: at most a chunk, plus the drift between two files whose records differ in length */
    size_t win = env_size("RPQ_CLI_FQ_WINDOW", 64u << 20), head = env_size("RPQ_CLI_FQ_HEAD", 32u << 20);
    if (!getenv("RPQ_CLI_FQ_WINDOW") && (uint64_t)chunk_bases * 15 > win) win = (size_t)chunk_bases * 15;
    if (!getenv("RPQ_CLI_FQ_HEAD") && (uint64_t)chunk_bases * 16 > head) head = (size_t)chunk_bases * 16;
    if (win + head >= (3ull << 30)) error_exit("chunk size too large for the streaming windows (< 4 GiB of text per file and call)");
    Feed fd[2];
    stamp("start");
    fd[0].rd.start(o.in1, head, win, 3);
    if (two) fd[1].rd.start(o.in2, head, win, 3);
    const int nf = two ? 2 : 1;
    /* the header is made from the first chunk (src/repaq.cpp:554-566): the first text holds it whole - 15 bytes of text per base of a
     * chunk (256-byte names on 20-base reads), or the whole input */
    stamp("windows allocated, readers running");
    for (int f = 0; f < nf; f++) { fd[f].first(); fd[f].at_least((size_t)chunk_bases * 15); }
    stamp("first windows read");

    char err[768];
    rpq_header h;
    const int hrc = rpq_make_header(fd[0].text, fd[0].len, two ? fd[1].text : NULL, two ? fd[1].len : 0, o.interleaved, chunk_bases, &h, err, sizeof err);
    StreamWriter out;
    out.start(o.out1);
    if (hrc == RPQ_NO_RECORDS) { out.finish(); for (int f = 0; f < nf; f++) fd[f].rd.stop(); return 0; }      /* no record: an empty output, as the reference leaves it */
    if (hrc) error_exit(err);
    { std::vector<char> hb(17 + 128); hb.resize(rpq_header_write(&h, (uint8_t*)hb.data(), hb.size())); out.put(std::move(hb)); }
    rpq_ctx* ctx = NULL;
    if (rpq_create(o.device, &ctx)) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    stamp("header made, context created");
    rpq_ctx* check = NULL;                                 /* -v / -f: the reference's codec4check */
    for (;;) {
        /* the window after this one must be there (or the input must have ended) before this one is encoded: only then is it
         * known whether the file ends within 1 MiB of this text, which is what decides Q13's flags for its chunks */
        Window* nxt[2] = {nullptr, nullptr};
        for (int f = 0; f < nf; f++) if (!fd[f].eof) { nxt[f] = fd[f].rd.get(fd[f].k + 1); fd[f].track(nxt[f]); }
        const bool final = fd[0].eof && (!two || fd[1].eof);
        uint64_t from[2] = {UINT64_MAX, UINT64_MAX}; bool tail[2] = {false, false};
        for (int f = 0; f < nf; f++) if (fd[f].size_known) nobreak_rule(fd[f].size, fd[f].rd.last_byte, from[f], tail[f]);
        if (!two && o.interleaved) { from[1] = from[0]; tail[1] = tail[0]; }
        rpq_encode_in in; memset(&in, 0, sizeof in);
        in.r1 = fd[0].text; in.r1_len = fd[0].len; in.r2 = two ? fd[1].text : NULL; in.r2_len = two ? fd[1].len : 0;
        in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.interleaved = o.interleaved; in.chunk_bases = chunk_bases;
        in.final = final;
        for (int f = 0; f < 2; f++) {
            const Feed& s = fd[two ? f : 0];
            in.file_offset[f] = s.file_pos;
            in.nobreak_from[f] = from[f] == UINT64_MAX ? UINT64_MAX : (from[f] > s.file_pos ? from[f] - s.file_pos : 0);
        }
        in.tail_flags = (uint16_t)((tail[0] ? RPQ_NO_LINE_BREAK_AT_END : 0) | (tail[1] ? RPQ_NO_LINE_BREAK_AT_END_R2 : 0));
        /* Paired files end together when they hold the same number of records - but their last windows need not: when one input has
         * ended and the other has not, the call is tried as the final one; it is, if it covers every record the ended file still had
         * (the pairs end with the shorter file, FastqReaderPair::read src/fastqreader.cpp:287-299); else it is repeated as an ordinary one */
        const bool one_sided = two && fd[0].eof != fd[1].eof;
        if (one_sided) in.final = 1;
        rpq_encode_out res;
        if (rpq_encode(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        bool last_call = final;
        if (one_sided) {
            const Feed& ended = fd[0].eof ? fd[0] : fd[1];
            const uint64_t used = fd[0].eof ? res.r1_consumed : res.r2_consumed;
            bool all = true;
            for (size_t q = (size_t)used; q < ended.len && all; q++) all = ended.text[q] == '\n' || ended.text[q] == '\r';
            if (all) last_call = true;
            else { in.final = 0; if (rpq_encode(ctx, &in, &res)) error_exit(rpq_last_error(ctx)); }
        }
        stamp("batch encoded");
        if (res.bytes) out.put(std::vector<char>((const char*)res.data, (const char*)res.data + res.bytes));
        if (o.verify && res.bytes) {
            /* the check of completeCheckAndOutput: decode what was just written, compare it read by read with what it was made from */
            if (!check && (rpq_create(o.device, &check) || rpq_set_header(check, &h))) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
            rpq_compare_in ci; memset(&ci, 0, sizeof ci);
            ci.rfq = res.data; ci.rfq_bytes = res.bytes; ci.rfq_mem = RPQ_MEM_HOST; ci.rfq_final = 1;
            ci.r1 = in.r1; ci.r1_len = in.final ? in.r1_len : res.r1_consumed; ci.r2 = in.r2; ci.r2_len = in.final ? in.r2_len : res.r2_consumed;
            ci.fq_mem = RPQ_MEM_HOST; ci.fq_final = 1;
            ci.fq_offset[0] = in.file_offset[0]; ci.fq_offset[1] = in.file_offset[1];
            rpq_compare_out co;
            if (o.interleaved) fprintf(stderr, "verify: --interleaved_in input is not checked\n");
            else if (rpq_compare(check, &ci, &co)) error_exit(rpq_last_error(check));
            else if (co.verdict == RPQ_CMP_RFQ_MORE || co.verdict == RPQ_CMP_FASTQ_MORE) error_exit("encoding error in chunk, the output will be wrong, quit now!");
            else if (co.verdict != RPQ_CMP_EQUAL)
                fprintf(stderr, "integrity check failure \nexpected: \n%.*s\ngot:\n%.*s\n", (int)co.fastq_field_len, co.fastq_field, (int)co.rfq_field_len, co.rfq_field);
        }
        if (last_call) break;
        /* what the call did not cover moves in front of the next window (or stays where it is when this input has ended) */
        const uint64_t used[2] = {res.r1_consumed, two ? res.r2_consumed : 0};
        bool progress = used[0] != 0;
        for (int f = 0; f < nf; f++) {
            Feed& s = fd[f];
            const size_t rest = s.len - (size_t)used[f];
            s.file_pos += used[f];
            if (nxt[f]) {
                if (rest > s.rd.head) error_exit("a chunk does not fit the streaming window; lower --chunk (or raise RPQ_CLI_FQ_HEAD)");
                char* dst = nxt[f]->base + s.rd.head - rest;
                memmove(dst, s.text + used[f], rest);
                /* two files whose records differ in length drift apart: the one that is ahead reads less next time */
                s.rd.release(s.k, rest > (4u << 20) ? s.rd.cap - (rest - (4u << 20)) : s.rd.cap);
                s.k++; s.cur = nxt[f]; s.text = dst; s.len = rest + nxt[f]->n; s.eof = nxt[f]->eof;
                progress = progress || nxt[f]->n != 0;
            } else { s.text += used[f]; s.len = rest; }
        }
        if (!progress) {
            /* nothing new to read on any side and no chunk came out: what is left is the file's tail */
            for (int f = 0; f < nf; f++) if (!fd[f].eof) error_exit("a chunk does not fit the streaming window; lower --chunk (or raise RPQ_CLI_FQ_WINDOW)");
        }
    }
    out.finish();
    stamp("output written");
    if (!getenv("RPQ_CLI_TEARDOWN")) _exit(0);             /* everything is on disk: unpinning windows and freeing device memory one by one only costs time */
    for (int f = 0; f < nf; f++) fd[f].rd.stop();
    if (check) rpq_destroy(check);
    rpq_destroy(ctx);
    stamp("done");
    return 0;
}

/* Repaq::decompress / decompressPE (src/repaq.cpp:262-413).  The .rfq is read in windows and decoded in batches of whole chunks, so
 * that device and pinned host memory stay bounded whatever the file size; two contexts take turns, the FASTQ of one batch is written
 * while the next one is decoded.  Whether a chunk is the LAST one of the file (the trailing-newline rule) is only known once the
 * bytes after it have failed to decode as a chunk, so the last chunk of a batch that is not the end of the file is held back and
 * decoded again as the first chunk of the next batch. */
static int do_decompress(const Opt& o) {
    const bool pe = !o.out2.empty();
    const size_t win = env_size("RPQ_CLI_RFQ_WINDOW", 16u << 20), head = env_size("RPQ_CLI_RFQ_HEAD", 32u << 20);
    StreamReader rd;
    rd.start(o.in1, head, win, 3);
    uint64_t k = 0;
    Window* cur = rd.get(0);
    char* text = cur->base + head; size_t len = cur->n; bool eof = cur->eof;
    StreamWriter w1, w2;
    if (len == 0 && eof) {
        /* an empty .rfq reads as a single-end header without chunks (RfqHeader::read keeps the constructor's values, src/rfqheader.cpp:7-43) */
        if (pe) error_exit("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
        w1.start(o.out1); w1.finish(); rd.stop();
        return 0;
    }
    while (len < 4096 && !eof) {                          /* the file header (17 + up to 128 bytes) in one piece */
        Window* nx = rd.get(k + 1);
        char* dst = nx->base + head - len;
        memmove(dst, text, len);
        rd.release(k, win);
        k++; cur = nx; text = dst; len += nx->n; eof = nx->eof;
    }
    char err[768]; rpq_header h; size_t used = 0;
    if (rpq_header_read((const uint8_t*)text, len, &h, &used, err, sizeof err)) error_exit(err);
    if (pe && !(h.flags & RPQ_PAIRED_END)) error_exit("The input RFQ file was encoded by single-end FASTQ, you should not specify <out2>");
    text += used; len -= used;
    rpq_ctx* ctx[2] = {NULL, NULL};
    for (int i = 0; i < 2; i++) { if (rpq_create(o.device, &ctx[i])) error_exit("no CUDA device: repaq_b200 has no CPU fallback"); if (rpq_set_header(ctx[i], &h)) error_exit(rpq_last_error(ctx[i])); }
    w1.start(o.out1);
    if (pe) w2.start(o.out2);
    uint64_t ticket[2][2] = {{0, 0}, {0, 0}};            /* writes that still read a context's result buffers */
    bool skip_first = false;                               /* decompressPE's `continue`: the chunk after a flagged one is never written */
    int turn = 0;
    for (;;) {
        Window* nxt = eof ? nullptr : rd.get(k + 1);
        const bool final = eof;
        rpq_ctx* c = ctx[turn];
        w1.wait(ticket[turn][0]); if (pe) w2.wait(ticket[turn][1]);     /* the writer is done with this context's previous result */
        rpq_decode_in in; memset(&in, 0, sizeof in);
        in.data = (const uint8_t*)text; in.bytes = len; in.mem = RPQ_MEM_HOST; in.out_mem = RPQ_MEM_HOST; in.split_pairs = pe;
        rpq_decode_out res;
        if (rpq_decode(c, &in, &res)) error_exit(rpq_last_error(c));
        size_t covered = 0;                                 /* bytes of the text that are done with */
        if (res.n_chunks >= (final ? 1u : 2u)) {
            const uint32_t n_keep = final ? res.n_chunks : res.n_chunks - 1;
            uint64_t a1 = 0, a2 = 0;
            for (uint32_t i = 0; i < n_keep; i++) {
                const rpq_chunk_info& ck = res.chunks[i];
                const bool last = final && i + 1 == res.n_chunks;
                const bool f1 = (ck.flags & RPQ_NO_LINE_BREAK_AT_END) != 0, f2 = (ck.flags & RPQ_NO_LINE_BREAK_AT_END_R2) != 0;
                if (skip_first) { skip_first = false; a1 += ck.out1_bytes; a2 += ck.out2_bytes; continue; }
                if (!pe) {
                    /* Repaq::decompress: only a flagged LAST chunk loses its final newline */
                    ticket[turn][0] = w1.put_borrowed(res.out1 + a1, (f1 && last && ck.out1_bytes) ? ck.out1_bytes - 1 : ck.out1_bytes);
                } else {
                    /* Repaq::decompressPE incl. its `continue` (src/repaq.cpp:395,405): after a flagged chunk that is not the last one,
                     * the rest of that chunk's output and the whole next chunk are never written */
                    bool skip_next = false;
                    if (f1) { if (last) ticket[turn][0] = w1.put_borrowed(res.out1 + a1, ck.out1_bytes ? ck.out1_bytes - 1 : 0); else { ticket[turn][0] = w1.put_borrowed(res.out1 + a1, ck.out1_bytes); skip_next = true; } }
                    else ticket[turn][0] = w1.put_borrowed(res.out1 + a1, ck.out1_bytes);
                    if (!skip_next) {
                        if (f2) { if (last) ticket[turn][1] = w2.put_borrowed(res.out2 + a2, ck.out2_bytes ? ck.out2_bytes - 1 : 0); else { ticket[turn][1] = w2.put_borrowed(res.out2 + a2, ck.out2_bytes); skip_next = true; } }
                        else ticket[turn][1] = w2.put_borrowed(res.out2 + a2, ck.out2_bytes);
                    }
                    skip_first = skip_next;
                }
                a1 += ck.out1_bytes; a2 += ck.out2_bytes;
            }
            covered = final ? len : (size_t)res.chunks[res.n_chunks - 1].offset;       /* the held-back chunk starts the next batch */
            turn ^= 1;
        } else if (final) break;                            /* what is left is not a chunk: the reference stops here too */
        if (final) break;
        const size_t rest = len - covered;
        if (rest > head) error_exit("a chunk does not fit the streaming window (raise RPQ_CLI_RFQ_HEAD)");
        char* dst = nxt->base + head - rest;
        memmove(dst, text + covered, rest);
        rd.release(k, win);
        k++; cur = nxt; text = dst; len = rest + nxt->n; eof = nxt->eof;
    }
    w1.finish(); if (pe) w2.finish();
    if (!getenv("RPQ_CLI_TEARDOWN")) _exit(0);
    rd.stop();
    rpq_destroy(ctx[0]); rpq_destroy(ctx[1]);
    return 0;
}

static std::vector<char> slurp(const std::string& path) {
    Source s; s.open(path);
    std::vector<char> b;
    size_t cap = 1 << 26, n = 0;
    b.resize(cap);
    for (;;) {
        const size_t got = s.read(b.data() + n, cap - n);
        n += got;
        if (n < cap) break;
        cap *= 2; b.resize(cap);
    }
    s.close();
    b.resize(n);
    return b;
}
static void spill(const std::string& path, const void* p, size_t n) { Sink s; s.open(path); s.write(p, n); s.close(); }

/* Repaq::compare / comparePE (src/repaq.cpp:36-233): the .rfq is decoded and checked read by read against the FASTQ file(s) on
 * the GPU, in batches of whole chunks against windows of FASTQ text; the report is reportCompareResult's (:235-259). */
static int do_compare(const Opt& o) {
    std::vector<char> rfq = slurp(o.rfq_compare), r1 = slurp(o.in1), r2;
    const bool pe = !o.in2.empty();
    if (pe) r2 = slurp(o.in2);
    char err[768]; rpq_header h; size_t used = 0;
    if (rfq.empty()) {
        /* an empty .rfq: a default header and no chunks (src/rfqheader.cpp:7-43); any header will do for zero chunks */
        memset(&h, 0, sizeof h); h.read_length_bytes = 1; h.flags = RPQ_ENCODE_QUAL_BY_COL; h.n_base_qual = '#'; h.overlap_shift = -24; h.qual_bins = 1; h.qual_buf[0] = 'F';
    } else if (rpq_header_read((const uint8_t*)rfq.data(), rfq.size(), &h, &used, err, sizeof err)) error_exit(err);
    rpq_ctx* ctx = NULL;
    if (rpq_create(o.device, &ctx)) error_exit("no CUDA device: repaq_b200 has no CPU fallback");
    if (rpq_set_header(ctx, &h)) error_exit(rpq_last_error(ctx));
    uint64_t FQ_WIN = 3ull << 30;                          /* < 4 GiB of text per file and call */
    uint64_t rfq_win = 256ull << 20;                       /* chunks decoded per call: ~2 GB of FASTQ at the usual ratios */
    if (const char* e = getenv("RPQ_CLI_RFQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) rfq_win = w; }      /* tests: small windows */
    if (const char* e = getenv("RPQ_CLI_FQ_WINDOW")) { const uint64_t w = strtoull(e, NULL, 10); if (w) FQ_WIN = w; }
    uint64_t at = used, a = 0, b = 0;
    unsigned long long fq_reads = 0, fq_bases = 0, rfq_reads = 0, rfq_bases = 0;
    bool passed = false; std::string msg;
    for (;;) {
        rpq_compare_in in; memset(&in, 0, sizeof in);
        const uint64_t nq = rfq.size() - at < rfq_win ? rfq.size() - at : rfq_win;
        const uint64_t n1 = r1.size() - a < FQ_WIN ? r1.size() - a : FQ_WIN, n2 = pe ? (r2.size() - b < FQ_WIN ? r2.size() - b : FQ_WIN) : 0;
        in.rfq = (const uint8_t*)rfq.data() + at; in.rfq_bytes = nq; in.rfq_mem = RPQ_MEM_HOST; in.rfq_final = at + nq == rfq.size();
        in.r1 = r1.data() + a; in.r1_len = n1; in.r2 = pe ? (r2.empty() ? "" : r2.data() + b) : NULL; in.r2_len = n2; in.fq_mem = RPQ_MEM_HOST;
        in.fq_final = (a + n1 == r1.size()) && (!pe || b + n2 == r2.size());
        in.fq_offset[0] = a; in.fq_offset[1] = b;
        rpq_compare_out res;
        if (rpq_compare(ctx, &in, &res)) error_exit(rpq_last_error(ctx));
        if (res.verdict == RPQ_CMP_NEED_FASTQ) {           /* these chunks decode to more reads than the window of text holds: more text, or fewer chunks */
            if (FQ_WIN < (3ull << 30)) { FQ_WIN = FQ_WIN * 2 < (3ull << 30) ? FQ_WIN * 2 : (3ull << 30); continue; }
            if (rfq_win <= (1ull << 20)) error_exit("compare: a batch of chunks does not fit the FASTQ window");
            rfq_win /= 2; continue;
        }
        fq_reads += res.fastq_reads; fq_bases += res.fastq_bases; rfq_reads += res.rfq_reads; rfq_bases += res.rfq_bases;
        const unsigned long long sr = pe ? rfq_reads / 2 : rfq_reads, sf = pe ? fq_reads / 2 : fq_reads;
        const char* unit = pe ? "pair" : "read";
        if (res.verdict >= RPQ_CMP_NAME && res.verdict <= RPQ_CMP_QUALITY) {
            static const char* what[4] = {"name", "sequence", "strand", "quality"};
            msg = std::string("The RFQ file and FASTQ file have different ") + what[res.verdict - RPQ_CMP_NAME] + " in the " + std::to_string(sr) + " " + unit + ". " +
                  std::string(res.rfq_field, res.rfq_field_len) + " | " + std::string(res.fastq_field, res.fastq_field_len);
            break;
        }
        if (res.verdict == RPQ_CMP_RFQ_MORE) {
            msg = "The RFQ file has more reads than the FASTQ file. The RFQ file has >= " + std::to_string(sr) + " " + unit + "s, while the FASTQ file only has " + std::to_string(sf) + " " + unit + "s";
            break;
        }
        if (res.verdict == RPQ_CMP_FASTQ_MORE) {
            msg = "The FASTQ file has more reads than the RFQ file. The FASTQ file has >= " + std::to_string(sf) + " " + unit + "s, while the RFQ file only has " + std::to_string(sr) + " " + unit + "s";
            break;
        }
        /* equal so far */
        a += res.r1_consumed; b += res.r2_consumed; at += res.rfq_consumed;
        if (in.rfq_final) at = rfq.size();                 /* whatever follows the last whole chunk is not a chunk: the reference stops there too */
        else if (res.rfq_consumed == 0) { if (rfq_win >= (2ull << 30)) error_exit("compare: a chunk does not fit the batch window"); rfq_win *= 2; continue; }
        if (at == rfq.size() && in.fq_final) { passed = true; break; }
        if (res.rfq_consumed == 0 && res.r1_consumed == 0 && !in.fq_final) {          /* the chunks are through and this window of text holds no whole record */
            if (FQ_WIN >= (3ull << 30)) error_exit("compare: a record does not fit the FASTQ window");
            FQ_WIN = FQ_WIN * 2 < (3ull << 30) ? FQ_WIN * 2 : (3ull << 30);
        }
    }
    std::string json = "{\n";
    json += passed ? "\t\"result\":\"passed\",\n" : "\t\"result\":\"failed\",\n";
    json += "\t\"msg\":\"" + msg + "\",\n";
    json += "\t\"fastq_reads\":" + std::to_string(fq_reads) + ",\n";
    json += "\t\"rfq_reads\":" + std::to_string(rfq_reads) + ",\n";
    json += "\t\"fastq_bases\":" + std::to_string(fq_bases) + ",\n";
    json += "\t\"rfq_bases\":" + std::to_string(rfq_bases) + "\n";
    json += "}\n";
    if (!o.json_compare.empty()) spill(o.json_compare, json.data(), json.size());
    fwrite(json.data(), 1, json.size(), stdout);
    rpq_destroy(ctx);
    return 0;
}

static void usage() {
    fprintf(stderr,
            "repaq_b200: repack FASTQ to a smaller binary file (.rfq) on a B200\nversion 0.5.1 (algorithm 2)\n"
            "usage: repaq_b200_cli [options]\n"
            "  -i, --in1                  input file name (.fq, .fq.gz; .rfq, .rfq.xz when decompressing)\n"
            "  -o, --out1                 output file name (.rfq, .rfq.xz; .fq, .fq.gz when decompressing)\n"
            "  -I, --in2                  read2 input file name when encoding paired-end FASTQ files\n"
            "  -O, --out2                 read2 output file name when decoding to paired-end FASTQ files\n"
            "  -c, --compress             compress input to output (the default mode)\n"
            "  -d, --decompress           decompress input to output\n"
            "  -k, --chunk                the chunk size (kilo bases) for encoding, default 1000=1000kb\n"
            "      --stdin                input from STDIN; add --interleaved_in for interleaved paired-end FASTQ\n"
            "      --stdout               write to STDOUT (paired-end data decode to interleaved FASTQ)\n"
            "      --interleaved_in       <in1> is an interleaved paired-end FASTQ\n"
            "  -v, --verify               verify the output stream (every chunk is decoded again and compared, on the GPU)\n"
            "  -f, --fast_verify          the reference verifies every tenth chunk here; this driver verifies every chunk, as -v\n"
            "  -p, --compare              compare <in1> (<in2>) with <rfq_to_compare> read by read\n"
            "  -r, --rfq_to_compare       the RFQ file to be compared with the input\n"
            "  -j, --json_compare_result  the file to store the comparison result (it is also printed on STDOUT)\n"
            "  -t, --thread               thread number for xz compression (default 1)\n"
            "  -z, --compression          xz compression level 1~9, default 3\n"
            "      --device=N             CUDA device (default 0)\n");
}

int main(int argc, char** argv) {
    if (argc == 1) { usage(); return 0; }
    if (argc == 2 && strcmp(argv[1], "--version") == 0) { printf("repaq 0.5.1\n"); return 0; }
    Opt o;
    std::vector<std::string> args(argv, argv + argc);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* lng) -> std::string {
            std::string pre = std::string("--") + lng + "=";
            if (a.compare(0, pre.size(), pre) == 0) return a.substr(pre.size());
            if (i + 1 >= argc) error_exit("option needs value: " + a);
            return argv[++i];
        };
        if (a == "-i" || a.compare(0, 6, "--in1=") == 0 || a == "--in1") o.in1 = val("in1");
        else if (a == "-I" || a.compare(0, 6, "--in2=") == 0 || a == "--in2") o.in2 = val("in2");
        else if (a == "-o" || a.compare(0, 7, "--out1=") == 0 || a == "--out1") o.out1 = val("out1");
        else if (a == "-O" || a.compare(0, 7, "--out2=") == 0 || a == "--out2") o.out2 = val("out2");
        else if (a == "-k" || a.compare(0, 8, "--chunk=") == 0 || a == "--chunk") o.k = atoi(val("chunk").c_str());
        else if (a == "-r" || a.compare(0, 17, "--rfq_to_compare=") == 0 || a == "--rfq_to_compare") o.rfq_compare = val("rfq_to_compare");
        else if (a == "-j" || a.compare(0, 22, "--json_compare_result=") == 0 || a == "--json_compare_result") o.json_compare = val("json_compare_result");
        else if (a == "-t" || a.compare(0, 9, "--thread=") == 0 || a == "--thread") o.threads = atoi(val("thread").c_str());
        else if (a == "-z" || a.compare(0, 14, "--compression=") == 0 || a == "--compression") o.level = atoi(val("compression").c_str());
        else if (a == "-p" || a == "--compare") o.compare = true;
        else if (a == "-v" || a == "--verify" || a == "-f" || a == "--fast_verify") o.verify = true;
        else if (a == "-c" || a == "--compress") o.compress = true;
        else if (a == "-d" || a == "--decompress") o.decompress = true;
        else if (a == "--interleaved_in") o.interleaved = true;
        else if (a == "--stdout") o.to_stdout = true;
        else if (a == "--stdin") o.from_stdin = true;
        else if (a.compare(0, 9, "--device=") == 0) o.device = atoi(a.c_str() + 9);
        else if (a == "-?" || a == "--help") { usage(); return 0; }
        else error_exit("unsupported option for the B200 driver: " + a);
    }
    if ((int)o.compress + (int)o.decompress + (int)o.compare > 1) error_exit("repaq can run in compress/decompress/compare mode, you can only choose any one mode.");
    if (!o.decompress && !o.compare) o.compress = true;                     /* compress is the default mode */
    o.threads = o.threads < 1 ? 1 : (o.threads > 16 ? 16 : o.threads);
    o.level = o.level < 1 ? 1 : (o.level > 9 ? 9 : o.level);
    if (o.compress && o.to_stdout && !o.out1.empty()) { fprintf(stderr, "Output to STDOUT, ignore --out1 = %s\n", o.out1.c_str()); o.out1.clear(); }
    if (o.decompress && o.from_stdin && !o.in1.empty()) { fprintf(stderr, "Input from STDIN, ignore --in1 = %s\n", o.in1.c_str()); o.in1.clear(); }
    if (o.compare && o.from_stdin && !o.rfq_compare.empty()) { fprintf(stderr, "Input from STDIN, ignore --rfq_to_compare = %s\n", o.rfq_compare.c_str()); o.rfq_compare.clear(); }
    /* Options::validate (src/options.cpp:36-111) */
    if (o.in1.empty()) {
        if (!o.in2.empty()) error_exit("read2 input is specified by <in2>, but read1 input is not specified by <in1>");
        if (o.from_stdin && !o.compare) o.in1 = "/dev/stdin"; else error_exit("Please specify input file by <in1>, or enable --stdin if you want to read STDIN");
    }
    if (o.out1.empty()) {
        if (!o.out2.empty()) error_exit("read2 output is specified by <out2>, but read1 output is not specified by <out1>");
        if (o.to_stdout) o.out1 = "/dev/stdout"; else if (!o.compare) error_exit("Please specify output file by <out1>, or enable --stdout if you want to read STDIN");
    }
    if (o.compress) {
        if (!o.out2.empty()) error_exit("In compress mode, only one RFQ output file is allowed, but you specified <out2>");
        if (is_fastq_name(o.out1)) error_exit("In compress mode, the output should not be a FASTQ file. Expect a .rfq or .rfq.xz file, but got " + o.out1);
        if (is_rfq_name(o.in1)) error_exit("In compress mode, the input should not be a RFQ file. Expect a .fq or .fq.gz file, but got " + o.in1);
        if (!o.in2.empty() && is_rfq_name(o.in2)) error_exit("In compress mode, the read2 input should not be a RFQ file. Expect a .fq or .fq.gz file, but got " + o.in2);
    }
    if (o.decompress) {
        if (!o.in2.empty()) error_exit("In decompress mode, only one RFQ input file is allowed, but you specified <in2>");
        if (is_fastq_name(o.in1)) error_exit("In decompress mode, the input should not be a FASTQ file. Expect a .rfq or .rfq.xz file, but got " + o.in1);
        if (is_rfq_name(o.out1)) error_exit("In decompress mode, the output should not be a RFQ file. Expect a .fq or .fq.gz file, but got " + o.out1);
        if (!o.out2.empty() && is_rfq_name(o.out2)) error_exit("In decompress mode, the read2 output should not be a RFQ file. Expect a .fq or .fq.gz file, but got " + o.out2);
    }
    if (o.compare) {
        if (o.from_stdin) o.rfq_compare = "/dev/stdin";
        if (o.rfq_compare.empty()) error_exit("In compare mode, you should specify the RFQ file to compare by <rfq_to_compare>");
        if (!o.out1.empty() || !o.out2.empty()) error_exit("In compare mode, you cannot specify the output by <out1> or <out2>");
    }
    const long long cs = (long long)(o.k < 100 ? 100 : o.k) * 1000;
    if (cs > 500000000) error_exit("chunk size cannot be greater than 500,000 kb");
    if ((ends_with(o.in1, ".xz") || ends_with(o.rfq_compare, ".xz")) && o.from_stdin) error_exit("STDIN cannot be read when the input is a .xz file");
    if (ends_with(o.out1, ".xz") && o.to_stdout) error_exit("STDOUT cannot be written when the output is a .xz file");

    /* ---- .xz: this program again, piped through the xz executable (src/main.cpp:133-178) */
    auto rerun = [&](const std::string& drop_value, const std::vector<std::string>& drop_flags, const std::string& prefix, const std::string& suffix) {
        std::string cmd = prefix;
        for (size_t i = 0; i < args.size(); i++) {
            const std::string& a = args[i];
            bool flag = false;
            for (const std::string& fl : drop_flags) if (a == fl) flag = true;
            if (flag) { i++; continue; }                                    /* the flag and its value */
            bool eq = false;
            for (const std::string& fl : drop_flags) if (fl.size() > 2 && a.compare(0, fl.size() + 1, fl + "=") == 0) eq = true;
            if (eq || a == drop_value) continue;
            cmd += "'" + a + "' ";
        }
        cmd += suffix;
        const int ret = system(cmd.c_str());
        if (ret != 0) error_exit("failed to call xz, please confirm that xz is installed in your system");
        return 0;
    };
    if (o.compress && ends_with(o.out1, ".xz")) {
        std::string xz = "--stdout | xz -z -c";
        if (o.threads > 1) xz += " -T" + std::to_string(o.threads);
        if (o.level <= 4) xz += " -" + std::to_string(o.level + 5);          /* equal to xz -6/7/8/9 */
        else {
            unsigned long long dict = (64ull * 1024 * 1024) << (o.level - 4);
            if (o.level == 9) dict = 1536ull * 1024 * 1024;
            xz += " --lzma2=\"dict=" + std::to_string(dict) + "\"";
        }
        if (o.level >= 4 && o.threads > 1) fprintf(stderr, "WARNING: when repaq compression level is >= 4, only single thread will be used for xz. Your options: compression = %d, thread = %d\n", o.level, o.threads);
        return rerun(o.out1, {"-o", "--out1"}, "", xz + " > '" + o.out1 + "'");
    }
    if (o.decompress && ends_with(o.in1, ".xz")) return rerun(o.in1, {"-i", "--in1"}, "xz -d -c '" + o.in1 + "' | ", "--stdin");
    if (o.compare && ends_with(o.rfq_compare, ".xz")) return rerun(o.rfq_compare, {"-r", "--rfq_to_compare"}, "xz -d -c '" + o.rfq_compare + "' | ", "--stdin");

    if (o.compare) return do_compare(o);
    if (o.compress) return do_compress(o);
    return do_decompress(o);
}
