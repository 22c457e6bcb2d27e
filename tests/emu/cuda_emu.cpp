/* cuda_emu.cpp - fiber scheduler of the SIMT emulator (TEST INFRASTRUCTURE ONLY, see cuda_emu.h). */
#define RPQ_EMU 1
#include "cuda_emu.h"

#include <mutex>

/* void emu_switch(void** save_sp, void* load_sp): save callee-saved registers on the current stack, publish the
 * stack pointer, adopt the other stack and resume there. */
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

thread_local Block* g_blk = nullptr;
thread_local uint32_t g_part = 0;

static void fiber_entry() {
    Block& b = blk();
    (*b.body)();
    /* thread exit: leave the warp and the block, wake nobody explicitly (waiters poll) */
    const int t = b.cur;
    b.done[t] = true;
    b.alive--;
    b.warps[t >> 5].alive &= ~(1u << (t & 31));
    emu_switch(&b.fiber_sp[t], b.sched_sp);
    abort();
}

static void run_block(Block& b) {
    g_blk = &b;
    const int n = b.nthreads;
    b.alive = n;
    memset(b.bar_gen, 0, sizeof(long long) * n);
    b.bar_buf_gen[0] = b.bar_buf_gen[1] = -1;
    for (int w = 0; w < (n + 31) / 32; w++) {
        Warp& W = b.warps[w];
        memset(&W, 0, sizeof W);
        W.buf_gen[0] = W.buf_gen[1] = -1;
        int lanes = std::min(32, n - 32 * w);
        W.alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1);
    }
    for (int t = 0; t < n; t++) {
        b.done[t] = false;
        /* initial frame: 6 callee-saved slots, then the return address = fiber_entry; keep (rsp+8) % 16 == 0 at entry */
        uintptr_t top = (uintptr_t)(b.stacks + (size_t)(t + 1) * kStackBytes);
        top &= ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                 /* fake return address of fiber_entry (alignment) */
        *--sp = (void*)&fiber_entry;
        for (int i = 0; i < 6; i++) *--sp = nullptr;
        b.fiber_sp[t] = (void*)sp;
    }
    int remaining = n;
    while (remaining > 0) {
        remaining = 0;
        for (int t = 0; t < n; t++) {
            if (b.done[t]) continue;
            b.cur = t;
            emu_switch(&b.sched_sp, b.fiber_sp[t]);
            if (!b.done[t]) remaining++;
        }
    }
    g_blk = nullptr;
}

static int n_workers() {
    const char* e = getenv("RPQ_EMU_THREADS");
    int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(n, 16));
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const long long total = (long long)grid.x * grid.y * grid.z;
    if (total == 0) return;
    assert(block.y == 1 && block.z == 1 && block.x <= (unsigned)kMaxThreads);
    std::atomic<long long> next{0};
    auto worker = [&]() {
        char* stacks = (char*)mmap(nullptr, (size_t)kMaxThreads * kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        assert(stacks != MAP_FAILED);
        Block* b = new Block();
        b->stacks = stacks;
        b->dyn_smem = (unsigned char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
        b->body = &body;
        b->bdim = block; b->gdim = grid; b->nthreads = (int)block.x;
        for (;;) {
            long long i = next.fetch_add(1);
            if (i >= total) break;
            b->bid = dim3((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((long long)grid.x * grid.y)));
            run_block(*b);
        }
        free(b->dyn_smem);
        delete b;
        munmap(stacks, (size_t)kMaxThreads * kStackBytes);
    };
    int nw = (int)std::min<long long>(n_workers(), total);
    if (nw <= 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nw; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

}  // namespace emu
