/*
 * cuda_emu.h - TEST INFRASTRUCTURE ONLY: a lock-step SIMT emulator for running the .cu kernels of
 * repaq_b200/csrc on a CPU, so kernel LOGIC can be unit-tested (and debugged with gdb/ASan) in the
 * GPU-less dev container.  It is compiled only into tests/emu/librepaq_emu.so (-DRPQ_EMU); the product
 * library librepaq_b200.so is built by nvcc without it and has no CPU path.
 *
 * Model: every CUDA thread is a fiber (own stack, hand-rolled x86-64 context switch); all fibers of a block
 * live on one OS thread, so `__shared__` becomes `static thread_local`; blocks are handed out in index order
 * to a few OS worker threads.  Warp collectives and __syncthreads are rendezvous points between fibers.
 * Limitations (the kernels are written to respect them): collectives must be called with a full mask by all
 * live lanes of the warp (exited lanes are fine); blockDim is 1-D.
 */
#pragma once
#ifndef RPQ_EMU
#error "cuda_emu.h is only for the -DRPQ_EMU test build"
#endif
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3_emu { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uchar4 { unsigned char x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef int cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

namespace emu {

extern "C" void emu_switch(void** save_sp, void* load_sp);

constexpr int kStackBytes = 128 * 1024;
constexpr int kMaxThreads = 1024;

struct Warp {
    uint64_t slot[2][32];
    uint32_t arrived[2];
    long long buf_gen[2];
    uint32_t alive;
    long long lane_gen[32];
};

struct Block {
    dim3 bid, bdim, gdim;
    int nthreads;
    void* sched_sp;
    void* fiber_sp[kMaxThreads];
    bool done[kMaxThreads];
    int cur;
    int alive;
    long long bar_gen[kMaxThreads];
    int bar_arrived[2];
    long long bar_buf_gen[2];
    Warp warps[kMaxThreads / 32];
    unsigned char* dyn_smem;
    const std::function<void()>* body;
    char* stacks;
};

extern thread_local Block* g_blk;
extern thread_local uint32_t g_part;   /* participants of the calling fiber's latest collective */

inline Block& blk() { return *g_blk; }
inline void yield() { Block& b = blk(); emu_switch(&b.fiber_sp[b.cur], b.sched_sp); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

struct TidProxy { struct V { unsigned x, y, z; }; };
inline unsigned tid_x() { return (unsigned)blk().cur; }

/* rendezvous of all live lanes of the calling lane's warp; returns the slot array of this round */
inline const uint64_t* warp_exchange(uint64_t v) {
    Block& b = blk();
    const int w = b.cur >> 5, l = b.cur & 31;
    Warp& W = b.warps[w];
    const long long g = W.lane_gen[l]++;
    const int buf = (int)(g & 1);
    if (W.buf_gen[buf] != g) { W.buf_gen[buf] = g; W.arrived[buf] = 0; }
    W.slot[buf][l] = v;
    W.arrived[buf] |= 1u << l;
    while ((W.arrived[buf] & W.alive) != W.alive) yield();
    g_part = W.arrived[buf];      /* the lanes that took part in this round (lanes may exit right after it) */
    return W.slot[buf];
}
inline uint32_t warp_alive() { return g_part; }

}  // namespace emu

struct emu_tid { operator unsigned() const { return emu::tid_x(); } };
struct emu_idx3 { emu_tid x; static constexpr unsigned y = 0, z = 0; };
struct emu_bid3 { struct X { operator unsigned() const { return emu::blk().bid.x; } } x; struct Y { operator unsigned() const { return emu::blk().bid.y; } } y; static constexpr unsigned z = 0; };
struct emu_bdim3 { struct X { operator unsigned() const { return emu::blk().bdim.x; } } x; static constexpr unsigned y = 1, z = 1; };
struct emu_gdim3 { struct X { operator unsigned() const { return emu::blk().gdim.x; } } x; struct Y { operator unsigned() const { return emu::blk().gdim.y; } } y; static constexpr unsigned z = 1; };
static const emu_idx3 threadIdx;
static const emu_bid3 blockIdx;
static const emu_bdim3 blockDim;
static const emu_gdim3 gridDim;
constexpr int warpSize = 32;

/* ---------------------------------------------------------------- synchronisation ---- */
static inline void __syncthreads() {
    emu::Block& b = emu::blk();
    const long long g = b.bar_gen[b.cur]++;
    const int buf = (int)(g & 1);
    if (b.bar_buf_gen[buf] != g) { b.bar_buf_gen[buf] = g; b.bar_arrived[buf] = 0; }
    b.bar_arrived[buf]++;
    while (b.bar_arrived[buf] < b.alive) emu::yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; emu::warp_exchange(0); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned __activemask() { emu::Block& b = emu::blk(); return b.warps[b.cur >> 5].alive; }

/* ---------------------------------------------------------------- warp collectives ---- */
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    (void)mask;
    const uint64_t* s = emu::warp_exchange(pred ? 1 : 0);
    unsigned alive = emu::warp_alive(), r = 0;
    for (int i = 0; i < 32; i++) if (((alive >> i) & 1) && s[i]) r |= 1u << i;
    return r;
}
static inline int __all_sync(unsigned mask, int pred) { unsigned r = __ballot_sync(mask, pred); unsigned a = emu::warp_alive(); return (r & a) == a; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    (void)mask; static_assert(sizeof(T) <= 8, "shfl");
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(raw);
    int lane = (int)(emu::tid_x() & 31);
    int base = lane & ~(width - 1);
    uint64_t r = s[base + (src & (width - 1))];
    T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    (void)mask; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(raw);
    int lane = (int)(emu::tid_x() & 31);
    int src = lane - (int)d;
    if (src < (lane & ~(width - 1))) src = lane;
    uint64_t r = s[src]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    (void)mask; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(raw);
    int lane = (int)(emu::tid_x() & 31);
    int src = lane + (int)d;
    if (src > (lane | (width - 1))) src = lane;
    uint64_t r = s[src]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    (void)mask; (void)width; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(raw);
    int lane = (int)(emu::tid_x() & 31);
    uint64_t r = s[lane ^ x]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    (void)mask; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t* s = emu::warp_exchange(raw);
    unsigned alive = emu::warp_alive(), r = 0;
    for (int i = 0; i < 32; i++) if (((alive >> i) & 1) && s[i] == raw) r |= 1u << i;
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { (void)mask; const uint64_t* s = emu::warp_exchange(v); unsigned a = emu::warp_alive(), r = 0; for (int i = 0; i < 32; i++) if ((a >> i) & 1) r += (unsigned)s[i]; return r; }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { (void)mask; const uint64_t* s = emu::warp_exchange(v); unsigned a = emu::warp_alive(), r = 0xffffffffu; for (int i = 0; i < 32; i++) if ((a >> i) & 1) r = std::min(r, (unsigned)s[i]); return r; }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { (void)mask; const uint64_t* s = emu::warp_exchange(v); unsigned a = emu::warp_alive(), r = 0; for (int i = 0; i < 32; i++) if ((a >> i) & 1) r = std::max(r, (unsigned)s[i]); return r; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { (void)mask; const uint64_t* s = emu::warp_exchange(v); unsigned a = emu::warp_alive(), r = 0; for (int i = 0; i < 32; i++) if ((a >> i) & 1) r |= (unsigned)s[i]; return r; }
static inline unsigned __reduce_and_sync(unsigned mask, unsigned v) { (void)mask; const uint64_t* s = emu::warp_exchange(v); unsigned a = emu::warp_alive(), r = 0xffffffffu; for (int i = 0; i < 32; i++) if ((a >> i) & 1) r &= (unsigned)s[i]; return r; }

/* ---------------------------------------------------------------- bit intrinsics ---- */
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if ((v >> i) & 1) r |= 1u << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (sh & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (sh & 31)) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned s = (sel >> (4 * i)) & 0xF; unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xFF; if (s & 8) byte = (byte & 0x80) ? 0xFF : 0; r |= byte << (8 * i); }
    return r;
}
/* find the n-th (0-based offset semantics of CUDA's __fns with base 0, positive offset n+1) set bit */
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset > 0) { for (unsigned i = base; i < 32; i++) if ((mask >> i) & 1) { if (--offset == 0) return i; } return 0xffffffffu; }
    if (offset < 0) { for (int i = (int)base; i >= 0; i--) if ((mask >> i) & 1) { if (++offset == 0) return (unsigned)i; } return 0xffffffffu; }
    return ((mask >> base) & 1) ? base : 0xffffffffu;
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r; }
static inline unsigned __vcmpne4(unsigned a, unsigned b) { return ~__vcmpeq4(a, b); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
using std::max;
using std::min;
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

/* ---------------------------------------------------------------- atomics ---- */
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <typename T> static inline T atomicMin(T* p, T v) { T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }
template <typename T> static inline T atomicMax(T* p, T v) { T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }

#define RPQ_EMU_SPIN_HINT() do { emu::yield(); std::this_thread::yield(); } while (0)
