/*
 * rpq_rt.h - the thin runtime layer under the kernels: memory, streams, launches, device-wide scans.
 *
 * Product build (nvcc, sm_100a): CUDA runtime + CUB.  There is no CPU path in the product.
 * -DRPQ_EMU (tests/emu only): the same kernel sources run under the lock-step SIMT emulator
 * tests/emu/cuda_emu.h so their logic can be unit-tested without a GPU.
 */
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef RPQ_EMU
#include "../../tests/emu/cuda_emu.h"
#define RPQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#define RPQ_DYN_SMEM(name) unsigned char* name = emu::blk().dyn_smem
#define RPQ_SPIN_HINT() RPQ_EMU_SPIN_HINT()
#else
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#define RPQ_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define RPQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define RPQ_SPIN_HINT() __nanosleep(20)
#endif

namespace rpq {

struct RtError { int code; char msg[256]; };

#ifdef RPQ_EMU
inline int rt_device_count() { return 1; }
inline int rt_set_device(int) { return 0; }
inline void* rt_malloc_device(size_t n) { return malloc(n ? n : 1); }
inline void rt_free_device(void* p) { free(p); }
inline void* rt_malloc_pinned(size_t n) { return malloc(n ? n : 1); }
inline void rt_free_pinned(void* p) { free(p); }
inline int rt_memcpy_h2d(void* d, const void* s, size_t n, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
inline int rt_memcpy_d2h(void* d, const void* s, size_t n, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
inline int rt_memcpy_d2d(void* d, const void* s, size_t n, cudaStream_t) { if (n) memmove(d, s, n); return 0; }
inline int rt_memset(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
inline int rt_stream_create(cudaStream_t* s) { *s = 0; return 0; }
inline int rt_stream_create_high_priority(cudaStream_t* s) { *s = 0; return 0; }
inline void rt_stream_destroy(cudaStream_t) {}
inline int rt_stream_sync(cudaStream_t) { return 0; }
inline int rt_last_error(char* buf, size_t n) { (void)buf; (void)n; return 0; }
inline int rt_sm_count() { return 8; }
inline bool rt_is_device_ptr(const void*) { return false; }
#else
inline int rt_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) { fprintf(stderr, "repaq_b200: CUDA error in %s: %s\n", what, cudaGetErrorString(e)); return -1; }
    return 0;
}
inline int rt_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
inline int rt_set_device(int d) { return rt_check(cudaSetDevice(d), "cudaSetDevice"); }
inline void* rt_malloc_device(size_t n) { void* p = nullptr; if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) return nullptr; return p; }
inline void rt_free_device(void* p) { if (p) cudaFree(p); }
inline void* rt_malloc_pinned(size_t n) { void* p = nullptr; if (cudaMallocHost(&p, n ? n : 1) != cudaSuccess) return nullptr; return p; }
inline void rt_free_pinned(void* p) { if (p) cudaFreeHost(p); }
inline int rt_memcpy_h2d(void* d, const void* s, size_t n, cudaStream_t st) { return n ? rt_check(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st), "H2D") : 0; }
inline int rt_memcpy_d2h(void* d, const void* s, size_t n, cudaStream_t st) { return n ? rt_check(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st), "D2H") : 0; }
inline int rt_memcpy_d2d(void* d, const void* s, size_t n, cudaStream_t st) { return n ? rt_check(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st), "D2D") : 0; }
inline int rt_memset(void* d, int v, size_t n, cudaStream_t st) { return n ? rt_check(cudaMemsetAsync(d, v, n, st), "memset") : 0; }
inline int rt_stream_create(cudaStream_t* s) { return rt_check(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking), "stream create"); }
/* a stream whose blocks are dispatched before the pending blocks of normal streams */
inline int rt_stream_create_high_priority(cudaStream_t* s) {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    return rt_check(cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest), "stream create");
}
inline void rt_stream_destroy(cudaStream_t s) { cudaStreamDestroy(s); }
inline int rt_stream_sync(cudaStream_t s) { return rt_check(cudaStreamSynchronize(s), "stream sync"); }
inline int rt_last_error(char* buf, size_t n) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    snprintf(buf, n, "CUDA: %s", cudaGetErrorString(e));
    return -1;
}
inline int rt_sm_count() { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }
inline bool rt_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
#endif

}  // namespace rpq

/* ---- events (device timing on the context's stream) and the one device-wide scan the path needs */
namespace rpq {
#ifdef RPQ_EMU
}  // namespace rpq
#include <chrono>
namespace rpq {
struct RtEvent { double t; };
inline void rt_event_create(RtEvent* e) { e->t = 0; }
inline void rt_event_destroy(RtEvent*) {}
inline void rt_event_record(RtEvent* e, cudaStream_t) { e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline float rt_event_ms(const RtEvent& a, const RtEvent& b) { return (float)(b.t - a.t); }
inline void rt_stream_wait_event(cudaStream_t, RtEvent*) {}
inline void rt_event_sync(RtEvent*) {}
inline size_t rt_scan_tmp_bytes(size_t) { return 16; }
inline int rt_inclusive_sum_u32_u64(const uint32_t* in, unsigned long long* out, size_t n, void*, size_t, cudaStream_t) {
    unsigned long long acc = 0;
    for (size_t i = 0; i < n; i++) { acc += in[i]; out[i] = acc; }
    return 0;
}
#else
struct RtEvent { cudaEvent_t e; };
inline void rt_event_create(RtEvent* e) { cudaEventCreate(&e->e); }
inline void rt_event_destroy(RtEvent* e) { cudaEventDestroy(e->e); }
inline void rt_event_record(RtEvent* e, cudaStream_t s) { cudaEventRecord(e->e, s); }
inline float rt_event_ms(const RtEvent& a, const RtEvent& b) { float ms = 0; cudaEventElapsedTime(&ms, a.e, b.e); return ms; }
inline void rt_stream_wait_event(cudaStream_t s, RtEvent* e) { cudaStreamWaitEvent(s, e->e, 0); }
inline void rt_event_sync(RtEvent* e) { cudaEventSynchronize(e->e); }
struct RtCastU64 { __host__ __device__ unsigned long long operator()(uint32_t v) const { return v; } };
inline size_t rt_scan_tmp_bytes(size_t n) {
    size_t bytes = 0;
    cub::TransformInputIterator<unsigned long long, RtCastU64, const uint32_t*> it(nullptr, RtCastU64());
    cub::DeviceScan::InclusiveSum(nullptr, bytes, it, (unsigned long long*)nullptr, (int)n);
    return bytes + 256;
}
inline int rt_inclusive_sum_u32_u64(const uint32_t* in, unsigned long long* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s) {
    cub::TransformInputIterator<unsigned long long, RtCastU64, const uint32_t*> it(in, RtCastU64());
    return rt_check(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, it, out, (int)n, s), "DeviceScan");
}
#endif
}  // namespace rpq
