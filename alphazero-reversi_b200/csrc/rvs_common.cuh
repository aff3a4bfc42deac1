// rvs_common.cuh -- error plumbing, launch accounting and host<->device staging shared by
// the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "../../include/rvs_b200.h"

namespace rvs {

constexpr int kNumSMs = 148;  // B200: grids are sized in multiples of this

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define RVS_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return ::rvs::fail(-100 - (int)_e, "%s failed: %s (%s:%d)", #expr,              \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

// launch + count + check.  Every kernel of this library goes through here so that
// rvs_launch_count() is the exact number of our kernels that ran.
#define RVS_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        ::rvs::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
        RVS_CUDA(cudaGetLastError());                                                        \
    } while (0)

// The same with programmatic dependent launch: the kernel may become resident while its predecessor in the stream is
// still running (after the predecessor's griddepcontrol.launch_dependents, or at its end) and MUST execute
// pdl_grid_wait() before it touches anything the predecessor reads or writes.  Hides the launch latency between the
// small kernels of an NN search wave (tree step -> first layer -> tower -> heads -> tree step ...).
#define RVS_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                               \
    do {                                                                                     \
        cudaLaunchConfig_t cfg_ = {};                                                        \
        cfg_.gridDim = dim3(grid_);                                                          \
        cfg_.blockDim = dim3(block_);                                                        \
        cfg_.dynamicSmemBytes = (smem_);                                                     \
        cfg_.stream = (stream_);                                                             \
        cudaLaunchAttribute attr_[1];                                                        \
        attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                    \
        attr_[0].val.programmaticStreamSerializationAllowed = 1;                             \
        cfg_.attrs = attr_;                                                                  \
        cfg_.numAttrs = 1;                                                                   \
        RVS_CUDA(cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__));                            \
        ::rvs::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
    } while (0)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// RVS_MEM_HOST_ASYNC is only honoured by rvs_engine_set_positions / rvs_engine_root_visits; everywhere else it
// is RVS_MEM_HOST (staged copies under the staging lock, outputs complete on return)
#define RVS_NORMALISE_MEM(mem)                                                                   \
    do {                                                                                         \
        if ((mem) == RVS_MEM_HOST_ASYNC) (mem) = RVS_MEM_HOST;                                   \
        if ((mem) != RVS_MEM_DEVICE && (mem) != RVS_MEM_HOST) return ::rvs::fail(-1, "bad mem %d", (int)(mem)); \
    } while (0)

inline int grid_for(int64_t n, int block, int ctas_per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Grow-only device scratch used to stage RVS_MEM_HOST arguments of the stateless calls.
// Guarded by g_stage_mu (stateless host calls are serialised).
extern std::mutex g_stage_mu;
int stage_get(int slot, size_t bytes, void** out);

// A bulk argument that is either used in place (device) or staged (host).
struct Arg {
    void* dev = nullptr;
    void* host = nullptr;
    size_t bytes = 0;
};

inline int arg_in(Arg& a, const void* p, size_t bytes, int mem, int slot, cudaStream_t s) {
    a.bytes = bytes;
    if (p == nullptr) { a.dev = nullptr; return 0; }
    if (mem == RVS_MEM_DEVICE) { a.dev = const_cast<void*>(p); return 0; }
    a.host = const_cast<void*>(p);
    int rc = stage_get(slot, bytes, &a.dev);
    if (rc) return rc;
    RVS_CUDA(cudaMemcpyAsync(a.dev, p, bytes, cudaMemcpyHostToDevice, s));
    return 0;
}
inline int arg_out(Arg& a, void* p, size_t bytes, int mem, int slot) {
    a.bytes = bytes;
    if (p == nullptr) { a.dev = nullptr; return 0; }
    if (mem == RVS_MEM_DEVICE) { a.dev = p; return 0; }
    a.host = p;
    return stage_get(slot, bytes, &a.dev);
}
inline int arg_back(Arg& a, cudaStream_t s) {
    if (a.host && a.dev) RVS_CUDA(cudaMemcpyAsync(a.host, a.dev, a.bytes, cudaMemcpyDeviceToHost, s));
    return 0;
}

}  // namespace rvs
