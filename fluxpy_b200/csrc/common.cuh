// common.cuh -- error handling, small device helpers (libfluxb200, sm_100a)
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace fluxb200 {

extern thread_local std::string g_last_error;

struct CudaError {
    std::string msg;
};

inline void set_error(const std::string &s) { g_last_error = s; }

#define FB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            char _buf[512];                                                                \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,     \
                     cudaGetErrorString(_e));                                              \
            throw fluxb200::CudaError{_buf};                                               \
        }                                                                                  \
    } while (0)

#define FB_REQUIRE(cond, text)                                                             \
    do {                                                                                   \
        if (!(cond)) throw fluxb200::CudaError{std::string(text)};                         \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        const size_t old = cap;
        if (p) FB_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        // a buffer that grows again grows by half (up to 8 GB extra): a few per cent at a time would free and
        // allocate gigabytes call after call
        size_t want = bytes + bytes / 8 + 256;
        if (old) want = std::max(want, old + std::min<size_t>(old / 2, (size_t)8 << 30));
        FB_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// grow-only pinned host buffer (small staging: per-sub-slab counts / totals)
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) FB_CUDA(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        FB_CUDA(cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped)); // kernels may store into it
        cap = want;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

} // namespace fluxb200
