// cuda_runtime.h (SIMT emulator) -- TEST INFRASTRUCTURE, not product code.
//
// A stand-in for the CUDA headers that lets g++ compile fluxpy_b200/csrc/*.cu(h) for the host, so
// that the kernels' own source can be executed on a machine without a GPU (tools/simt/README.md).
// Every CUDA thread is a fiber; the 32 fibers of a warp meet at every *_sync intrinsic and every
// thread of a block at __syncthreads(), so warp-collective code runs with CUDA's semantics and a
// collective that not every named lane reaches is reported as a deadlock instead of passing
// silently.  Nothing under fluxpy_b200/ knows about this file; only tests/ builds and loads the
// emulated library (tools/simt/build_emu.py).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#define FB_EMU 1

// ---- qualifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define EMU_NOINLINE __attribute__((noinline)) // build_emu.py rewrites __noinline__ (libstdc++ uses that token)
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
// extern __shared__ T name[];  is rewritten by build_emu.py into  T *name = emu::dyn_smem<T>();

// ---- vector types -------------------------------------------------------------------------------
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {
// ---- engine (simt_engine.cpp) ---------------------------------------------------------------------
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lane_linear; // thread index within the block
};
extern thread_local ThreadCtx tctx;
// All live lanes of the calling lane's warp exchange one 64-bit value.  Returns the 32 slots (valid until
// the collective after the next one) and the mask of lanes that took part.
const uint64_t *warp_exchange(uint64_t v, unsigned *present);
void block_barrier();
void *dyn_smem_raw();
template <class T> inline T *dyn_smem() { return reinterpret_cast<T *>(dyn_smem_raw()); }
void run_grid(dim3 grid, dim3 block, size_t smem, void (*invoke)(void *), void *closure);
int num_workers();
[[noreturn]] void fail(const char *msg);

template <class T> inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle of a wide type");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T> inline T from_bits(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}
inline int lane_id() { return tctx.lane_linear & 31; }
} // namespace emu

#define threadIdx (emu::tctx.tid)
#define blockIdx (emu::tctx.bid)
#define blockDim (emu::tctx.bdim)
#define gridDim (emu::tctx.gdim)
constexpr int warpSize = 32;

// ---- warp collectives -----------------------------------------------------------------------------
inline void emu_check_mask(unsigned mask, unsigned present) {
    // the kernels here name the full warp; every named lane that is still alive must have arrived
    if ((mask & present) != present) emu::fail("warp collective: a participating lane is not named in the mask");
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned present;
    const uint64_t *s = emu::warp_exchange(emu::to_bits(v), &present);
    emu_check_mask(mask, present);
    const int lane = emu::lane_id();
    const int l = (lane & ~(width - 1)) | (src & (width - 1));
    return ((present >> l) & 1u) ? emu::from_bits<T>(s[l]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int o, int width = 32) {
    unsigned present;
    const uint64_t *s = emu::warp_exchange(emu::to_bits(v), &present);
    emu_check_mask(mask, present);
    const int l = emu::lane_id() ^ o;
    return (l < 32 && ((present >> l) & 1u)) ? emu::from_bits<T>(s[l]) : v;
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned o, int width = 32) {
    unsigned present;
    const uint64_t *s = emu::warp_exchange(emu::to_bits(v), &present);
    emu_check_mask(mask, present);
    const int l = emu::lane_id() - (int)o;
    return (l >= 0 && ((present >> l) & 1u)) ? emu::from_bits<T>(s[l]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned o, int width = 32) {
    unsigned present;
    const uint64_t *s = emu::warp_exchange(emu::to_bits(v), &present);
    emu_check_mask(mask, present);
    const int l = emu::lane_id() + (int)o;
    return (l < 32 && ((present >> l) & 1u)) ? emu::from_bits<T>(s[l]) : v;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned present;
    const uint64_t *s = emu::warp_exchange(pred ? 1u : 0u, &present);
    emu_check_mask(mask, present);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((present >> l) & 1u) && s[l]) r |= 1u << l;
    return r & mask;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    unsigned present;
    const uint64_t mine = emu::to_bits(v);
    const uint64_t *s = emu::warp_exchange(mine, &present);
    emu_check_mask(mask, present);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((present >> l) & 1u) && s[l] == mine) r |= 1u << l;
    return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    unsigned present;
    emu::warp_exchange(0, &present);
    emu_check_mask(mask, present);
}
inline void __syncthreads() { emu::block_barrier(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---- bit intrinsics -------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
// position of the offset-th set bit of mask counting from bit `base` (inclusive): upwards for offset > 0,
// downwards for offset < 0; offset == 0 -> base if that bit is set; 0xffffffff when there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
    if (offset > 0) {
        for (unsigned b = base; b < 32; ++b)
            if (((mask >> b) & 1u) && --offset == 0) return b;
    } else {
        for (int b = (int)base; b >= 0; --b)
            if (((mask >> b) & 1u) && ++offset == 0) return (unsigned)b;
    }
    return 0xffffffffu;
}
inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
inline float __int_as_float(int i) { return emu::from_bits<float>(emu::to_bits(i)); }
inline float __uint_as_float(unsigned i) { return emu::from_bits<float>(emu::to_bits(i)); }
inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
inline double __longlong_as_double(long long i) { return emu::from_bits<double>(emu::to_bits(i)); }

// ---- arithmetic with explicit rounding (the translation unit is compiled with -ffp-contract=off) ----
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __double2float_rn(double d) { return (float)d; }
using std::isfinite;
using std::isnan;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---- memory ---------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
template <class T> inline void __stcg(T *p, T v) { *p = v; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

// ---- atomics (blocks run on several host threads) -----------------------------------------------------
template <class T> inline T emu_atomic_add(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return emu_atomic_add(p, v); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return emu_atomic_add(p, v); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return emu_atomic_add(p, v); }
template <class T> inline T emu_atomic_fadd(T *p, T v) {
    using U = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;
    U *q = reinterpret_cast<U *>(p);
    U old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
    while (true) {
        T cur;
        memcpy(&cur, &old, sizeof(T));
        const T nv = cur + v;
        U nu;
        memcpy(&nu, &nv, sizeof(T));
        if (__atomic_compare_exchange_n(q, &old, nu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return cur;
    }
}
inline float atomicAdd(float *p, float v) { return emu_atomic_fadd(p, v); }
inline double atomicAdd(double *p, double v) { return emu_atomic_fadd(p, v); }
template <class T, class F> inline T emu_atomic_rmw(T *p, F f) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (!__atomic_compare_exchange_n(p, &old, f(old), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMin(unsigned *p, unsigned v) { return emu_atomic_rmw(p, [v](unsigned o) { return o < v ? o : v; }); }
inline unsigned atomicMax(unsigned *p, unsigned v) { return emu_atomic_rmw(p, [v](unsigned o) { return o > v ? o : v; }); }
inline int atomicMin(int *p, int v) { return emu_atomic_rmw(p, [v](int o) { return o < v ? o : v; }); }
inline int atomicMax(int *p, int v) { return emu_atomic_rmw(p, [v](int o) { return o > v ? o : v; }); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    return emu_atomic_rmw(p, [v](unsigned long long o) { return o > v ? o : v; });
}
inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ---- runtime API: one synchronous "device", every stream executes immediately ------------------------
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorMemoryAllocation = 2;
struct emuStream {
    int priority;
};
typedef emuStream *cudaStream_t;
struct emuEvent {
    std::chrono::steady_clock::time_point t;
};
typedef emuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2;
typedef void (*cudaHostFn_t)(void *);

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : 101; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
    *v = a == cudaDevAttrMultiProcessorCount ? emu::num_workers() : 227 * 1024;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
// Device allocations carry a 256-byte red zone on either side (checked when the block is freed: a kernel
// that wrote out of bounds aborts the process with a message) and start out poisoned, not zeroed.
namespace emu {
constexpr size_t kRedZone = 256;
constexpr unsigned char kRedByte = 0xA5;
struct AllocHeader {
    size_t bytes;
    uint64_t magic;
};
inline void check_redzones(void *user) {
    unsigned char *base = static_cast<unsigned char *>(user) - kRedZone;
    AllocHeader h;
    memcpy(&h, base, sizeof(h));
    if (h.magic != 0xFB200E5Dull) fail("cudaFree of a pointer cudaMalloc did not return (or its header was overwritten)");
    for (size_t k = sizeof(h); k < kRedZone; ++k)
        if (base[k] != kRedByte) fail("out-of-bounds write BELOW a device allocation");
    const unsigned char *tail = static_cast<unsigned char *>(user) + h.bytes;
    for (size_t k = 0; k < kRedZone; ++k)
        if (tail[k] != kRedByte) fail("out-of-bounds write ABOVE a device allocation");
}
} // namespace emu
inline cudaError_t cudaMalloc(void **p, size_t bytes) {
    const size_t total = (bytes + 2 * emu::kRedZone + 255) / 256 * 256;
    unsigned char *base = static_cast<unsigned char *>(aligned_alloc(256, total));
    if (!base) return cudaErrorMemoryAllocation;
    memset(base, emu::kRedByte, emu::kRedZone);
    const emu::AllocHeader h{bytes, 0xFB200E5Dull};
    memcpy(base, &h, sizeof(h));
    memset(base + emu::kRedZone, 0xCD, bytes); // device memory is uninitialised: code relying on zeros shows up
    memset(base + emu::kRedZone + bytes, emu::kRedByte, emu::kRedZone);
    *p = base + emu::kRedZone;
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    emu::check_redzones(p);
    free(static_cast<unsigned char *>(p) - emu::kRedZone);
    return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) {
    *p = aligned_alloc(256, (bytes + 255) / 256 * 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emuStream{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int prio) { *s = new emuStream{prio}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void *arg) { fn(arg); return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- kernel launch:  k<<<grid, block, smem, stream>>>(args...)  ->  emu::Launch(grid, block, smem, stream)(k, args...)
namespace emu {
struct Launch {
    dim3 grid, block;
    size_t smem;
    Launch(dim3 g, dim3 b, size_t s = 0, cudaStream_t = nullptr) : grid(g), block(b), smem(s) {}
    template <class... P, class... A> void operator()(void (*kernel)(P...), A &&...args) {
        auto body = [&]() { kernel(static_cast<P>(args)...); };
        using B = decltype(body);
        run_grid(grid, block, smem, [](void *c) { (*static_cast<B *>(c))(); }, &body);
    }
};
} // namespace emu
