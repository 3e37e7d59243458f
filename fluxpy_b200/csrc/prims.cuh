// prims.cuh -- hand-written device primitives: LSD radix sort of (key, value)
// pairs (K2) and single-pass block scans (K5).  No CUB/Thrust.
#pragma once
#include "common.cuh"

namespace fluxb200 {

// ---------------------------------------------------------------------------
// exclusive scan, one CTA walking the array (sizes here: <= a few million)
// ---------------------------------------------------------------------------
template <class Tin, class Tout>
__global__ void __launch_bounds__(1024) scan_exclusive_kernel(const Tin *in, Tout *out, int64_t n,
                                                              Tout *total) { // in may alias out
    __shared__ Tout warp_sums[32];
    __shared__ Tout carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t idx = base + tid;
        Tout v = idx < n ? (Tout)in[idx] : (Tout)0;
        Tout x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Tout y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            Tout w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Tout y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w; // inclusive over warps
        }
        __syncthreads();
        const Tout carry = carry_s;
        const Tout warp_off = warp ? warp_sums[warp - 1] : (Tout)0;
        if (idx < n) out[idx] = carry + warp_off + x - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
    if (tid == 0) {
        if (total) *total = carry_s;
    }
}

template <class Tin, class Tout>
inline void scan_exclusive(const Tin *in, Tout *out, int64_t n, Tout *total, cudaStream_t st) {
    scan_exclusive_kernel<Tin, Tout><<<1, 1024, 0, st>>>(in, out, n, total);
    FB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable; keys uint64, values uint32
// ---------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads)
    radix_hist_kernel(const uint64_t *__restrict__ keys, int n, int shift,
                      uint32_t *__restrict__ hist, int nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int e = base + it * kSortThreads + threadIdx.x;
        if (e < n) atomicAdd(&h[(keys[e] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
    radix_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                         uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, int n, int shift,
                         const uint32_t *__restrict__ offs, int nblocks) {
    __shared__ uint32_t base_s[256];
    __shared__ uint32_t wcnt[kSortThreads / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    base_s[tid] = offs[tid * nblocks + blockIdx.x];
    const int base = blockIdx.x * kSortTile;
    for (int it = 0; it < kSortItems; ++it) {
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) wcnt[w][tid] = 0;
        __syncthreads();
        const int e = base + it * kSortThreads + tid;
        const bool valid = e < n;
        uint64_t k = 0;
        uint32_t v = 0;
        unsigned d = 0x1000u + lane; // invalid lanes: unique groups
        if (valid) {
            k = kin[e];
            v = vin[e];
            d = (unsigned)((k >> shift) & 255u);
        }
        const unsigned grp = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(grp & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[warp][d] = __popc(grp);
        __syncthreads();
        if (valid) {
            uint32_t off = base_s[d] + rank;
            for (int w = 0; w < warp; ++w) off += wcnt[w][d];
            kout[off] = k;
            vout[off] = v;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) tot += wcnt[w][tid];
        base_s[tid] += tot;
        __syncthreads();
    }
}

// Sorts (keys, vals) by the low `bits` bits of the keys.  keys/vals are
// overwritten with the result; tmp buffers must hold n elements each; hist
// must hold 256 * ceil(n / kSortTile) uint32.
struct RadixSorter {
    DevBuf ktmp, vtmp, hist;
    int launches = 0;
    void sort(uint64_t *keys, uint32_t *vals, int n, int bits, cudaStream_t st) {
        if (n <= 1) return;
        const int nblocks = (int)ceil_div(n, kSortTile);
        ktmp.reserve(sizeof(uint64_t) * (size_t)n);
        vtmp.reserve(sizeof(uint32_t) * (size_t)n);
        hist.reserve(sizeof(uint32_t) * 256 * (size_t)nblocks);
        uint64_t *ka = keys, *kb = ktmp.as<uint64_t>();
        uint32_t *va = vals, *vb = vtmp.as<uint32_t>();
        int passes = (bits + 7) / 8;
        if (passes & 1) ++passes; // even number of passes: result lands in keys/vals
        for (int p = 0; p < passes; ++p) {
            const int shift = 8 * p;
            radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(ka, n, shift, hist.as<uint32_t>(),
                                                                nblocks);
            scan_exclusive<uint32_t, uint32_t>(hist.as<uint32_t>(), hist.as<uint32_t>(),
                                               256 * (int64_t)nblocks, nullptr, st);
            radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(ka, va, kb, vb, n, shift,
                                                                   hist.as<uint32_t>(), nblocks);
            FB_CUDA(cudaGetLastError());
            launches += 3;
            uint64_t *tk = ka; ka = kb; kb = tk;
            uint32_t *tv = va; va = vb; vb = tv;
        }
    }
    void release() {
        ktmp.release();
        vtmp.release();
        hist.release();
    }
};

} // namespace fluxb200
