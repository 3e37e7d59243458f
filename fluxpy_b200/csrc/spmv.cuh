// spmv.cuh -- row-sharded CSR SpMV / Jacobi step on the device-resident slab
// (SURVEY section 8f, row N2).  Replaces the `FF @ x` of the reference's Jacobi
// radiosity iteration (src/flux/solve.py:25-45) and of compute_steady_state_temp
// (src/flux/model.py:8-24) on the slab a rank owns.  HBM-bound: every stored
// entry (data + index) is read once per product; the iterate (8 bytes x
// columns) stays in L2.
#pragma once
#include "common.cuh"

namespace fluxb200 {

constexpr int kSpmvThreads = 256;

// y[r] = (E ? E[r] : 0) + sum_k data[k] * (rho ? rho[col]*x[col] : rho_s*x[col]),  accumulated in double.
// One CTA per row (rows of a form-factor matrix hold ~n/2 entries).  When
// diffmax != nullptr also |y[r] - x[row_offset + r]| is max-reduced into it
// (the convergence test of solve.py:41).
template <class T, class IDX>
__global__ void __launch_bounds__(kSpmvThreads)
    csr_jacobi_kernel(const int64_t *__restrict__ indptr, const IDX *__restrict__ indices,
                      const T *__restrict__ data, int m, const double *__restrict__ E,
                      const double *__restrict__ rho, double rho_s, const double *__restrict__ x,
                      double *__restrict__ y, unsigned long long *__restrict__ diffmax, int64_t row_offset) {
    __shared__ double warp_sum[kSpmvThreads / 32];
    const int r = blockIdx.x;
    if (r >= m) return;
    const int64_t b = indptr[r], e = indptr[r + 1];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int64_t k = b + threadIdx.x;
    // four independent loads in flight per thread
    for (; k + 3 * kSpmvThreads < e; k += 4 * kSpmvThreads) {
        const IDX c0 = indices[k], c1 = indices[k + kSpmvThreads], c2 = indices[k + 2 * kSpmvThreads],
                  c3 = indices[k + 3 * kSpmvThreads];
        const T d0 = data[k], d1 = data[k + kSpmvThreads], d2 = data[k + 2 * kSpmvThreads],
                d3 = data[k + 3 * kSpmvThreads];
        double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
        if (rho) {
            x0 *= rho[c0];
            x1 *= rho[c1];
            x2 *= rho[c2];
            x3 *= rho[c3];
        }
        acc0 = fma((double)d0, x0, acc0);
        acc1 = fma((double)d1, x1, acc1);
        acc2 = fma((double)d2, x2, acc2);
        acc3 = fma((double)d3, x3, acc3);
    }
    for (; k < e; k += kSpmvThreads) {
        const IDX c0 = indices[k];
        double x0 = x[c0];
        if (rho) x0 *= rho[c0];
        acc0 = fma((double)data[k], x0, acc0);
    }
    double acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kSpmvThreads / 32; ++w) s += warp_sum[w];
        if (!rho) s *= rho_s;
        const double out = (E ? E[r] : 0.0) + s;
        y[r] = out;
        if (diffmax) {
            const double d = fabs(out - x[row_offset + r]);
            // non-negative doubles order like their bit patterns; NaN (divergence) must win
            const unsigned long long bits = (d != d) ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(d);
            atomicMax(diffmax, bits);
        }
    }
}

} // namespace fluxb200
