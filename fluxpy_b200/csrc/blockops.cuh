// blockops.cuh -- operations on a device-resident CSR slab that feed the
// reference's hierarchical compression:
//   * sub-block extraction  spmat[row_inds, :][:, col_inds]
//     (reference src/flux/compressed_form_factors.py:562);
//   * products with a thin dense matrix, A @ X and A^T @ X, the two kernels of a
//     randomised range finder for the SVD leaves
//     (reference src/flux/compressed_form_factors.py:388-405 -> src/flux/linalg.py:8-50,
//      where ARPACK `svds` does the same products one vector at a time on the CPU).
#pragma once
#include "common.cuh"

namespace fluxb200 {

constexpr int kBlockThreads = 256;

// counts[r'] = number of entries of source row rows[r'] whose column is selected
template <class IDX>
__global__ void __launch_bounds__(kBlockThreads)
    extract_count_kernel(const int64_t *__restrict__ indptr, const IDX *__restrict__ indices,
                         const int *__restrict__ rows, int mr, const int *__restrict__ newpos,
                         int64_t *__restrict__ counts) {
    __shared__ unsigned warp_cnt[kBlockThreads / 32];
    const int r = blockIdx.x;
    if (r >= mr) return;
    const int64_t b = indptr[rows[r]], e = indptr[rows[r] + 1];
    unsigned c = 0;
    for (int64_t k = b + threadIdx.x; k < e; k += kBlockThreads) c += newpos[indices[k]] >= 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < kBlockThreads / 32; ++w) t += warp_cnt[w];
        counts[r] = t;
    }
}

// order-preserving copy of the selected entries (source order = ascending source column)
template <class T, class IDX>
__global__ void __launch_bounds__(kBlockThreads)
    extract_fill_kernel(const int64_t *__restrict__ indptr, const IDX *__restrict__ indices,
                        const T *__restrict__ data, const int *__restrict__ rows, int mr,
                        const int *__restrict__ newpos, const int64_t *__restrict__ out_indptr,
                        IDX *__restrict__ out_indices, T *__restrict__ out_data) {
    __shared__ unsigned warp_cnt[kBlockThreads / 32];
    const int r = blockIdx.x;
    if (r >= mr) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = indptr[rows[r]], e = indptr[rows[r] + 1];
    // every warp owns a contiguous segment of the source row (multiple of 32 entries)
    const int64_t groups = (e - b + 31) / 32, gper = (groups + kBlockThreads / 32 - 1) / (kBlockThreads / 32);
    const int64_t kb = min(e, b + warp * gper * 32), ke = min(e, b + (warp + 1) * gper * 32);
    unsigned c = 0;
    for (int64_t k0 = kb; k0 < ke; k0 += 32) {
        const int64_t k = k0 + lane;
        const bool keep = k < ke && newpos[indices[k]] >= 0;
        c += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) warp_cnt[warp] = c;
    __syncthreads();
    int64_t off = out_indptr[r];
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    for (int64_t k0 = kb; k0 < ke; k0 += 32) {
        const int64_t k = k0 + lane;
        int np = -1;
        if (k < ke) np = newpos[indices[k]];
        const unsigned bal = __ballot_sync(0xffffffffu, np >= 0);
        if (np >= 0) {
            const int64_t dst = off + __popc(bal & ((1u << lane) - 1u));
            out_indices[dst] = (IDX)np;
            out_data[dst] = data[k];
        }
        off += __popc(bal);
    }
}

// Y (m x k) = A @ X (n x k), row-major, k <= 32: lane = right-hand-side column,
// every warp walks a segment of the row's entries, X rows are read coalesced.
template <class T, class IDX>
__global__ void __launch_bounds__(kBlockThreads)
    csr_matmat_kernel(const int64_t *__restrict__ indptr, const IDX *__restrict__ indices,
                      const T *__restrict__ data, int m, const double *__restrict__ X, int k,
                      double *__restrict__ Y) {
    __shared__ double part[kBlockThreads / 32][32];
    const int r = blockIdx.x;
    if (r >= m) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = indptr[r], e = indptr[r + 1];
    double acc = 0.0;
    for (int64_t q0 = b + warp * 32; q0 < e; q0 += kBlockThreads) {
        // the 32 lanes load 32 consecutive entries, then broadcast them one by one
        const int64_t q = q0 + lane;
        const IDX cmine = q < e ? indices[q] : (IDX)0;
        const double dmine = q < e ? (double)data[q] : 0.0;
        const int cnt = (int)min((int64_t)32, e - q0);
        for (int j = 0; j < cnt; ++j) {
            const IDX c = __shfl_sync(0xffffffffu, cmine, j);
            const double d = __shfl_sync(0xffffffffu, dmine, j);
            if (lane < k) acc = fma(d, X[(size_t)c * k + lane], acc);
        }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && lane < k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) s += part[w][lane];
        Y[(size_t)r * k + lane] = s;
    }
}

// Y (n x k) += A^T @ X (m x k): scatter with fp64 atomics (Y zeroed by the caller)
template <class T, class IDX>
__global__ void __launch_bounds__(kBlockThreads)
    csr_rmatmat_kernel(const int64_t *__restrict__ indptr, const IDX *__restrict__ indices,
                       const T *__restrict__ data, int m, const double *__restrict__ X, int k,
                       double *__restrict__ Y) {
    const int r = blockIdx.x;
    if (r >= m) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = indptr[r], e = indptr[r + 1];
    const double xr = lane < k ? X[(size_t)r * k + lane] : 0.0;
    for (int64_t q0 = b + warp * 32; q0 < e; q0 += kBlockThreads) {
        const int64_t q = q0 + lane;
        const IDX cmine = q < e ? indices[q] : (IDX)0;
        const double dmine = q < e ? (double)data[q] : 0.0;
        const int cnt = (int)min((int64_t)32, e - q0);
        for (int j = 0; j < cnt; ++j) {
            const IDX c = __shfl_sync(0xffffffffu, cmine, j);
            const double d = __shfl_sync(0xffffffffu, dmine, j);
            if (lane < k) atomicAdd(&Y[(size_t)c * k + lane], d * xr);
        }
    }
}

} // namespace fluxb200
