// selftest.cu -- checks of the SIMT emulator itself (tools/simt), compiled by tests/test_emu_engine.py
// through the same source rewrite as the library.  Each mode prints "ok" or aborts.
#include <cuda_runtime.h>
#include <string>
#include <vector>

__global__ void reduce_kernel(const int *in, int n, int *block_sums, unsigned long long *total) {
    __shared__ int warp_sums[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = i < n ? in[i] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += warp_sums[w];
        block_sums[blockIdx.x] = s;
        atomicAdd(total, (unsigned long long)s);
    }
}

// ballot / match_any / shfl_up prefix / __fns / dynamic shared memory, one warp per row of 32 values
__global__ void warp_ops_kernel(const int *in, unsigned *ballots, unsigned *groups, int *prefix, unsigned *nth) {
    extern __shared__ int stage[];
    const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int v = in[row * 32 + lane];
    stage[threadIdx.x] = v;
    __syncwarp();
    const unsigned b = __ballot_sync(0xffffffffu, v & 1);
    const unsigned g = __match_any_sync(0xffffffffu, stage[threadIdx.x ^ 1] % 5);
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 0) ballots[row] = b;
    groups[row * 32 + lane] = g;
    prefix[row * 32 + lane] = x;
    nth[row * 32 + lane] = __fns(b, 0, lane + 1);
}

__global__ void divergent_collective_kernel(int *out) {
    // a full-mask collective that only half of the warp reaches: undefined on hardware, a reported deadlock here
    if ((threadIdx.x & 31) < 16) out[threadIdx.x] = __shfl_sync(0xffffffffu, (int)threadIdx.x, 0);
    __syncthreads();
}

__global__ void missed_barrier_kernel(int *out) {
    // half of every warp waits at the block barrier, the other half at a warp collective that needs them
    if ((threadIdx.x & 31) < 16) __syncthreads();
    else __syncwarp();
    out[threadIdx.x] = 1;
}

__global__ void oob_kernel(int *buf, int n) { buf[n + threadIdx.x] = 1; }

#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            printf("FAILED: %s (line %d)\n", #c, __LINE__);             \
            return 1;                                                   \
        }                                                               \
    } while (0)

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "ops";
    if (mode == "ops") {
        const int n = 100000, B = 256, G = (n + B - 1) / B;
        std::vector<int> h(n);
        long long want = 0;
        for (int i = 0; i < n; ++i) want += (h[i] = (i * 2654435761u) % 1000);
        int *d, *bs;
        unsigned long long *tot;
        cudaMalloc((void **)&d, sizeof(int) * n);
        cudaMalloc((void **)&bs, sizeof(int) * G);
        cudaMalloc((void **)&tot, sizeof(*tot));
        cudaMemcpy(d, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
        cudaMemset(tot, 0, sizeof(*tot));
        reduce_kernel<<<G, B>>>(d, n, bs, tot);
        unsigned long long got = 0;
        cudaMemcpy(&got, tot, sizeof(got), cudaMemcpyDeviceToHost);
        CHECK((long long)got == want);
        const int rows = 64;
        unsigned *bal, *grp, *nth;
        int *pre;
        cudaMalloc((void **)&bal, sizeof(unsigned) * rows);
        cudaMalloc((void **)&grp, sizeof(unsigned) * rows * 32);
        cudaMalloc((void **)&nth, sizeof(unsigned) * rows * 32);
        cudaMalloc((void **)&pre, sizeof(int) * rows * 32);
        warp_ops_kernel<<<rows / 4, 128, sizeof(int) * 128>>>(d, bal, grp, pre, nth);
        std::vector<unsigned> hb(rows), hg(rows * 32), hn(rows * 32);
        std::vector<int> hp(rows * 32);
        cudaMemcpy(hb.data(), bal, sizeof(unsigned) * rows, cudaMemcpyDeviceToHost);
        cudaMemcpy(hg.data(), grp, sizeof(unsigned) * rows * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(hn.data(), nth, sizeof(unsigned) * rows * 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(hp.data(), pre, sizeof(int) * rows * 32, cudaMemcpyDeviceToHost);
        for (int r = 0; r < rows; ++r) {
            unsigned b = 0;
            int run = 0;
            for (int l = 0; l < 32; ++l) b |= (unsigned)(h[r * 32 + l] & 1) << l;
            CHECK(hb[r] == b);
            int seen = 0;
            for (int l = 0; l < 32; ++l) {
                run += h[r * 32 + l];
                CHECK(hp[r * 32 + l] == run);
                unsigned g = 0;
                for (int k = 0; k < 32; ++k) g |= (unsigned)(h[r * 32 + (k ^ 1)] % 5 == h[r * 32 + (l ^ 1)] % 5) << k;
                CHECK(hg[r * 32 + l] == g);
                unsigned want_nth = 0xffffffffu; // position of the (l+1)-th set bit of b
                int c = 0;
                for (int k = 0; k < 32; ++k)
                    if (((b >> k) & 1u) && ++c == l + 1) { want_nth = (unsigned)k; break; }
                CHECK(hn[r * 32 + l] == want_nth);
                (void)seen;
            }
        }
        cudaFree(d); cudaFree(bs); cudaFree(tot); cudaFree(bal); cudaFree(grp); cudaFree(nth); cudaFree(pre);
        printf("ok\n");
        return 0;
    }
    int *buf;
    cudaMalloc((void **)&buf, sizeof(int) * 64);
    cudaMemset(buf, 0, sizeof(int) * 64);
    if (mode == "divergent") divergent_collective_kernel<<<1, 64>>>(buf);
    if (mode == "barrier") missed_barrier_kernel<<<1, 64>>>(buf);
    if (mode == "oob") {
        oob_kernel<<<1, 32>>>(buf, 64);
        cudaFree(buf);
    }
    printf("not detected\n");
    return 0;
}
