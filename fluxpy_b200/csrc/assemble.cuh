// assemble.cuh -- K4 fused cull + occlusion, K6 order-preserving CSR fill, and
// the visibility / occlusion query kernels.
//
// Replaces the Python row loop of get_form_factor_matrix (reference
// src/flux/form_factors.py:45-70) and the per-row Embree stream call
// (src/flux/shape.py:349-398).
#pragma once
#include "common.cuh"
#include "trace.cuh"

namespace fluxb200 {

constexpr int kTraceWarps = 8;
constexpr int kTraceThreads = kTraceWarps * 32;
constexpr int kChunkCols = 1024; // sorted columns per warp work unit = 32 ballot words

#define FB_PI 3.141592653589793

// Un-normalised numerator max(0, n_i.d) * max(0, -n_j.d), d = p_j - p_i, in
// double from the shape model's own P, N (form_factors.py:46-47 evaluated
// directly); explicit FMA chain = the oracle's dot3d.
template <class T>
__device__ __forceinline__ double numerator(const Real4<T> &Pi, const Real4<T> &Ni, const Real4<T> &Pj,
                                            const Real4<T> &Nj, double &dx, double &dy, double &dz) {
    dx = __dsub_rn((double)Pj.x, (double)Pi.x);
    dy = __dsub_rn((double)Pj.y, (double)Pi.y);
    dz = __dsub_rn((double)Pj.z, (double)Pi.z);
    double a = __fma_rn((double)Ni.x, dx, __fma_rn((double)Ni.y, dy, __dmul_rn((double)Ni.z, dz)));
    double b = -__fma_rn((double)Nj.x, dx, __fma_rn((double)Nj.y, dy, __dmul_rn((double)Nj.z, dz)));
    a = a > 0.0 ? a : 0.0;
    b = b > 0.0 ? b : 0.0;
    return __dmul_rn(a, b);
}

template <class T> __device__ __forceinline__ bool survives_cull(double num, T eps);
template <> __device__ __forceinline__ bool survives_cull<float>(double num, float eps) {
    const float v = __double2float_rn(num); // abs(row_data) > eps in the shape model's dtype
    return v > eps || -v > eps;
}
template <> __device__ __forceinline__ bool survives_cull<double>(double num, double eps) {
    return num > eps || -num > eps;
}

template <class T> __device__ __forceinline__ Real4<T> load_real4(const Real4<T> *p);
template <> __device__ __forceinline__ Real4<float> load_real4<float>(const Real4<float> *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    return Real4<float>{v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Real4<double> load_real4<double>(const Real4<double> *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return Real4<double>{a.x, a.y, b.x, b.y};
}

template <class T> struct TraceArgs {
    const Real4<T> *faceP, *faceN;  // per face, original order
    const int *rows;                // m face ids (I)
    const int *face_leaf;           // leaf position of every face
    const int *node_up, *leaf_up;   // (parent node << 1 | slot) of every internal node / leaf, -1 = none
    const int2 *node_range;         // first / last leaf position under every internal node
    const Real4<T> *colP, *colN;    // n columns gathered in leaf (Morton) order
    const int *col_face, *col_leaf; // face id / leaf position of sorted column s
    int m, n, nwords;               // nwords = ceil(n/32)
    int nchunks;                    // ceil(n / kChunkCols)
    T eps;
    const float4 *nodes, *tri;
    int ninternal, ntop, nfaces;
    uint32_t *bits;                 // m x nwords visibility words, sorted-column order
    uint32_t *row_counts;           // m
    unsigned long long *tested;     // [0] rays traced, [1] work-unit counter of this launch; kHor: [2] batches,
                                    // [3] batches walked without the zone records, [4] rays whose walk started at the zone node
    int *error_flag;
    float scale;                    // largest |coordinate| of the mesh (pads of the shaft filter)
    int shaft_filter;               // 0: test every record of the per-unit list (A/B check)
    // horizon skip (kHor instantiation only; horizon.cuh) -- appended, so the layout above is unchanged
    const float2 *hz;               // per face: (hor, rmin) from the current P, N
    const int *zone_node;           // per leaf: its zone's node (-1: the leaf alone)
    const float4 *colH;             // per sorted column: (hor, rmin, zone node, up code of the zone node)
    int zone_leaves;                // Z
    float pert;                     // ray perturbation + Pluecker slack, in length units
    // second-generation kernel (trace2.cuh)
    const float4 *chunk_info;       // 3 float4 per chunk of 1024 sorted columns (chunk_info_kernel)
    int2 *lost;                     // (row of this launch, sorted column) of rays to be resolved by resolve_lost_kernel
    unsigned *lost_count;           // entries appended to `lost`
    unsigned lost_cap;
};

// explicit shared-space loads / stores from a 32-bit shared address kept in a register: the
// generic-pointer form makes ptxas rebuild the address (thread id, cluster CTA id, window base)
// inside the hot loops
#ifndef FB_EMU
typedef uint32_t smem_addr_t;
__device__ __forceinline__ float4 lds_f4(smem_addr_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 lds_i2(smem_addr_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_i1(smem_addr_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_i1(smem_addr_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
#else // host build of the kernels for the SIMT emulator (tools/simt, test infrastructure): plain pointers
typedef size_t smem_addr_t;
inline float4 lds_f4(smem_addr_t addr) { return *reinterpret_cast<const float4 *>(addr); }
inline int2 lds_i2(smem_addr_t addr) { return *reinterpret_cast<const int2 *>(addr); }
inline int lds_i1(smem_addr_t addr) { return *reinterpret_cast<const volatile int *>(addr); }
inline void sts_i1(smem_addr_t addr, int v) { *reinterpret_cast<volatile int *>(addr) = v; }
#endif

// Loop-iteration counters of the trace kernel, SIMT-emulator builds only (tools/simt): the kernel is bound by
// instruction issue, so these counts (per 32-ray batch) are what a variant is judged by before GPU time
// is spent on it.  [0] batches, [1] phase-A iterations, [2] / [3] phase-B / phase-C iterations of the
// slowest lane, [4] converged exact-test iterations, [5] / [6] phase-B / phase-C lane-iterations.
#ifdef FB_EMU
namespace emu_stats {
inline unsigned long long iters[8];
inline int warp_max(int v) {
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
} // namespace emu_stats
#define FB_COUNT(k, v) __atomic_fetch_add(&emu_stats::iters[k], (unsigned long long)(v), __ATOMIC_RELAXED)
#endif

constexpr int kLeafCap = 8; // deferred candidate triangles per lane
#ifndef FB_TRACE_MIN_BLOCKS
#define FB_TRACE_MIN_BLOCKS 4 // resident CTAs per SM the register budget is cut for (64 registers)
#endif

// K4, persistent and warp-centric.  A work unit is one (row, chunk of 1024
// Morton-ordered columns); every warp of the grid pulls units from one global
// counter, consecutive units being consecutive chunks of the same row (so the
// warps in flight share the source face and neighbouring targets).  Per unit:
//   list     the child records hanging off the root -> source-leaf path (and
//            off the shared part of the target side) go to shared memory once;
//   phase 1  geometric cull: 32 coalesced float4 column loads per lane, the
//            survivors as 32 ballot words (lane k keeps word k); the list is
//            then shaft-filtered against the hull of the source and the
//            chunk's targets;
//   phase 2  survivors are compacted 32 at a time (prefix of popcounts +
//            find-nth-set-bit) so every lane of a batch holds a ray.  A: the
//            listed records in a warp-uniform loop; B: from the target leaf
//            upwards, the sibling record of every level; C: top-down, with a
//            per-lane stack, only the subtrees that were hit.  Triangles whose
//            box and fitted slab are hit go to a per-lane list in shared memory
//            and are Pluecker-tested in a converged loop; occluded rays clear
//            their bit in the warp's shared words;
//   phase 3  the 32 final words go out as one coalesced 128-byte store.
// kTop = true is the plain variant (top of the tree staged in shared memory,
// top-down traversal from the root only); it is kept as the measured
// alternative and as an independent check of the path walk.
// kHor = true adds the horizon skip (horizon.cuh): a batch whose rays all leave above the source
// face's near-zone horizon walks the list without the records inside zone(i); a ray that arrives
// above its target's horizon starts the upward walk at zone(j)'s node instead of the target leaf.
template <class T, bool kTop, bool kHor = false>
__global__ void __launch_bounds__(kTraceThreads, FB_TRACE_MIN_BLOCKS) trace_kernel(const TraceArgs<T> A) {
    static_assert(!(kTop && kHor), "the horizon skip belongs to the path-walk variant");
    extern __shared__ float4 smem_top[];
    __shared__ uint32_t words_s[kTraceWarps][32];
    __shared__ int leaf_s[kLeafCap][kTraceThreads];
    __shared__ float4 path_s[kTop ? 1 : kTraceWarps][kTop ? 1 : 3 * kStackDepth];
    __shared__ int2 range_s[kTop ? 1 : kTraceWarps][kTop ? 1 : kStackDepth];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    if (kTop) {
        for (int k = threadIdx.x; k < 6 * A.ntop; k += kTraceThreads) smem_top[k] = __ldg(A.nodes + k);
        __syncthreads();
    }
    const BvhView bvh{A.nodes, smem_top, A.tri, kTop ? A.ntop : 0, A.ninternal, A.nfaces, A.error_flag};
    smem_addr_t path_base = (smem_addr_t)__cvta_generic_to_shared(&path_s[warp][0]);
    smem_addr_t range_base = (smem_addr_t)__cvta_generic_to_shared(&range_s[warp][0]);
    smem_addr_t leaf_base = (smem_addr_t)__cvta_generic_to_shared(&leaf_s[0][tid]);
    asm volatile("" : "+r"(path_base), "+r"(range_base), "+r"(leaf_base)); // keep them: do not rematerialise
    const unsigned total_units = (unsigned)A.m * (unsigned)A.nchunks;
    unsigned long long tested = 0;
    unsigned hc_batches = 0, hc_src = 0, hc_tgt = 0; // kHor counters (lane 0 / per lane)

    while (true) {
        unsigned unit = 0;
        if (lane == 0) unit = (unsigned)atomicAdd(A.tested + 1, 1ull);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= total_units) break;
        const int r = (int)(unit / (unsigned)A.nchunks), c = (int)(unit - (unsigned)r * (unsigned)A.nchunks);
        const int i = A.rows[r];
        const Real4<T> Pi = load_real4<T>(A.faceP + i), Ni = load_real4<T>(A.faceN + i);
        const int s0 = c * kChunkCols;
        // Every ray of this unit starts on triangle i, so every traversal would walk
        // root -> leaf(i) and test, level by level, the subtree hanging off that path.
        // List those child records once (the source leaf's own record first, then the
        // sibling at every level up to the root): the batches below test them in a
        // uniform loop with broadcast shared-memory loads instead of descending from the root.
        int npath = 0;
        int cref = -0x7fffffff; // common ancestor of the chunk's targets when usable (see below)
        int xdrop = -1;         // list entry of X when every target of the chunk is under it (cref valid)
        if (!kTop && A.ninternal > 0) {
            const int ileaf = A.face_leaf[i];
            int code = A.leaf_up[ileaf];
            bool own = true; // first entry: the leaf's own record, then siblings
            while (code >= 0) {
                const int p = code >> 1, slot = code & 1;
                if (own) {
                    if (lane < 3) path_s[warp][3 * npath + lane] = __ldg(A.nodes + 6 * (size_t)p + 3 * slot + lane);
                    if (lane == 3) range_s[warp][npath] = make_int2(ileaf, ileaf);
                    ++npath;
                    own = false;
                }
                const float4 *rec = A.nodes + 6 * (size_t)p + 3 * (1 - slot);
                if (lane < 3) path_s[warp][3 * npath + lane] = __ldg(rec + lane);
                if (lane == 3) {
                    const int ref = rec_ref(__ldg(rec + 2));
                    range_s[warp][npath] = ref < 0 ? make_int2(~ref, ~ref) : A.node_range[ref];
                }
                ++npath;
                code = A.node_up[p];
            }
            __syncwarp();
            // Target side, shared part: all targets of this chunk lie under their common
            // ancestor C.  When one source-path record X holds the whole chunk, the siblings
            // between X and C are the same for every ray of the unit: append them to the list
            // (they go through the shaft filter and the uniform loop), and let the per-ray
            // upward walk stop at C instead of X.
            const int leaf_lo = A.col_leaf[s0], leaf_hi = A.col_leaf[min(A.n, s0 + kChunkCols) - 1];
            int xe = -1;
            for (int e = 0; e < npath; ++e)
                if (range_s[warp][e].x <= leaf_lo && leaf_hi <= range_s[warp][e].y) xe = e;
            if (xe >= 0 && leaf_lo != leaf_hi) {
                const int xr = rec_ref(path_s[warp][3 * xe + 2]);
                int c = A.leaf_up[leaf_lo] >> 1;
                while (!(A.node_range[c].x <= leaf_lo && leaf_hi <= A.node_range[c].y)) c = A.node_up[c] >> 1;
                int cur = c, n2 = npath;
                bool complete = true;
                while (cur != xr) {
                    const int up = A.node_up[cur];
                    if (up < 0 || n2 >= kStackDepth) {
                        complete = false;
                        break;
                    }
                    const int pp = up >> 1, slot = up & 1;
                    const float4 *rec = A.nodes + 6 * (size_t)pp + 3 * (1 - slot);
                    if (lane < 3) path_s[warp][3 * n2 + lane] = __ldg(rec + lane);
                    if (lane == 3) {
                        const int ref = rec_ref(__ldg(rec + 2));
                        range_s[warp][n2] = ref < 0 ? make_int2(~ref, ~ref) : A.node_range[ref];
                    }
                    ++n2;
                    cur = pp;
                }
                if (complete) {
                    npath = n2;
                    cref = c;
                    xdrop = xe;
                }
                __syncwarp();
            }
        }
        // horizon skip, per unit: leaf range of zone(i), the source face's horizon, and whether the
        // upward walk of this unit's rays may start at the target's zone node (the walk's end C must
        // lie above every zone: a node of more than Z leaves)
        int zlo = 0, zhi = -1, nout = 0;
        float hor_i = INFINITY;
        bool tskip_unit = false;
        if constexpr (kHor) {
            if (A.ninternal > 0) {
                const int ileaf = A.face_leaf[i];
                const int zn = A.zone_node[ileaf];
                zlo = zhi = ileaf;
                if (zn >= 0) {
                    const int2 zr = A.node_range[zn];
                    zlo = zr.x;
                    zhi = zr.y;
                }
                hor_i = __ldg(A.hz + i).x;
                if (cref != -0x7fffffff) {
                    const int2 cr = A.node_range[cref];
                    tskip_unit = cr.y - cr.x + 1 > A.zone_leaves;
                }
            }
        }
        // ---- phase 1: cull -----------------------------------------------------
        uint32_t myword = 0;
        float bl0 = INFINITY, bl1 = INFINITY, bl2 = INFINITY, bh0 = -INFINITY, bh1 = -INFINITY, bh2 = -INFINITY;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int s = s0 + k * 32 + lane;
            bool keep = false;
            if (s < A.n) {
                const Real4<T> Pj = load_real4<T>(A.colP + s), Nj = load_real4<T>(A.colN + s);
                if (!kTop) { // bounding box of the chunk's target centroids (shaft filter below)
                    bl0 = fminf(bl0, (float)Pj.x); bh0 = fmaxf(bh0, (float)Pj.x);
                    bl1 = fminf(bl1, (float)Pj.y); bh1 = fmaxf(bh1, (float)Pj.y);
                    bl2 = fminf(bl2, (float)Pj.z); bh2 = fmaxf(bh2, (float)Pj.z);
                }
                double dx, dy, dz;
                double num = numerator<T>(Pi, Ni, Pj, Nj, dx, dy, dz);
                if (A.col_face[s] == i) num = 0.0; // row_data[i == J] = 0
                keep = survives_cull<T>(num, A.eps);
            }
            const uint32_t w = __ballot_sync(0xffffffffu, keep);
            if (lane == k) myword = w;
        }
        // ---- shaft filter of the source-path list --------------------------------------
        // Every ray of this unit lies in the convex hull of the source centroid and the
        // chunk's target centroids.  A source-path record whose (padded) box or fitted slab is
        // separated from that hull along x, y, z or its own slab direction cannot be hit by
        // any of them: drop it for the whole unit.  Records holding targets always stay.
        // The list is then partitioned in place: [records that passed] [records that did not] and, when
        // every target of the chunk lies under one record X (cref valid), X itself is taken out -- its
        // inside is covered by the upward walk -- so that the uniform loop over the first nsel (or, for
        // a batch that needs the unfiltered list, nall) records is a plain walk with no per-record checks.
        int nsel = 0, nall = npath;
        if (!kTop && npath > 0) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                bl0 = fminf(bl0, __shfl_xor_sync(0xffffffffu, bl0, o)); bh0 = fmaxf(bh0, __shfl_xor_sync(0xffffffffu, bh0, o));
                bl1 = fminf(bl1, __shfl_xor_sync(0xffffffffu, bl1, o)); bh1 = fmaxf(bh1, __shfl_xor_sync(0xffffffffu, bh1, o));
                bl2 = fminf(bl2, __shfl_xor_sync(0xffffffffu, bl2, o)); bh2 = fmaxf(bh2, __shfl_xor_sync(0xffffffffu, bh2, o));
            }
            const int leaf_lo = A.col_leaf[s0], leaf_hi = A.col_leaf[min(A.n, s0 + kChunkCols) - 1];
            const float px = (float)Pi.x, py = (float)Pi.y, pz = (float)Pi.z;
            const float pad = 3e-5f * A.scale + 1e-4f * fmaxf(fmaxf(bh0 - bl0, bh1 - bl1), bh2 - bl2) + 2e-3f;
            const float h0l = fminf(px, bl0) - pad, h0h = fmaxf(px, bh0) + pad;
            const float h1l = fminf(py, bl1) - pad, h1h = fmaxf(py, bh1) + pad;
            const float h2l = fminf(pz, bl2) - pad, h2h = fmaxf(pz, bh2) + pad;
            static_assert(kStackDepth <= 64, "the list is partitioned two entries per lane");
            float4 ra[2], rb2[2], rc[2];
            int2 rr[2];
            bool keepr[2], other[2];
            bool inzone[2] = {false, false}; // kHor: a kept record inside zone(i), other than the source's own
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = h * 32 + lane;
                keepr[h] = other[h] = false;
                if (e < npath) {
                    const float4 a = path_s[warp][3 * e], b = path_s[warp][3 * e + 1], cc = path_s[warp][3 * e + 2];
                    const int2 rg = range_s[warp][e];
                    ra[h] = a; rb2[h] = b; rc[h] = cc; rr[h] = rg;
                    bool keep = true;
                    if (A.shaft_filter && (rg.y < leaf_lo || rg.x > leaf_hi)) { // holds no target of this chunk
                        // record layout: a = (lo.x, hi.x, lo.y, hi.y), b = (lo.z, hi.z, slab_min, slab_max), cc = (dir | ref)
                        if (a.x > h0h || a.y < h0l || a.z > h1h || a.w < h1l || b.x > h2h || b.y < h2l) keep = false;
                        // extent of the hull along the slab direction
                        const float sp = cc.x * px + cc.y * py + cc.z * pz;
                        const float lo_s = fminf(cc.x * bl0, cc.x * bh0) + fminf(cc.y * bl1, cc.y * bh1) + fminf(cc.z * bl2, cc.z * bh2);
                        const float hi_s = fmaxf(cc.x * bl0, cc.x * bh0) + fmaxf(cc.y * bl1, cc.y * bh1) + fmaxf(cc.z * bl2, cc.z * bh2);
                        const float spad = pad * (fabsf(cc.x) + fabsf(cc.y) + fabsf(cc.z));
                        if (fminf(sp, lo_s) - spad > b.w || fmaxf(sp, hi_s) + spad < b.z) keep = false;
                    }
                    if (e != xdrop) {
                        keepr[h] = keep;
                        other[h] = !keep;
                        if constexpr (kHor) inzone[h] = keep && e != 0 && rg.x >= zlo && rg.y <= zhi;
                    }
                }
            }
            const uint32_t kb0 = __ballot_sync(0xffffffffu, keepr[0]), kb1 = __ballot_sync(0xffffffffu, keepr[1]);
            const uint32_t ob0 = __ballot_sync(0xffffffffu, other[0]), ob1 = __ballot_sync(0xffffffffu, other[1]);
            const uint32_t lt = (1u << lane) - 1u;
            nsel = __popc(kb0) + __popc(kb1);
            nall = nsel + __popc(ob0) + __popc(ob1);
            uint32_t zb0 = 0, zb1 = 0;
            if constexpr (kHor) { // three classes: [kept, outside zone(i)] [kept, inside] [not kept]
                zb0 = __ballot_sync(0xffffffffu, inzone[0]);
                zb1 = __ballot_sync(0xffffffffu, inzone[1]);
                nout = nsel - __popc(zb0) - __popc(zb1);
            }
            __syncwarp(); // every lane holds its two records: the list can be rewritten
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int d = -1;
                if constexpr (kHor) {
                    const uint32_t k0 = kb0 & ~zb0, k1 = kb1 & ~zb1;
                    if (inzone[h]) d = nout + (h ? __popc(zb0) : 0) + __popc((h ? zb1 : zb0) & lt);
                    else if (keepr[h]) d = (h ? __popc(k0) : 0) + __popc((h ? k1 : k0) & lt);
                    else if (other[h]) d = nsel + (h ? __popc(ob0) : 0) + __popc((h ? ob1 : ob0) & lt);
                } else if (keepr[h]) d = (h ? __popc(kb0) : 0) + __popc((h ? kb1 : kb0) & lt);
                else if (other[h]) d = nsel + (h ? __popc(ob0) : 0) + __popc((h ? ob1 : ob0) & lt);
                if (d >= 0) {
                    path_s[warp][3 * d] = ra[h];
                    path_s[warp][3 * d + 1] = rb2[h];
                    path_s[warp][3 * d + 2] = rc[h];
                    range_s[warp][d] = rr[h];
                }
            }
            __syncwarp();
        }
        // ---- phase 2: trace the survivors, 32 rays per batch ---------------------
        uint32_t incl = __popc(myword);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        tested += (lane == 0) ? total : 0;
        words_s[warp][lane] = myword;
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t want = base + lane; // index of my survivor within the chunk
            // smallest k with incl[k] > want (binary search over the lanes' prefix)
            int k = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, incl, k + step - 1);
                if (v <= want) k += step;
            }
            const uint32_t wk = __shfl_sync(0xffffffffu, myword, k);
            const uint32_t before = __shfl_sync(0xffffffffu, incl - __popc(myword), k);
            // ---- ray set-up and the target's own hit distance (converged) ----------
            Ray ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
            int bit = 0, tleaf = -1, tface = 0;
            float tj = 0.f;
            bool active = false, blocked = false, overshoot = false;
            int scol = 0;       // kHor: my column and the centroid distance of my ray
            float dist_h = 0.f;
            if (want < total) {
                bit = __fns(wk, 0, (int)(want - before) + 1);
                const int s = s0 + k * 32 + bit;
                if constexpr (kHor) scol = s;
                const Real4<T> Pj = load_real4<T>(A.colP + s);
                if (setup_ray(Pi, Pj, ray)) { // else masked pair: "vis by default" (shape.py:392)
                    tleaf = A.col_leaf[s];
                    tface = A.col_face[s];
                    if (target_hit_t(bvh, ray, tleaf, tj)) {
                        active = A.ninternal > 0;
                        // The shaft filter assumes the ray ends at the target centroid.  A ray that
                        // grazes its target has an ill-conditioned hit distance and may run on well
                        // past the centroid: such a batch uses the unfiltered list.
                        const float ex = (float)Pj.x - (float)Pi.x, ey = (float)Pj.y - (float)Pi.y,
                                    ez = (float)Pj.z - (float)Pi.z;
                        if constexpr (kHor) {
                            dist_h = sqrtf(ex * ex + ey * ey + ez * ez);
                            overshoot = tj * 1.000002f > dist_h + (1e-5f * A.scale + 1e-3f);
                        } else
                            overshoot = tj * 1.000002f > sqrtf(ex * ex + ey * ey + ez * ez) + (1e-5f * A.scale + 1e-3f);
                    } else {
                        blocked = true; // the ray misses its own target: closest hit is not j
                    }
                }
            }
            // ---- traversal with deferred leaf tests -----------------------------------
            // Every lane walks its own ray (no votes inside the loops: lanes that finish a
            // phase early wait at its end).  Triangles whose box and fitted slab are hit go
            // to a per-lane list in shared memory and are Pluecker-tested by the whole warp
            // in one converged loop at the end of the batch; a lane whose list is full
            // tests the newcomer on the spot (rare).
            const RayBox rb = make_raybox(ray);
            const float tmax = tj * 1.000002f;
            int stack[kStackDepth];
            int sp = 0, node = 0, nl = 0;
            auto push_leaf = [&](int leaf) {
                if (nl < kLeafCap) sts_i1(leaf_base + (uint32_t)(nl++) * (uint32_t)(sizeof(int) * kTraceThreads), leaf);
                else if (leaf_occludes_cold(A.tri, ray, tj, leaf, tface)) {
                    blocked = true;
                    active = false;
                }
            };
            auto push_node = [&](int ref) {
                if (sp < kStackDepth) stack[sp++] = ref;
                else *A.error_flag = 1;
            };
            // (the tree depth was checked against kStackDepth when it was built)
            if (!kTop) {
                // phase A: the records along the source path, same for all lanes.  The one
                // sibling subtree that holds the target (X) is not tested: its inside is
                // covered by phase B.
                int xref = ~tleaf;
                bool xbig = false; // kHor: the record that holds my target is larger than any zone
                const bool fullpath = __any_sync(0xffffffffu, active && overshoot);
                int nuse = fullpath ? nall : nsel;
                if constexpr (kHor) {
                    // source end: every ray of the batch leaves above the horizon of zone(i) -> no triangle of
                    // the zone other than i can be met: the records inside the zone are not walked
                    bool clear = true; // a lane without a ray does not object
                    if (active) {
                        const float si = (float)Ni.x * ray.dx + (float)Ni.y * ray.dy + (float)Ni.z * ray.dz;
                        clear = si > hor_i;
                    }
                    if (__all_sync(0xffffffffu, clear) && !fullpath) {
                        nuse = nout;
                        ++hc_src;
                    }
                    ++hc_batches;
                }
#ifdef FB_EMU
                if (lane == 0) FB_COUNT(0, 1), FB_COUNT(1, nuse);
                int emu_nb = 0;
#endif
                const float tmax_a = active ? tmax : -1.0f; // a lane without a ray never hits
                if (cref != -0x7fffffff) { // X is not in the list: nothing to check per record
                    smem_addr_t addr = path_base;
                    for (int ks = 0; ks < nuse; ++ks, addr += 48) {
                        const float4 a = lds_f4(addr), b = lds_f4(addr + 16), cc = lds_f4(addr + 32);
                        if (child_hit(ray, rb, a, b, cc, tmax_a)) {
                            const int ref = rec_ref(cc);
                            if (ref < 0) push_leaf(~ref);
                            else push_node(ref);
                        }
                    }
                } else { // the chunk straddles several records: every lane skips the one holding its target
                    smem_addr_t addr = path_base, raddr = range_base;
                    for (int ks = 0; ks < nuse; ++ks, addr += 48, raddr += 8) {
                        const float4 a = lds_f4(addr), b = lds_f4(addr + 16), cc = lds_f4(addr + 32);
                        const int2 rg = lds_i2(raddr);
                        const int ref = rec_ref(cc);
                        if (tleaf >= rg.x && tleaf <= rg.y) {
                            xref = ref;
                            if constexpr (kHor) xbig = rg.y - rg.x + 1 > A.zone_leaves;
                        } else if (child_hit(ray, rb, a, b, cc, tmax_a)) {
                            if (ref < 0) push_leaf(~ref);
                            else push_node(ref);
                        }
                    }
                }
                // phase B: from the target leaf up to X (or the chunk's common ancestor), the
                // sibling at every level -- one 48-byte record per level instead of a two-child
                // node per level from the root.  `code` = (parent << 1 | my slot), so the
                // sibling is record code ^ 1 of the node array.
                const int stop = cref != -0x7fffffff ? cref : xref;
                int cur = ~tleaf, code = tleaf >= 0 ? A.leaf_up[tleaf] : -1;
                if constexpr (kHor) {
                    // target end: the ray arrives above the horizon of zone(j) and the tested interval does
                    // not run on past p_j as far as the zone's nearest other triangle -> nothing in zone(j)
                    // except j can be met: the walk starts at the zone's node
                    if (active && (tskip_unit || xbig)) { // the walk ends (at C, or at my X) above every zone
                        const float4 h = __ldg(A.colH + scol);
                        const Real4<T> Nj = load_real4<T>(A.colN + scol);
                        const float st = -((float)Nj.x * ray.dx + (float)Nj.y * ray.dy + (float)Nj.z * ray.dz);
                        const float beyond = tmax - (dist_h - ray_eps()); // ideal hit: dist - 1e-3 along the ray
                        const int zn = __float_as_int(h.z);
                        if (zn >= 0 && st > h.x && beyond + 2.0f * A.pert < h.y) {
                            cur = zn;
                            code = __float_as_int(h.w);
                            ++hc_tgt;
                        }
                    }
                }
                while (active && cur != stop && code >= 0) {
                    const float4 *rec = A.nodes + 3 * (size_t)(unsigned)(code ^ 1);
                    const float4 a = __ldg(rec), b = __ldg(rec + 1), cc = __ldg(rec + 2);
                    cur = code >> 1;
                    code = A.node_up[cur];
#ifdef FB_EMU
                    ++emu_nb;
#endif
                    if (child_hit(ray, rb, a, b, cc, tmax)) {
                        const int ref = rec_ref(cc);
                        if (ref < 0) push_leaf(~ref);
                        else push_node(ref);
                    }
                }
                // phase C: the subtrees that were actually hit, top-down
                if (active) {
                    if (sp > 0) node = stack[--sp];
                    else active = false;
                }
#ifdef FB_EMU
                {
                    FB_COUNT(5, emu_nb);
                    const int mb = emu_stats::warp_max(emu_nb);
                    if (lane == 0) FB_COUNT(2, mb);
                }
#endif
            }
#ifdef FB_EMU
            int emu_c = 0;
#endif
            while (active) {
#ifdef FB_EMU
                ++emu_c;
#endif
                float4 q[6];
                load_node<kTop>(bvh, node, q);
                const bool h0 = child_hit(ray, rb, q[0], q[1], q[2], tmax);
                const bool h1 = child_hit(ray, rb, q[3], q[4], q[5], tmax);
                const int r0 = rec_ref(q[2]), r1 = rec_ref(q[5]);
                if (h0 && r0 < 0 && ~r0 != tleaf) push_leaf(~r0);
                if (h1 && r1 < 0 && ~r1 != tleaf) push_leaf(~r1);
                const bool i0 = h0 && r0 >= 0, i1 = h1 && r1 >= 0;
                if (i0 && i1) stack[sp++] = r1;
                node = i0 ? r0 : r1;
                if (!(i0 || i1)) {
                    if (sp > 0) node = stack[--sp];
                    else active = false;
                }
            }
#ifdef FB_EMU
            if (!kTop) {
                FB_COUNT(6, emu_c);
                const int mc = emu_stats::warp_max(emu_c);
                if (lane == 0) FB_COUNT(3, mc);
            }
#endif
            { // converged exact tests of the listed candidates
                int mx = nl;
#pragma unroll
                for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#ifdef FB_EMU
                if (lane == 0) FB_COUNT(4, mx);
#endif
                for (int q = 0; q < mx; ++q)
                    if (q < nl && !blocked)
                        blocked = leaf_occludes(bvh, ray, tj, lds_i1(leaf_base + (uint32_t)q * (uint32_t)(sizeof(int) * kTraceThreads)), tface);
            }
            if (blocked) atomicAnd(&words_s[warp][k], ~(1u << bit));
        }
        __syncwarp();
        // ---- phase 3: publish -----------------------------------------------------
        const uint32_t fin = words_s[warp][lane];
        const int wi = c * 32 + lane;
        if (wi < A.nwords) A.bits[(size_t)r * A.nwords + wi] = fin;
        unsigned count = __popc(fin);
#pragma unroll
        for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(&A.row_counts[r], count);
        __syncwarp();
    }
    if (lane == 0 && tested) atomicAdd(A.tested, tested);
    if constexpr (kHor) {
#pragma unroll
        for (int o = 16; o; o >>= 1) hc_tgt += __shfl_xor_sync(0xffffffffu, hc_tgt, o);
        if (lane == 0) {
            atomicAdd(A.tested + 2, (unsigned long long)hc_batches);
            atomicAdd(A.tested + 3, (unsigned long long)hc_src);
            atomicAdd(A.tested + 4, (unsigned long long)hc_tgt);
        }
    }
}

// ---------------------------------------------------------------------------
// K6: CSR fill in two kernels.  K4 leaves the visibility words of a row in
// sorted-column (BVH-leaf) order; the CSR wants the caller's column order with
// ascending positions (form_factors.py:52, 69).
//
// K6a un-permutes: one CTA stages R rows of leaf-order words in shared memory
// and reads rank_of_pos once per column for all R rows (the permutation is the
// same for every row), so the m*n lookups are shared-memory reads; 32 ballots
// give each lane one J-order word, stored coalesced.
// K6b emits: one CTA per 8 rows (a warp each).  Per group of 1024 columns the CTA
// gathers the columns' P, N, A into shared memory once for its 8 rows; the lanes
// of a warp expand the set bits of their word of the row into a list of column
// positions, then the warp walks that list 32 entries at a time -- every lane
// computes one stored entry (no lanes idling on zeros) and data / indices go out
// as consecutive, coalesced stores.  Values are recomputed in fp64 and rounded
// once to the model dtype (form_factors.py:62-64).
// ---------------------------------------------------------------------------
constexpr int kFillThreads = 256;
constexpr int kFillWarps = kFillThreads / 32;

// kSmem = false (R = 1): rows too long for shared memory are looked up in global memory
template <int R, bool kSmem>
__global__ void __launch_bounds__(kFillThreads)
unpermute_kernel(const uint32_t *__restrict__ bits, const int *__restrict__ rank_of_pos, int m, int n,
                 int nwords, uint32_t *__restrict__ jbits, uint32_t *__restrict__ gcount) {
    extern __shared__ uint32_t rows_smem[]; // R x nwords, leaf order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = blockIdx.x * R;
    const int nr = min(R, m - row0);
    const uint32_t *rows_s = kSmem ? rows_smem : bits + (size_t)row0 * nwords;
    if (kSmem) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t *dst = rows_smem + (size_t)r * nwords;
            const uint32_t *src = bits + (size_t)(row0 + r) * nwords;
            for (int k = threadIdx.x; k < nwords; k += kFillThreads) dst[k] = r < nr ? src[k] : 0u;
        }
        __syncthreads();
    }
    const int ngroups = (nwords + 31) / 32;
    for (int g = warp; g < ngroups; g += kFillWarps) {
        uint32_t mine[R];
#pragma unroll
        for (int r = 0; r < R; ++r) mine[r] = 0u;
        const int q0 = g * 1024;
        const int tiles = min(32, (n - q0 + 31) / 32);
        for (int t0 = 0; t0 < tiles; t0 += 8) {
            int s[8]; // eight independent loads of the permutation in flight per lane
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + (t0 + u) * 32 + lane;
                s[u] = (t0 + u < tiles && q < n) ? __ldg(rank_of_pos + q) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const bool set = s[u] >= 0 && ((rows_s[(size_t)r * nwords + (s[u] >> 5)] >> (s[u] & 31)) & 1u);
                    const uint32_t w = __ballot_sync(0xffffffffu, set);
                    if (lane == t0 + u) mine[r] = w;
                }
            }
        }
        const int wi = g * 32 + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nr) { // warp-uniform
                if (wi < nwords) jbits[(size_t)(row0 + r) * nwords + wi] = mine[r];
                uint32_t c = __popc(mine[r]);
#pragma unroll
                for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) gcount[(size_t)(row0 + r) * ngroups + g] = c;
            }
        }
    }
}

template <class T> struct FillArgs {
    const Real4<T> *faceP, *faceN;
    const int *rows;        // m
    const int *cols;        // n face ids in original J order
    int m, n, nwords;
    const uint32_t *jbits;  // m x nwords visibility words, ORIGINAL column order
    const uint32_t *gcount; // m x ceil(nwords/32) entries per (row, group of 1024 columns)
    const int64_t *indptr;  // m + 1 (device, int64), local to this launch's rows
    int64_t out_base;       // position of row 0's first entry in data / indices
    T *data;
    void *indices;          // int32 or int64; NULL: the caller expands jbits itself
    int index_width;
};

// One CTA = 8 consecutive rows (one per warp) x one segment of the column groups (blockIdx.y; the
// row's entry count before the segment comes from K6a's per-group counts).  Per group of 1024 columns
// the CTA gathers P, N, A of the columns into shared memory once --
// ALREADY CONVERTED TO DOUBLE (round 1 converted per stored entry: 13 F2F per entry kept the XU pipe 53 %
// busy and the kernel at 20 % of the HBM write bandwidth) -- the gather through `cols` and the L2 reads
// are shared by the 8 rows; then every warp emits its row's entries of the half, two entries per lane
// and iteration (independent fp64 chains).
#ifndef FB_EMIT_PER_LANE
#define FB_EMIT_PER_LANE 2
#endif
constexpr int kFillGroup = 1024;
constexpr int kFillHalf = 1024; // columns staged at a time (= one group; 512 halved the lanes of the bit expansion)
// staged columns, structure of arrays: px[512] py[512] pz[512] area[512] nx[512] ny[512] nz[512] (doubles).
// (An array of 64-byte structs put every lane's 16-byte pieces on two bank groups: 16-way conflicts, 5.1 ms per
// slab against 2.3 ms for the round-1 kernel -- profiles/r02c_*.)
struct EmitCol { double px, py, pz, area, nx, ny, nz; };
template <class T> constexpr size_t emit_smem_bytes() { return sizeof(double) * 7 * kFillHalf; }

// F_ij = max(0, n_i.d) max(0, -n_j.d) A_j / (pi r^4), d = p_j - p_i, in fp64 from the model-dtype inputs
// (form_factors.py:46-47 evaluated directly, :62-64); the operation order of numerator<T>() above
__device__ __forceinline__ double form_factor_value(double pix, double piy, double piz, double nix, double niy,
                                                    double niz, const EmitCol &c) {
    const double dx = __dsub_rn(c.px, pix), dy = __dsub_rn(c.py, piy), dz = __dsub_rn(c.pz, piz);
    double a = __fma_rn(nix, dx, __fma_rn(niy, dy, __dmul_rn(niz, dz)));
    double b = -__fma_rn(c.nx, dx, __fma_rn(c.ny, dy, __dmul_rn(c.nz, dz)));
    // max(0, a) * max(0, b) = a * b when both are positive and +0 otherwise (finite inputs): one combined
    // test and one select of the product instead of two fp64 maxima (DSETP.MAX + FSEL + SEL + moves were 14 of
    // the kernel's ~117 instructions per entry in the r02i capture)
    double num = __dmul_rn(a, b); // (j == i: d == 0, so num == 0 and r2 == 0 -> 0, as row_data[i == J] = 0)
    if (!(a > 0.0 && b > 0.0)) num = 0.0;
    const double r2 = __fma_rn(dx, dx, __fma_rn(dy, dy, __dmul_rn(dz, dz)));
    const double sden = __dmul_rn(FB_PI, __dmul_rn(r2, r2));                  // :62
    return sden == 0.0 ? 0.0 : __ddiv_rn(__dmul_rn(num, c.area), sden);     // :63-64
}

template <class T>
__global__ void __launch_bounds__(kFillThreads) emit_kernel(const FillArgs<T> A) {
    extern __shared__ __align__(16) unsigned char stage_raw[];
    double *col_s = reinterpret_cast<double *>(stage_raw);
    auto staged = [&](int c) {
        return EmitCol{col_s[c], col_s[kFillHalf + c], col_s[2 * kFillHalf + c], col_s[3 * kFillHalf + c],
                       col_s[4 * kFillHalf + c], col_s[5 * kFillHalf + c], col_s[6 * kFillHalf + c]};
    };
    __shared__ uint16_t list_s[kFillWarps][kFillHalf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * kFillWarps + warp;
    // blockIdx.y = segment of the column groups (enough CTAs for a small row slab)
    const int ngroups = (A.nwords + 31) / 32;
    const int gper = (ngroups + (int)gridDim.y - 1) / (int)gridDim.y;
    const int g_begin = blockIdx.y * gper, g_end = min(ngroups, g_begin + gper);
    if (g_begin >= g_end) return;
    const bool live = r < A.m && A.indptr[r + 1] != A.indptr[r];
    const uint32_t *jb = A.jbits + (size_t)(live ? r : 0) * A.nwords;
    int64_t off = 0;
    if (live) { // entries of the row before my segment
        uint32_t before = 0;
        for (int g = lane; g < g_begin; g += 32) before += A.gcount[(size_t)r * ngroups + g];
#pragma unroll
        for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        off = A.out_base + A.indptr[r] + before;
    }
    const int i = live ? A.rows[r] : 0;
    const Real4<T> Pi = load_real4<T>(A.faceP + i), Ni = load_real4<T>(A.faceN + i);
    const double pix = (double)Pi.x, piy = (double)Pi.y, piz = (double)Pi.z;
    const double nix = (double)Ni.x, niy = (double)Ni.y, niz = (double)Ni.z;
    constexpr int kTilesPerGroup = kFillGroup / kFillHalf, kTileWords = kFillHalf / 32;
    // (Fetching the next tile's columns into registers ahead of the emission was measured, r02m: long-scoreboard
    // stalls 2.6 -> 1.0 warps per issue, short-scoreboard 1.9 -> 2.9, 2.30 ms against 2.26 -- the kernel waits on
    // its fp64 chains and shared-memory reads, not on the gather.  Not kept.)
    const int gh_begin = kTilesPerGroup * g_begin, gh_end = kTilesPerGroup * g_end;
    for (int gh = gh_begin; gh < gh_end; ++gh) {
        const int q0 = gh * kFillHalf; // first column of this tile
        if (q0 >= A.n) break;          // (uniform: the last tile of the last group may be empty)
        __syncthreads(); // the previous half's columns are no longer read
        for (int c = threadIdx.x; c < kFillHalf; c += kFillThreads) {
            const int q = q0 + c;
            if (q < A.n) {
                const int j = A.cols[q];
                const Real4<T> Pj = load_real4<T>(A.faceP + j), Nj = load_real4<T>(A.faceN + j);
                col_s[c] = (double)Pj.x;
                col_s[kFillHalf + c] = (double)Pj.y;
                col_s[2 * kFillHalf + c] = (double)Pj.z;
                col_s[3 * kFillHalf + c] = (double)Pj.w;
                col_s[4 * kFillHalf + c] = (double)Nj.x;
                col_s[5 * kFillHalf + c] = (double)Nj.y;
                col_s[6 * kFillHalf + c] = (double)Nj.z;
            }
        }
        __syncthreads();
        if (!live) continue;
        const int wi = gh * kTileWords + lane;
        uint32_t word = (lane < kTileWords && wi < A.nwords) ? jb[wi] : 0u;
        const int c = __popc(word);
        int incl = c;
#pragma unroll
        for (int o = 1; o < kTileWords; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, kTileWords - 1);
        if (total == 0) continue;
        int p = incl - c;
        while (word) { // ascending positions of my word's set bits
            list_s[warp][p++] = (uint16_t)(lane * 32 + (__ffs(word) - 1));
            word &= word - 1;
        }
        __syncwarp();
        for (int e = lane; e < total; e += 32 * FB_EMIT_PER_LANE) {
            // FB_EMIT_PER_LANE independent entries per lane and iteration (2; 3 and 4 measured the same 3.14 ms of
            // fill per slab, r02o: the kernel waits on its fp64 chains and on shared-memory reads -- wait 2.8,
            // short-scoreboard 1.9 warps per issue -- whatever the number of chains the compiler is offered)
            int cc[FB_EMIT_PER_LANE];
            double vv[FB_EMIT_PER_LANE];
#pragma unroll
            for (int u = 0; u < FB_EMIT_PER_LANE; ++u) cc[u] = (int)list_s[warp][min(e + 32 * u, total - 1)];
#pragma unroll
            for (int u = 0; u < FB_EMIT_PER_LANE; ++u) vv[u] = form_factor_value(pix, piy, piz, nix, niy, niz, staged(cc[u]));
            // streaming stores: the CSR is written once and must not push the mesh and BVH, which a
            // concurrently running trace kernel lives on, out of L2
            const int64_t dst = off + e;
#pragma unroll
            for (int u = 0; u < FB_EMIT_PER_LANE; ++u)
                if (e + 32 * u < total) {
                    __stcs(A.data + dst + 32 * u, (T)vv[u]);
                    if (A.indices) {
                        if (A.index_width == 4) __stcs(reinterpret_cast<int *>(A.indices) + dst + 32 * u, q0 + cc[u]);
                        else __stcs(reinterpret_cast<long long *>(A.indices) + dst + 32 * u, (long long)(q0 + cc[u]));
                    }
                }
        }
        __syncwarp();
        off += total;
    }
}

// ---------------------------------------------------------------------------
// query kernels (TrimeshShapeModel hooks)
// ---------------------------------------------------------------------------
template <class T>
__global__ void visibility_kernel(const Real4<T> *__restrict__ faceP, const int *__restrict__ rows, int m,
                                  const int *__restrict__ cols, int n, const int *__restrict__ face_leaf,
                                  const float4 *nodes, const float4 *tri, int ninternal, int nf,
                                  int *error_flag, int bruteforce, uint8_t *__restrict__ vis) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * n) return;
    const int p = (int)(idx / n), q = (int)(idx - (int64_t)p * n);
    const int i = rows[p], j = cols[q];
    const Real4<T> Pi = load_real4<T>(faceP + i), Pj = load_real4<T>(faceP + j);
    Ray ray;
    bool visible = true;
    if (setup_ray(Pi, Pj, ray)) {
        if (bruteforce) visible = target_visible_bruteforce(tri, nf, ray, face_leaf[j], j);
        else {
            const BvhView bvh{nodes, nullptr, tri, 0, ninternal, nf, error_flag};
            visible = target_visible(bvh, ray, face_leaf[j], j);
        }
    }
    vis[idx] = visible ? 1 : 0;
}

// origin P[i] + eps*N[i] (shape.py:410), direction D as given; any hit in [0, inf]
template <class T>
__global__ void occluded_kernel(const Real4<T> *__restrict__ faceP, const Real4<T> *__restrict__ faceN,
                                const int *__restrict__ rows, int m, const T *__restrict__ D, int nd,
                                int mode, const float4 *nodes, const float4 *tri, int ninternal, int nf,
                                int *error_flag, uint8_t *__restrict__ occ) {
    const int cols = mode == 2 ? nd : 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m * cols) return;
    const int p = (int)(idx / cols), c = (int)(idx - (int64_t)p * cols);
    const int i = rows[p];
    const Real4<T> Pi = load_real4<T>(faceP + i), Ni = load_real4<T>(faceN + i);
    const T eps = (T)ray_eps();
    const T *d = mode == 0 ? D : (mode == 1 ? D + 3 * (size_t)p : D + 3 * (size_t)c);
    Ray ray;
    ray.ox = (float)rn_add<T>(Pi.x, rn_mul<T>(eps, Ni.x));
    ray.oy = (float)rn_add<T>(Pi.y, rn_mul<T>(eps, Ni.y));
    ray.oz = (float)rn_add<T>(Pi.z, rn_mul<T>(eps, Ni.z));
    ray.dx = (float)d[0];
    ray.dy = (float)d[1];
    ray.dz = (float)d[2];
    const BvhView bvh{nodes, nullptr, tri, 0, ninternal, nf, error_flag};
    occ[idx] = occluded_anyhit(bvh, ray, __int_as_float(0x7f800000), -1, 0x7fffffff) ? 1 : 0;
}

__global__ void intersect1_kernel(float ox, float oy, float oz, float dx, float dy, float dz,
                                  const float4 *nodes, const float4 *tri, int ninternal, int nf,
                                  int *error_flag, int *face_out, float *t_out) {
    Ray ray{ox, oy, oz, dx, dy, dz};
    const BvhView bvh{nodes, nullptr, tri, 0, ninternal, nf, error_flag};
    float t = __int_as_float(0x7f800000);
    int face;
    closest_hit(bvh, ray, t, face);
    *face_out = face;
    *t_out = t;
}

// ---- per-call preparation -----------------------------------------------------
__global__ void col_keys_kernel(const int *__restrict__ cols, int n, const int *__restrict__ face_leaf,
                                uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    keys[q] = (uint64_t)(uint32_t)face_leaf[cols[q]];
    vals[q] = (uint32_t)q;
}

// sorted position s -> original position pos[s]; gathers the column arrays
template <class T>
__global__ void col_gather_kernel(const uint32_t *__restrict__ pos, const int *__restrict__ cols, int n,
                                  const int *__restrict__ face_leaf, const Real4<T> *__restrict__ faceP,
                                  const Real4<T> *__restrict__ faceN, Real4<T> *__restrict__ colP,
                                  Real4<T> *__restrict__ colN, int *__restrict__ col_face,
                                  int *__restrict__ col_leaf, int *__restrict__ rank_of_pos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int q = (int)pos[s];
    const int f = cols[q];
    colP[s] = faceP[f];
    Real4<T> Nf = faceN[f];
    // .w: 8 * 2^-24 * max|N_k| -- the target's share of the float32 cull's error bound (trace2.cuh: cull_keep)
    Nf.w = (T)(4.76837158e-7f * fmaxf(fmaxf(fabsf((float)Nf.x), fabsf((float)Nf.y)), fabsf((float)Nf.z)));
    colN[s] = Nf;
    col_face[s] = f;
    col_leaf[s] = face_leaf[f];
    rank_of_pos[q] = s;
}

// interleave host-side P (nf x 3), N (nf x 3), A (nf) into the packed arrays; *changed gets bit 0 when a
// P or N value differs from the one it replaces (the horizons of horizon.cuh depend on them), bit 1 for A
template <class T>
__global__ void pack_face_kernel(const T *__restrict__ P, const T *__restrict__ N, const T *__restrict__ A,
                                 int nf, Real4<T> *__restrict__ faceP, Real4<T> *__restrict__ faceN,
                                 int *__restrict__ changed) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    bool diff = false;
    if (P) {
        const T x = P[3 * (size_t)f], y = P[3 * (size_t)f + 1], z = P[3 * (size_t)f + 2];
        diff = !(faceP[f].x == x && faceP[f].y == y && faceP[f].z == z); // NaN counts as a change
        faceP[f].x = x;
        faceP[f].y = y;
        faceP[f].z = z;
    }
    bool adiff = false;
    if (A) {
        adiff = !(faceP[f].w == A[f]);
        faceP[f].w = A[f];
    }
    if (N) {
        const T x = N[3 * (size_t)f], y = N[3 * (size_t)f + 1], z = N[3 * (size_t)f + 2];
        diff = diff || !(faceN[f].x == x && faceN[f].y == y && faceN[f].z == z);
        faceN[f].x = x;
        faceN[f].y = y;
        faceN[f].z = z;
        faceN[f].w = (T)0;
    }
    if (diff) atomicOr(reinterpret_cast<unsigned *>(changed), 1u);  // P or N: horizons and prepared column sets are stale
    if (adiff) atomicOr(reinterpret_cast<unsigned *>(changed), 2u); // A: the prepared column sets only
}

template <class T>
__global__ void unpack_face_kernel(const Real4<T> *__restrict__ faceP, const Real4<T> *__restrict__ faceN,
                                   int nf, T *__restrict__ P, T *__restrict__ N, T *__restrict__ A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    if (P) {
        P[3 * (size_t)f] = faceP[f].x;
        P[3 * (size_t)f + 1] = faceP[f].y;
        P[3 * (size_t)f + 2] = faceP[f].z;
    }
    if (A) A[f] = faceP[f].w;
    if (N) {
        N[3 * (size_t)f] = faceN[f].x;
        N[3 * (size_t)f + 1] = faceN[f].y;
        N[3 * (size_t)f + 2] = faceN[f].z;
    }
}

__global__ void counts_to_i64_kernel(const uint32_t *__restrict__ c, int m, int64_t *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) out[r] = (int64_t)c[r];
}

// Row counts and the sub-slab's entry count straight into page-locked host memory (zero-copy
// stores): a cudaMemcpy of a few KB would queue on the copy engine behind the previous sub-slab's
// multi-megabyte copy-out and stall the host loop that waits for these numbers.
__global__ void publish_counts_kernel(const uint32_t *__restrict__ c, int m, const int64_t *__restrict__ total,
                                      uint32_t *host_counts, int64_t *host_total) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) host_counts[r] = c[r];
    if (r == 0) *host_total = *total;
    __threadfence_system();
}

__global__ void indptr_to_i32_kernel(const int64_t *__restrict__ in, int n, int32_t *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = (int32_t)in[r];
}

} // namespace fluxb200
