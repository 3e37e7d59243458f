// trace2.cuh -- K4, second generation of the fused cull + occlusion kernel.
//
// Same contract as trace_kernel (assemble.cuh): replaces the row loop of get_form_factor_matrix
// (reference src/flux/form_factors.py:45-60) and EmbreeTrimeshShapeModel._get_visibility
// (src/flux/shape.py:349-398); every visibility bit is identical to the oracle's (the set of
// triangles that reach the exact Pluecker test is unchanged).
//
// What differs from the first generation, each item aimed at a measured cost of that kernel
// (profiles/r02a_*: 2063 warp instructions per 32-ray batch; 8 % of them in the fp64 cull, 6 % in the
// survivor compaction, 5 % building the per-unit record list with 7 of 32 lanes; 4.3e8 local-memory loads
// per slab from the per-lane traversal stacks and register spills, long-scoreboard the top stall):
//
//  * float32 models: the cull decides in float32 with a rigorous error bound and evaluates the fp64
//    numerator only for pairs whose float32 interval straddles eps (one in ~1e4).
//  * survivor compaction: the cull runs two 32-column tiles at a time and appends the survivors to a
//    small FIFO in shared memory; a batch starts whenever 32 are waiting (no prefix search, no select).
//  * the per-unit record list is built by all lanes (one walks the `up` chain, 32 load records); data
//    that is the same for every row of a chunk (bounding box of the target centroids, leaf range,
//    common ancestor) comes from chunk_info_kernel, once per call.
//  * the per-lane traversal stack and the per-lane list of candidate triangles live in shared memory
//    (8 + 6 entries per lane); unit state, the source face and the counters too: the hot loops run
//    without local memory and without spills at 64 registers.  The footprint is 39.8 KB per CTA, so that four
//    CTAs leave 92 KB of the SM's 256 KB to L1 (51.8 KB per CTA with 16-entry stacks measured 1.7 % slower).
//  * a ray that grazes many triangles (more candidates than its list holds) puts the surplus into a list
//    shared by the warp; those are Pluecker-tested by ALL lanes at the end of the batch, 32 at a time, the
//    owner's ray fetched by shuffles (the first generation tested each of them on the spot with one lane
//    active: 7 % of its instructions on the bench mesh).
//  * a stack or list that still overflows never loses a result: the ray's column is flagged and traced
//    again on its own from the root at the end of the unit (cold, out of line, outside every hot loop).
//
// Measured and dropped on the way (profiles/r02b_*): phase C as a warp-shared work queue (every hit node an
// item in a shared-memory ring, rounds of 32 items with the rays in a shared pool).  Lane utilisation of the
// traversal rose from 61 % to 88 %, but a round cost 210 instructions against 106 for a per-lane visit of two
// children (pool loads, ray-box set-up, ballot-aggregated pushes, a 77 KB kernel that stalled on
// instruction fetch): 39.2 ms per slab against 34.5 ms for the first generation.
#pragma once
#include "assemble.cuh"

namespace fluxb200 {

#ifndef FB_KPATH
#define FB_KPATH 36
#endif
#ifndef FB_KSTK
#define FB_KSTK 8
#endif
#ifndef FB_KCAND
#define FB_KCAND 6
#endif
#ifndef FB_KOVF
#define FB_KOVF 64
#endif
#ifndef FB_C_FORM
#define FB_C_FORM 0 // loop form of phase C (tuning variants)
#endif
#ifndef FB_T2_PACKED
// packed FP32 pairs (FFMA2 / FMUL2 / FADD2) in the box + slab test: bit 0 phases A / B, bit 1 phase C.  Same
// results either way (trace.cuh child_hit); measured on a 4096-row slab of the 200k-face crater (r02k):
// 0: 30.90 ms, 1: 34.88 ms, 2: 30.56 ms, 3: 33.18 ms -- the pairs pay in the two-child test of phase C, whose
// operands arrive as aligned register quads from LDG.128, and cost in phases A / B.
#define FB_T2_PACKED 2
#endif
constexpr int kPathCap = FB_KPATH; // records of the per-unit list (source path + shared target side)
constexpr int kStk = FB_KSTK;      // traversal stack entries per lane (shared memory)
constexpr int kCand = FB_KCAND;    // candidate triangles per lane (shared memory)
constexpr int kOvf = FB_KOVF;      // ... and per warp for the lanes whose own list is full (grazing rays)
constexpr int kFifoCap = 128;  // survivors between the cull and the batches (power of two; < 32 + 2 tiles)
constexpr int kNodeBits = 26;  // (host check: node ids below 2^26)

struct __align__(16) WarpShared {
    float4 path[3 * kPathCap];  // records, 3 float4 each (layout: trace.cuh child_hit)
    int2 range[kPathCap];       // leaf range of every record
    int stack[kStk][32];        // per-lane traversal stack: [level][lane]
    int cand[kCand][32];        // per-lane candidate triangles (leaf positions): [k][lane]
    uint32_t ovf[kOvf];         // candidates of lanes whose own list is full: (lane << 26) | leaf position
    uint32_t recs[64];          // unit set-up: record index (node << 1 | slot) of every list entry
    uint32_t words[32];         // the unit's visibility words
    uint32_t redo[32];          // columns whose ray overflowed its stack / list: traced again at the end of the unit
    uint16_t fifo[kFifoCap];    // cull survivors waiting for a batch (column within the chunk)
    // unit state and counters live here, not in registers: the kernel is register-bound (64 per thread)
    double src[8];              // the source face's P (xyz, A) and N (xyz, 0) in the model dtype
    int u[8];                   // kU* below
    unsigned long long tested;  // cull survivors = rays of this warp
    unsigned ctr[6];            // kC* below
    uint32_t novf, ovfhit;      // entries of ovf[] in this batch; lanes whose ray one of them occludes
    uint32_t pad_[2];
};
enum { kUnsel = 0, kUnall, kUnout, kUcref, kUtskip, kUhor, kUrow, kUchunk };
enum { kCbatches = 0, kCsrc, kCtgt, kCrounds, kCitems, kCcold };
static_assert(sizeof(WarpShared) % 16 == 0, "warp blocks stay 16-byte aligned");
constexpr size_t trace2_smem_bytes() { return sizeof(WarpShared) * kTraceWarps; }

// Per chunk of 1024 leaf-ordered columns, the same for every row: info[3c] = (bbox lo of the
// target centroids | first leaf), info[3c+1] = (bbox hi | last leaf), info[3c+2] = (common
// ancestor node or -1, its leaf count, -, -).  One warp per chunk.
template <class T>
__global__ void chunk_info_kernel(const Real4<T> *__restrict__ colP, const int *__restrict__ col_leaf, int n,
                                  int nchunks, const int *__restrict__ leaf_up, const int *__restrict__ node_up,
                                  const int2 *__restrict__ node_range, int ninternal,
                                  float4 *__restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int c = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= nchunks) return;
    const int s0 = c * kChunkCols, s1 = min(n, s0 + kChunkCols);
    float l0 = INFINITY, l1 = INFINITY, l2 = INFINITY, h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY;
    for (int s = s0 + lane; s < s1; s += 32) {
        const Real4<T> P = colP[s];
        l0 = fminf(l0, (float)P.x); h0 = fmaxf(h0, (float)P.x);
        l1 = fminf(l1, (float)P.y); h1 = fmaxf(h1, (float)P.y);
        l2 = fminf(l2, (float)P.z); h2 = fmaxf(h2, (float)P.z);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, o)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, o));
        l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, o)); h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, o));
        l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, o)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
    }
    if (lane == 0) {
        const int leaf_lo = col_leaf[s0], leaf_hi = col_leaf[s1 - 1];
        int anc = -1, size = 0;
        if (ninternal > 0 && leaf_lo != leaf_hi) {
            anc = leaf_up[leaf_lo] >> 1;
            while (!(node_range[anc].x <= leaf_lo && leaf_hi <= node_range[anc].y)) anc = node_up[anc] >> 1;
            size = node_range[anc].y - node_range[anc].x + 1;
        }
        info[3 * (size_t)c] = make_float4(l0, l1, l2, __int_as_float(leaf_lo));
        info[3 * (size_t)c + 1] = make_float4(h0, h1, h2, __int_as_float(leaf_hi));
        info[3 * (size_t)c + 2] = make_float4(__int_as_float(anc), __int_as_float(size), 0.f, 0.f);
    }
}

// out of line, last resort, once per unit and only if a second pass overflowed too: the flagged columns' rays,
// each on its own with the generic traversal from the root (trace.cuh: target_visible)
template <class T>
__device__ __noinline__ void redo_unit_cold(const TraceArgs<T> &A, const Real4<T> Pi, int s0, uint32_t flagged,
                                            uint32_t *word) {
    const BvhView bvh{A.nodes, nullptr, A.tri, 0, A.ninternal, A.nfaces, A.error_flag};
    const int lane = threadIdx.x & 31;
    while (flagged) {
        const int bit = __ffs(flagged) - 1;
        flagged &= flagged - 1;
        const int s = s0 + lane * 32 + bit;
        Ray ray;
        if (setup_ray(Pi, load_real4<T>(A.colP + s), ray) && !target_visible(bvh, ray, A.col_leaf[s], A.col_face[s]))
            *word &= ~(1u << bit);
    }
}

// The rays K4 could not finish (a traversal stack or candidate list overflowed: rays that graze very many
// triangles), one thread each with the generic traversal from the root.  An occluded one loses its bit and
// its row's count goes down by one.  Grid-stride over *count (no host round trip to size the launch).
template <class T>
__global__ void resolve_lost_kernel(const TraceArgs<T> A) {
    const BvhView bvh{A.nodes, nullptr, A.tri, 0, A.ninternal, A.nfaces, A.error_flag};
    const unsigned n = min(*A.lost_count, A.lost_cap);
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int2 e = A.lost[k];
        const int i = A.rows[e.x], s = e.y;
        Ray ray;
        if (setup_ray(load_real4<T>(A.faceP + i), load_real4<T>(A.colP + s), ray) &&
            !target_visible(bvh, ray, A.col_leaf[s], A.col_face[s])) {
            const uint32_t bit = 1u << (s & 31);
            const uint32_t old = atomicAnd(&A.bits[(size_t)e.x * A.nwords + (s >> 5)], ~bit);
            if (old & bit) atomicAdd(&A.row_counts[e.x], 0xffffffffu); // -1 (mod 2^32)
        }
    }
}

// ---- the cull of one 32-column tile: keep = abs(num) > eps in the model dtype (form_factors.py:46-52) ----
// float64 models: the numerator in fp64, as K6b recomputes it.
__device__ __forceinline__ bool cull_keep(const Real4<double> &Pi, const Real4<double> &Ni, const Real4<double> &Pj,
                                          const Real4<double> &Nj, double eps, float, float, float) {
    double dx, dy, dz;
    return survives_cull<double>(numerator<double>(Pi, Ni, Pj, Nj, dx, dy, dz), eps);
}
// float32 models.  The exact rule is v = float(num64) > eps with num64 the fp64 numerator of float
// inputs.  In float32: d~ = fl(Pj - Pi), a~ = fl(Ni.d~), b~ = fl(-Nj.d~); with u = 2^-24 and
// L = |d~x| + |d~y| + |d~z|,  |a~ - a| <= 4.5 u max|Ni| L  (one rounding of d, three of the dot
// product), the same for b; so |a~+ b~+ - a+ b+| <= a~+ eb + b~+ ea + ea eb + u a~+ b~+.  With
// g = 8 u in ea = g max|Ni| L, eb = g max|Nj| L the bound E holds with a factor 1.7 to spare and
// covers the roundings of E itself.  Sure keep: num~ - E > eps (1 + 1e-6); sure cull: num~ + E <
// eps (1 - 1e-6) (the 1e-6 covers the roundings of v and of num~).  Anything else -- including
// every NaN / inf, for which both comparisons are false -- takes the fp64 rule.  `gi` = g max|Ni|,
// Nj.w = g max|Nj| (col_gather_kernel).
__device__ __forceinline__ bool cull_keep(const Real4<float> &Pi, const Real4<float> &Ni, const Real4<float> &Pj,
                                          const Real4<float> &Nj, float eps, float gi, float eps_hi, float eps_lo) {
    const float dx = Pj.x - Pi.x, dy = Pj.y - Pi.y, dz = Pj.z - Pi.z;
    const float a = fmaf(Ni.x, dx, fmaf(Ni.y, dy, Ni.z * dz));
    const float b = -fmaf(Nj.x, dx, fmaf(Nj.y, dy, Nj.z * dz));
    const float L = fabsf(dx) + fabsf(dy) + fabsf(dz);
    const float ea = gi * L, eb = Nj.w * L;
    const float ap = fmaxf(a, 0.f), bp = fmaxf(b, 0.f);
    const float num = ap * bp;
    const float E = fmaf(ap, eb, fmaf(bp, ea, ea * eb));
    if (num - E > eps_hi) return true;
    if (num + E < eps_lo) return false;
    double ddx, ddy, ddz;
    return survives_cull<float>(numerator<float>(Pi, Ni, Pj, Nj, ddx, ddy, ddz), eps);
}

// the source face's P / N in the warp's shared block (16-byte accesses: the block is 16-byte aligned)
__device__ __forceinline__ void put_real4(double *dst, const Real4<float> &v) {
    *reinterpret_cast<float4 *>(dst) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void put_real4(double *dst, const Real4<double> &v) {
    reinterpret_cast<double2 *>(dst)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2 *>(dst)[1] = make_double2(v.z, v.w);
}
template <class T> __device__ __forceinline__ Real4<T> get_real4(const double *src);
template <> __device__ __forceinline__ Real4<float> get_real4<float>(const double *src) {
    const float4 v = *reinterpret_cast<const float4 *>(src);
    return Real4<float>{v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Real4<double> get_real4<double>(const double *src) {
    const double2 a = reinterpret_cast<const double2 *>(src)[0], b = reinterpret_cast<const double2 *>(src)[1];
    return Real4<double>{a.x, a.y, b.x, b.y};
}

// (Measured and dropped, r02o: the cull's column loads -- 32 KB per unit, every byte used once by this SM -- through
// L2 only (ld.global.cg) so that L1 keeps BVH records: 28.92 / 28.88 ms against 28.92 ms per slab.)
// make_raybox (trace.cuh) with the hardware's approximate reciprocal (1 ulp) instead of three IEEE divisions
// (~30 instructions per batch): the box test only has to be conservative, and its final comparison carries a
// guard band of 2e-6 relative -- twenty times the error introduced here.
__device__ __forceinline__ RayBox make_raybox_fast(const Ray &r) {
    auto inv = [](float d) {
        const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
        float q;
#ifndef FB_EMU
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(a));
#else
        q = 1.0f / a;
#endif
        return q;
    };
    RayBox rb;
    rb.ix = inv(r.dx);
    rb.iy = inv(r.dy);
    rb.iz = inv(r.dz);
    rb.ox = r.ox * rb.ix;
    rb.oy = r.oy * rb.iy;
    rb.oz = r.oz * rb.iz;
    rb.xod = f2_pack(r.ox, r.dx);
    rb.yod = f2_pack(r.oy, r.dy);
    rb.zod = f2_pack(r.oz, r.dz);
    return rb;
}

#ifdef FB_EMU
#define FB_COUNT2(k, v) FB_COUNT(k, v)
#define FB_CHECK(cond) do { if (!(cond)) { fprintf(stderr, "trace2 check failed line %d: %s\n", __LINE__, #cond); abort(); } } while (0)
#else
#define FB_COUNT2(k, v) ((void)0)
#define FB_CHECK(cond) ((void)0)
#endif

template <class T>
__global__ void __launch_bounds__(kTraceThreads, FB_TRACE_MIN_BLOCKS) trace2_kernel(const TraceArgs<T> A) {
    extern __shared__ float4 smem_dyn[];
    const int lane = threadIdx.x & 31;
    WarpShared *const W = reinterpret_cast<WarpShared *>(smem_dyn) + (threadIdx.x >> 5);
    const uint32_t lt = (1u << lane) - 1u;
    const BvhView bvh{A.nodes, nullptr, A.tri, 0, A.ninternal, A.nfaces, A.error_flag};
    const bool hor = A.hz != nullptr;
    const unsigned total_units = (unsigned)A.m * (unsigned)A.nchunks;
    if (lane < 6) W->ctr[lane] = 0u;
    if (lane == 0) W->tested = 0ull;
    __syncwarp();

    // ---- one batch: up to 32 survivors (column ix of the chunk each) of the current unit --------------
    auto batch = [&](const int ix, const bool valid) {
        const int s0 = W->u[kUchunk] * kChunkCols;
        const int cref = W->u[kUcref];
        // ---- ray set-up and the target's own hit distance (converged) --------------------------
        Ray ray = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
        int tleaf = -1, tface = 0;
        float tj = 0.f, dist_h = 0.f;
        bool active = false, blocked = false, overshoot = false;
        if (valid) {
            const int s = s0 + ix;
            const Real4<T> Pi = get_real4<T>(&W->src[0]);
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
                    dist_h = sqrtf(ex * ex + ey * ey + ez * ez);
                    overshoot = tj * 1.000002f > dist_h + (1e-5f * A.scale + 1e-3f);
                } else {
                    blocked = true; // the ray misses its own target: closest hit is not j
                }
            }
        }
        if (lane == 0) W->novf = 0u, W->ovfhit = 0u;
        __syncwarp();
        const RayBox rb = make_raybox_fast(ray);
        const float tmax = tj * 1.000002f;
        // per-lane stack of hit nodes and list of candidate triangles, both in shared memory: [entry][lane]
        smem_addr_t stk = (smem_addr_t)__cvta_generic_to_shared(&W->stack[0][lane]);
        smem_addr_t cnd = (smem_addr_t)__cvta_generic_to_shared(&W->cand[0][lane]);
        asm volatile("" : "+r"(stk), "+r"(cnd)); // keep them: do not rebuild the addresses at every push
        int sp = 0, nl = 0;
        bool lost = false; // something did not fit: the column is traced again, on its own, at the end of the unit
        auto push = [&](int ref) {
            if (ref < 0) {
                if (nl < kCand) sts_i1(cnd + (uint32_t)(nl++) * 128u, ~ref);
                else { // my list is full: the warp's shared list, tested by all lanes at the end of the batch
                    const uint32_t at = atomicAdd(&W->novf, 1u);
                    if (at < (uint32_t)kOvf) W->ovf[at] = ((uint32_t)lane << kNodeBits) | (uint32_t)~ref;
                    else lost = true, FB_COUNT2(2, 1);
                }
            } else {
                if (sp < kStk) sts_i1(stk + (uint32_t)(sp++) * 128u, ref);
                else lost = true, FB_COUNT2(7, 1);
            }
        };
        // ---- phase A: the listed records, the same for all lanes ---------------------------------
        int xref = ~tleaf;
        bool xbig = false; // the record that holds my target is larger than any zone
        {
            const bool fullpath = __any_sync(0xffffffffu, active && overshoot);
            int nuse = fullpath ? W->u[kUnall] : W->u[kUnsel];
            if (hor) {
                // source end: every ray of the batch leaves above the horizon of zone(i) -> no triangle
                // of the zone other than i can be met: the records inside the zone are not walked
                bool clear = true; // a lane without a ray does not object
                if (active) {
                    const Real4<T> Ni = get_real4<T>(&W->src[4]);
                    const float si = (float)Ni.x * ray.dx + (float)Ni.y * ray.dy + (float)Ni.z * ray.dz;
                    clear = si > __int_as_float(W->u[kUhor]);
                }
                const bool skip = __all_sync(0xffffffffu, clear) && !fullpath;
                if (skip) nuse = W->u[kUnout];
                if (lane == 0) W->ctr[kCbatches] += 1u, W->ctr[kCsrc] += skip ? 1u : 0u;
            }
            if (lane == 0) FB_COUNT2(0, 1), FB_COUNT2(1, nuse);
            const float tmax_a = active ? tmax : -1.0f; // a lane without a ray never hits
            smem_addr_t addr = (smem_addr_t)__cvta_generic_to_shared(&W->path[0]);
            if (cref != -0x7fffffff) { // X is not in the list: nothing to check per record
                for (int ks = 0; ks < nuse; ++ks, addr += 48) {
                    const float4 a = lds_f4(addr), b = lds_f4(addr + 16), cc = lds_f4(addr + 32);
                    if (child_hit<FB_T2_PACKED & 1>(ray, rb, a, b, cc, tmax_a)) push(rec_ref(cc));
                }
            } else { // the chunk straddles several records: every lane skips the one holding its target
                smem_addr_t raddr = (smem_addr_t)__cvta_generic_to_shared(&W->range[0]);
                for (int ks = 0; ks < nuse; ++ks, addr += 48, raddr += 8) {
                    const float4 a = lds_f4(addr), b = lds_f4(addr + 16), cc = lds_f4(addr + 32);
                    const int2 rg = lds_i2(raddr);
                    const int ref = rec_ref(cc);
                    if (tleaf >= rg.x && tleaf <= rg.y) {
                        xref = ref;
                        xbig = rg.y - rg.x + 1 > A.zone_leaves;
                    } else if (child_hit<FB_T2_PACKED & 1>(ray, rb, a, b, cc, tmax_a))
                        push(ref);
                }
            }
        }
        // ---- phase B: from the target leaf up to X (or C), the sibling at every level -- one 48-byte
        // record per level.  `code` = (parent << 1 | my slot), so the sibling is record code ^ 1.
        bool tskipped = false;
        {
            const int stop = cref != -0x7fffffff ? cref : xref;
            int cur = ~tleaf, code = tleaf >= 0 ? A.leaf_up[tleaf] : -1;
            if (hor && active && (W->u[kUtskip] || xbig)) {
                // target end: the ray arrives above the horizon of zone(j) and the tested interval does
                // not run on past p_j as far as the zone's nearest other triangle -> nothing in zone(j)
                // except j can be met: the walk starts at the zone's node
                const float4 h = __ldg(A.colH + s0 + ix);
                const Real4<T> Nj = load_real4<T>(A.colN + s0 + ix);
                const float st = -((float)Nj.x * ray.dx + (float)Nj.y * ray.dy + (float)Nj.z * ray.dz);
                const float beyond = tmax - (dist_h - ray_eps()); // ideal hit: dist - 1e-3 along the ray
                const int zn = __float_as_int(h.z);
                if (zn >= 0 && st > h.x && beyond + 2.0f * A.pert < h.y) {
                    cur = zn;
                    code = __float_as_int(h.w);
                    tskipped = true;
                }
            }
            while (active && cur != stop && code >= 0) {
                const float4 *rec = A.nodes + 3 * (size_t)(unsigned)(code ^ 1);
                const float4 a = __ldg(rec), b = __ldg(rec + 1), cc = __ldg(rec + 2);
                cur = code >> 1;
                code = A.node_up[cur];
                if (child_hit<FB_T2_PACKED & 1>(ray, rb, a, b, cc, tmax)) push(rec_ref(cc));
            }
        }
        // ---- phase C: the subtrees that were actually hit, top-down, every lane on its own -------------
#if FB_C_FORM == 1
        { // the loop form of the first-generation kernel: one flag, one select, pop or stop when nothing was hit
            int node = 0;
            bool walking = active && sp > 0;
            if (walking) node = lds_i1(stk + (uint32_t)(--sp) * 128u);
            while (walking) {
                float4 q[6];
                load_node<false>(bvh, node, q);
                const bool h0 = child_hit<(FB_T2_PACKED >> 1) & 1>(ray, rb, q[0], q[1], q[2], tmax);
                const bool h1 = child_hit<(FB_T2_PACKED >> 1) & 1>(ray, rb, q[3], q[4], q[5], tmax);
                const int r0 = rec_ref(q[2]), r1 = rec_ref(q[5]);
                if (h0 && r0 < 0 && ~r0 != tleaf) push(r0);
                if (h1 && r1 < 0 && ~r1 != tleaf) push(r1);
                const bool i0 = h0 && r0 >= 0, i1 = h1 && r1 >= 0;
                if (i0 && i1) push(r1);
                node = i0 ? r0 : r1;
                if (!(i0 || i1)) {
                    if (sp > 0) node = lds_i1(stk + (uint32_t)(--sp) * 128u);
                    else walking = false;
                }
            }
        }
#else
        if (active && sp > 0) {
            int node = lds_i1(stk + (uint32_t)(--sp) * 128u);
            while (true) {
                float4 q[6];
                load_node<false>(bvh, node, q);
                const bool h0 = child_hit<(FB_T2_PACKED >> 1) & 1>(ray, rb, q[0], q[1], q[2], tmax);
                const bool h1 = child_hit<(FB_T2_PACKED >> 1) & 1>(ray, rb, q[3], q[4], q[5], tmax);
                const int r0 = rec_ref(q[2]), r1 = rec_ref(q[5]);
                if (h0 && r0 < 0 && ~r0 != tleaf) push(r0);
                if (h1 && r1 < 0 && ~r1 != tleaf) push(r1);
                const bool i0 = h0 && r0 >= 0, i1 = h1 && r1 >= 0;
                if (i0 && i1) push(r1);
                if (i0) node = r0;
                else if (i1) node = r1;
                else if (sp > 0) node = lds_i1(stk + (uint32_t)(--sp) * 128u);
                else break;
            }
        }
#endif
        // (Measured and dropped, profiles/r02d_*: the same loop written for predication -- pushes as selects, one
        // backward branch -- ran 34.5 ms per slab against 32.5 ms for this form.)
        if (hor) {
            const uint32_t ts = __ballot_sync(0xffffffffu, tskipped);
            if (lane == 0) W->ctr[kCtgt] += (unsigned)__popc(ts);
        }
        { // converged exact tests of the listed candidates
            int mx = nl;
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            for (int k = 0; k < mx; ++k)
                if (k < nl && !blocked) blocked = leaf_occludes(bvh, ray, tj, lds_i1(cnd + (uint32_t)k * 128u), tface);
        }
        __syncwarp();
        { // the shared surplus list: every lane takes one entry and tests it against its owner's ray
            const int n = min((int)W->novf, kOvf);
            for (int base = 0; base < n; base += 32) {
                const bool have = base + lane < n;
                const uint32_t e = have ? W->ovf[base + lane] : 0u;
                const int src = (int)(e >> kNodeBits);
                Ray rs;
                rs.ox = __shfl_sync(0xffffffffu, ray.ox, src); rs.oy = __shfl_sync(0xffffffffu, ray.oy, src);
                rs.oz = __shfl_sync(0xffffffffu, ray.oz, src); rs.dx = __shfl_sync(0xffffffffu, ray.dx, src);
                rs.dy = __shfl_sync(0xffffffffu, ray.dy, src); rs.dz = __shfl_sync(0xffffffffu, ray.dz, src);
                const float tjs = __shfl_sync(0xffffffffu, tj, src);
                const int tfs = __shfl_sync(0xffffffffu, tface, src);
                if (have && leaf_occludes(bvh, rs, tjs, (int)(e & ((1u << kNodeBits) - 1u)), tfs)) atomicOr(&W->ovfhit, 1u << src);
            }
            if (n > 0) {
                __syncwarp();
                if ((W->ovfhit >> lane) & 1u) blocked = true;
            }
        }
        if (lost && !blocked) {
            // to the launch's list of unresolved rays (resolve_lost_kernel traces them, one thread each, after
            // this kernel); should that list be full: flagged for the in-kernel cold pass at the end of the unit
            const unsigned at = atomicAdd(A.lost_count, 1u);
            if (at < A.lost_cap) A.lost[at] = make_int2(W->u[kUrow], s0 + ix);
            else atomicOr(&W->redo[ix >> 5], 1u << (ix & 31));
            atomicAdd(&W->ctr[kCcold], 1u);
        }
        if (blocked) atomicAnd(&W->words[ix >> 5], ~(1u << (ix & 31)));
        __syncwarp();
    };

    while (true) {
        unsigned unit = 0;
        if (lane == 0) unit = (unsigned)atomicAdd(A.tested + 1, 1ull);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= total_units) break;
        const int r = (int)(unit / (unsigned)A.nchunks), c = (int)(unit - (unsigned)r * (unsigned)A.nchunks);
        const int i = A.rows[r];
        {
            const Real4<T> Pi = load_real4<T>(A.faceP + i), Ni = load_real4<T>(A.faceN + i);
            if (lane == 0) {
                put_real4(&W->src[0], Pi);
                put_real4(&W->src[4], Ni);
                W->u[kUrow] = r;
                W->u[kUchunk] = c;
            }
        }
        const float4 ci0 = __ldg(A.chunk_info + 3 * (size_t)c), ci1 = __ldg(A.chunk_info + 3 * (size_t)c + 1),
                     ci2 = __ldg(A.chunk_info + 3 * (size_t)c + 2);
        const int leaf_lo = __float_as_int(ci0.w), leaf_hi = __float_as_int(ci1.w);
        // ---- the unit's record list ---------------------------------------------------------------
        // Every ray of the unit starts on triangle i: list the child records hanging off the root ->
        // leaf(i) path once (the leaf's own record, then the sibling at every level); when one of them
        // (X) holds the whole chunk, the siblings between X and the chunk's common ancestor C follow.
        {
            int npath = 0, nsel = 0, nall = 0, nout = 0;
            int cref = -0x7fffffff; // C when usable
            int zlo = 0, zhi = -1;
            float hor_i = INFINITY;
            bool tskip_unit = false;
            if (A.ninternal > 0) {
                const int ileaf = A.face_leaf[i];
                uint32_t *recs = W->recs;  // record index (node << 1 | slot) of every entry
                {
                    int code = A.leaf_up[ileaf];
                    if (lane == 0) recs[0] = (uint32_t)code;
                    npath = 1;
                    while (code >= 0 && npath < kPathCap) { // (the host checked depth + 1 <= kPathCap)
                        if (lane == 0) recs[npath] = (uint32_t)(code ^ 1);
                        ++npath;
                        code = A.node_up[code >> 1];
                    }
                }
                __syncwarp();
                float4 ra[2] = {}, rb2[2] = {}, rc[2] = {};
                int2 rr[2];
                auto load_entries = [&](int first, int last) { // entries [first, last): 32 at a time, straight to registers
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = h * 32 + lane;
                        if (e >= first && e < last) {
                            const float4 *rec = A.nodes + 3 * (size_t)recs[e];
                            ra[h] = __ldg(rec);
                            rb2[h] = __ldg(rec + 1);
                            rc[h] = __ldg(rec + 2);
                            const int ref = rec_ref(rc[h]);
                            rr[h] = ref < 0 ? make_int2(~ref, ~ref) : A.node_range[ref];
                        }
                    }
                };
                rr[0] = rr[1] = make_int2(0x7fffffff, -1);
                load_entries(0, npath);
                int xdrop = -1; // X, taken out of the list when every target of the chunk is under it
                if (leaf_lo != leaf_hi) {
                    const uint32_t m0 = __ballot_sync(0xffffffffu, lane < npath && rr[0].x <= leaf_lo && leaf_hi <= rr[0].y);
                    const uint32_t m1 = __ballot_sync(0xffffffffu, 32 + lane < npath && rr[1].x <= leaf_lo && leaf_hi <= rr[1].y);
                    const int xe = m1 ? 63 - __clz(m1) : (m0 ? 31 - __clz(m0) : -1);
                    if (xe >= 0) {
                        const int xr = __shfl_sync(0xffffffffu, __float_as_int(xe < 32 ? rc[0].w : rc[1].w), xe & 31);
                        int cur = __float_as_int(ci2.x), n2 = npath;
                        bool complete = true;
                        while (cur != xr) {
                            const int up = A.node_up[cur];
                            if (up < 0 || n2 >= kPathCap) {
                                complete = false;
                                break;
                            }
                            if (lane == 0) recs[n2] = (uint32_t)(up ^ 1);
                            ++n2;
                            cur = up >> 1;
                        }
                        if (complete) {
                            __syncwarp();
                            load_entries(npath, n2);
                            npath = n2;
                            cref = __float_as_int(ci2.x);
                            xdrop = xe;
                        }
                    }
                }
                if (hor) { // horizon skip: leaf range of zone(i), the source face's horizon (horizon.cuh)
                    const int zn = A.zone_node[ileaf];
                    zlo = zhi = ileaf;
                    if (zn >= 0) {
                        const int2 zr = A.node_range[zn];
                        zlo = zr.x;
                        zhi = zr.y;
                    }
                    hor_i = __ldg(A.hz + i).x;
                    tskip_unit = cref != -0x7fffffff && __float_as_int(ci2.y) > A.zone_leaves;
                }
                // ---- shaft filter + partition: [kept, outside zone(i)] [kept, inside] [not kept]; X dropped ----
                // Every ray of the unit lies in the convex hull of the source centroid and the chunk's target
                // centroids: a record separated from that hull along x, y, z or its slab direction cannot be hit.
                const Real4<T> Pi = load_real4<T>(A.faceP + i);
                const float px = (float)Pi.x, py = (float)Pi.y, pz = (float)Pi.z;
                const float pad = 3e-5f * A.scale + 1e-4f * fmaxf(fmaxf(ci1.x - ci0.x, ci1.y - ci0.y), ci1.z - ci0.z) + 2e-3f;
                const float h0l = fminf(px, ci0.x) - pad, h0h = fmaxf(px, ci1.x) + pad;
                const float h1l = fminf(py, ci0.y) - pad, h1h = fmaxf(py, ci1.y) + pad;
                const float h2l = fminf(pz, ci0.z) - pad, h2h = fmaxf(pz, ci1.z) + pad;
                bool keepr[2], other[2], inzone[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = h * 32 + lane;
                    keepr[h] = other[h] = inzone[h] = false;
                    if (e < npath) {
                        const float4 a = ra[h], b = rb2[h], cc = rc[h];
                        const int2 rg = rr[h];
                        bool keep = true;
                        if (A.shaft_filter && (rg.y < leaf_lo || rg.x > leaf_hi)) { // holds no target of this chunk
                            // record layout: a = (lo.x, hi.x, lo.y, hi.y), b = (lo.z, hi.z, slab_min, slab_max), cc = (dir | ref)
                            if (a.x > h0h || a.y < h0l || a.z > h1h || a.w < h1l || b.x > h2h || b.y < h2l) keep = false;
                            const float sp = cc.x * px + cc.y * py + cc.z * pz;
                            const float lo_s = fminf(cc.x * ci0.x, cc.x * ci1.x) + fminf(cc.y * ci0.y, cc.y * ci1.y) + fminf(cc.z * ci0.z, cc.z * ci1.z);
                            const float hi_s = fmaxf(cc.x * ci0.x, cc.x * ci1.x) + fmaxf(cc.y * ci0.y, cc.y * ci1.y) + fmaxf(cc.z * ci0.z, cc.z * ci1.z);
                            const float spad = pad * (fabsf(cc.x) + fabsf(cc.y) + fabsf(cc.z));
                            if (fminf(sp, lo_s) - spad > b.w || fmaxf(sp, hi_s) + spad < b.z) keep = false;
                        }
                        if (e != xdrop) {
                            keepr[h] = keep;
                            other[h] = !keep;
                            inzone[h] = hor && keep && e != 0 && rg.x >= zlo && rg.y <= zhi;
                        }
                    }
                }
                const uint32_t kb0 = __ballot_sync(0xffffffffu, keepr[0]), kb1 = __ballot_sync(0xffffffffu, keepr[1]);
                const uint32_t ob0 = __ballot_sync(0xffffffffu, other[0]), ob1 = __ballot_sync(0xffffffffu, other[1]);
                const uint32_t zb0 = __ballot_sync(0xffffffffu, inzone[0]), zb1 = __ballot_sync(0xffffffffu, inzone[1]);
                nsel = __popc(kb0) + __popc(kb1);
                nall = nsel + __popc(ob0) + __popc(ob1);
                nout = nsel - __popc(zb0) - __popc(zb1);
                const uint32_t k0 = kb0 & ~zb0, k1 = kb1 & ~zb1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int d = -1;
                    if (inzone[h]) d = nout + (h ? __popc(zb0) : 0) + __popc((h ? zb1 : zb0) & lt);
                    else if (keepr[h]) d = (h ? __popc(k0) : 0) + __popc((h ? k1 : k0) & lt);
                    else if (other[h]) d = nsel + (h ? __popc(ob0) : 0) + __popc((h ? ob1 : ob0) & lt);
                    if (d >= 0) {
                        FB_CHECK(d < kPathCap);
                        W->path[3 * d] = ra[h];
                        W->path[3 * d + 1] = rb2[h];
                        W->path[3 * d + 2] = rc[h];
                        W->range[d] = rr[h];
                    }
                }
            }
            if (lane == 0) {
                W->u[kUnsel] = nsel;
                W->u[kUnall] = nall;
                W->u[kUnout] = nout;
                W->u[kUcref] = cref;
                W->u[kUtskip] = tskip_unit ? 1 : 0;
                W->u[kUhor] = __float_as_int(hor_i);
            }
            __syncwarp();
        }
        // ---- cull (two 32-column tiles at a time) -> survivor FIFO -> batches of 32 rays ---------------
        {
            uint32_t fhead = 0;
            int fcnt = 0;
            unsigned nsurv = 0;
            const int s0 = c * kChunkCols;
            for (int k = 0; k < 32; k += 2) {
                {
                    const Real4<T> Pi = get_real4<T>(&W->src[0]);
                    const Real4<T> Ni = get_real4<T>(&W->src[4]);
                    float gi = 0.f, eps_hi = 0.f, eps_lo = 0.f;
                    if (sizeof(T) == 4) {
                        const float e = (float)A.eps;
                        gi = 4.76837158e-7f * fmaxf(fmaxf(fabsf((float)Ni.x), fabsf((float)Ni.y)), fabsf((float)Ni.z));
                        eps_hi = e + fabsf(e) * 1e-6f + 2.4e-38f;
                        eps_lo = e - fabsf(e) * 1e-6f;
                    }
                    const int sa = s0 + k * 32 + lane, sb = sa + 32;
                    bool keep0 = false, keep1 = false;
                    // (the diagonal needs no test: j == i gives d == 0 and a zero numerator)
                    if (sb < A.n) { // both tiles inside: four loads in flight
                        const Real4<T> Pa = load_real4<T>(A.colP + sa), Na = load_real4<T>(A.colN + sa);
                        const Real4<T> Pb = load_real4<T>(A.colP + sb), Nb = load_real4<T>(A.colN + sb);
                        keep0 = cull_keep(Pi, Ni, Pa, Na, A.eps, gi, eps_hi, eps_lo);
                        keep1 = cull_keep(Pi, Ni, Pb, Nb, A.eps, gi, eps_hi, eps_lo);
                    } else if (sa < A.n) {
                        const Real4<T> Pa = load_real4<T>(A.colP + sa), Na = load_real4<T>(A.colN + sa);
                        keep0 = cull_keep(Pi, Ni, Pa, Na, A.eps, gi, eps_hi, eps_lo);
                    }
                    const uint32_t w0 = __ballot_sync(0xffffffffu, keep0), w1 = __ballot_sync(0xffffffffu, keep1);
                    if (lane == 0) {
                        W->words[k] = w0, W->words[k + 1] = w1;
                        W->redo[k] = W->redo[k + 1] = 0u;
                    }
                    const int n0 = __popc(w0);
                    if (keep0) W->fifo[(fhead + fcnt + __popc(w0 & lt)) & (kFifoCap - 1)] = (uint16_t)(k * 32 + lane);
                    if (keep1) W->fifo[(fhead + fcnt + n0 + __popc(w1 & lt)) & (kFifoCap - 1)] = (uint16_t)(k * 32 + 32 + lane);
                    fcnt += n0 + __popc(w1);
                    nsurv += (unsigned)(n0 + __popc(w1));
                }
                __syncwarp();
                while (fcnt >= 32 || (k == 30 && fcnt > 0)) {
                    const int nin = min(32, fcnt);
                    const bool valid = lane < nin;
                    const int ix = valid ? (int)W->fifo[(fhead + lane) & (kFifoCap - 1)] : 0;
                    fhead += nin;
                    fcnt -= nin;
                    batch(ix, valid);
                }
            }
            if (lane == 0) W->tested += nsurv;
        }
        // ---- publish -----------------------------------------------------------------------------------
        __syncwarp();
        {
            uint32_t fin = W->words[lane];
            const int rr = W->u[kUrow], cc = W->u[kUchunk];
            const uint32_t again = W->redo[lane] & fin;
            if (__any_sync(0xffffffffu, again != 0u))
                redo_unit_cold<T>(A, get_real4<T>(&W->src[0]), cc * kChunkCols, again, &fin);
            const int wi = cc * 32 + lane;
            if (wi < A.nwords) A.bits[(size_t)rr * A.nwords + wi] = fin;
            unsigned count = __popc(fin);
#pragma unroll
            for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
            if (lane == 0 && count) atomicAdd(&A.row_counts[rr], count);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (W->tested) atomicAdd(A.tested, W->tested);
        atomicAdd(A.tested + 2, (unsigned long long)W->ctr[kCbatches]);
        atomicAdd(A.tested + 3, (unsigned long long)W->ctr[kCsrc]);
        atomicAdd(A.tested + 4, (unsigned long long)W->ctr[kCtgt]);
        atomicAdd(A.tested + 7, (unsigned long long)W->ctr[kCcold]);
    }
}

} // namespace fluxb200
