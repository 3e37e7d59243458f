// fluxb200.cu -- host side of libfluxb200.so: the opaque mesh handle and the
// C ABI declared in include/fluxb200.h.  All device work of one handle is
// enqueued on the handle's own stream.
#include "../../include/fluxb200.h"
#include <algorithm>
#include <chrono>
#include <functional>
#include <memory>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "assemble.cuh"
#include "common.cuh"
#include "lbvh.cuh"
#include "prims.cuh"
#include "spmv.cuh"
#include "blockops.cuh"
#include "horizon.cuh"
#include "trace.cuh"
#include "trace2.cuh"
#include "host_expand.h"


namespace fluxb200 {
thread_local std::string g_last_error;
}
using namespace fluxb200;

// One prepared column set J of an assembly: the columns sorted by leaf (Morton) position and everything gathered
// in that order.  The handle keeps the last few (LRU): CompressedFormFactorMatrix assembles 16 / 64 root blocks
// over 4 / 8 column parts (reference src/flux/compressed_form_factors.py:551-567), and the sort + gathers of a
// part are the same for every row part.
struct ColSet {
    DevBuf cols, colP, colN, col_face, col_leaf, rank_of_pos, chunk_info, colH;
    std::vector<int64_t> key; // the caller's J (empty when arange)
    bool arange = false, valid = false;
    size_t n = 0;
    uint64_t face_epoch = 0, tree_epoch = 0, used = 0;
    void release() {
        for (DevBuf *b : {&cols, &colP, &colN, &col_face, &col_leaf, &rank_of_pos, &chunk_info, &colH}) b->release();
        valid = false;
    }
};

struct fluxb200_mesh {
    int device = 0;
    int dtype = FLUXB200_F32;
    size_t nv = 0, nf = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    int num_sms = 148;
    int max_smem_optin = 48 * 1024;

    DevBuf V, F, V32, faceP, faceN; // geometry
    // LBVH
    DevBuf keys, vals, left, right, parent, first, last, box, slab, flags, pre, flag_by_pre, top_before,
        scene, scalars, nodes, tri, face_leaf, node_up, leaf_up, node_range;
    RadixSorter sorter;
    int nnodes = 0, ninternal = 0, ntop = 0, max_depth = 0;
    int top_nodes_opt = 0;   // measured: L1 already serves the top of the tree (profiles/)
    int slab_limit_opt = 1 << 30;
    int blocks_per_sm = 4;
    int shaft_filter_opt = 1;
    // horizon skip of the trace kernel (horizon.cuh).  On by default since round 2: bit-identical CSRs over the
    // whole gpu tier on a B200 with the skip on and off, trace kernel 49.2 -> 34.5 ms per 4096-row slab of the
    // 200k-face crater (profiles/r02a_*); the horizons cost 6.0 ms once per mesh / change of P, N.
    int horizon_skip_opt = 1;
    int horizon_zone_opt = 1023; // Z: leaves per near zone (below the 1024-column chunk: the upward walk must end above every zone)
    bool hz_dirty = true;       // P, N or the tree changed since the horizons were computed
    DevBuf hz, zone_node, zone_up;
    // trace kernel generation: 2 = warp-shared traversal queue (trace2.cuh), 1 = per-lane stacks (assemble.cuh)
    int trace_variant_opt = 2;
    DevBuf lost; // lost: [0] count, then (row, column) pairs of rays for resolve_lost_kernel
    float ms_build = 0.f;
    float scene_h[7] = {};

    // per-call state
    DevBuf rows, ckeys, cvals, bits, row_counts, counts64, indptr, indptr32, tested, out_data, out_indices, qtmp,
        qout, jbits;
    // prepared column sets: [0] is the scratch set of the query kernels, [1..] the LRU cache of assemblies
    std::vector<std::unique_ptr<ColSet>> colsets;
    ColSet *cs = nullptr;          // the set the current call works on
    uint64_t face_epoch = 1;       // bumped when P or N changes on the device
    uint64_t tree_epoch = 1;       // bumped when the tree (or an option baked into per-column data) changes
    uint64_t use_clock = 0;
    int colset_cache_opt = 8;      // cached column sets (0: prepare every call)
    int64_t colset_hits = 0, colset_misses = 0;
    // streaming assembly (double-buffered sub-slabs, second stream for fill + D2H)
    cudaStream_t copy_stream = nullptr; // fill kernels
    cudaStream_t d2h_stream = nullptr;  // copy-out (its own stream: fill(k+1) must not queue behind D2H(k))
    static constexpr int kSlots = 3;
    cudaEvent_t slot_free[kSlots] = {};
    cudaEvent_t d2h_done[kSlots] = {};
    std::vector<cudaEvent_t> sub_events;
    std::vector<cudaEvent_t> tl_events; // FLUXB200_TIMELINE: fill / copy-out brackets per sub-slab
    DevBuf sbits[kSlots], scounts[kSlots], scounts64[kSlots], sindptr[kSlots], stage_data[kSlots], stage_idx[kSlots],
        sjbits[kSlots];
    HostBuf h_nnz, h_counts;
    int64_t dev_capacity_hint = 0;
    int out_index_width = 0; // index width of the library-owned device CSR (0: none)
    int sub_rows_opt = 512;
    int ramp_opt = 1; // host output: ramp of short first sub-slabs + fill(k) before trace(k+1) (0: round-2 r02i behaviour)
    // host copy-out: ship the J-order visibility words instead of the column indices and let a few
    // host threads write the indices (host_expand.cpp)
    static constexpr int kHostSlots = 8; // ring of page-locked word buffers (sub-slabs in flight on the host)
    HostBuf h_jbits;
    HostBuf h_data[kHostSlots]; // destination 3 (pageable output): page-locked staging of a sub-slab's values
    HostExpander expander;
    int host_expand_opt = 1;
    int host_threads_opt = 0; // 0 = automatic
    int fill_rows_opt = 0;   // rows per CTA of the un-permute kernel: 0 = as many as fit, -1 = no shared memory
    size_t m = 0, n = 0;
    int nwords = 0;
    double eps = 0;
    int64_t nnz = 0;
    bool have_count = false;
    fluxb200_ff_stats stats = {};

    size_t esize() const { return dtype == FLUXB200_F64 ? 8 : 4; }
};

struct fluxb200_csr {
    int device = 0;
    int dtype = FLUXB200_F32;
    int index_width = 4;
    int64_t m = 0, n = 0, nnz = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {};
    DevBuf indptr, indices, data, scratch;
    float last_ms = 0.f;
};

namespace {

template <class F> int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const CudaError &e) {
        set_error(e.msg);
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

inline int blocks_for(int64_t n, int threads) { return (int)std::max<int64_t>(1, ceil_div(n, threads)); }

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        FB_CUDA(cudaSetDevice(dev));
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

// int64 host index array -> validated int32 (NULL = arange)
std::vector<int> to_i32(const int64_t *a, size_t len, size_t bound, const char *what) {
    std::vector<int> out(len);
    if (!a) {
        if (len == 0) return out;
        FB_REQUIRE(len == bound, std::string(what) + ": NULL index array means arange(num_faces)");
        for (size_t k = 0; k < len; ++k) out[k] = (int)k;
        return out;
    }
    for (size_t k = 0; k < len; ++k) {
        const int64_t v = a[k];
        if (v < 0 || (size_t)v >= bound)
            throw CudaError{std::string(what) + ": index out of range"};
        out[k] = (int)v;
    }
    return out;
}

template <class T> void build_geometry(fluxb200_mesh *M) {
    const int nf = (int)M->nf;
    if (nf)
        face_geometry_kernel<T><<<blocks_for(nf, 256), 256, 0, M->stream>>>(
            M->V.as<T>(), M->F.as<int>(), nf, M->faceP.as<Real4<T>>(), M->faceN.as<Real4<T>>());
    const size_t n3 = 3 * M->nv;
    if (n3)
        vertices_to_f32_kernel<T><<<blocks_for((int64_t)n3, 256), 256, 0, M->stream>>>(
            M->V.as<T>(), n3, M->V32.as<float>());
    FB_CUDA(cudaGetLastError());
}

void bvh_build(fluxb200_mesh *M) {
    const int n = (int)M->nf;
    cudaStream_t st = M->stream;
    M->nnodes = n ? 2 * n - 1 : 0;
    M->ninternal = n > 1 ? n - 1 : 0;
    M->ntop = 0;
    M->max_depth = 0;
    M->scalars.reserve(sizeof(int) * 8);
    FB_CUDA(cudaMemsetAsync(M->scalars.p, 0, sizeof(int) * 8, st));
    if (n == 0) return;
    const int nn = 2 * n - 1;
    M->keys.reserve(sizeof(uint64_t) * n);
    M->vals.reserve(sizeof(uint32_t) * n);
    M->left.reserve(sizeof(int) * n);
    M->right.reserve(sizeof(int) * n);
    M->first.reserve(sizeof(int) * n);
    M->last.reserve(sizeof(int) * n);
    M->parent.reserve(sizeof(int) * nn);
    M->box.reserve(sizeof(float) * 9 * nn);
    M->slab.reserve(sizeof(unsigned) * 2 * nn);
    M->flags.reserve(sizeof(int) * n);
    M->pre.reserve(sizeof(int) * n);
    M->flag_by_pre.reserve(sizeof(int) * n);
    M->top_before.reserve(sizeof(int) * n);
    M->scene.reserve(sizeof(unsigned) * 8);
    M->nodes.reserve(sizeof(float4) * 6 * std::max(n - 1, 1));
    M->tri.reserve(sizeof(float4) * 3 * n);
    M->face_leaf.reserve(sizeof(int) * n);
    M->node_up.reserve(sizeof(int) * n);
    M->leaf_up.reserve(sizeof(int) * n);
    M->node_range.reserve(sizeof(int2) * n);
    FB_CUDA(cudaMemsetAsync(M->leaf_up.p, 0xff, sizeof(int) * n, st)); // single-face mesh: no parent

    FB_CUDA(cudaEventRecord(M->ev[0], st));
    const unsigned scene_init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    FB_CUDA(cudaMemcpyAsync(M->scene.p, scene_init, sizeof(scene_init), cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaMemsetAsync(M->parent.p, 0xff, sizeof(int) * nn, st));
    FB_CUDA(cudaMemsetAsync(M->flags.p, 0, sizeof(int) * n, st));
    const int B = 256, G = blocks_for(n, B);
    scene_bounds_kernel<<<G, B, 0, st>>>(M->V32.as<float>(), M->F.as<int>(), n, M->scene.as<unsigned>());
    morton_kernel<<<G, B, 0, st>>>(M->V32.as<float>(), M->F.as<int>(), n, M->scene.as<unsigned>(),
                                   M->keys.as<uint64_t>(), M->vals.as<uint32_t>());
    FB_CUDA(cudaGetLastError());
    M->sorter.sort(M->keys.as<uint64_t>(), M->vals.as<uint32_t>(), n, 63, st);
    if (n > 1)
        karras_kernel<<<blocks_for(n - 1, B), B, 0, st>>>(M->keys.as<uint64_t>(), n, M->left.as<int>(),
                                                          M->right.as<int>(), M->parent.as<int>(),
                                                          M->first.as<int>(), M->last.as<int>());
    refit_kernel<<<G, B, 0, st>>>(M->V32.as<float>(), M->F.as<int>(), n, M->vals.as<uint32_t>(),
                                  M->left.as<int>(), M->right.as<int>(), M->parent.as<int>(),
                                  M->scene.as<unsigned>(), M->box.as<float>(), M->flags.as<int>(),
                                  M->tri.as<float4>(), M->face_leaf.as<int>());
    slab_init_kernel<<<blocks_for(nn, B), B, 0, st>>>(nn, M->box.as<float>(), M->slab.as<unsigned>());
    slab_extent_kernel<<<G, B, 0, st>>>(n, M->tri.as<float4>(), M->parent.as<int>(), M->first.as<int>(),
                                        M->last.as<int>(), M->box.as<float>(), M->slab_limit_opt,
                                        M->slab.as<unsigned>());
    int *scal = M->scalars.as<int>();
    if (n > 1) {
        // "top" = the (at most top_nodes) internal nodes with the largest subtrees: an
        // upward-closed set, so a threshold on the leaf count selects it
        int threshold = n;
        if (M->top_nodes_opt > 0 && n - 1 > 0) {
            std::vector<int> fi(n - 1), la(n - 1);
            FB_CUDA(cudaMemcpyAsync(fi.data(), M->first.p, sizeof(int) * (n - 1), cudaMemcpyDeviceToHost, st));
            FB_CUDA(cudaMemcpyAsync(la.data(), M->last.p, sizeof(int) * (n - 1), cudaMemcpyDeviceToHost, st));
            FB_CUDA(cudaStreamSynchronize(st));
            for (int k = 0; k < n - 1; ++k) fi[k] = la[k] - fi[k] + 1; // leaf counts
            const int budget = M->top_nodes_opt;
            if (n - 1 <= budget) threshold = 0;
            else {
                std::nth_element(fi.begin(), fi.begin() + budget, fi.end(), std::greater<int>());
                threshold = fi[budget]; // counts strictly above the (budget+1)-th largest: <= budget nodes
            }
        }
        preorder_kernel<<<blocks_for(nn, B), B, 0, st>>>(n, M->left.as<int>(), M->parent.as<int>(),
                                                         M->first.as<int>(), M->last.as<int>(),
                                                         threshold,
                                                         M->pre.as<int>(), M->flag_by_pre.as<int>(), scal + 1);
        scan_exclusive<int, int>(M->flag_by_pre.as<int>(), M->top_before.as<int>(), n - 1, scal + 0, st);
        flatten_kernel<<<blocks_for(n - 1, B), B, 0, st>>>(n, M->left.as<int>(), M->right.as<int>(),
                                                           M->pre.as<int>(), M->top_before.as<int>(),
                                                           M->flag_by_pre.as<int>(), scal + 0,
                                                           M->box.as<float>(), M->slab.as<unsigned>(),
                                                           M->scene.as<unsigned>(), M->nodes.as<float4>(),
                                                           M->node_up.as<int>(), M->leaf_up.as<int>(),
                                                           M->first.as<int>(), M->last.as<int>(),
                                                           M->node_range.as<int2>());
    }
    M->hz_dirty = true;
    ++M->tree_epoch; // every prepared column set is stale
    if (M->horizon_skip_opt) {
        M->zone_node.reserve(sizeof(int) * n);
        M->zone_up.reserve(sizeof(int) * n);
        zone_kernel<<<G, B, 0, st>>>(n, M->leaf_up.as<int>(), M->node_up.as<int>(), M->node_range.as<int2>(),
                                     M->horizon_zone_opt, M->zone_node.as<int>(), M->zone_up.as<int>());
    }
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaEventRecord(M->ev[1], st));
    int h[2];
    unsigned sc[8];
    FB_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaMemcpyAsync(sc, M->scene.p, sizeof(sc), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    M->ntop = h[0];
    M->max_depth = h[1];
    FB_REQUIRE(M->ntop <= std::max(M->top_nodes_opt, 1), "internal: top-node budget exceeded");
    FB_REQUIRE(M->max_depth < kStackDepth - 2,
               "LBVH deeper than the traversal stack (too many coincident face centroids)");
    for (int k = 0; k < 7; ++k) {
        const unsigned u = sc[k];
        const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        memcpy(&M->scene_h[k], &v, 4);
    }
    FB_CUDA(cudaEventElapsedTime(&M->ms_build, M->ev[0], M->ev[1]));
}

// traversal-stack overflow cannot happen (depth is checked at build time); the
// device flag is a second line of defence, read after every query
void check_error_flag(fluxb200_mesh *M) {
    int flag = 0;
    FB_CUDA(cudaMemcpyAsync(&flag, M->scalars.as<int>() + 6, sizeof(int), cudaMemcpyDeviceToHost, M->stream));
    FB_CUDA(cudaStreamSynchronize(M->stream));
    FB_REQUIRE(flag == 0, "internal: BVH traversal stack overflow");
}

template <class T> void set_face_data(fluxb200_mesh *M, const void *P, const void *N, const void *A) {
    const size_t nf = M->nf;
    if (!nf || (!P && !N && !A)) return;
    cudaStream_t st = M->stream;
    M->qtmp.reserve(sizeof(T) * 7 * nf);
    T *dP = M->qtmp.as<T>(), *dN = dP + 3 * nf, *dA = dN + 3 * nf;
    if (P) FB_CUDA(cudaMemcpyAsync(dP, P, sizeof(T) * 3 * nf, cudaMemcpyHostToDevice, st));
    if (N) FB_CUDA(cudaMemcpyAsync(dN, N, sizeof(T) * 3 * nf, cudaMemcpyHostToDevice, st));
    if (A) FB_CUDA(cudaMemcpyAsync(dA, A, sizeof(T) * nf, cudaMemcpyHostToDevice, st));
    int *changed = M->scalars.as<int>() + 7; // scalars: [0] ntop, [1] depth, [6] traversal error flag, [7] this
    FB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
    pack_face_kernel<T><<<blocks_for((int64_t)nf, 256), 256, 0, st>>>(
        P ? dP : nullptr, N ? dN : nullptr, A ? dA : nullptr, (int)nf, M->faceP.as<Real4<T>>(),
        M->faceN.as<Real4<T>>(), changed);
    FB_CUDA(cudaGetLastError());
    int h_changed = 1; // derived data (horizons, prepared column sets) is redone only when P or N really changed
    FB_CUDA(cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    if (h_changed & 1) M->hz_dirty = true;
    if (h_changed) ++M->face_epoch; // (the areas travel in the gathered colP.w: a new A invalidates the copies too)
}

template <class T> void get_face_data(fluxb200_mesh *M, void *P, void *N, void *A) {
    const size_t nf = M->nf;
    if (!nf || (!P && !N && !A)) return;
    cudaStream_t st = M->stream;
    M->qtmp.reserve(sizeof(T) * 7 * nf);
    T *dP = M->qtmp.as<T>(), *dN = dP + 3 * nf, *dA = dN + 3 * nf;
    unpack_face_kernel<T><<<blocks_for((int64_t)nf, 256), 256, 0, st>>>(
        M->faceP.as<Real4<T>>(), M->faceN.as<Real4<T>>(), (int)nf, dP, dN, dA);
    FB_CUDA(cudaGetLastError());
    if (P) FB_CUDA(cudaMemcpyAsync(P, dP, sizeof(T) * 3 * nf, cudaMemcpyDeviceToHost, st));
    if (N) FB_CUDA(cudaMemcpyAsync(N, dN, sizeof(T) * 3 * nf, cudaMemcpyDeviceToHost, st));
    if (A) FB_CUDA(cudaMemcpyAsync(A, dA, sizeof(T) * nf, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
}

ColSet *scratch_colset(fluxb200_mesh *M) {
    if (M->colsets.empty()) M->colsets.emplace_back(new ColSet());
    return M->colsets[0].get();
}

void upload_rows(fluxb200_mesh *M, const int64_t *I, size_t m) {
    FB_REQUIRE(m < (1ull << 31), "index sets must have fewer than 2^31 entries");
    const std::vector<int> rows = to_i32(I, m, M->nf, "I");
    M->rows.reserve(sizeof(int) * std::max<size_t>(m, 1));
    if (m) FB_CUDA(cudaMemcpyAsync(M->rows.p, rows.data(), sizeof(int) * m, cudaMemcpyHostToDevice, M->stream));
    FB_CUDA(cudaStreamSynchronize(M->stream)); // the host vector goes out of scope
}

void upload_cols(fluxb200_mesh *M, ColSet *C, const int64_t *J, size_t n) {
    FB_REQUIRE(n < (1ull << 31), "index sets must have fewer than 2^31 entries");
    const std::vector<int> cols = to_i32(J, n, M->nf, "J");
    C->cols.reserve(sizeof(int) * std::max<size_t>(n, 1));
    if (n) FB_CUDA(cudaMemcpyAsync(C->cols.p, cols.data(), sizeof(int) * n, cudaMemcpyHostToDevice, M->stream));
    FB_CUDA(cudaStreamSynchronize(M->stream));
}

// query kernels: both index sets, columns into the scratch set (never cached)
void upload_index_sets(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n) {
    ColSet *C = scratch_colset(M);
    C->valid = false;
    M->cs = C;
    upload_rows(M, I, m);
    upload_cols(M, C, J, n);
}

// The prepared column set for J: a cached one (same J, same face data, same tree) or the least recently used
// entry, to be refilled by the caller (returns false).
bool find_colset(fluxb200_mesh *M, const int64_t *J, size_t n) {
    scratch_colset(M);
    const size_t cap = (size_t)std::max(M->colset_cache_opt, 1);
    ColSet *lru = nullptr;
    for (size_t k = 1; k < M->colsets.size(); ++k) {
        ColSet *C = M->colsets[k].get();
        if (M->colset_cache_opt > 0 && C->valid && C->n == n && C->face_epoch == M->face_epoch &&
            C->tree_epoch == M->tree_epoch && C->arange == (J == nullptr) &&
            (J == nullptr || memcmp(C->key.data(), J, sizeof(int64_t) * n) == 0)) {
            C->used = ++M->use_clock;
            M->cs = C;
            ++M->colset_hits;
            return true;
        }
        if (!lru || C->used < lru->used) lru = C;
    }
    ++M->colset_misses;
    if (M->colsets.size() < cap + 1) {
        M->colsets.emplace_back(new ColSet());
        lru = M->colsets.back().get();
    }
    lru->valid = false;
    lru->n = n;
    lru->arange = J == nullptr;
    if (J) lru->key.assign(J, J + n);
    else lru->key.clear();
    lru->face_epoch = M->face_epoch;
    lru->tree_epoch = M->tree_epoch;
    lru->used = ++M->use_clock;
    M->cs = lru;
    return false;
}

// ---- per-call pieces shared by the two-phase and the streaming assembly ---------

// what a float32 ray can be off the ideal one, plus the slack of the Pluecker edge tests: 16 ulps of the
// largest coordinate (tools/k4_horizon_check.py)
inline float horizon_pert(const fluxb200_mesh *M) { return 16.0f * 1.1920929e-7f * M->scene_h[6]; }

// index sets -> device; columns sorted by leaf (Morton) position and gathered
template <class T> int prepare_call(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                                    double eps) {
    cudaStream_t st = M->stream;
    M->have_count = false;
    M->stats = fluxb200_ff_stats{};
    M->stats.pairs_all = (int64_t)m * (int64_t)n;
    M->m = m;
    M->n = n;
    M->eps = eps;
    M->nwords = (int)ceil_div((int64_t)n, 32);
    int launches = 0;
    upload_rows(M, I, m);
    M->stats.h2d_bytes = (int64_t)(sizeof(int) * m);
    M->tested.reserve(sizeof(unsigned long long) * 8); // [0] rays, [1] work-unit counter, [2..4] horizon-skip counters
    FB_CUDA(cudaMemsetAsync(M->tested.p, 0, sizeof(unsigned long long) * 8, st));
    const bool hor = M->horizon_skip_opt && M->ninternal > 0 && M->ntop == 0;
    if (hor && M->hz_dirty && m && n) { // per-face horizons from the shape model's CURRENT P, N (before the lookup:
        const int nf = (int)M->nf;      // a set cached under this face epoch must see these horizons)
        M->hz.reserve(sizeof(float2) * (size_t)nf);
        horizon_kernel<T><<<blocks_for((int64_t)nf * 32, 256), 256, 0, st>>>(
            nf, M->faceP.as<Real4<T>>(), M->faceN.as<Real4<T>>(), M->face_leaf.as<int>(),
            M->zone_node.as<int>(), M->node_range.as<int2>(), M->tri.as<float4>(), horizon_pert(M),
            M->hz.as<float2>());
        M->hz_dirty = false;
        ++launches;
    }
    if (find_colset(M, J, n)) return launches; // sorted, gathered and annotated by an earlier call: nothing to do
    ColSet *C = M->cs;
    upload_cols(M, C, J, n);
    M->stats.h2d_bytes += (int64_t)(sizeof(int) * n);
    if (m && n) {
        M->ckeys.reserve(sizeof(uint64_t) * n);
        M->cvals.reserve(sizeof(uint32_t) * n);
        C->colP.reserve(sizeof(Real4<T>) * n);
        C->colN.reserve(sizeof(Real4<T>) * n);
        C->col_face.reserve(sizeof(int) * n);
        C->col_leaf.reserve(sizeof(int) * n);
        C->rank_of_pos.reserve(sizeof(int) * n);
        const int B = 256;
        const int l0 = M->sorter.launches;
        col_keys_kernel<<<blocks_for((int64_t)n, B), B, 0, st>>>(C->cols.as<int>(), (int)n,
                                                                 M->face_leaf.as<int>(),
                                                                 M->ckeys.as<uint64_t>(),
                                                                 M->cvals.as<uint32_t>());
        int bits = 1;
        while ((1ull << bits) < M->nf) ++bits;
        M->sorter.sort(M->ckeys.as<uint64_t>(), M->cvals.as<uint32_t>(), (int)n, bits, st);
        col_gather_kernel<T><<<blocks_for((int64_t)n, B), B, 0, st>>>(
            M->cvals.as<uint32_t>(), C->cols.as<int>(), (int)n, M->face_leaf.as<int>(),
            M->faceP.as<Real4<T>>(), M->faceN.as<Real4<T>>(), C->colP.as<Real4<T>>(),
            C->colN.as<Real4<T>>(), C->col_face.as<int>(), C->col_leaf.as<int>(),
            C->rank_of_pos.as<int>());
        FB_CUDA(cudaGetLastError());
        launches += 2 + (M->sorter.launches - l0);
        { // per-chunk data shared by every row (trace2.cuh)
            const int nchunks = (int)ceil_div((int64_t)n, kChunkCols);
            C->chunk_info.reserve(sizeof(float4) * 3 * (size_t)nchunks);
            chunk_info_kernel<T><<<blocks_for((int64_t)nchunks * 32, B), B, 0, st>>>(
                C->colP.as<Real4<T>>(), C->col_leaf.as<int>(), (int)n, nchunks, M->leaf_up.as<int>(),
                M->node_up.as<int>(), M->node_range.as<int2>(), M->ninternal, C->chunk_info.as<float4>());
            FB_CUDA(cudaGetLastError());
            ++launches;
        }
        if (hor) {
            C->colH.reserve(sizeof(float4) * n);
            col_horizon_kernel<<<blocks_for((int64_t)n, B), B, 0, st>>>(
                C->col_face.as<int>(), C->col_leaf.as<int>(), (int)n, M->hz.as<float2>(), M->zone_node.as<int>(),
                M->zone_up.as<int>(), C->colH.as<float4>());
            FB_CUDA(cudaGetLastError());
            ++launches;
        }
        C->valid = true;
    }
    return launches;
}

// K4 for rows [row0, row0 + mr) of the uploaded index set
template <class T> void launch_trace(fluxb200_mesh *M, size_t row0, size_t mr, uint32_t *bits,
                                     uint32_t *row_counts, cudaStream_t st) {
    TraceArgs<T> A{};
    A.faceP = M->faceP.as<Real4<T>>();
    A.faceN = M->faceN.as<Real4<T>>();
    A.rows = M->rows.as<int>() + row0;
    A.face_leaf = M->face_leaf.as<int>();
    A.node_up = M->node_up.as<int>();
    A.leaf_up = M->leaf_up.as<int>();
    A.node_range = M->node_range.as<int2>();
    A.colP = M->cs->colP.as<Real4<T>>();
    A.colN = M->cs->colN.as<Real4<T>>();
    A.col_face = M->cs->col_face.as<int>();
    A.col_leaf = M->cs->col_leaf.as<int>();
    A.m = (int)mr;
    A.n = (int)M->n;
    A.nwords = M->nwords;
    A.eps = (T)M->eps;
    A.nodes = M->nodes.as<float4>();
    A.tri = M->tri.as<float4>();
    A.ninternal = M->ninternal;
    A.nfaces = (int)M->nf;
    A.ntop = M->ntop;
    A.bits = bits;
    A.row_counts = row_counts;
    A.tested = M->tested.as<unsigned long long>();
    A.error_flag = M->scalars.as<int>() + 6;
    A.scale = M->scene_h[6];
    A.shaft_filter = M->shaft_filter_opt;
    const bool hor = M->horizon_skip_opt && M->ninternal > 0 && M->ntop == 0;
    A.hz = hor ? M->hz.as<float2>() : nullptr;
    A.zone_node = hor ? M->zone_node.as<int>() : nullptr;
    A.colH = hor ? M->cs->colH.as<float4>() : nullptr;
    A.zone_leaves = M->horizon_zone_opt;
    A.pert = horizon_pert(M);
    A.chunk_info = M->cs->chunk_info.as<float4>();
    A.nchunks = (int)ceil_div((int64_t)M->n, kChunkCols);
    FB_REQUIRE((int64_t)mr * A.nchunks < (1ll << 32) - 65536, "too many work units for one launch");
    const size_t smem = sizeof(float4) * 6 * (size_t)M->ntop;
    FB_REQUIRE((int)smem + 16384 <= M->max_smem_optin, "top_nodes does not fit in shared memory");
    if (M->ntop > 0)
        FB_CUDA(cudaFuncSetAttribute(trace_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)std::max<size_t>(smem, 1024)));
    FB_CUDA(cudaMemsetAsync(M->tested.as<unsigned long long>() + 1, 0, sizeof(unsigned long long), st));
    const int64_t units = (int64_t)mr * A.nchunks;
    const int grid = (int)std::min<int64_t>((int64_t)M->num_sms * M->blocks_per_sm,
                                            std::max<int64_t>(1, ceil_div(units, kTraceWarps)));
    // second generation: needs the source path (depth + 1 records) and some of the target side in its list,
    // and node ids in 26 bits; anything else takes the first-generation kernel
    const bool gen2 = M->trace_variant_opt == 2 && M->ntop == 0 && M->max_depth + 1 + 8 <= kPathCap &&
                      M->ninternal < (1 << kNodeBits);
    if (gen2) {
        // rays the kernel cannot finish in its bounded shared-memory stacks go to a list and are traced by
        // resolve_lost_kernel right after it (typically 0.05 % of the rays: those grazing very many triangles)
        const unsigned lost_cap = (unsigned)std::min<int64_t>(1 << 22, std::max<int64_t>(1024, (int64_t)mr * 256));
        M->lost.reserve(16 + sizeof(int2) * (size_t)lost_cap);
        A.lost_count = M->lost.as<unsigned>();
        A.lost = reinterpret_cast<int2 *>(M->lost.as<unsigned char>() + 16);
        A.lost_cap = lost_cap;
        FB_CUDA(cudaMemsetAsync(A.lost_count, 0, sizeof(unsigned), st));
        FB_CUDA(cudaFuncSetAttribute(trace2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)trace2_smem_bytes()));
        // shared memory for exactly the resident CTAs; the rest of the 256 KB stays L1 (BVH records, triangles)
        const int carve = (int)std::min<size_t>(100, ((size_t)M->blocks_per_sm * (trace2_smem_bytes() + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        FB_CUDA(cudaFuncSetAttribute(trace2_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        trace2_kernel<T><<<grid, kTraceThreads, trace2_smem_bytes(), st>>>(A);
        resolve_lost_kernel<T><<<M->num_sms * 4, 128, 0, st>>>(A);
    } else if (M->ntop > 0) trace_kernel<T, true><<<grid, kTraceThreads, smem, st>>>(A);
    else if (hor) trace_kernel<T, false, true><<<grid, kTraceThreads, 0, st>>>(A);
    else trace_kernel<T, false><<<grid, kTraceThreads, 0, st>>>(A);
    FB_CUDA(cudaGetLastError());
}

// K6 for rows [row0, row0 + mr): local indptr (mr + 1) -> data / indices at out_base.
// K6a (un-permute R rows per CTA into `jbits`, J-order words) then K6b (emit).  `jbits`
// is a scratch of the stream the call is enqueued on, or the caller's own buffer when it
// wants the J-order words as a result (indices == NULL: the host expands them).
template <int R, bool kSmem>
void launch_unpermute(const uint32_t *bits, const int *rank_of_pos, int mr, int n, int nwords, uint32_t *jbits,
                      uint32_t *gcount, size_t smem, cudaStream_t st) {
    FB_CUDA(cudaFuncSetAttribute(unpermute_kernel<R, kSmem>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)std::max<size_t>(smem, 1024)));
    unpermute_kernel<R, kSmem><<<(unsigned)ceil_div(mr, R), kFillThreads, smem, st>>>(bits, rank_of_pos, mr, n,
                                                                                       nwords, jbits, gcount);
}

// words of J-order scratch launch_fill needs for mr rows: the words themselves + the per-group counts
inline size_t fill_scratch_words(size_t mr, int nwords) {
    return mr * (size_t)std::max(nwords, 1) + mr * (size_t)ceil_div(std::max(nwords, 1), 32);
}

template <class T> void launch_fill(fluxb200_mesh *M, size_t row0, size_t mr, const uint32_t *bits,
                                    const int64_t *indptr_local, int64_t out_base, T *data, void *indices,
                                    int index_width, uint32_t *jbits, cudaStream_t st) {
    const int nwords = M->nwords, n = (int)M->n;
    const size_t row_bytes = sizeof(uint32_t) * (size_t)nwords;
    uint32_t *gcount = jbits + mr * (size_t)nwords;
    // rows per CTA: four CTAs per SM when >= 2 rows fit in a quarter of the shared memory, else fewer CTAs
    const size_t budget_q = (size_t)M->max_smem_optin / 4 - 1024, budget_full = (size_t)M->max_smem_optin - 1024;
    size_t R = budget_q / row_bytes;
    if (R < 2) R = (budget_full / 2) / row_bytes;
    if (R < 2) R = budget_full / row_bytes;
    if (M->fill_rows_opt > 0) R = std::min<size_t>(budget_full / row_bytes, (size_t)M->fill_rows_opt); // A/B and tests
    if (M->fill_rows_opt < 0) R = 0;
    const int *rank = M->cs->rank_of_pos.as<int>();
    if (R >= 8) launch_unpermute<8, true>(bits, rank, (int)mr, n, nwords, jbits, gcount, 8 * row_bytes, st);
    else if (R >= 4) launch_unpermute<4, true>(bits, rank, (int)mr, n, nwords, jbits, gcount, 4 * row_bytes, st);
    else if (R >= 2) launch_unpermute<2, true>(bits, rank, (int)mr, n, nwords, jbits, gcount, 2 * row_bytes, st);
    else if (R == 1) launch_unpermute<1, true>(bits, rank, (int)mr, n, nwords, jbits, gcount, row_bytes, st);
    else launch_unpermute<1, false>(bits, rank, (int)mr, n, nwords, jbits, gcount, 0, st);
    FB_CUDA(cudaGetLastError());
    FillArgs<T> A;
    A.faceP = M->faceP.as<Real4<T>>();
    A.faceN = M->faceN.as<Real4<T>>();
    A.rows = M->rows.as<int>() + row0;
    A.cols = M->cs->cols.as<int>();
    A.m = (int)mr;
    A.n = (int)M->n;
    A.nwords = nwords;
    A.jbits = jbits;
    A.gcount = gcount;
    A.indptr = indptr_local;
    A.out_base = out_base;
    A.data = data;
    A.indices = indices;
    A.index_width = index_width;
    // enough CTAs (8 rows x one column segment each) for about four waves of four resident CTAs per SM
    const int64_t row_blocks = ceil_div((int64_t)mr, kFillWarps), ngroups = ceil_div(nwords, 32);
    const int64_t segs = std::max<int64_t>(1, std::min<int64_t>(ngroups, ceil_div(16 * (int64_t)M->num_sms, row_blocks)));
    FB_CUDA(cudaFuncSetAttribute(emit_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)emit_smem_bytes<T>()));
    emit_kernel<T><<<dim3((unsigned)row_blocks, (unsigned)segs), kFillThreads, emit_smem_bytes<T>(), st>>>(A);
    FB_CUDA(cudaGetLastError());
}
constexpr int kFillLaunches = 2; // kernels per launch_fill

// ---- two-phase API: count (all rows, bits kept on the device), then fill ---------
template <class T> void ff_count(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                                 double eps, int64_t *row_counts) {
    cudaStream_t st = M->stream;
    FB_CUDA(cudaEventRecord(M->ev[0], st));
    int launches = prepare_call<T>(M, I, m, J, n, eps);
    M->row_counts.reserve(sizeof(uint32_t) * std::max<size_t>(m, 1));
    M->counts64.reserve(sizeof(int64_t) * std::max<size_t>(m, 1));
    M->indptr.reserve(sizeof(int64_t) * (m + 1));
    FB_CUDA(cudaMemsetAsync(M->row_counts.p, 0, sizeof(uint32_t) * std::max<size_t>(m, 1), st));
    FB_CUDA(cudaMemsetAsync(M->indptr.p, 0, sizeof(int64_t) * (m + 1), st));
    if (m && n) M->bits.reserve(sizeof(uint32_t) * m * (size_t)M->nwords);
    FB_CUDA(cudaEventRecord(M->ev[1], st));
    if (m && n) {
        launch_trace<T>(M, 0, m, M->bits.as<uint32_t>(), M->row_counts.as<uint32_t>(), st);
        launches += 1;
        M->stats.trace_launches = 1;
    }
    FB_CUDA(cudaEventRecord(M->ev[2], st));
    if (m) {
        counts_to_i64_kernel<<<blocks_for((int64_t)m, 256), 256, 0, st>>>(M->row_counts.as<uint32_t>(),
                                                                         (int)m, M->counts64.as<int64_t>());
        scan_exclusive<int64_t, int64_t>(M->counts64.as<int64_t>(), M->indptr.as<int64_t>(), (int64_t)m,
                                         M->indptr.as<int64_t>() + m, st);
        launches += 2;
    }
    FB_CUDA(cudaEventRecord(M->ev[3], st));
    int64_t nnz = 0;
    unsigned long long tested = 0;
    FB_CUDA(cudaMemcpyAsync(&nnz, M->indptr.as<int64_t>() + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaMemcpyAsync(&tested, M->tested.p, sizeof(tested), cudaMemcpyDeviceToHost, st));
    if (row_counts && m)
        FB_CUDA(cudaMemcpyAsync(row_counts, M->counts64.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    M->nnz = nnz;
    M->stats.nnz = nnz;
    M->stats.pairs_tested = (int64_t)tested;
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_prepare, M->ev[0], M->ev[1]));
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_trace, M->ev[1], M->ev[2]));
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_scan, M->ev[2], M->ev[3]));
    M->stats.kernel_launches = launches;
    M->stats.d2h_bytes = (int64_t)(sizeof(int64_t) + sizeof(tested) + (row_counts ? sizeof(int64_t) * m : 0));
    check_error_flag(M);
    M->have_count = true;
}

template <class T> void ff_fill(fluxb200_mesh *M, int index_width, int destination, void *indptr,
                                void *indices, void *data) {
    FB_REQUIRE(M->have_count, "fluxb200_ff_fill: no preceding fluxb200_ff_count on this handle");
    FB_REQUIRE(index_width == 4 || index_width == 8, "index_width must be 4 or 8");
    FB_REQUIRE(destination >= 0 && destination <= 2, "destination must be 0, 1 or 2");
    // int32 must hold the column positions always, and the entry offsets only where indptr is emitted in the
    // index dtype (host / caller buffers); the library's own device CSR keeps an int64 indptr
    if (index_width == 4)
        FB_REQUIRE(M->n < (1ull << 31) && (destination == 2 || M->nnz < (1ll << 31)), "int32 indices cannot hold this matrix");
    cudaStream_t st = M->stream;
    const size_t m = M->m, n = M->n;
    const int64_t nnz = M->nnz;
    T *d_data;
    void *d_indices;
    if (destination == 1) {
        d_data = reinterpret_cast<T *>(data);
        d_indices = indices;
    } else {
        M->out_data.reserve(sizeof(T) * std::max<int64_t>(nnz, 1));
        M->out_indices.reserve((size_t)index_width * std::max<int64_t>(nnz, 1));
        d_data = M->out_data.as<T>();
        d_indices = M->out_indices.p;
    }
    FB_CUDA(cudaEventRecord(M->ev[0], st));
    int launches = 0;
    if (m && n && nnz) {
        // pieces of <= 4096 rows bound the J-order scratch (the pieces run back to back on one stream)
        const size_t piece = std::min<size_t>(m, 4096);
        M->jbits.reserve(sizeof(uint32_t) * fill_scratch_words(piece, M->nwords));
        for (size_t r0 = 0; r0 < m; r0 += piece) {
            const size_t mr = std::min(piece, m - r0);
            launch_fill<T>(M, r0, mr, M->bits.as<uint32_t>() + r0 * (size_t)M->nwords, M->indptr.as<int64_t>() + r0,
                           0, d_data, d_indices, index_width, M->jbits.as<uint32_t>(), st);
            launches += kFillLaunches;
        }
    }
    void *d_indptr = M->indptr.p;
    if (index_width == 4) {
        M->indptr32.reserve(sizeof(int32_t) * (m + 1));
        indptr_to_i32_kernel<<<blocks_for((int64_t)m + 1, 256), 256, 0, st>>>(
            M->indptr.as<int64_t>(), (int)(m + 1), M->indptr32.as<int32_t>());
        d_indptr = M->indptr32.p;
        ++launches;
    }
    FB_CUDA(cudaEventRecord(M->ev[1], st));
    if (destination == 0) {
        FB_CUDA(cudaMemcpyAsync(indptr, d_indptr, (size_t)index_width * (m + 1), cudaMemcpyDeviceToHost, st));
        if (nnz) {
            FB_CUDA(cudaMemcpyAsync(indices, d_indices, (size_t)index_width * nnz, cudaMemcpyDeviceToHost, st));
            FB_CUDA(cudaMemcpyAsync(data, d_data, sizeof(T) * nnz, cudaMemcpyDeviceToHost, st));
        }
    } else if (destination == 1) {
        FB_CUDA(cudaMemcpyAsync(indptr, d_indptr, (size_t)index_width * (m + 1), cudaMemcpyDeviceToDevice, st));
    }
    FB_CUDA(cudaEventRecord(M->ev[2], st));
    FB_CUDA(cudaStreamSynchronize(st));
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_fill, M->ev[0], M->ev[1]));
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_d2h, M->ev[1], M->ev[2]));
    M->stats.kernel_launches += launches;
    if (destination == 0)
        M->stats.d2h_bytes += (int64_t)((size_t)index_width * (m + 1) + ((size_t)index_width + sizeof(T)) * (size_t)nnz);
    M->out_index_width = destination == 2 ? index_width : 0;
}

// ---- streaming assembly: row sub-slabs pipelined over three streams ----------------
// compute stream: trace(k) -> counts(k) -> local indptr(k) -> nnz(k), counts(k) to pinned host
// fill stream   : fill(k)             (submitted before trace(k+1), see below)
// copy-out      : D2H(k)              (runs under trace(k+1); bits / staging triple-buffered)
// Returns false when `capacity` entries do not suffice (stats.nnz = entries needed).
template <class T> bool ff_assemble(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                                    double eps, int index_width, int destination, void *indptr,
                                    void *indices, void *data, int64_t capacity, int64_t *row_counts) {
    FB_REQUIRE(index_width == 4 || index_width == 8, "index_width must be 4 or 8");
    FB_REQUIRE(destination == 0 || destination == 2 || destination == 3,
               "destination must be 0 (page-locked host buffers), 2 (library device buffers) or 3 (ordinary host buffers)");
    // 3 = host output into ORDINARY (pageable) memory: page-locking a multi-gigabyte result costs more than
    // assembling it (about 0.5 s per GB), so the values go through page-locked staging slots and host threads
    // move them on; the column indices are written by host threads anyway.  Otherwise identical to 0.
    const bool staged = destination == 3;
    if (staged) destination = 0;
    if (index_width == 4) FB_REQUIRE(n < (1ull << 31), "int32 indices cannot hold this many columns");
    cudaStream_t s0 = M->stream, s1 = M->copy_stream, s2 = M->d2h_stream;
    FB_CUDA(cudaEventRecord(M->ev[0], s0));
    int launches = prepare_call<T>(M, I, m, J, n, eps);
    FB_CUDA(cudaEventRecord(M->ev[1], s0));
    // host output: small sub-slabs so that copy-out overlaps tracing; device-resident output has
    // nothing to overlap, so it uses pieces 8x larger (fewer host round trips, bounded bit buffers)
    const size_t sub_want = destination == 0 ? (size_t)M->sub_rows_opt : (size_t)M->sub_rows_opt * 8;
    const size_t sub = std::max<size_t>(1, std::min<size_t>(sub_want, std::max<size_t>(m, 1)));
    // sub-slab boundaries: a short first piece and a doubling ramp (the copy-out, which on a host with ~85 GB/s
    // of memory write bandwidth is the longest of the three pipelines, starts 2 ms into the call instead of 9:
    // profiles/r02k_timeline.md), full-size pieces, then a geometrically shrinking tail so that the last
    // fill + copy-out (which no tracing overlaps) is short
    std::vector<size_t> bounds{0};
    {
        size_t pos = 0;
        const size_t min_piece = std::min(sub, std::max<size_t>(32, sub / 4)); // never above the slot size
        if (destination == 0 && M->ramp_opt)
            for (size_t piece = min_piece; piece < sub && m - pos >= 4 * sub; piece *= 2) {
                pos += piece;
                bounds.push_back(pos);
            }
        while (pos < m) {
            const size_t rem = m - pos;
            size_t piece = sub;
            if (destination == 0 && rem <= 2 * sub) piece = std::max(min_piece, (rem + 1) / 2);
            piece = std::min(std::min(piece, sub), rem);
            if (rem - piece < min_piece && rem <= sub) piece = rem; // no crumb at the end (a 32-row launch costs 0.9 ms)
            pos += piece;
            bounds.push_back(pos);
        }
    }
    const size_t nsub = bounds.size() - 1;
    // per-slot (double-buffered) device state
    for (int b = 0; b < fluxb200_mesh::kSlots; ++b) {
        M->sbits[b].reserve(sizeof(uint32_t) * sub * (size_t)std::max(M->nwords, 1));
        M->scounts[b].reserve(sizeof(uint32_t) * sub);
        M->scounts64[b].reserve(sizeof(int64_t) * sub);
        M->sindptr[b].reserve(sizeof(int64_t) * (sub + 1));
        M->sjbits[b].reserve(sizeof(uint32_t) * fill_scratch_words(sub, M->nwords));
    }
    // host output: the column indices travel as J-order visibility words (n/8 bytes per row instead
    // of 4 or 8 bytes per entry) and host threads expand them while the next sub-slabs are traced
    const bool expand = destination == 0 && (M->host_expand_opt != 0 || staged) && n > 0;
    const size_t slot_words = sub * (size_t)std::max(M->nwords, 1);
    std::vector<std::unique_ptr<ExpandTask>> tasks(expand ? nsub : 0);
    struct TaskGuard { // no worker may outlive the buffers it writes to, whatever path leaves this frame
        fluxb200_mesh *M;
        std::vector<std::unique_ptr<ExpandTask>> &tasks;
        ~TaskGuard() {
            cudaStreamSynchronize(M->copy_stream);
            cudaStreamSynchronize(M->d2h_stream); // every callback that will ever fire has fired
            for (auto &t : tasks)
                if (t && t->queued.load()) M->expander.wait(t.get());
        }
    } task_guard{M, tasks};
    if (expand) {
        int nthreads = M->host_threads_opt;
        if (nthreads <= 0) { // share the host cores with the other ranks of a torchrun job
            const char *lws = getenv("LOCAL_WORLD_SIZE");
            const int ranks = std::max(1, lws ? atoi(lws) : 1);
            const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
            nthreads = std::min(8, std::max(1, hw / ranks - 1)); // one core per rank stays with the submitting thread
        }
        M->expander.start(nthreads);
        M->h_jbits.reserve(sizeof(uint32_t) * slot_words * fluxb200_mesh::kHostSlots);
    }
    M->h_nnz.reserve(sizeof(int64_t) * std::max<size_t>(nsub, 1));
    M->h_counts.reserve(sizeof(uint32_t) * std::max<size_t>(m, 1));
    while (M->sub_events.size() < 3 * nsub) {
        cudaEvent_t e;
        FB_CUDA(cudaEventCreate(&e));
        M->sub_events.push_back(e);
    }
    // FLUXB200_TIMELINE=<file>: per-sub-slab device timestamps of trace / fill / copy-out (diagnostic)
    const char *tl_path = getenv("FLUXB200_TIMELINE");
    std::vector<double> tl_host(tl_path ? 3 * nsub : 0);
    const auto tl_clock0 = std::chrono::steady_clock::now();
    auto tl_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl_clock0).count(); };
    if (tl_path)
        while (M->tl_events.size() < 4 * nsub) {
            cudaEvent_t e;
            FB_CUDA(cudaEventCreate(&e));
            M->tl_events.push_back(e);
        }
    if (destination == 2) {
        if (capacity <= 0)
            // first call: 0.70 of the dense size (+ 12.5 % in the buffer) holds every slab of the bench meshes with
            // the 25 % headroom below, so the buffers are allocated once -- growing them in a later call is a
            // cudaFree + cudaMalloc of gigabytes, 5-45 ms in the middle of somebody's loop (r02n); 0.62 for
            // results above 16 GB (the resident half of a 200k-face matrix on 2 GPUs must still fit)
            capacity = M->dev_capacity_hint > 0
                           ? M->dev_capacity_hint
                           : std::max<int64_t>(1024, (int64_t)(((double)m * (double)n * 8.0 > 16e9 ? 0.62 : 0.70) *
                                                               (double)m * (double)n));
        capacity = std::min<int64_t>(capacity, std::max<int64_t>((int64_t)m * (int64_t)n, 1));
        M->out_data.reserve(sizeof(T) * (size_t)capacity);
        M->out_indices.reserve((size_t)index_width * (size_t)capacity);
    }
    int64_t *h_nnz = M->h_nnz.as<int64_t>(), *h_nnz_dev = nullptr;
    uint32_t *h_counts = M->h_counts.as<uint32_t>(), *h_counts_dev = nullptr;
    FB_CUDA(cudaHostGetDevicePointer((void **)&h_nnz_dev, h_nnz, 0));
    FB_CUDA(cudaHostGetDevicePointer((void **)&h_counts_dev, h_counts, 0));
    int64_t total = 0, d2h_bytes = 0;
    bool overflow = false;
    float ms_fill = 0.f;

    auto enqueue_trace = [&](size_t k) {
        const int b = (int)(k % fluxb200_mesh::kSlots);
        const size_t row0 = bounds[k], mr = bounds[k + 1] - row0;
        if (k >= (size_t)fluxb200_mesh::kSlots)
            FB_CUDA(cudaStreamWaitEvent(s0, M->slot_free[b], 0)); // fill(k - kSlots) is done with slot b
        FB_CUDA(cudaMemsetAsync(M->scounts[b].p, 0, sizeof(uint32_t) * mr, s0));
        if (tl_path) tl_host[3 * k] = tl_now();
        FB_CUDA(cudaEventRecord(M->sub_events[3 * k], s0));
        if (n) launch_trace<T>(M, row0, mr, M->sbits[b].as<uint32_t>(), M->scounts[b].as<uint32_t>(), s0);
        FB_CUDA(cudaEventRecord(M->sub_events[3 * k + 1], s0));
        counts_to_i64_kernel<<<blocks_for((int64_t)mr, 256), 256, 0, s0>>>(M->scounts[b].as<uint32_t>(), (int)mr,
                                                                          M->scounts64[b].as<int64_t>());
        scan_exclusive<int64_t, int64_t>(M->scounts64[b].as<int64_t>(), M->sindptr[b].as<int64_t>(), (int64_t)mr,
                                         M->sindptr[b].as<int64_t>() + mr, s0);
        publish_counts_kernel<<<blocks_for((int64_t)mr, 256), 256, 0, s0>>>(
            M->scounts[b].as<uint32_t>(), (int)mr, M->sindptr[b].as<int64_t>() + mr, h_counts_dev + row0, h_nnz_dev + k);
        FB_CUDA(cudaGetLastError());
        FB_CUDA(cudaEventRecord(M->sub_events[3 * k + 2], s0));
        d2h_bytes += (int64_t)(sizeof(int64_t) + sizeof(uint32_t) * mr);
        launches += (n ? 1 : 0) + 3;
    };
    auto finish = [&](size_t k) {
        const int b = (int)(k % fluxb200_mesh::kSlots);
        const size_t row0 = bounds[k], mr = bounds[k + 1] - row0;
        FB_CUDA(cudaEventSynchronize(M->sub_events[3 * k + 2]));
        if (tl_path) tl_host[3 * k + 1] = tl_now();
        const int64_t nnz_k = h_nnz[k], off = total;
        total += nnz_k;
        // host output in int32: the offsets must fit too.  Reported as "capacity too small" with the entry count,
        // so that the caller retries with int64 (the device-resident CSR keeps an int64 indptr: no limit there)
        if (destination == 0 && index_width == 4 && total >= (1ll << 31)) overflow = true;
        if (total > capacity) overflow = true;
        FB_CUDA(cudaStreamWaitEvent(s1, M->sub_events[3 * k + 2], 0));
        bool slot_recorded = false;
        if (!overflow && nnz_k) {
            T *d_data;
            void *d_idx;
            int64_t base;
            if (destination == 2) {
                d_data = M->out_data.as<T>();
                d_idx = M->out_indices.p;
                base = off;
            } else {
                M->stage_data[b].reserve(sizeof(T) * (size_t)nnz_k);
                if (!expand) M->stage_idx[b].reserve((size_t)index_width * (size_t)nnz_k);
                d_data = M->stage_data[b].as<T>();
                d_idx = M->stage_idx[b].p;
                base = 0;
            }
            if (destination == 0 && k >= (size_t)fluxb200_mesh::kSlots)
                FB_CUDA(cudaStreamWaitEvent(s1, M->d2h_done[b], 0)); // the slot's staging has left the device
            if (tl_path) FB_CUDA(cudaEventRecord(M->tl_events[4 * k], s1));
            launch_fill<T>(M, row0, mr, M->sbits[b].as<uint32_t>(), M->sindptr[b].as<int64_t>(), base, d_data,
                           expand ? nullptr : d_idx, index_width, M->sjbits[b].as<uint32_t>(), s1);
            if (tl_path) FB_CUDA(cudaEventRecord(M->tl_events[4 * k + 1], s1));
            launches += kFillLaunches;
            FB_CUDA(cudaEventRecord(M->slot_free[b], s1)); // bits / indptr of the slot are consumed
            slot_recorded = true;
            // Host output: trace(k+1) starts when fill(k) has finished.  Left to itself the persistent trace
            // kernel moves in behind the fill's FIRST kernel and the emit kernel only gets its SMs back when
            // that trace retires, so every copy-out started a whole sub-slab late (r02k timeline).
            if (destination == 0 && M->ramp_opt) FB_CUDA(cudaStreamWaitEvent(s0, M->slot_free[b], 0));
            if (destination == 0) {
                FB_CUDA(cudaStreamWaitEvent(s2, M->slot_free[b], 0));
                if (tl_path) FB_CUDA(cudaEventRecord(M->tl_events[4 * k + 2], s2));
                char *h_stage = nullptr;
                if (staged) { // (the slot is free: its last user, sub-slab k - kHostSlots, is waited for just below)
                    if (k >= (size_t)fluxb200_mesh::kHostSlots && tasks[k - fluxb200_mesh::kHostSlots])
                        M->expander.wait(tasks[k - fluxb200_mesh::kHostSlots].get());
                    HostBuf &hb = M->h_data[k % fluxb200_mesh::kHostSlots];
                    hb.reserve(sizeof(T) * (size_t)nnz_k);
                    h_stage = hb.as<char>();
                }
                FB_CUDA(cudaMemcpyAsync(staged ? (void *)h_stage : (void *)((char *)data + sizeof(T) * (size_t)off), d_data,
                                        sizeof(T) * (size_t)nnz_k, cudaMemcpyDeviceToHost, s2));
                d2h_bytes += (int64_t)(sizeof(T) * (size_t)nnz_k);
                d2h_bytes += expand ? (int64_t)(sizeof(uint32_t) * mr * (size_t)M->nwords)
                                    : (int64_t)((size_t)index_width * (size_t)nnz_k);
                if (expand) {
                    // the host slot is free once the sub-slab that used it last has been expanded
                    if (k >= (size_t)fluxb200_mesh::kHostSlots && tasks[k - fluxb200_mesh::kHostSlots])
                        M->expander.wait(tasks[k - fluxb200_mesh::kHostSlots].get());
                    uint32_t *h_words = M->h_jbits.as<uint32_t>() + (k % fluxb200_mesh::kHostSlots) * slot_words;
                    FB_CUDA(cudaMemcpyAsync(h_words, M->sjbits[b].p, sizeof(uint32_t) * mr * (size_t)M->nwords,
                                            cudaMemcpyDeviceToHost, s2));
                    tasks[k].reset(new ExpandTask());
                    ExpandTask *t = tasks[k].get();
                    t->words = h_words;
                    t->nwords = M->nwords;
                    t->mr = mr;
                    t->offs.resize(mr + 1);
                    t->offs[0] = off;
                    for (size_t r = 0; r < mr; ++r) t->offs[r + 1] = t->offs[r] + (int64_t)h_counts[row0 + r];
                    t->indices = indices;
                    t->index_width = index_width;
                    if (staged) {
                        t->copy_src = h_stage;
                        t->copy_dst = (char *)data + sizeof(T) * (size_t)off;
                        t->copy_bytes = sizeof(T) * (size_t)nnz_k;
                    }
                    t->pieces = (int)std::max<size_t>(1, std::min<size_t>(mr, 2 * (size_t)M->expander.threads()));
                    t->pending.store(t->pieces);
                    t->owner = &M->expander;
                    FB_CUDA(cudaLaunchHostFunc(s2, [](void *arg) {
                        ExpandTask *t = reinterpret_cast<ExpandTask *>(arg);
                        t->queued.store(1);
                        t->owner->submit(t);
                    }, t));
                } else {
                    FB_CUDA(cudaMemcpyAsync((char *)indices + (size_t)index_width * (size_t)off, d_idx,
                                            (size_t)index_width * (size_t)nnz_k, cudaMemcpyDeviceToHost, s2));
                }
                if (tl_path) FB_CUDA(cudaEventRecord(M->tl_events[4 * k + 3], s2));
                FB_CUDA(cudaEventRecord(M->d2h_done[b], s2));
            }
        }
        if (tl_path) tl_host[3 * k + 2] = tl_now();
        if (!slot_recorded) {
            FB_CUDA(cudaEventRecord(M->slot_free[b], s1));
            if (destination == 0) { // keep d2h_done[b] a valid "slot is free" marker for sub-slab k + kSlots
                FB_CUDA(cudaStreamWaitEvent(s2, M->slot_free[b], 0));
                FB_CUDA(cudaEventRecord(M->d2h_done[b], s2));
            }
        }
    };

    FB_CUDA(cudaEventRecord(M->ev[4], s1));
    // The persistent trace kernel owns every CTA slot while it runs, so fill(k) is put on the
    // (higher-priority) copy stream BEFORE trace(k+1) is submitted: the fill takes the idle
    // machine first, trace(k+1) moves in as its CTAs retire, D2H(k) runs under trace(k+1).
    // Cost: one host round trip (tens of microseconds) of idle GPU per sub-slab.
    for (size_t k = 0; k < nsub; ++k) {
        enqueue_trace(k);
        finish(k);
    }
    FB_CUDA(cudaEventRecord(M->slot_free[0], s2)); // (reused as a plain marker) copy-out done -> span end on s1
    FB_CUDA(cudaStreamWaitEvent(s1, M->slot_free[0], 0));
    FB_CUDA(cudaEventRecord(M->ev[5], s1));
    FB_CUDA(cudaEventRecord(M->ev[2], s0));
    unsigned long long tested = 0;
    FB_CUDA(cudaMemcpyAsync(&tested, M->tested.p, sizeof(tested), cudaMemcpyDeviceToHost, s0));
    FB_CUDA(cudaStreamSynchronize(s0));
    FB_CUDA(cudaStreamSynchronize(s1));
    FB_CUDA(cudaStreamSynchronize(s2));
    for (auto &t : tasks) {
        if (!t) continue;
        FB_REQUIRE(t->queued.load(), "internal: a copy-out callback did not run");
        M->expander.wait(t.get());
        FB_REQUIRE(t->mismatch.load() == 0, "internal: visibility words and row counts disagree");
    }
    if (tl_path && destination == 0 && !overflow) {
        if (FILE *f = fopen(tl_path, "w")) {
            fprintf(f, "k,rows,host_submit_trace,host_trace_done,host_fill_submitted,trace0,trace1,scan1,fill0,fill1,d2h0,d2h1\n");
            auto rel = [&](cudaEvent_t e) {
                float ms = -1.f;
                if (cudaEventElapsedTime(&ms, M->ev[0], e) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
                return ms;
            };
            for (size_t k = 0; k < nsub; ++k)
                fprintf(f, "%zu,%zu,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f\n", k, bounds[k + 1] - bounds[k],
                        tl_host[3 * k], tl_host[3 * k + 1], tl_host[3 * k + 2], rel(M->sub_events[3 * k]),
                        rel(M->sub_events[3 * k + 1]), rel(M->sub_events[3 * k + 2]), rel(M->tl_events[4 * k]),
                        rel(M->tl_events[4 * k + 1]), rel(M->tl_events[4 * k + 2]), rel(M->tl_events[4 * k + 3]));
            fprintf(f, "# host total %.3f ms\n", tl_now());
            fclose(f);
        }
    }
    // leave the handle's main stream ordered after the copy stream
    M->nnz = total;
    M->stats.nnz = total;
    M->stats.pairs_tested = (int64_t)tested;
    // (row slabs of one mesh differ by +-15 % in their entry counts; a short buffer costs a whole re-trace)
    M->dev_capacity_hint = std::max<int64_t>(M->dev_capacity_hint, total + total / 4);
    float ms = 0.f;
    M->stats.ms_trace = 0.f;
    for (size_t k = 0; k < nsub; ++k) {
        FB_CUDA(cudaEventElapsedTime(&ms, M->sub_events[3 * k], M->sub_events[3 * k + 1]));
        M->stats.ms_trace += ms;
        FB_CUDA(cudaEventElapsedTime(&ms, M->sub_events[3 * k + 1], M->sub_events[3 * k + 2]));
        M->stats.ms_scan += ms;
    }
    FB_CUDA(cudaEventElapsedTime(&M->stats.ms_prepare, M->ev[0], M->ev[1]));
    FB_CUDA(cudaEventElapsedTime(&ms_fill, M->ev[4], M->ev[5]));
    M->stats.ms_fill = ms_fill; // span of the copy stream: fills + D2H, overlapped with tracing
    M->stats.ms_d2h = 0.f;
    M->stats.trace_launches = (int)nsub;
    M->stats.kernel_launches = launches;
    M->stats.d2h_bytes = d2h_bytes + (int64_t)sizeof(tested);
    if (row_counts)
        for (size_t r = 0; r < m; ++r) row_counts[r] = (int64_t)h_counts[r];
    check_error_flag(M);
    if (overflow) return false;
    if (destination == 0 && indptr) { // global indptr on the host, in the index dtype
        int64_t run = 0;
        if (index_width == 4) {
            int32_t *ip = reinterpret_cast<int32_t *>(indptr);
            ip[0] = 0;
            for (size_t r = 0; r < m; ++r) ip[r + 1] = (int32_t)(run += h_counts[r]);
        } else {
            int64_t *ip = reinterpret_cast<int64_t *>(indptr);
            ip[0] = 0;
            for (size_t r = 0; r < m; ++r) ip[r + 1] = (run += h_counts[r]);
        }
    }
    M->out_index_width = 0;
    if (destination == 2) { // device indptr for downstream device consumers
        M->out_index_width = index_width;
        M->indptr.reserve(sizeof(int64_t) * (m + 1));
        std::vector<int64_t> ip(m + 1, 0);
        for (size_t r = 0; r < m; ++r) ip[r + 1] = ip[r] + h_counts[r];
        FB_CUDA(cudaMemcpyAsync(M->indptr.p, ip.data(), sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice, s0));
        FB_CUDA(cudaStreamSynchronize(s0));
        M->have_count = true; // fluxb200_ff_device_csr is valid
    }
    return true;
}

template <class T> void visibility(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                                   uint8_t *vis, int brute) {
    if (!m || !n) return;
    cudaStream_t st = M->stream;
    upload_index_sets(M, I, m, J, n);
    const int64_t total = (int64_t)m * (int64_t)n;
    M->qout.reserve((size_t)total);
    FB_REQUIRE(ceil_div(total, 128) < (1ll << 31), "visibility: too many pairs for one call");
    visibility_kernel<T><<<blocks_for(total, 128), 128, 0, st>>>(
        M->faceP.as<Real4<T>>(), M->rows.as<int>(), (int)m, M->cs->cols.as<int>(), (int)n,
        M->face_leaf.as<int>(), M->nodes.as<float4>(), M->tri.as<float4>(), M->ninternal, (int)M->nf,
        M->scalars.as<int>() + 6, brute,
        M->qout.as<uint8_t>());
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaMemcpyAsync(vis, M->qout.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    check_error_flag(M);
}

template <class T> void is_occluded(fluxb200_mesh *M, const int64_t *I, size_t m, const void *D, size_t nd,
                                    int mode, uint8_t *occ) {
    FB_REQUIRE(mode >= 0 && mode <= 2, "is_occluded: mode must be 0, 1 or 2");
    if (mode == 0) FB_REQUIRE(nd == 1, "is_occluded: mode 0 takes one direction");
    if (mode == 1) FB_REQUIRE(nd == m, "is_occluded: mode 1 needs one direction per face of I");
    if (!m || !nd) return;
    cudaStream_t st = M->stream;
    upload_index_sets(M, I, m, nullptr, 0);
    const int64_t total = (int64_t)m * (mode == 2 ? (int64_t)nd : 1);
    M->qout.reserve((size_t)total);
    M->qtmp.reserve(sizeof(T) * 3 * nd);
    FB_CUDA(cudaMemcpyAsync(M->qtmp.p, D, sizeof(T) * 3 * nd, cudaMemcpyHostToDevice, st));
    occluded_kernel<T><<<blocks_for(total, 128), 128, 0, st>>>(
        M->faceP.as<Real4<T>>(), M->faceN.as<Real4<T>>(), M->rows.as<int>(), (int)m, M->qtmp.as<T>(),
        (int)nd, mode, M->nodes.as<float4>(), M->tri.as<float4>(), M->ninternal, (int)M->nf,
        M->scalars.as<int>() + 6, M->qout.as<uint8_t>());
    FB_CUDA(cudaGetLastError());
    FB_CUDA(cudaMemcpyAsync(occ, M->qout.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    check_error_flag(M);
}

} // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
#define DISPATCH(M, fn, ...)                                   \
    do {                                                       \
        if ((M)->dtype == FLUXB200_F64) fn<double>(__VA_ARGS__); \
        else fn<float>(__VA_ARGS__);                           \
    } while (0)

extern "C" {

const char *fluxb200_last_error(void) { return g_last_error.c_str(); }
int fluxb200_abi_version(void) { return FLUXB200_ABI_VERSION; }

int fluxb200_device_count(int *count) {
    return guarded([&] {
        FB_REQUIRE(count, "count is NULL");
        FB_CUDA(cudaGetDeviceCount(count));
    });
}

int fluxb200_mesh_create(const void *V, size_t nv, const int64_t *F, size_t nf, int dtype_code, int device,
                         fluxb200_mesh **out) {
    fluxb200_mesh *M = nullptr;
    int rc = guarded([&] {
        FB_REQUIRE(out, "out is NULL");
        FB_REQUIRE(dtype_code == FLUXB200_F32 || dtype_code == FLUXB200_F64,
                   "unsupported dtype (float32 and float64 only)");
        FB_REQUIRE((V || !nv) && (F || !nf), "V / F is NULL");
        FB_REQUIRE(nf < (1ull << 30) && nv < (1ull << 31), "mesh too large");
        int ndev = 0;
        FB_CUDA(cudaGetDeviceCount(&ndev));
        FB_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
        DeviceGuard guard(device);
        M = new fluxb200_mesh();
        M->device = device;
        M->dtype = dtype_code;
        M->nv = nv;
        M->nf = nf;
        int prio_lo = 0, prio_hi = 0;
        FB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        FB_CUDA(cudaStreamCreateWithPriority(&M->stream, cudaStreamNonBlocking, prio_lo));
        // fill + copies run under the persistent trace kernel: let them win free CTA slots
        FB_CUDA(cudaStreamCreateWithPriority(&M->copy_stream, cudaStreamNonBlocking, prio_hi));
        FB_CUDA(cudaStreamCreateWithPriority(&M->d2h_stream, cudaStreamNonBlocking, prio_hi));
        for (auto &e : M->slot_free) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : M->d2h_done) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : M->ev) FB_CUDA(cudaEventCreate(&e));
        FB_CUDA(cudaDeviceGetAttribute(&M->num_sms, cudaDevAttrMultiProcessorCount, device));
        FB_CUDA(cudaDeviceGetAttribute(&M->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        std::vector<int> F32(3 * nf);
        for (size_t k = 0; k < 3 * nf; ++k) {
            if (F[k] < 0 || (size_t)F[k] >= nv) throw CudaError{"F: vertex index out of range"};
            F32[k] = (int)F[k];
        }
        const size_t es = M->esize();
        M->V.reserve(es * 3 * std::max<size_t>(nv, 1));
        M->F.reserve(sizeof(int) * 3 * std::max<size_t>(nf, 1));
        M->V32.reserve(sizeof(float) * 3 * std::max<size_t>(nv, 1));
        M->faceP.reserve(es * 4 * std::max<size_t>(nf, 1));
        M->faceN.reserve(es * 4 * std::max<size_t>(nf, 1));
        if (nv) FB_CUDA(cudaMemcpyAsync(M->V.p, V, es * 3 * nv, cudaMemcpyHostToDevice, M->stream));
        if (nf) FB_CUDA(cudaMemcpyAsync(M->F.p, F32.data(), sizeof(int) * 3 * nf, cudaMemcpyHostToDevice, M->stream));
        FB_CUDA(cudaStreamSynchronize(M->stream));
        DISPATCH(M, build_geometry, M);
        bvh_build(M);
        *out = M;
    });
    if (rc && M) {
        fluxb200_mesh_destroy(M);
    }
    return rc;
}

int fluxb200_mesh_destroy(fluxb200_mesh *M) {
    if (!M) return 0;
    return guarded([&] {
        DeviceGuard guard(M->device);
        if (M->stream) cudaStreamSynchronize(M->stream);
        DevBuf *bufs[] = {&M->V, &M->F, &M->V32, &M->faceP, &M->faceN, &M->keys, &M->vals, &M->left, &M->right,
                          &M->parent, &M->first, &M->last, &M->box, &M->slab, &M->flags, &M->pre, &M->flag_by_pre,
                          &M->top_before, &M->scene, &M->scalars, &M->nodes, &M->tri, &M->face_leaf, &M->node_up, &M->leaf_up, &M->node_range, &M->rows,
                          &M->ckeys, &M->cvals, &M->lost, &M->bits, &M->row_counts, &M->counts64, &M->indptr, &M->indptr32,
                          &M->tested, &M->out_data, &M->out_indices, &M->qtmp, &M->qout, &M->jbits,
                          &M->hz, &M->zone_node, &M->zone_up};
        for (DevBuf *b : bufs) b->release();
        for (auto &C : M->colsets) C->release();
        M->colsets.clear();
        for (int k = 0; k < fluxb200_mesh::kSlots; ++k) {
            M->sbits[k].release(); M->scounts[k].release(); M->scounts64[k].release(); M->sindptr[k].release();
            M->stage_data[k].release(); M->stage_idx[k].release();
            M->sjbits[k].release();
            if (M->slot_free[k]) cudaEventDestroy(M->slot_free[k]);
            if (M->d2h_done[k]) cudaEventDestroy(M->d2h_done[k]);
        }
        M->h_nnz.release(); M->h_counts.release(); M->h_jbits.release();
        for (auto &hb : M->h_data) hb.release();
        for (auto &e : M->sub_events) cudaEventDestroy(e);
        for (auto &e : M->tl_events) cudaEventDestroy(e);
        if (M->copy_stream) { cudaStreamSynchronize(M->copy_stream); cudaStreamDestroy(M->copy_stream); }
        if (M->d2h_stream) { cudaStreamSynchronize(M->d2h_stream); cudaStreamDestroy(M->d2h_stream); }
        M->sorter.release();
        for (auto &e : M->ev)
            if (e) cudaEventDestroy(e);
        if (M->stream) cudaStreamDestroy(M->stream);
        delete M;
    });
}

int fluxb200_mesh_set_face_data(fluxb200_mesh *M, const void *P, const void *N, const void *A) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        DISPATCH(M, set_face_data, M, P, N, A);
    });
}

int fluxb200_mesh_get_face_data(fluxb200_mesh *M, void *P, void *N, void *A) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        DISPATCH(M, get_face_data, M, P, N, A);
    });
}

int fluxb200_bvh_build(fluxb200_mesh *M) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        M->have_count = false;
        bvh_build(M);
    });
}

int fluxb200_bvh_info_get(fluxb200_mesh *M, fluxb200_bvh_info *info) {
    return guarded([&] {
        FB_REQUIRE(M && info, "NULL argument");
        info->num_faces = (int64_t)M->nf;
        info->num_nodes = M->ninternal;
        info->num_top_nodes = M->ntop;
        info->max_depth = M->max_depth;
        info->ms_build = M->ms_build;
        for (int k = 0; k < 3; ++k) {
            info->scene_lo[k] = M->scene_h[k];
            info->scene_hi[k] = M->scene_h[3 + k];
        }
    });
}

int fluxb200_bvh_export(fluxb200_mesh *M, float *nodes, int32_t *leaf_face) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        if (!M->nf) return;
        if (nodes)
            FB_CUDA(cudaMemcpyAsync(nodes, M->nodes.p, sizeof(float) * 24 * (size_t)M->ninternal,
                                    cudaMemcpyDeviceToHost, M->stream));
        if (leaf_face)
            FB_CUDA(cudaMemcpyAsync(leaf_face, M->vals.p, sizeof(int32_t) * M->nf, cudaMemcpyDeviceToHost,
                                    M->stream));
        FB_CUDA(cudaStreamSynchronize(M->stream));
        if (nodes) { // the documented export layout per child: (lo.xyz, ref) (hi.xyz, slab_min) (slab_dir.xyz, slab_max);
            // on the device the records are stored pair-aligned for the packed box / slab test (trace.cuh child_hit)
            for (size_t k = 0; k < 2 * (size_t)M->ninternal; ++k) {
                float *q = nodes + 12 * k;
                const float d[12] = {q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10], q[11]};
                q[0] = d[0]; q[1] = d[2]; q[2] = d[4]; q[3] = d[11];   // lo.xyz | ref
                q[4] = d[1]; q[5] = d[3]; q[6] = d[5]; q[7] = d[6];    // hi.xyz | slab_min
                q[8] = d[8]; q[9] = d[9]; q[10] = d[10]; q[11] = d[7]; // slab_dir | slab_max
            }
        }
    });
}

int fluxb200_ff_count(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n, double eps,
                      int64_t *row_counts, fluxb200_ff_stats *stats) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        DISPATCH(M, ff_count, M, I, m, J, n, eps, row_counts);
        if (stats) *stats = M->stats;
    });
}

int fluxb200_ff_fill(fluxb200_mesh *M, int index_width, int destination, void *indptr, void *indices,
                     void *data, fluxb200_ff_stats *stats) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        DISPATCH(M, ff_fill, M, index_width, destination, indptr, indices, data);
        if (stats) *stats = M->stats;
    });
}

int fluxb200_ff_assemble(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n, double eps,
                         int index_width, int destination, void *indptr, void *indices, void *data,
                         int64_t capacity, int64_t *row_counts, fluxb200_ff_stats *stats) {
    bool ok = true;
    int rc = guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        if (M->dtype == FLUXB200_F64)
            ok = ff_assemble<double>(M, I, m, J, n, eps, index_width, destination, indptr, indices, data,
                                     capacity, row_counts);
        else
            ok = ff_assemble<float>(M, I, m, J, n, eps, index_width, destination, indptr, indices, data,
                                    capacity, row_counts);
        if (stats) *stats = M->stats;
    });
    if (rc) return rc;
    if (!ok) {
        set_error("capacity too small for the assembled matrix (stats.nnz entries needed)");
        return FLUXB200_OVERFLOW;
    }
    return 0;
}

int fluxb200_host_alloc(size_t bytes, void **ptr) {
    return guarded([&] {
        FB_REQUIRE(ptr, "ptr is NULL");
        FB_CUDA(cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocPortable));
    });
}

int fluxb200_host_free(void *ptr) {
    return guarded([&] {
        if (ptr) FB_CUDA(cudaFreeHost(ptr));
    });
}

int fluxb200_ff_device_csr(fluxb200_mesh *M, void **indptr, void **indices, void **data, int64_t *nnz) {
    return guarded([&] {
        FB_REQUIRE(M && M->have_count, "no assembled matrix on this handle");
        if (indptr) *indptr = M->indptr.p;
        if (indices) *indices = M->out_indices.p;
        if (data) *data = M->out_data.p;
        if (nnz) *nnz = M->nnz;
    });
}

// ---- device-resident CSR slab: detach, products, download -----
int fluxb200_ff_detach_csr(fluxb200_mesh *M, fluxb200_csr **out) {
    return guarded([&] {
        FB_REQUIRE(M && out, "NULL argument");
        FB_REQUIRE(M->have_count && M->out_index_width != 0, "no device-resident matrix on this handle "
                   "(run fluxb200_ff_assemble / fluxb200_ff_fill with destination 2 first)");
        DeviceGuard guard(M->device);
        fluxb200_csr *C = new fluxb200_csr();
        C->device = M->device;
        C->dtype = M->dtype;
        C->index_width = M->out_index_width;
        C->m = (int64_t)M->m;
        C->n = (int64_t)M->n;
        C->nnz = M->nnz;
        FB_CUDA(cudaStreamCreateWithFlags(&C->stream, cudaStreamNonBlocking));
        for (auto &e : C->ev) FB_CUDA(cudaEventCreate(&e));
        std::swap(C->indptr, M->indptr);
        std::swap(C->indices, M->out_indices);
        std::swap(C->data, M->out_data);
        M->have_count = false;
        M->out_index_width = 0;
        M->dev_capacity_hint = 0;
        *out = C;
    });
}

int fluxb200_csr_destroy(fluxb200_csr *C) {
    if (!C) return 0;
    return guarded([&] {
        DeviceGuard guard(C->device);
        if (C->stream) cudaStreamSynchronize(C->stream);
        C->indptr.release();
        C->indices.release();
        C->data.release();
        C->scratch.release();
        for (auto &e : C->ev)
            if (e) cudaEventDestroy(e);
        if (C->stream) cudaStreamDestroy(C->stream);
        delete C;
    });
}

int fluxb200_csr_info(fluxb200_csr *C, int64_t *m, int64_t *n, int64_t *nnz, int *dtype_code, int *index_width,
                      float *last_ms) {
    return guarded([&] {
        FB_REQUIRE(C, "csr is NULL");
        if (m) *m = C->m;
        if (n) *n = C->n;
        if (nnz) *nnz = C->nnz;
        if (dtype_code) *dtype_code = C->dtype;
        if (index_width) *index_width = C->index_width;
        if (last_ms) *last_ms = C->last_ms;
    });
}

int fluxb200_csr_to_host(fluxb200_csr *C, void *indptr, void *indices, void *data) {
    return guarded([&] {
        FB_REQUIRE(C, "csr is NULL");
        DeviceGuard guard(C->device);
        const size_t es = C->dtype == FLUXB200_F64 ? 8 : 4;
        FB_REQUIRE(!(indptr && C->index_width == 4 && C->nnz >= (1ll << 31)),
                   "this slab has 2^31 or more entries: its int32 index dtype cannot hold the row offsets on the host");
        if (indptr) { // int64 on the device; narrowed on the host side if asked
            std::vector<int64_t> ip((size_t)C->m + 1);
            FB_CUDA(cudaMemcpyAsync(ip.data(), C->indptr.p, sizeof(int64_t) * ip.size(), cudaMemcpyDeviceToHost, C->stream));
            FB_CUDA(cudaStreamSynchronize(C->stream));
            if (C->index_width == 4)
                for (size_t k = 0; k < ip.size(); ++k) reinterpret_cast<int32_t *>(indptr)[k] = (int32_t)ip[k];
            else
                memcpy(indptr, ip.data(), sizeof(int64_t) * ip.size());
        }
        if (indices && C->nnz)
            FB_CUDA(cudaMemcpyAsync(indices, C->indices.p, (size_t)C->index_width * (size_t)C->nnz, cudaMemcpyDeviceToHost, C->stream));
        if (data && C->nnz)
            FB_CUDA(cudaMemcpyAsync(data, C->data.p, es * (size_t)C->nnz, cudaMemcpyDeviceToHost, C->stream));
        FB_CUDA(cudaStreamSynchronize(C->stream));
    });
}

int fluxb200_csr_jacobi_step(fluxb200_csr *C, const double *E_dev, const double *rho_dev, double rho_scalar,
                             const double *x_dev, double *y_dev, double *diffmax_host, int64_t row_offset) {
    return guarded([&] {
        FB_REQUIRE(C && x_dev && y_dev, "NULL argument");
        DeviceGuard guard(C->device);
        if (diffmax_host) FB_REQUIRE(row_offset >= 0 && row_offset + C->m <= C->n, "row_offset outside the iterate");
        C->scratch.reserve(sizeof(unsigned long long));
        unsigned long long *dm = diffmax_host ? C->scratch.as<unsigned long long>() : nullptr;
        if (dm) FB_CUDA(cudaMemsetAsync(dm, 0, sizeof(unsigned long long), C->stream));
        FB_CUDA(cudaEventRecord(C->ev[0], C->stream));
        if (C->m) {
            const unsigned grid = (unsigned)C->m;
#define FB_SPMV(TT, II)                                                                                   \
    csr_jacobi_kernel<TT, II><<<grid, kSpmvThreads, 0, C->stream>>>(                                      \
        C->indptr.as<int64_t>(), C->indices.as<II>(), C->data.as<TT>(), (int)C->m, E_dev, rho_dev, rho_scalar, \
        x_dev, y_dev, dm, row_offset)
            if (C->dtype == FLUXB200_F64) {
                if (C->index_width == 4) FB_SPMV(double, int32_t);
                else FB_SPMV(double, int64_t);
            } else {
                if (C->index_width == 4) FB_SPMV(float, int32_t);
                else FB_SPMV(float, int64_t);
            }
#undef FB_SPMV
            FB_CUDA(cudaGetLastError());
        }
        FB_CUDA(cudaEventRecord(C->ev[1], C->stream));
        unsigned long long bits = 0;
        if (dm) FB_CUDA(cudaMemcpyAsync(&bits, dm, sizeof(bits), cudaMemcpyDeviceToHost, C->stream));
        FB_CUDA(cudaStreamSynchronize(C->stream));
        FB_CUDA(cudaEventElapsedTime(&C->last_ms, C->ev[0], C->ev[1]));
        if (diffmax_host) memcpy(diffmax_host, &bits, sizeof(double));
    });
}

// ---- sub-block extraction and thin dense products on a resident slab ----
int fluxb200_csr_extract(fluxb200_csr *C, const int64_t *rows, size_t mr, const int64_t *cols, size_t nc,
                         fluxb200_csr **out) {
    fluxb200_csr *B = nullptr;
    int rc = guarded([&] {
        FB_REQUIRE(C && out && (rows || !mr) && (cols || !nc), "NULL argument");
        FB_REQUIRE(mr < (1ull << 31) && nc < (1ull << 31), "index sets too large");
        DeviceGuard guard(C->device);
        std::vector<int> hrows(mr), newpos((size_t)C->n, -1);
        for (size_t k = 0; k < mr; ++k) {
            FB_REQUIRE(rows[k] >= 0 && rows[k] < C->m, "row index out of range");
            hrows[k] = (int)rows[k];
        }
        for (size_t k = 0; k < nc; ++k) {
            FB_REQUIRE(cols[k] >= 0 && cols[k] < C->n, "column index out of range");
            FB_REQUIRE(newpos[(size_t)cols[k]] < 0, "repeated column index");
            newpos[(size_t)cols[k]] = (int)k;
        }
        B = new fluxb200_csr();
        B->device = C->device;
        B->dtype = C->dtype;
        B->index_width = C->index_width;
        B->m = (int64_t)mr;
        B->n = (int64_t)nc;
        FB_CUDA(cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking));
        for (auto &e : B->ev) FB_CUDA(cudaEventCreate(&e));
        cudaStream_t st = C->stream;
        DevBuf drows, dnew, dcounts;
        drows.reserve(sizeof(int) * std::max<size_t>(mr, 1));
        dnew.reserve(sizeof(int) * std::max<size_t>((size_t)C->n, 1));
        dcounts.reserve(sizeof(int64_t) * std::max<size_t>(mr, 1));
        B->indptr.reserve(sizeof(int64_t) * (mr + 1));
        if (mr) FB_CUDA(cudaMemcpyAsync(drows.p, hrows.data(), sizeof(int) * mr, cudaMemcpyHostToDevice, st));
        if (C->n) FB_CUDA(cudaMemcpyAsync(dnew.p, newpos.data(), sizeof(int) * (size_t)C->n, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemsetAsync(B->indptr.p, 0, sizeof(int64_t) * (mr + 1), st));
        FB_CUDA(cudaEventRecord(C->ev[0], st));
        if (mr) {
            if (C->index_width == 4)
                extract_count_kernel<int32_t><<<(unsigned)mr, kBlockThreads, 0, st>>>(
                    C->indptr.as<int64_t>(), C->indices.as<int32_t>(), drows.as<int>(), (int)mr, dnew.as<int>(),
                    dcounts.as<int64_t>());
            else
                extract_count_kernel<int64_t><<<(unsigned)mr, kBlockThreads, 0, st>>>(
                    C->indptr.as<int64_t>(), C->indices.as<int64_t>(), drows.as<int>(), (int)mr, dnew.as<int>(),
                    dcounts.as<int64_t>());
            scan_exclusive<int64_t, int64_t>(dcounts.as<int64_t>(), B->indptr.as<int64_t>(), (int64_t)mr,
                                             B->indptr.as<int64_t>() + mr, st);
        }
        int64_t nnz = 0;
        FB_CUDA(cudaMemcpyAsync(&nnz, B->indptr.as<int64_t>() + mr, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        FB_CUDA(cudaStreamSynchronize(st));
        B->nnz = nnz;
        const size_t es = C->dtype == FLUXB200_F64 ? 8 : 4;
        B->indices.reserve((size_t)C->index_width * std::max<int64_t>(nnz, 1));
        B->data.reserve(es * std::max<int64_t>(nnz, 1));
        if (mr && nnz) {
#define FB_EXTRACT(TT, II)                                                                                     \
    extract_fill_kernel<TT, II><<<(unsigned)mr, kBlockThreads, 0, st>>>(                                       \
        C->indptr.as<int64_t>(), C->indices.as<II>(), C->data.as<TT>(), drows.as<int>(), (int)mr, dnew.as<int>(), \
        B->indptr.as<int64_t>(), B->indices.as<II>(), B->data.as<TT>())
            if (C->dtype == FLUXB200_F64) {
                if (C->index_width == 4) FB_EXTRACT(double, int32_t);
                else FB_EXTRACT(double, int64_t);
            } else {
                if (C->index_width == 4) FB_EXTRACT(float, int32_t);
                else FB_EXTRACT(float, int64_t);
            }
#undef FB_EXTRACT
            FB_CUDA(cudaGetLastError());
        }
        FB_CUDA(cudaEventRecord(C->ev[1], st));
        FB_CUDA(cudaStreamSynchronize(st));
        FB_CUDA(cudaEventElapsedTime(&C->last_ms, C->ev[0], C->ev[1]));
        drows.release();
        dnew.release();
        dcounts.release();
        *out = B;
    });
    if (rc && B) fluxb200_csr_destroy(B);
    return rc;
}

int fluxb200_csr_matmat(fluxb200_csr *C, const double *X_dev, int k, double *Y_dev, int transpose) {
    return guarded([&] {
        FB_REQUIRE(C && X_dev && Y_dev, "NULL argument");
        FB_REQUIRE(k >= 1 && k <= 32, "matmat: 1 <= k <= 32 right-hand sides per call");
        DeviceGuard guard(C->device);
        cudaStream_t st = C->stream;
        FB_CUDA(cudaEventRecord(C->ev[0], st));
        if (transpose) FB_CUDA(cudaMemsetAsync(Y_dev, 0, sizeof(double) * (size_t)C->n * k, st));
        if (C->m) {
            const unsigned grid = (unsigned)C->m;
#define FB_MM(KERNEL, TT, II)                                                                          \
    KERNEL<TT, II><<<grid, kBlockThreads, 0, st>>>(C->indptr.as<int64_t>(), C->indices.as<II>(),       \
                                                   C->data.as<TT>(), (int)C->m, X_dev, k, Y_dev)
#define FB_MM_DISPATCH(KERNEL)                                  \
    do {                                                        \
        if (C->dtype == FLUXB200_F64) {                         \
            if (C->index_width == 4) FB_MM(KERNEL, double, int32_t); \
            else FB_MM(KERNEL, double, int64_t);                \
        } else {                                                \
            if (C->index_width == 4) FB_MM(KERNEL, float, int32_t);  \
            else FB_MM(KERNEL, float, int64_t);                 \
        }                                                       \
    } while (0)
            if (transpose) FB_MM_DISPATCH(csr_rmatmat_kernel);
            else FB_MM_DISPATCH(csr_matmat_kernel);
#undef FB_MM_DISPATCH
#undef FB_MM
            FB_CUDA(cudaGetLastError());
        }
        FB_CUDA(cudaEventRecord(C->ev[1], st));
        FB_CUDA(cudaStreamSynchronize(st));
        FB_CUDA(cudaEventElapsedTime(&C->last_ms, C->ev[0], C->ev[1]));
    });
}

int fluxb200_visibility(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                        uint8_t *vis) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        M->have_count = false;
        DISPATCH(M, visibility, M, I, m, J, n, vis, 0);
    });
}

int fluxb200_visibility_bruteforce(fluxb200_mesh *M, const int64_t *I, size_t m, const int64_t *J, size_t n,
                                   uint8_t *vis) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        M->have_count = false;
        DISPATCH(M, visibility, M, I, m, J, n, vis, 1);
    });
}

int fluxb200_is_occluded(fluxb200_mesh *M, const int64_t *I, size_t m, const void *D, size_t nd, int mode,
                         uint8_t *occluded) {
    return guarded([&] {
        FB_REQUIRE(M, "mesh is NULL");
        DeviceGuard guard(M->device);
        M->have_count = false;
        DISPATCH(M, is_occluded, M, I, m, D, nd, mode, occluded);
    });
}

int fluxb200_intersect1(fluxb200_mesh *M, const double x[3], const double d[3], int *hit, int64_t *face,
                        double *t, double xt[3]) {
    return guarded([&] {
        FB_REQUIRE(M && x && d && hit, "NULL argument");
        DeviceGuard guard(M->device);
        *hit = 0;
        if (!M->nf) return;
        M->scalars.reserve(sizeof(int) * 8);
        int *dface = M->scalars.as<int>() + 4;
        float *dt = reinterpret_cast<float *>(M->scalars.as<int>() + 5);
        intersect1_kernel<<<1, 1, 0, M->stream>>>((float)x[0], (float)x[1], (float)x[2], (float)d[0],
                                                  (float)d[1], (float)d[2], M->nodes.as<float4>(),
                                                  M->tri.as<float4>(), M->ninternal, (int)M->nf,
                                                  M->scalars.as<int>() + 6, dface, dt);
        FB_CUDA(cudaGetLastError());
        int hface;
        float ht;
        FB_CUDA(cudaMemcpyAsync(&hface, dface, sizeof(int), cudaMemcpyDeviceToHost, M->stream));
        FB_CUDA(cudaMemcpyAsync(&ht, dt, sizeof(float), cudaMemcpyDeviceToHost, M->stream));
        FB_CUDA(cudaStreamSynchronize(M->stream));
        if (hface >= 0) {
            *hit = 1;
            if (face) *face = hface;
            if (t) *t = ht;
            if (xt)
                for (int k = 0; k < 3; ++k) xt[k] = x[k] + (double)ht * d[k];
        }
    });
}

int fluxb200_slab_plan(size_t m, int nranks, const int64_t *weights, int64_t *starts) {
    return guarded([&] {
        FB_REQUIRE(nranks > 0 && starts, "bad arguments");
        starts[0] = 0;
        if (!weights) {
            for (int r = 1; r <= nranks; ++r) starts[r] = (int64_t)(((__int128)m * r) / nranks);
            return;
        }
        long double total = 0;
        for (size_t k = 0; k < m; ++k) total += (long double)(weights[k] > 0 ? weights[k] : 0) + 1;
        long double acc = 0;
        size_t k = 0;
        for (int r = 1; r < nranks; ++r) {
            const long double goal = total * r / nranks;
            while (k < m && acc + (long double)(weights[k] > 0 ? weights[k] : 0) + 1 <= goal) {
                acc += (long double)(weights[k] > 0 ? weights[k] : 0) + 1;
                ++k;
            }
            starts[r] = (int64_t)k;
        }
        starts[nranks] = (int64_t)m;
    });
}

int fluxb200_expand_words(const uint32_t *words, size_t nwords, int index_width, void *out, int64_t *count) {
    return guarded([&] {
        FB_REQUIRE((words || !nwords) && out && count, "NULL argument");
        FB_REQUIRE(index_width == 4 || index_width == 8, "index_width must be 4 or 8");
        FB_REQUIRE(nwords < (1ull << 26), "too many words");
        *count = expand_words(words, (int)nwords, index_width, out);
    });
}

int fluxb200_expand_rows(const uint32_t *words, size_t nwords, size_t mr, const int64_t *offs, int index_width,
                         void *indices, int nthreads) {
    return guarded([&] {
        FB_REQUIRE((words || !nwords || !mr) && offs && (indices || offs[mr] == offs[0]), "NULL argument");
        FB_REQUIRE(index_width == 4 || index_width == 8, "index_width must be 4 or 8");
        FB_REQUIRE(nwords < (1ull << 26) && nthreads >= 0 && nthreads <= 64, "argument out of range");
        if (!mr) return;
        HostExpander pool; // the same worker pool fluxb200_ff_assemble feeds from its copy-out callbacks
        if (nthreads == 0) nthreads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        pool.start(nthreads);
        ExpandTask t;
        t.words = words;
        t.nwords = (int)nwords;
        t.mr = mr;
        t.offs.assign(offs, offs + mr + 1);
        t.indices = indices;
        t.index_width = index_width;
        t.pieces = (int)std::max<size_t>(1, std::min<size_t>(mr, 2 * (size_t)nthreads));
        t.pending.store(t.pieces);
        t.owner = &pool;
        pool.submit(&t);
        pool.wait(&t);
        FB_REQUIRE(t.mismatch.load() == 0, "a row's set bits do not match its CSR row length");
    });
}

int fluxb200_mesh_stream(fluxb200_mesh *M, void **stream) {
    return guarded([&] {
        FB_REQUIRE(M && stream, "NULL argument");
        *stream = (void *)M->stream;
    });
}

int fluxb200_trace_counters(fluxb200_mesh *M, int64_t out[8]) {
    return guarded([&] {
        FB_REQUIRE(M && out, "NULL argument");
        DeviceGuard guard(M->device);
        unsigned long long h[8] = {};
        if (M->tested.p) {
            FB_CUDA(cudaMemcpyAsync(h, M->tested.p, sizeof(h), cudaMemcpyDeviceToHost, M->stream));
            FB_CUDA(cudaStreamSynchronize(M->stream));
        }
        out[0] = (int64_t)h[0];
        out[1] = (int64_t)h[2];
        out[2] = (int64_t)h[3];
        out[3] = (int64_t)h[4];
        out[4] = (int64_t)h[5];
        out[5] = (int64_t)h[6];
        out[6] = (int64_t)h[7];
        out[7] = M->colset_hits;
    });
}

#ifdef FB_EMU
// SIMT-emulator builds only (tools/simt): read and reset the trace kernel's loop-iteration counters
int fluxb200_emu_trace_iterations(int64_t out[8]) {
    for (int k = 0; k < 8; ++k) out[k] = (int64_t)__atomic_exchange_n(&emu_stats::iters[k], 0ull, __ATOMIC_RELAXED);
    return 0;
}
#endif

int fluxb200_set_option(fluxb200_mesh *M, const char *name, int64_t value) {
    return guarded([&] {
        FB_REQUIRE(M && name, "NULL argument");
        DeviceGuard guard(M->device);
        const std::string s(name);
        if (s == "slab_limit") {
            FB_REQUIRE(value >= 0 && value < (1ll << 31), "slab_limit out of range");
            M->slab_limit_opt = (int)value;
            M->have_count = false;
            bvh_build(M);
        } else if (s == "shaft_filter") {
            M->shaft_filter_opt = value ? 1 : 0;
        } else if (s == "horizon_skip") {
            M->horizon_skip_opt = value ? 1 : 0;
            M->have_count = false;
            bvh_build(M); // the zone table belongs to the tree
        } else if (s == "horizon_zone") {
            FB_REQUIRE(value >= 1 && value <= 65536, "horizon_zone out of range (1..65536 leaves)");
            M->horizon_zone_opt = (int)value;
            M->have_count = false;
            bvh_build(M);
        } else if (s == "colset_cache") {
            FB_REQUIRE(value >= 0 && value <= 64, "colset_cache out of range (0..64 prepared column sets)");
            M->colset_cache_opt = (int)value;
            for (size_t k = 1; k < M->colsets.size(); ++k) M->colsets[k]->release();
            M->colsets.resize(std::min<size_t>(M->colsets.size(), 1));
        } else if (s == "trace_variant") {
            FB_REQUIRE(value == 1 || value == 2, "trace_variant must be 1 or 2");
            M->trace_variant_opt = (int)value;
        } else if (s == "blocks_per_sm") {
            FB_REQUIRE(value >= 1 && value <= 8, "blocks_per_sm out of range");
            M->blocks_per_sm = (int)value;
        } else if (s == "top_nodes") {
            FB_REQUIRE(value >= 0 && value * 96 + 16384 <= M->max_smem_optin, "top_nodes out of range");
            M->top_nodes_opt = (int)value;
            M->have_count = false;
            bvh_build(M);
        } else if (s == "host_expand") {
            M->host_expand_opt = value ? 1 : 0;
        } else if (s == "host_threads") {
            FB_REQUIRE(value >= 0 && value <= 64, "host_threads out of range");
            M->host_threads_opt = (int)value;
        } else if (s == "fill_rows") {
            FB_REQUIRE(value >= -1 && value <= 8, "fill_rows out of range");
            M->fill_rows_opt = (int)value;
        } else if (s == "pipeline_ramp") {
            M->ramp_opt = value ? 1 : 0;
        } else if (s == "sub_rows") {
            FB_REQUIRE(value >= 1 && value <= (1 << 20), "sub_rows out of range");
            M->sub_rows_opt = (int)value;
        } else {
            throw CudaError{"unknown option " + s};
        }
    });
}

} // extern "C"
