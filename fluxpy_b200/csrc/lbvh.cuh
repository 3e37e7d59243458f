// lbvh.cuh -- K0 face geometry, K1 Morton codes, K3 Karras hierarchy + refit +
// fitted slabs + flattening into 96-byte two-child nodes.
//
// Replaces the Embree scene build of EmbreeTrimeshShapeModel._make_scene
// (reference src/flux/shape.py:296-344) and the NumPy face geometry helpers
// (src/flux/shape.py:16-45).
#pragma once
#include "common.cuh"
#include "prims.cuh"
#include <float.h>

namespace fluxb200 {

// Node format: see flatten_kernel.  Every child carries, besides its AABB, a
// SLAB fitted to the surface it bounds: the unit direction of the subtree's
// area-weighted normal and the extent of its vertices along it.  Centroid-to-
// centroid rays graze the surface near both ends; the AABB of a sloped patch
// is mostly empty space above the surface, the fitted slab is as thin as the
// patch is flat, so most grazing false positives are rejected.

template <class T> struct Real4 { T x, y, z, w; };
template <> struct __align__(16) Real4<float> { float x, y, z, w; };
template <> struct __align__(32) Real4<double> { double x, y, z, w; };

// ---- K0: face geometry in the array dtype, NumPy operation order ------------
template <class T> __device__ __forceinline__ T rn_mul(T a, T b);
template <> __device__ __forceinline__ float rn_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double rn_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T rn_add(T a, T b);
template <> __device__ __forceinline__ float rn_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double rn_add<double>(double a, double b) { return __dadd_rn(a, b); }
template <class T> __device__ __forceinline__ T rn_sub(T a, T b);
template <> __device__ __forceinline__ float rn_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double rn_sub<double>(double a, double b) { return __dsub_rn(a, b); }
template <class T> __device__ __forceinline__ T rn_div(T a, T b);
template <> __device__ __forceinline__ float rn_div<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double rn_div<double>(double a, double b) { return __ddiv_rn(a, b); }
template <class T> __device__ __forceinline__ T rn_sqrt(T a);
template <> __device__ __forceinline__ float rn_sqrt<float>(float a) { return __fsqrt_rn(a); }
template <> __device__ __forceinline__ double rn_sqrt<double>(double a) { return __dsqrt_rn(a); }

// P = V[F].mean(axis=1); C = cross(v1-v0, v2-v0); N = C/|C|; A = |C|/2
// (shape.py:16-45), written into the packed per-face arrays
//   faceP[f] = (P.x, P.y, P.z, A)   faceN[f] = (N.x, N.y, N.z, 0)
template <class T>
__global__ void face_geometry_kernel(const T *__restrict__ V, const int *__restrict__ F, int nf,
                                     Real4<T> *__restrict__ faceP, Real4<T> *__restrict__ faceN) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const T *v0 = V + 3 * (size_t)F[3 * f], *v1 = V + 3 * (size_t)F[3 * f + 1],
            *v2 = V + 3 * (size_t)F[3 * f + 2];
    T p[3], a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        p[k] = rn_div<T>(rn_add<T>(rn_add<T>(v0[k], v1[k]), v2[k]), (T)3);
        a[k] = rn_sub<T>(v1[k], v0[k]);
        b[k] = rn_sub<T>(v2[k], v0[k]);
    }
    T c[3];
    c[0] = rn_sub<T>(rn_mul<T>(a[1], b[2]), rn_mul<T>(a[2], b[1]));
    c[1] = rn_sub<T>(rn_mul<T>(a[2], b[0]), rn_mul<T>(a[0], b[2]));
    c[2] = rn_sub<T>(rn_mul<T>(a[0], b[1]), rn_mul<T>(a[1], b[0]));
    const T nrm = rn_sqrt<T>(rn_add<T>(rn_add<T>(rn_mul<T>(c[0], c[0]), rn_mul<T>(c[1], c[1])),
                                       rn_mul<T>(c[2], c[2])));
    Real4<T> P4{p[0], p[1], p[2], rn_div<T>(nrm, (T)2)};
    Real4<T> N4{rn_div<T>(c[0], nrm), rn_div<T>(c[1], nrm), rn_div<T>(c[2], nrm), (T)0};
    faceP[f] = P4;
    faceN[f] = N4;
}

// float32 vertex buffer handed to the ray tracer (shape.py:319-325)
template <class T>
__global__ void vertices_to_f32_kernel(const T *__restrict__ V, size_t n3, float *__restrict__ V32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) V32[i] = (float)V[i];
}

// ---- K1: triangle boxes, scene bounds, Morton codes ---------------------------
__device__ __forceinline__ unsigned flt_ord(float f) { // order-preserving float -> uint
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_flt(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// scene[0..2] = min centroid, [3..5] = max centroid, [6] = max |coordinate|
__global__ void scene_bounds_kernel(const float *__restrict__ V32, const int *__restrict__ F, int nf,
                                    unsigned *__restrict__ scene) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, mx = 0.f;
    if (f < nf) {
        const float *v0 = V32 + 3 * (size_t)F[3 * f], *v1 = V32 + 3 * (size_t)F[3 * f + 1],
                    *v2 = V32 + 3 * (size_t)F[3 * f + 2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float l = fminf(fminf(v0[k], v1[k]), v2[k]), h = fmaxf(fmaxf(v0[k], v1[k]), v2[k]);
            const float c = 0.5f * (l + h);
            lo[k] = hi[k] = c;
            mx = fmaxf(mx, fmaxf(fabsf(l), fabsf(h)));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&scene[k], flt_ord(lo[k]));
            atomicMax(&scene[3 + k], flt_ord(hi[k]));
        }
        atomicMax(&scene[6], flt_ord(mx));
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) { // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(const float *__restrict__ V32, const int *__restrict__ F, int nf,
                              const unsigned *__restrict__ scene, uint64_t *__restrict__ keys,
                              uint32_t *__restrict__ vals) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const float *v0 = V32 + 3 * (size_t)F[3 * f], *v1 = V32 + 3 * (size_t)F[3 * f + 1],
                *v2 = V32 + 3 * (size_t)F[3 * f + 2];
    uint64_t code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float l = fminf(fminf(v0[k], v1[k]), v2[k]), h = fmaxf(fmaxf(v0[k], v1[k]), v2[k]);
        const double c = 0.5f * (l + h);
        // one scale for all axes (cubical Morton cells): a flat terrain does not
        // waste a third of the bits on its thin axis
        const double slo = ord_flt(scene[k]);
        const double ext = fmax(fmax((double)ord_flt(scene[3]) - ord_flt(scene[0]),
                                     (double)ord_flt(scene[4]) - ord_flt(scene[1])),
                                (double)ord_flt(scene[5]) - ord_flt(scene[2]));
        double u = ext > 0 ? (c - slo) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        const uint64_t q = (uint64_t)fmin(u * 2097152.0, 2097151.0);
        code |= spread21(q) << (2 - k); // x is the most significant of each triple
    }
    keys[f] = code;
    vals[f] = (uint32_t)f;
}

// A triangle whose two edge vectors are (numerically) parallel -- collinear vertices, a repeated vertex, a
// zero-length edge.  The per-triangle Pluecker test of the arithmetic contract has no well-defined plane for
// it: its edge functions are rounding noise and the "hit" distance it reports is unrelated to where the
// triangle is, so no bounding volume is conservative for it.  Such leaves get the whole scene as their box and
// an open slab (every ray reaches them and the exact test decides, as in the contract's brute force), and a
// near zone that holds one has no horizon (horizon.cuh).  sin(angle between the edges) <= 1e-4.
__device__ __forceinline__ bool degenerate_triangle(float e1x, float e1y, float e1z, float e2x, float e2y, float e2z) {
    const float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const float nn = nx * nx + ny * ny + nz * nz;
    const float l1 = e1x * e1x + e1y * e1y + e1z * e1z, l2 = e2x * e2x + e2y * e2y + e2z * e2z;
    return !(nn > 1e-8f * l1 * l2);
}

// ---- K3: Karras (2012) hierarchy ---------------------------------------------
// node ids: internal i in [0, n-2] (root = 0), leaf k -> (n-1) + k
__device__ __forceinline__ int lbvh_delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void karras_kernel(const uint64_t *__restrict__ keys, int n, int *__restrict__ left,
                              int *__restrict__ right, int *__restrict__ parent,
                              int *__restrict__ first, int *__restrict__ last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1) + gamma : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    first[i] = lo;
    last[i] = hi;
    if (i == 0) parent[0] = -1;
}

// leaf boxes (padded) + triangles in leaf order + bottom-up refit of boxes and
// of the area-weighted normal sums (the direction of each node's fitted slab)
__global__ void refit_kernel(const float *__restrict__ V32, const int *__restrict__ F, int n,
                             const uint32_t *__restrict__ leaf_face, const int *__restrict__ left,
                             const int *__restrict__ right, const int *__restrict__ parent,
                             const unsigned *__restrict__ scene, float *__restrict__ box /*9 per node*/,
                             int *__restrict__ flags, float4 *__restrict__ tri,
                             int *__restrict__ face_leaf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int f = (int)leaf_face[k];
    face_leaf[f] = k;
    const float *v0 = V32 + 3 * (size_t)F[3 * f], *v1 = V32 + 3 * (size_t)F[3 * f + 1],
                *v2 = V32 + 3 * (size_t)F[3 * f + 2];
    tri[3 * (size_t)k + 0] = make_float4(v0[0], v0[1], v0[2], __int_as_float(f));
    tri[3 * (size_t)k + 1] = make_float4(v1[0], v1[1], v1[2], 0.f);
    tri[3 * (size_t)k + 2] = make_float4(v2[0], v2[1], v2[2], 0.f);
    // padding: the Pluecker test accepts rays that pass a few ulps (of the
    // largest coordinate) outside a triangle; boxes must not be tighter
    const float pad = 1.0e-6f * ord_flt(scene[6]) + 1e-30f;
    float b[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        b[c] = fminf(fminf(v0[c], v1[c]), v2[c]) - pad;
        b[3 + c] = fmaxf(fmaxf(v0[c], v1[c]), v2[c]) + pad;
    }
    const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    b[6] = e1y * e2z - e1z * e2y; // area normal (2 * area * unit normal)
    b[7] = e1z * e2x - e1x * e2z;
    b[8] = e1x * e2y - e1y * e2x;
    if (degenerate_triangle(e1x, e1y, e1z, e2x, e2y, e2z)) { // never culled: the exact test decides
        const float big = 2.0f * ord_flt(scene[6]) + 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            b[c] = -big;
            b[3 + c] = big;
            b[6 + c] = 0.f;
        }
    }
    int node = (n - 1) + k;
#pragma unroll
    for (int c = 0; c < 9; ++c) box[9 * (size_t)node + c] = b[c];
    if (n == 1) return;
    int p = parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&flags[p], 1) == 0) return; // first child to arrive stops
        const int lc = left[p], rc = right[p];
        volatile const float *bl = box + 9 * (size_t)lc, *br = box + 9 * (size_t)rc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            b[c] = fminf(bl[c], br[c]);
            b[3 + c] = fmaxf(bl[3 + c], br[3 + c]);
            b[6 + c] = bl[6 + c] + br[6 + c];
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) box[9 * (size_t)p + c] = b[c];
        p = parent[p];
    }
}

// unit slab direction of every node from its area-normal sum (faces of either
// orientation count the same way only up to sign -- a folded subtree may cancel
// to ~0, then any fixed direction is as good as another) + slab extent init
__global__ void slab_init_kernel(int nn, float *__restrict__ box, unsigned *__restrict__ slab /*2 per node*/) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nn) return;
    float *b = box + 9 * (size_t)x;
    const float ax = b[6], ay = b[7], az = b[8];
    const float l = sqrtf(ax * ax + ay * ay + az * az);
    if (l > 1e-30f && isfinite(l)) {
        b[6] = ax / l;
        b[7] = ay / l;
        b[8] = az / l;
    } else {
        b[6] = 0.f;
        b[7] = 0.f;
        b[8] = 1.f;
    }
    slab[2 * (size_t)x] = 0xffffffffu; // ordered-uint min
    slab[2 * (size_t)x + 1] = 0u;      // ordered-uint max
}

// every leaf projects its three vertices on the slab direction of itself and of
// each ancestor holding at most `limit` leaves (larger nodes keep an open slab)
__global__ void slab_extent_kernel(int n, const float4 *__restrict__ tri, const int *__restrict__ parent,
                                   const int *__restrict__ first, const int *__restrict__ last,
                                   const float *__restrict__ box, int limit, unsigned *__restrict__ slab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p0 = tri[3 * (size_t)k], p1 = tri[3 * (size_t)k + 1], p2 = tri[3 * (size_t)k + 2];
    int x = (n - 1) + k;
    if (degenerate_triangle(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z, p2.x - p0.x, p2.y - p0.y, p2.z - p0.z)) {
        while (x >= 0) { // open slab for the leaf and every ancestor
            atomicMin(&slab[2 * (size_t)x], flt_ord(-FLT_MAX));
            atomicMax(&slab[2 * (size_t)x + 1], flt_ord(FLT_MAX));
            x = parent[x];
        }
        return;
    }
    while (x >= 0) {
        if (x < n - 1 && last[x] - first[x] + 1 > limit) break; // ancestors only get larger
        const float *b = box + 9 * (size_t)x;
        const float nx = b[6], ny = b[7], nz = b[8];
        const float d0 = nx * p0.x + ny * p0.y + nz * p0.z;
        const float d1 = nx * p1.x + ny * p1.y + nz * p1.z;
        const float d2 = nx * p2.x + ny * p2.y + nz * p2.z;
        atomicMin(&slab[2 * (size_t)x], flt_ord(fminf(fminf(d0, d1), d2)));
        atomicMax(&slab[2 * (size_t)x + 1], flt_ord(fmaxf(fmaxf(d0, d1), d2)));
        x = parent[x];
    }
}

// pre-order index among INTERNAL nodes and "top" flag, by walking to the root
__global__ void preorder_kernel(int n, const int *__restrict__ left, const int *__restrict__ parent,
                                const int *__restrict__ first, const int *__restrict__ last,
                                int top_leaf_threshold, int *__restrict__ pre,
                                int *__restrict__ flag_by_pre, int *__restrict__ max_depth) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = 2 * n - 1;
    if (x >= nn) return;
    int idx = 0, depth = 0, c = x;
    while (true) {
        const int p = parent[c];
        if (p < 0) break;
        const int lc = left[p];
        if (lc == c) idx += 1;
        else idx += 1 + ((lc >= n - 1) ? 0 : last[lc] - first[lc]); // internals of the left sibling
        c = p;
        ++depth;
    }
    atomicMax(max_depth, depth);
    if (x >= n - 1) return; // leaves are not stored as nodes
    pre[x] = idx;
    flag_by_pre[idx] = (last[x] - first[x] + 1) > top_leaf_threshold ? 1 : 0;
}

// 96-byte internal node = two children, each three float4:
//   (lo.x, hi.x, lo.y, hi.y)  (lo.z, hi.z, slab_min, slab_max)  (slab_dir.xyz, bits(ref))
// ref >= 0: internal node index, ref < 0: triangle ~ref (leaf order).
// Order: the ntop internal nodes with the largest subtrees first (pre-order
// among themselves; the prefix staged in shared memory), then the rest in
// pre-order.
__global__ void flatten_kernel(int n, const int *__restrict__ left, const int *__restrict__ right,
                               const int *__restrict__ pre, const int *__restrict__ top_before,
                               const int *__restrict__ flag_by_pre, const int *__restrict__ ntop_p,
                               const float *__restrict__ box, const unsigned *__restrict__ slab,
                               const unsigned *__restrict__ scene, float4 *__restrict__ nodes,
                               int *__restrict__ node_up, int *__restrict__ leaf_up,
                               const int *__restrict__ first, const int *__restrict__ last,
                               int2 *__restrict__ node_range) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n - 1) return;
    const int ntop = *ntop_p;
    auto final_id = [&](int y) {
        const int p = pre[y];
        return flag_by_pre[p] ? top_before[p] : ntop + (p - top_before[p]);
    };
    const float pad = 8.0e-6f * ord_flt(scene[6]) + 1e-30f;
    const int me = final_id(x);
    if (x == 0) node_up[me] = -1; // root
    node_range[me] = make_int2(first[x], last[x]); // leaf positions covered by the node
    const int ch[2] = {left[x], right[x]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int y = ch[c];
        const int ref = (y >= n - 1) ? ~(y - (n - 1)) : final_id(y);
        // "up" link of the child: (parent node << 1) | slot -- lets a warp list the
        // siblings along the root -> source-triangle path (see trace_kernel)
        if (y >= n - 1) leaf_up[y - (n - 1)] = (me << 1) | c;
        else node_up[ref] = (me << 1) | c;
        const float *b = box + 9 * (size_t)y;
        const unsigned umin = slab[2 * (size_t)y], umax = slab[2 * (size_t)y + 1];
        float smin = -INFINITY, smax = INFINITY;
        if (umin <= umax) { // extents were written (node within the slab limit)
            smin = ord_flt(umin) - pad;
            smax = ord_flt(umax) + pad;
        }
        // (lo.x, hi.x, lo.y, hi.y) (lo.z, hi.z, slab_min, slab_max) (slab_dir.xyz | ref): see trace.cuh child_hit
        nodes[6 * (size_t)me + 3 * c + 0] = make_float4(b[0], b[3], b[1], b[4]);
        nodes[6 * (size_t)me + 3 * c + 1] = make_float4(b[2], b[5], smin, smax);
        nodes[6 * (size_t)me + 3 * c + 2] = make_float4(b[6], b[7], b[8], __int_as_float(ref));
    }
}

} // namespace fluxb200
