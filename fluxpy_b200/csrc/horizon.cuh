// horizon.cuh -- per-face horizons of the near zone ("horizon skip" of the trace kernel, K4).
//
// Centroid-to-centroid rays leave and arrive above the local relief, yet every ray tests the small
// subtrees next to its source triangle and walks the lowest levels above its target leaf
// (profiles/r01b_k4_model.md: ~8 records and ~7 levels per ray that are practically never entered).
// That can be decided once per face instead of once per record:
//
//   zone(f)  = the largest ancestor of leaf(f) holding at most Z leaves;
//   hor[f]   = sup over every point x of every OTHER triangle of zone(f) of  n_f.(x - p_f)/|x - p_f|
//              (for a unit n_f: the sine of the elevation of x above the face's tangent plane), plus the
//              perturbation a float32 ray can have against the ideal one;
//   rmin[f]  = the smallest distance from p_f to another triangle of zone(f).
//
// A ray p_i -> p_j with unit direction d can only meet a triangle g of zone(i) at a point x whose direction
// from p_i IS d, so  n_i.d > hor[i]  excludes every triangle of zone(i) except i itself; the same with
// -d at the target end for the part of the ray before p_j, and  (how far the tested interval runs on past
// p_j) < rmin[j]  for the part after it.  Both are statements about the geometry, not approximations, and
// they hold for ANY vector n_f (the shape model's normals are user-mutable, reference
// src/flux/shape.py:55-112): hor is recomputed from the current P, N whenever they change.
//
// The exact supremum over a triangle: the three vertices, the stationary point of every edge (a linear
// equation), +inf when the line p_f + t n_f (t > 0) pierces the triangle.  Checked by brute force on the
// CPU (tools/k4_horizon_check.py: 0 violations in 1.4e9 Pluecker tests).
#pragma once
#include "common.cuh"
#include "lbvh.cuh"

namespace fluxb200 {

// zone_node[k]: flattened id of the largest ancestor of leaf k with at most zone_leaves leaves (-1: none,
// the zone is the leaf alone); zone_up[k]: the (parent << 1 | slot) code of that node (of the leaf if none)
__global__ void zone_kernel(int n, const int *__restrict__ leaf_up, const int *__restrict__ node_up,
                            const int2 *__restrict__ node_range, int zone_leaves, int *__restrict__ zone_node,
                            int *__restrict__ zone_up) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int code = leaf_up[k], zn = -1, zup = code;
    while (code >= 0) {
        const int p = code >> 1;
        const int2 r = node_range[p];
        if (r.y - r.x + 1 > zone_leaves) break;
        zn = p;
        zup = node_up[p];
        code = zup;
    }
    zone_node[k] = zn;
    zone_up[k] = zup;
}

// sup of n.(x - p)/|x - p| over the triangle (a, b, c) and the smallest |x - p|, in double
__device__ __forceinline__ void triangle_elevation(const double p[3], const double n[3], const float4 &a,
                                                   const float4 &b, const float4 &c, double &sup, double &rmin) {
    const double q[3][3] = {{(double)a.x - p[0], (double)a.y - p[1], (double)a.z - p[2]},
                            {(double)b.x - p[0], (double)b.y - p[1], (double)b.z - p[2]},
                            {(double)c.x - p[0], (double)c.y - p[1], (double)c.z - p[2]}};
    double best = -INFINITY, rm = INFINITY;
    auto take = [&](double e) { best = (e > best || e != e) ? e : best; }; // NaN sticks: such a face never skips
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double *A = q[k], *B = q[(k + 1) % 3];
        const double gamma = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
        const double la = sqrt(gamma);
        rm = fmin(rm, la);
        const double alpha = n[0] * A[0] + n[1] * A[1] + n[2] * A[2];
        if (la > 0.0) take(alpha / la);
        const double E[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
        // along the edge: f(s) = (alpha + beta s) / sqrt(gamma + 2 delta s + eps s^2); f' = 0 is linear in s
        const double beta = n[0] * E[0] + n[1] * E[1] + n[2] * E[2];
        const double delta = A[0] * E[0] + A[1] * E[1] + A[2] * E[2];
        const double eps = E[0] * E[0] + E[1] * E[1] + E[2] * E[2];
        const double den = beta * delta - alpha * eps;
        if (den != 0.0) {
            const double s = (alpha * delta - beta * gamma) / den;
            if (s > 0.0 && s < 1.0) {
                const double l2 = gamma + 2.0 * delta * s + eps * s * s;
                if (l2 > 0.0) take((alpha + beta * s) / sqrt(l2));
            }
        }
        if (eps > 0.0) { // closest point of the edge
            const double s = fmin(fmax(-delta / eps, 0.0), 1.0);
            rm = fmin(rm, sqrt(fmax(gamma + 2.0 * delta * s + eps * s * s, 0.0)));
        }
    }
    // interior: the half-line p + t n pierces the triangle (the functional is maximal there), and the
    // foot of p on the triangle's plane (closest point)
    const double *A = q[0];
    const double E1[3] = {q[1][0] - A[0], q[1][1] - A[1], q[1][2] - A[2]};
    const double E2[3] = {q[2][0] - A[0], q[2][1] - A[1], q[2][2] - A[2]};
    const double m[3] = {E1[1] * E2[2] - E1[2] * E2[1], E1[2] * E2[0] - E1[0] * E2[2], E1[0] * E2[1] - E1[1] * E2[0]};
    const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    if (degenerate_triangle((float)E1[0], (float)E1[1], (float)E1[2], (float)E2[0], (float)E2[1], (float)E2[2]))
        best = INFINITY; // no well-defined plane: the exact test's hits on it obey no geometry (lbvh.cuh)
    if (mm > 0.0) {
        const double mA = m[0] * A[0] + m[1] * A[1] + m[2] * A[2];
        const double d11 = E1[0] * E1[0] + E1[1] * E1[1] + E1[2] * E1[2];
        const double d12 = E1[0] * E2[0] + E1[1] * E2[1] + E1[2] * E2[2];
        const double d22 = E2[0] * E2[0] + E2[1] * E2[1] + E2[2] * E2[2];
        const double det = d11 * d22 - d12 * d12;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const double *dir = pass ? m : n;
            const double md = m[0] * dir[0] + m[1] * dir[1] + m[2] * dir[2];
            if (md == 0.0 || !(det > 0.0)) continue;
            const double t = mA / md;
            if (!pass && !(t > 0.0)) continue;
            const double X[3] = {t * dir[0] - A[0], t * dir[1] - A[1], t * dir[2] - A[2]};
            const double x1 = X[0] * E1[0] + X[1] * E1[1] + X[2] * E1[2];
            const double x2 = X[0] * E2[0] + X[1] * E2[1] + X[2] * E2[2];
            const double u = (x1 * d22 - x2 * d12) / det, w = (x2 * d11 - x1 * d12) / det;
            if (u >= -1e-9 && w >= -1e-9 && u + w <= 1.0 + 1e-9) {
                if (!pass) best = INFINITY;
                else rm = fmin(rm, fabs(t) * sqrt(mm));
            }
        }
    }
    sup = best;
    rmin = rm;
}

// One warp per face.  hz[f] = (hor, rmin).  pert = the displacement (in length units) a float32 ray can have
// against the ideal one plus the slack of the Pluecker edge tests.  If the traced ray meets triangle g, a point
// x of g lies within pert of a point y of the ideal ray; the unit directions of x and y seen from p_f differ
// by at most 2 pert / max(|x - p_f|, |y - p_f|) <= 2 pert / rmin(g), at either end of the ray.  So a ray
// whose n_f.d exceeds  sup_g + 2 pert |n_f| / rmin(g)  for every g of the zone meets none of them
// (checked by brute force: tools/k4_horizon_check.py, formula 1).
template <class T>
__global__ void __launch_bounds__(256)
    horizon_kernel(int nf, const Real4<T> *__restrict__ faceP, const Real4<T> *__restrict__ faceN,
                   const int *__restrict__ face_leaf, const int *__restrict__ zone_node,
                   const int2 *__restrict__ node_range, const float4 *__restrict__ tri, float pert,
                   float2 *__restrict__ hz) {
    const int lane = threadIdx.x & 31;
    const int f = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= nf) return; // warp-uniform
    const int leaf = face_leaf[f];
    const int zn = zone_node[leaf];
    int lo = leaf, hi = leaf;
    if (zn >= 0) {
        const int2 r = node_range[zn];
        lo = r.x;
        hi = r.y;
    }
    const double p[3] = {(double)faceP[f].x, (double)faceP[f].y, (double)faceP[f].z};
    const double n[3] = {(double)faceN[f].x, (double)faceN[f].y, (double)faceN[f].z};
    const double nlen = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double best = -INFINITY, rzone = INFINITY;
    for (int k = lo + lane; k <= hi; k += 32) {
        if (k == leaf) continue;
        double sup, rm;
        triangle_elevation(p, n, __ldg(tri + 3 * (size_t)k), __ldg(tri + 3 * (size_t)k + 1),
                           __ldg(tri + 3 * (size_t)k + 2), sup, rm);
        const double e = rm > 0.0 ? sup + 2.0 * (double)pert * nlen / rm : INFINITY;
        best = (e > best || e != e) ? e : best;
        rzone = fmin(rzone, rm);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double e = __shfl_xor_sync(0xffffffffu, best, o);
        best = (e > best || e != e) ? e : best;
        rzone = fmin(rzone, __shfl_xor_sync(0xffffffffu, rzone, o));
    }
    if (lane == 0) {
        // the kernel evaluates n.d in float32: a margin of a few float32 ulps of |n|, and round up / down
        best += 1.0e-5 * nlen;
        float h = (float)best, r = (float)rzone;
        if ((double)h < best) h = __uint_as_float(__float_as_uint(h) + (h >= 0.f ? 1u : 0xffffffffu));
        if ((double)r > rzone) r = __uint_as_float(__float_as_uint(r) - 1u); // r > 0 here
        if (!(nlen < INFINITY)) h = INFINITY;                                 // NaN / inf normal: never skip
        hz[f] = make_float2(h, r);
    }
}

// per column (leaf order): (hor, rmin, zone node, up code of the zone node)
__global__ void col_horizon_kernel(const int *__restrict__ col_face, const int *__restrict__ col_leaf, int n,
                                   const float2 *__restrict__ hz, const int *__restrict__ zone_node,
                                   const int *__restrict__ zone_up, float4 *__restrict__ colH) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float2 h = hz[col_face[s]];
    const int leaf = col_leaf[s];
    colH[s] = make_float4(h.x, h.y, __int_as_float(zone_node[leaf]), __int_as_float(zone_up[leaf]));
}

} // namespace fluxb200
