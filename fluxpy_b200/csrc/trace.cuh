// trace.cuh -- ray set-up, Pluecker triangle test, box + fitted-slab test, generic BVH traversal.
//
// Device restatement of the semantics of EmbreeTrimeshShapeModel._get_visibility
// / _is_occluded (reference src/flux/shape.py:349-421) on Embree's robust-mode
// closest-hit kernels.  The arithmetic of the ray set-up and of the triangle
// test is pinned operation by operation (the "arithmetic contract"):
// every multiply/add goes through a round-to-nearest intrinsic so nvcc cannot
// contract or reassociate it, fused operations appear only as explicit FMAs.
// The box / slab test is NOT part of the contract: it only has to be conservative.
#pragma once
#include "common.cuh"
#include "lbvh.cuh"

namespace fluxb200 {

struct Ray {
    float ox, oy, oz, dx, dy, dz;
};

// 1e3*np.finfo(np.float32).resolution as NumPy 2 evaluates it: float32 0x3A83126F
__device__ __forceinline__ float ray_eps() { return __uint_as_float(0x3A83126Fu); }

// shape.py:357-362, 380 in float32.  false: pair masked out ("vis by default").
__device__ __forceinline__ bool setup_ray(const Real4<float> &Pi, const Real4<float> &Pj, Ray &r) {
    const float eps = ray_eps();
    const float dx = __fsub_rn(Pj.x, Pi.x), dy = __fsub_rn(Pj.y, Pi.y), dz = __fsub_rn(Pj.z, Pi.z);
    const float nrm = __fsqrt_rn(
        __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    if (!(nrm > eps)) return false;
    r.dx = __fdiv_rn(dx, nrm);
    r.dy = __fdiv_rn(dy, nrm);
    r.dz = __fdiv_rn(dz, nrm);
    r.ox = __fadd_rn(Pi.x, __fmul_rn(eps, r.dx));
    r.oy = __fadd_rn(Pi.y, __fmul_rn(eps, r.dy));
    r.oz = __fadd_rn(Pi.z, __fmul_rn(eps, r.dz));
    return true;
}

// same in float64, rounded to the float32 ray buffers at the end
__device__ __forceinline__ bool setup_ray(const Real4<double> &Pi, const Real4<double> &Pj, Ray &r) {
    const double eps = (double)ray_eps();
    const double dx = __dsub_rn(Pj.x, Pi.x), dy = __dsub_rn(Pj.y, Pi.y), dz = __dsub_rn(Pj.z, Pi.z);
    const double nrm = __dsqrt_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    if (!(nrm > eps)) return false;
    const double Dx = __ddiv_rn(dx, nrm), Dy = __ddiv_rn(dy, nrm), Dz = __ddiv_rn(dz, nrm);
    r.dx = __double2float_rn(Dx);
    r.dy = __double2float_rn(Dy);
    r.dz = __double2float_rn(Dz);
    r.ox = __double2float_rn(__dadd_rn(Pi.x, __dmul_rn(eps, Dx)));
    r.oy = __double2float_rn(__dadd_rn(Pi.y, __dmul_rn(eps, Dy)));
    r.oz = __double2float_rn(__dadd_rn(Pi.z, __dmul_rn(eps, Dz)));
    return true;
}

__device__ __forceinline__ float c_msub(float a, float b, float c) { return __fmaf_rn(a, b, -c); }
__device__ __forceinline__ float c_dot(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}

// Embree robust-mode (Pluecker) ray/triangle test; true and t when the ray hits
// with tnear <= t <= tfar.
__device__ __forceinline__ bool pluecker_hit(const Ray &r, float tnear, float tfar, const float4 &p0,
                                             const float4 &p1, const float4 &p2, float &t_out) {
    const float v0x = __fsub_rn(p0.x, r.ox), v0y = __fsub_rn(p0.y, r.oy), v0z = __fsub_rn(p0.z, r.oz);
    const float v1x = __fsub_rn(p1.x, r.ox), v1y = __fsub_rn(p1.y, r.oy), v1z = __fsub_rn(p1.z, r.oz);
    const float v2x = __fsub_rn(p2.x, r.ox), v2y = __fsub_rn(p2.y, r.oy), v2z = __fsub_rn(p2.z, r.oz);
    const float e0x = __fsub_rn(v2x, v0x), e0y = __fsub_rn(v2y, v0y), e0z = __fsub_rn(v2z, v0z);
    const float e1x = __fsub_rn(v0x, v1x), e1y = __fsub_rn(v0y, v1y), e1z = __fsub_rn(v0z, v1z);
    const float e2x = __fsub_rn(v1x, v2x), e2y = __fsub_rn(v1y, v2y), e2z = __fsub_rn(v1z, v2z);
    float sx, sy, sz, cx, cy, cz;
    // U = dot(cross(e0, v2+v0), D)
    sx = __fadd_rn(v2x, v0x); sy = __fadd_rn(v2y, v0y); sz = __fadd_rn(v2z, v0z);
    cx = c_msub(e0y, sz, __fmul_rn(e0z, sy));
    cy = c_msub(e0z, sx, __fmul_rn(e0x, sz));
    cz = c_msub(e0x, sy, __fmul_rn(e0y, sx));
    const float U = c_dot(cx, cy, cz, r.dx, r.dy, r.dz);
    sx = __fadd_rn(v0x, v1x); sy = __fadd_rn(v0y, v1y); sz = __fadd_rn(v0z, v1z);
    cx = c_msub(e1y, sz, __fmul_rn(e1z, sy));
    cy = c_msub(e1z, sx, __fmul_rn(e1x, sz));
    cz = c_msub(e1x, sy, __fmul_rn(e1y, sx));
    const float V = c_dot(cx, cy, cz, r.dx, r.dy, r.dz);
    sx = __fadd_rn(v1x, v2x); sy = __fadd_rn(v1y, v2y); sz = __fadd_rn(v1z, v2z);
    cx = c_msub(e2y, sz, __fmul_rn(e2z, sy));
    cy = c_msub(e2z, sx, __fmul_rn(e2x, sz));
    cz = c_msub(e2x, sy, __fmul_rn(e2y, sx));
    const float W = c_dot(cx, cy, cz, r.dx, r.dy, r.dz);
    const float UVW = __fadd_rn(__fadd_rn(U, V), W);
    const float eps = __fmul_rn(1.1920929e-7f, fabsf(UVW));
    const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
    if (!(mn >= -eps || mx <= eps)) return false;
    // stable triangle normal
    const float ab_x = __fmul_rn(e0z, e1y), ab_y = __fmul_rn(e0x, e1z), ab_z = __fmul_rn(e0y, e1x);
    const float bc_x = __fmul_rn(e1z, e2y), bc_y = __fmul_rn(e1x, e2z), bc_z = __fmul_rn(e1y, e2x);
    const float cabx = c_msub(e0y, e1z, ab_x), caby = c_msub(e0z, e1x, ab_y), cabz = c_msub(e0x, e1y, ab_z);
    const float cbcx = c_msub(e1y, e2z, bc_x), cbcy = c_msub(e1z, e2x, bc_y), cbcz = c_msub(e1x, e2y, bc_z);
    const float Ngx = fabsf(ab_x) < fabsf(bc_x) ? cabx : cbcx;
    const float Ngy = fabsf(ab_y) < fabsf(bc_y) ? caby : cbcy;
    const float Ngz = fabsf(ab_z) < fabsf(bc_z) ? cabz : cbcz;
    const float dn = c_dot(Ngx, Ngy, Ngz, r.dx, r.dy, r.dz);
    const float den = __fadd_rn(dn, dn);
    const float Tn = c_dot(v0x, v0y, v0z, Ngx, Ngy, Ngz);
    const float T = __fadd_rn(Tn, Tn);
    if (den == 0.0f) return false;
    const float t = __fdiv_rn(T, den);
    if (!(tnear <= t && t <= tfar)) return false;
    t_out = t;
    return true;
}

// ---------------------------------------------------------------------------
// BVH traversal.  NOT part of the arithmetic contract: it only has to be
// conservative (never skip a triangle the Pluecker test would accept); which
// triangles get tested is the only thing it decides.
// ---------------------------------------------------------------------------
constexpr int kStackDepth = 64;

struct BvhView {
    const float4 *nodes;  // global, 6 float4 per internal node (see lbvh.cuh flatten_kernel)
    const float4 *top;    // shared-memory copy of the first ntop nodes (or nullptr)
    const float4 *tri;    // 3 float4 per triangle, leaf order; tri[3k].w = face id bits
    int ntop;
    int ninternal;        // number of internal nodes (0: mesh with < 2 faces)
    int nfaces;
    int *error_flag;      // set to 1 on traversal stack overflow
};

// ---- packed FP32 pairs (sm_100: FFMA2 / FADD2 / FMUL2, one issue slot for two IEEE operations; a scalar
// operand is broadcast by the hardware when both halves of a pair are the same register) -----------------
#ifndef FB_EMU
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
    f2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
    f2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
#else // SIMT-emulator build (tools/simt): two plain floats
struct f2_t { float lo, hi; };
inline f2_t f2_pack(float lo, float hi) { return f2_t{lo, hi}; }
inline void f2_unpack(f2_t v, float &lo, float &hi) { lo = v.lo; hi = v.hi; }
inline f2_t f2_fma(f2_t a, f2_t b, f2_t c) { return f2_t{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
inline f2_t f2_mul(f2_t a, f2_t b) { return f2_t{a.lo * b.lo, a.hi * b.hi}; }
inline f2_t f2_add(f2_t a, f2_t b) { return f2_t{a.lo + b.lo, a.hi + b.hi}; }
#endif
__device__ __forceinline__ f2_t f2_both(float x) { return f2_pack(x, x); }

// per-ray constants of the box / slab tests
struct RayBox {
    float ix, iy, iz;    // 1/d, |d| clamped away from zero
    float ox, oy, oz;    // o/d
    f2_t xod, yod, zod;  // (o.x, d.x) (o.y, d.y) (o.z, d.z): the two dot products of the slab test run as pairs
};
__device__ __forceinline__ RayBox make_raybox(const Ray &r) {
    auto inv = [](float d) {
        const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
        return 1.0f / a;
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

// A child record is three float4, laid out so that every pair the test works on is an aligned register pair:
//   a = (lo.x, hi.x, lo.y, hi.y)   b = (lo.z, hi.z, slab_min, slab_max)   c = (slab_dir.xyz | ref)
// ref >= 0: internal node index, ref < 0: triangle ~ref (leaf order).
__device__ __forceinline__ int rec_ref(const float4 &c) { return __float_as_int(c.w); }

// AABB slab test on [0, tmax] intersected with the fitted-slab interval smin <= n.(o + t d) <= smax.  Boxes
// and slab extents are padded at build time; the final comparison carries one more relative guard band.
// kPacked: 16 of the FMA-pipe operations run as 8 packed instructions; the scalar form performs the same IEEE
// operations in the same order, so both give the same answer bit for bit (which one is faster depends on the
// kernel around it: profiles/r02_summary.md section 1).
#ifndef FB_PACKED
#define FB_PACKED 1
#endif
template <int kPacked = FB_PACKED>
__device__ __forceinline__ bool child_hit(const Ray &r, const RayBox &rb, const float4 &a, const float4 &b,
                                          const float4 &c, float tmax) {
    float no, nd, s0, s1, x0, x1, y0, y1, z0, z1;
    if (kPacked) {
        // (n.o, n.d) as a pair: c.z * (oz, dz), then + c.y * (oy, dy), then + c.x * (ox, dx)
        f2_t nod = f2_mul(f2_both(c.z), rb.zod);
        nod = f2_fma(f2_both(c.y), rb.yod, nod);
        nod = f2_fma(f2_both(c.x), rb.xod, nod);
        f2_unpack(nod, no, nd);
    } else {
        no = fmaf(c.x, r.ox, fmaf(c.y, r.oy, c.z * r.oz));
        nd = fmaf(c.x, r.dx, fmaf(c.y, r.dy, c.z * r.dz));
    }
    float rn; // approximate reciprocal (MUFU.RCP): +-inf when the ray runs parallel to the slab
#ifndef FB_EMU
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(nd));
#else
    rn = 1.0f / nd; // SIMT-emulator build (tools/simt); the box / slab test only has to be conservative
#endif
    if (kPacked) { // ((smin, smax) - no) * rn
        f2_unpack(f2_mul(f2_add(f2_pack(b.z, b.w), f2_both(-no)), f2_both(rn)), s0, s1);
    } else {
        s0 = (b.z - no) * rn;
        s1 = (b.w - no) * rn;
    }
    // parallel ray: (smin-no), (smax-no) of opposite sign -> (-inf, +inf), no clipping;
    // same sign -> both +inf or both -inf -> empty.  NaN (0*inf) is dropped by fmin/fmax.
    float tn = fmaxf(fminf(s0, s1), 0.0f);
    float tf = fminf(fmaxf(s0, s1), tmax);
    if (kPacked) {
        f2_unpack(f2_fma(f2_pack(a.x, a.y), f2_both(rb.ix), f2_both(-rb.ox)), x0, x1);
        f2_unpack(f2_fma(f2_pack(a.z, a.w), f2_both(rb.iy), f2_both(-rb.oy)), y0, y1);
        f2_unpack(f2_fma(f2_pack(b.x, b.y), f2_both(rb.iz), f2_both(-rb.oz)), z0, z1);
    } else {
        x0 = fmaf(a.x, rb.ix, -rb.ox), x1 = fmaf(a.y, rb.ix, -rb.ox);
        y0 = fmaf(a.z, rb.iy, -rb.oy), y1 = fmaf(a.w, rb.iy, -rb.oy);
        z0 = fmaf(b.x, rb.iz, -rb.oz), z1 = fmaf(b.y, rb.iz, -rb.oz);
    }
    tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tn));
    tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tf));
    return tn <= tf * 1.000002f;
}

template <bool kTop = true>
__device__ __forceinline__ void load_node(const BvhView &bvh, int node, float4 (&q)[6]) {
    if (kTop && node < bvh.ntop) { // staged in shared memory
        const float4 *p = bvh.top + 6 * node;
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = p[k];
    } else {
        const float4 *p = bvh.nodes + 6 * (size_t)node;
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = __ldg(p + k);
    }
}

// exact test of one candidate triangle (leaf position `leaf`) for the occlusion
// question "hit with 0 <= t and (t < tlimit, or t == tlimit and face > target_face)"
__device__ __forceinline__ bool leaf_occludes(const BvhView &bvh, const Ray &r, float tlimit, int leaf,
                                              int target_face) {
    const float4 p0 = __ldg(bvh.tri + 3 * (size_t)leaf);
    const float4 p1 = __ldg(bvh.tri + 3 * (size_t)leaf + 1);
    const float4 p2 = __ldg(bvh.tri + 3 * (size_t)leaf + 2);
    float t;
    if (!pluecker_hit(r, 0.0f, tlimit, p0, p1, p2, t)) return false;
    return t < tlimit || __float_as_int(p0.w) > target_face;
}

// out-of-line copy for the rare lane whose deferred-candidate list is full (K4)
__device__ __noinline__ bool leaf_occludes_cold(const float4 *tri, Ray r, float tlimit, int leaf,
                                                int target_face) {
    const float4 p0 = __ldg(tri + 3 * (size_t)leaf);
    const float4 p1 = __ldg(tri + 3 * (size_t)leaf + 1);
    const float4 p2 = __ldg(tri + 3 * (size_t)leaf + 2);
    float t;
    if (!pluecker_hit(r, 0.0f, tlimit, p0, p1, p2, t)) return false;
    return t < tlimit || __float_as_int(p0.w) > target_face;
}

// r-th (0-based) set bit of w (w has more than r set bits): five popcount halvings, no loop
__device__ __forceinline__ int nth_set_bit(uint32_t w, int r) {
    int pos = 0;
    int c = __popc(w & 0xffffu);
    if (r >= c) { r -= c; pos = 16; }
    c = __popc((w >> pos) & 0xffu);
    if (r >= c) { r -= c; pos += 8; }
    c = __popc((w >> pos) & 0xfu);
    if (r >= c) { r -= c; pos += 4; }
    c = __popc((w >> pos) & 0x3u);
    if (r >= c) { r -= c; pos += 2; }
    c = (int)((w >> pos) & 1u);
    if (r >= c) pos += 1;
    return pos;
}

// Generic any-hit traversal with immediate leaf tests (query kernels).
// With tlimit = t of the target triangle this is "closest hit != target" of the
// oracle's index-ordered closest-hit definition.  tlimit = +inf, target_leaf =
// -1: plain occlusion query.
// `start`: the internal node the walk begins at (0 = root: the whole tree)
__device__ __forceinline__ bool occluded_anyhit(const BvhView &bvh, const Ray &r, float tlimit,
                                                int target_leaf, int target_face, int start = 0) {
    if (bvh.nfaces == 0) return false;
    if (bvh.ninternal == 0) // single triangle
        return target_leaf != 0 && leaf_occludes(bvh, r, tlimit, 0, target_face);
    const RayBox rb = make_raybox(r);
    const float tmax = tlimit * 1.000002f;
    int stack[kStackDepth];
    int sp = 0, node = start;
    while (true) {
        float4 q[6];
        load_node(bvh, node, q);
        int next = -1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (child_hit(r, rb, q[3 * c], q[3 * c + 1], q[3 * c + 2], tmax)) {
                const int ref = rec_ref(q[3 * c + 2]);
                if (ref < 0) {
                    const int leaf = ~ref;
                    if (leaf != target_leaf && leaf_occludes(bvh, r, tlimit, leaf, target_face)) return true;
                } else if (next < 0) {
                    next = ref;
                } else if (sp < kStackDepth) {
                    stack[sp++] = ref;
                } else {
                    *bvh.error_flag = 1;
                }
            }
        }
        if (next < 0) {
            if (sp == 0) return false;
            next = stack[--sp];
        }
        node = next;
    }
}

// visibility of target triangle (leaf position tleaf, face id tface) along ray r
__device__ __forceinline__ bool target_hit_t(const BvhView &bvh, const Ray &r, int tleaf, float &tj) {
    const float4 p0 = __ldg(bvh.tri + 3 * (size_t)tleaf);
    const float4 p1 = __ldg(bvh.tri + 3 * (size_t)tleaf + 1);
    const float4 p2 = __ldg(bvh.tri + 3 * (size_t)tleaf + 2);
    return pluecker_hit(r, 0.0f, __int_as_float(0x7f800000), p0, p1, p2, tj);
}

__device__ __forceinline__ bool target_visible(const BvhView &bvh, const Ray &r, int tleaf, int tface) {
    float tj;
    if (!target_hit_t(bvh, r, tleaf, tj)) return false;
    return !occluded_anyhit(bvh, r, tj, tleaf, tface);
}

// the same definitions without the tree (test hook)
__device__ __forceinline__ bool target_visible_bruteforce(const float4 *tri, int nf, const Ray &r,
                                                          int tleaf, int tface) {
    float tj;
    if (!pluecker_hit(r, 0.0f, __int_as_float(0x7f800000), __ldg(tri + 3 * (size_t)tleaf),
                      __ldg(tri + 3 * (size_t)tleaf + 1), __ldg(tri + 3 * (size_t)tleaf + 2), tj))
        return false;
    for (int k = 0; k < nf; ++k) {
        if (k == tleaf) continue;
        const float4 p0 = __ldg(tri + 3 * (size_t)k);
        float t;
        if (pluecker_hit(r, 0.0f, tj, p0, __ldg(tri + 3 * (size_t)k + 1), __ldg(tri + 3 * (size_t)k + 2), t))
            if (t < tj || __float_as_int(p0.w) > tface) return false;
    }
    return true;
}

// closest hit (index-ordered definition) for intersect1
__device__ __forceinline__ bool closest_hit(const BvhView &bvh, const Ray &r, float &t_best, int &face) {
    face = -1;
    if (bvh.nfaces == 0) return false;
    auto try_leaf = [&](int leaf) {
        const float4 p0 = __ldg(bvh.tri + 3 * (size_t)leaf);
        float t;
        if (pluecker_hit(r, 0.0f, t_best, p0, __ldg(bvh.tri + 3 * (size_t)leaf + 1),
                         __ldg(bvh.tri + 3 * (size_t)leaf + 2), t)) {
            const int f = __float_as_int(p0.w);
            if (t < t_best || face < 0 || f > face) {
                t_best = t;
                face = f;
            }
        }
    };
    if (bvh.ninternal == 0) {
        try_leaf(0);
        return face >= 0;
    }
    const RayBox rb = make_raybox(r);
    int stack[kStackDepth];
    int sp = 0, node = 0;
    while (true) {
        float4 q[6];
        load_node(bvh, node, q);
        int next = -1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (child_hit(r, rb, q[3 * c], q[3 * c + 1], q[3 * c + 2], t_best * 1.000002f)) {
                const int ref = rec_ref(q[3 * c + 2]);
                if (ref < 0) try_leaf(~ref);
                else if (next < 0) next = ref;
                else if (sp < kStackDepth) stack[sp++] = ref;
                else *bvh.error_flag = 1;
            }
        }
        if (next < 0) {
            if (sp == 0) break;
            next = stack[--sp];
        }
        node = next;
    }
    return face >= 0;
}

} // namespace fluxb200
