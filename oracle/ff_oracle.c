/*
 * ff_oracle.c -- CPU ORACLE for the fluxpy form-factor assembly hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The shipped path (fluxpy_b200/, libfluxb200.so) never
 * links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm for the path
 *
 *     flux.form_factors.get_form_factor_matrix       src/flux/form_factors.py:11-72
 *       -> TrimeshShapeModel.get_visibility_1_to_N   src/flux/shape.py:165-170
 *         -> EmbreeTrimeshShapeModel._get_visibility src/flux/shape.py:349-398
 *           -> embree.Scene.intersect1M              (third party, NOT in /root/reference)
 *     EmbreeTrimeshShapeModel._is_occluded           src/flux/shape.py:400-421
 *
 * Third-party dependency holding the ray/triangle arithmetic: Embree 3.13.2
 * (.travis.yml:20) through sampotter/python-embree (unpinned HEAD,
 * .travis.yml:25).  Neither is vendored nor installable here (no network), so
 * the closest-hit kernel below RESTATES Embree's published robust-mode
 * algorithm (scene flag RTC_SCENE_FLAG_ROBUST, src/flux/shape.py:316 ->
 * Pluecker-coordinate triangle test of kernels/geometry/
 * triangle_intersector_pluecker.h, "stable" geometric normal, depth
 * t = T/den, accept tnear <= t <= tfar, no back-face culling).
 *
 * PARITY STATUS.  The numerical (F_ij) half and the ray set-up half are pinned
 * against the reference's own Python, executed in the build container by
 * oracle/make_golden.py (golden files under tests/golden/).  The ray/triangle
 * half is "parity unpinned" against real Embree binaries: it is anchored only
 * on the reference's own structural tests (tests/test_shape.py:16-112,
 * tests/test_form_factors.py:18-57) and on brute-force self-consistency.
 *
 * THE ARITHMETIC CONTRACT.  Every floating-point operation below is written
 * out with explicit rounding points (compile with -ffp-contract=off; fused
 * operations appear only as explicit fmaf()/fma() calls).  The CUDA product
 * path is an independent implementation of the same contract, so results are
 * comparable bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_INVALID_ID 0xFFFFFFFFu /* embree.INVALID_GEOMETRY_ID */

/* eps of src/flux/shape.py:351,402: 1e3*np.finfo(np.float32).resolution.  Under
 * NumPy 2 this is the float32 scalar 0x3A83126F = 0.0010000000474974513. */
static float ray_eps_f32(void) {
    union { uint32_t u; float f; } c;
    c.u = 0x3A83126Fu;
    return c.f;
}
float oracle_ray_eps(void) { return ray_eps_f32(); }

/* ------------------------------------------------------------------------ */
/* Scene: float32 vertex buffer + uint32 index buffer (shape.py:319-333)     */
/* ------------------------------------------------------------------------ */

typedef struct {
    float lo[3], hi[3];
    int left, right; /* children (internal) */
    int first, count; /* triangle range into perm (leaf when count > 0) */
} bvh_node;

typedef struct oracle_scene {
    size_t nf;
    float *tri;    /* nf * 9: p0 p1 p2 of every face, float32 */
    /* median-split BVH, used only to make the oracle finish in seconds; the
     * brute-force mode (use_bvh = 0) is the definition of the result */
    bvh_node *nodes;
    int nnodes;
    uint32_t *perm;
} oracle_scene;

/* --- the Pluecker test ---------------------------------------------------- */

/* STUDY VARIANTS (tests/test_oracle_golden.py::test_embree_arithmetic_variants_*): the contract -- what the CUDA
 * path is compared with bit for bit -- is variant (0, 0).  Real Embree differs from it in two places that no
 * reference test pins: the hit distance is T * rcp(den) with rcp = the hardware's 12-bit reciprocal estimate plus
 * one Newton step (embree3 common/math/math.h, common/simd/vfloat4_sse2.h), not an exact division; and among
 * equal-t hits the winner depends on Embree's BVH order, not on the face index.  These switches let a test bound
 * what either can change: g_tmode 1 = r*(2 - r*den) (SSE2 build), 2 = r + r*(1 - den*r) with FMA (AVX2 build);
 * g_tie 1 = among equal t the SMALLEST index wins. */
#include <immintrin.h>
static int g_tmode = 0, g_tie = 0;
void oracle_set_study_variant(int tmode, int tie) {
    g_tmode = tmode;
    g_tie = tie;
}
static inline float hit_distance(float T, float den) {
    if (g_tmode == 0) return T / den;
    const __m128 a = _mm_set_ss(den);
    const __m128 r = _mm_rcp_ss(a);
    float rcp;
    if (g_tmode == 1) rcp = _mm_cvtss_f32(_mm_mul_ss(r, _mm_sub_ss(_mm_set_ss(2.0f), _mm_mul_ss(r, a))));
    else rcp = _mm_cvtss_f32(_mm_add_ss(r, _mm_mul_ss(r, _mm_fnmadd_ss(a, r, _mm_set_ss(1.0f)))));
    return rcp * T;
}

static inline float msubf(float a, float b, float c) { return fmaf(a, b, -c); } /* a*b - c */
static inline float dot3f(const float a[3], const float b[3]) {
    return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2]));
}
static inline void cross3f(const float a[3], const float b[3], float r[3]) {
    r[0] = msubf(a[1], b[2], a[2] * b[1]);
    r[1] = msubf(a[2], b[0], a[0] * b[2]);
    r[2] = msubf(a[0], b[1], a[1] * b[0]);
}

/* Returns 1 and *t_out when the ray (org, dir, [tnear, tfar]) hits triangle p. */
static inline int pluecker_hit(const float org[3], const float dir[3], float tnear,
                               float tfar, const float *p, float *t_out) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3], s[3], c[3];
    for (int k = 0; k < 3; ++k) {
        v0[k] = p[k] - org[k];
        v1[k] = p[3 + k] - org[k];
        v2[k] = p[6 + k] - org[k];
    }
    for (int k = 0; k < 3; ++k) {
        e0[k] = v2[k] - v0[k];
        e1[k] = v0[k] - v1[k];
        e2[k] = v1[k] - v2[k];
    }
    for (int k = 0; k < 3; ++k) s[k] = v2[k] + v0[k];
    cross3f(e0, s, c);
    const float U = dot3f(c, dir);
    for (int k = 0; k < 3; ++k) s[k] = v0[k] + v1[k];
    cross3f(e1, s, c);
    const float V = dot3f(c, dir);
    for (int k = 0; k < 3; ++k) s[k] = v1[k] + v2[k];
    cross3f(e2, s, c);
    const float W = dot3f(c, dir);
    const float UVW = (U + V) + W;
    const float eps = FLT_EPSILON * fabsf(UVW);
    const float mn = fminf(fminf(U, V), W);
    const float mx = fmaxf(fmaxf(U, V), W);
    if (!(mn >= -eps || mx <= eps)) return 0;

    /* "stable" triangle normal: per component, take e0 x e1 or e1 x e2,
     * whichever has the smaller leading product */
    const float ab_x = e0[2] * e1[1], ab_y = e0[0] * e1[2], ab_z = e0[1] * e1[0];
    const float bc_x = e1[2] * e2[1], bc_y = e1[0] * e2[2], bc_z = e1[1] * e2[0];
    const float cab[3] = {msubf(e0[1], e1[2], ab_x), msubf(e0[2], e1[0], ab_y),
                          msubf(e0[0], e1[1], ab_z)};
    const float cbc[3] = {msubf(e1[1], e2[2], bc_x), msubf(e1[2], e2[0], bc_y),
                          msubf(e1[0], e2[1], bc_z)};
    float Ng[3];
    Ng[0] = (fabsf(ab_x) < fabsf(bc_x)) ? cab[0] : cbc[0];
    Ng[1] = (fabsf(ab_y) < fabsf(bc_y)) ? cab[1] : cbc[1];
    Ng[2] = (fabsf(ab_z) < fabsf(bc_z)) ? cab[2] : cbc[2];

    const float dn = dot3f(Ng, dir);
    const float den = dn + dn;
    const float Tn = dot3f(v0, Ng);
    const float T = Tn + Tn;
    if (den == 0.0f) return 0;
    const float t = hit_distance(T, den);
    if (!(tnear <= t && t <= tfar)) return 0;
    *t_out = t;
    return 1;
}

/* Closest hit over all triangles, visited in index order, a later triangle
 * replacing an earlier one at equal t (accept t <= tfar, then tfar = t).  This
 * order is the oracle's definition; any acceleration structure must reproduce
 * it: the winner is the smallest t, and among equal t the largest index. */
static void closest_hit_brute(const oracle_scene *s, const float org[3], const float dir[3],
                              float tnear, float *tfar, uint32_t *prim) {
    float best = *tfar;
    uint32_t id = ORACLE_INVALID_ID;
    for (size_t k = 0; k < s->nf; ++k) {
        float t;
        if (pluecker_hit(org, dir, tnear, best, s->tri + 9 * k, &t)) {
            if (g_tie && id != ORACLE_INVALID_ID && t == best) continue; /* study variant: the first of equal t stays */
            best = t;
            id = (uint32_t)k;
        }
    }
    *tfar = best;
    *prim = id;
}

/* --- acceleration: median-split BVH with a double-precision slab test ------ */

/* A triangle whose edge vectors are numerically parallel (collinear vertices, a repeated vertex) has no
 * well-defined plane: the Pluecker test's edge functions are rounding noise and the distance it reports is
 * unrelated to where the triangle is, so no box is conservative for it.  The brute force is the definition;
 * the accelerator never culls such a triangle (found by tools/simt/fuzz.py). */
static int tri_degenerate(const float *p) {
    const float e1[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, e2[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const float l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    return !(nn > 1e-8f * l1 * l2);
}

static void tri_bounds(const float *p, float lo[3], float hi[3]) {
    if (tri_degenerate(p)) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = -FLT_MAX / 4;
            hi[k] = FLT_MAX / 4;
        }
        return;
    }
    for (int k = 0; k < 3; ++k) {
        lo[k] = fminf(fminf(p[k], p[3 + k]), p[6 + k]);
        hi[k] = fmaxf(fmaxf(p[k], p[3 + k]), p[6 + k]);
    }
}

typedef struct { float c; uint32_t id; } sort_item;
static int cmp_item(const void *a, const void *b) {
    const sort_item *x = (const sort_item *)a, *y = (const sort_item *)b;
    if (x->c < y->c) return -1;
    if (x->c > y->c) return 1;
    return (x->id > y->id) - (x->id < y->id);
}

static int build_rec(oracle_scene *s, int first, int count, sort_item *scratch) {
    int me = s->nnodes++;
    bvh_node *nd = &s->nodes[me];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int q = first; q < first + count; ++q) {
        float l[3], h[3];
        tri_bounds(s->tri + 9 * (size_t)s->perm[q], l, h);
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], l[k]);
            hi[k] = fmaxf(hi[k], h[k]);
            float c = 0.5f * (l[k] + h[k]);
            clo[k] = fminf(clo[k], c);
            chi[k] = fmaxf(chi[k], c);
        }
    }
    for (int k = 0; k < 3; ++k) {
        /* generous padding: the Pluecker test tolerates rays a few ulps outside
         * the triangle, the boxes must never be tighter than that */
        float pad = 1e-4f * (hi[k] - lo[k]) + 1e-5f * fmaxf(fabsf(lo[k]), fabsf(hi[k])) + 1e-30f;
        nd->lo[k] = lo[k] - pad;
        nd->hi[k] = hi[k] + pad;
    }
    nd->left = nd->right = -1;
    nd->first = first;
    nd->count = count;
    if (count <= 4) return me;
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    for (int q = 0; q < count; ++q) {
        float l[3], h[3];
        uint32_t id = s->perm[first + q];
        tri_bounds(s->tri + 9 * (size_t)id, l, h);
        scratch[q].c = 0.5f * (l[ax] + h[ax]);
        scratch[q].id = id;
    }
    qsort(scratch, (size_t)count, sizeof(sort_item), cmp_item);
    for (int q = 0; q < count; ++q) s->perm[first + q] = scratch[q].id;
    int half = count / 2;
    nd->count = 0;
    int l = build_rec(s, first, half, scratch);
    int r = build_rec(s, first + half, count - half, scratch);
    s->nodes[me].left = l;
    s->nodes[me].right = r;
    return me;
}

static int slab_hit(const bvh_node *nd, const double o[3], const double inv[3], double tnear,
                    double tfar) {
    double t0 = tnear, t1 = tfar;
    for (int k = 0; k < 3; ++k) {
        double a = ((double)nd->lo[k] - o[k]) * inv[k];
        double b = ((double)nd->hi[k] - o[k]) * inv[k];
        if (a != a || b != b) continue; /* 0 * inf: origin on a slab plane of a flat axis */
        double n = a < b ? a : b, f = a < b ? b : a;
        if (n > t0) t0 = n;
        if (f < t1) t1 = f;
    }
    return t0 <= t1 * (1.0 + 1e-6) + 1e-30;
}

static void closest_hit_bvh(const oracle_scene *s, const float org[3], const float dir[3],
                            float tnear, float *tfar, uint32_t *prim) {
    float best = *tfar;
    uint32_t id = ORACLE_INVALID_ID;
    double o[3], inv[3];
    for (int k = 0; k < 3; ++k) {
        o[k] = org[k];
        inv[k] = 1.0 / (double)dir[k];
    }
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const bvh_node *nd = &s->nodes[stack[--sp]];
        if (!slab_hit(nd, o, inv, tnear, best)) continue;
        if (nd->count > 0) {
            for (int q = nd->first; q < nd->first + nd->count; ++q) {
                uint32_t k = s->perm[q];
                float t;
                /* same winner as the index-ordered brute force: smaller t, or
                 * equal t and larger index */
                if (pluecker_hit(org, dir, tnear, best, s->tri + 9 * (size_t)k, &t)) {
                    if (t < best || id == ORACLE_INVALID_ID || (g_tie ? k < id : k > id)) {
                        best = t;
                        id = k;
                    }
                }
            }
        } else {
            stack[sp++] = nd->left;
            stack[sp++] = nd->right;
        }
    }
    *tfar = best;
    *prim = id;
}

oracle_scene *oracle_scene_create(const float *verts, size_t nv, const uint32_t *faces, size_t nf) {
    (void)nv;
    oracle_scene *s = (oracle_scene *)calloc(1, sizeof(*s));
    s->nf = nf;
    s->tri = (float *)malloc(sizeof(float) * 9 * (nf ? nf : 1));
    for (size_t f = 0; f < nf; ++f)
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < 3; ++k) s->tri[9 * f + 3 * c + k] = verts[3 * (size_t)faces[3 * f + c] + k];
    s->nodes = (bvh_node *)malloc(sizeof(bvh_node) * (2 * (nf ? nf : 1)));
    s->perm = (uint32_t *)malloc(sizeof(uint32_t) * (nf ? nf : 1));
    for (size_t f = 0; f < nf; ++f) s->perm[f] = (uint32_t)f;
    s->nnodes = 0;
    if (nf) {
        sort_item *scratch = (sort_item *)malloc(sizeof(sort_item) * nf);
        build_rec(s, 0, (int)nf, scratch);
        free(scratch);
    }
    return s;
}

void oracle_scene_destroy(oracle_scene *s) {
    if (!s) return;
    free(s->tri);
    free(s->nodes);
    free(s->perm);
    free(s);
}

static inline void closest_hit(const oracle_scene *s, int use_bvh, const float org[3],
                               const float dir[3], float tnear, float *tfar, uint32_t *prim) {
    if (s->nf == 0) {
        *prim = ORACLE_INVALID_ID;
        return;
    }
    if (use_bvh) closest_hit_bvh(s, org, dir, tnear, tfar, prim);
    else closest_hit_brute(s, org, dir, tnear, tfar, prim);
}

/* Threads an OpenMP region with this request really gets (bench.py prints it next to
 * `cores`: under torchrun OMP_NUM_THREADS=1 is exported and a default-sized team is one thread). */
int oracle_team_size(int nthreads) {
    int got = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp single
        got = omp_get_num_threads();
    }
#endif
    (void)nthreads;
    return got;
}

/* rtcIntersect1M stand-in: stream of n single rays (shape.py:375-390).  On a
 * hit tfar <- t, prim_id <- triangle, geom_id <- 0; otherwise untouched. */
void oracle_intersect1M(const oracle_scene *s, size_t n, const float *org, const float *dir,
                        const float *tnear, float *tfar, uint32_t *prim_id, uint32_t *geom_id,
                        int use_bvh, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long long q = 0; q < (long long)n; ++q) {
        float tf = tfar[q];
        uint32_t prim;
        closest_hit(s, use_bvh, org + 3 * q, dir + 3 * q, tnear[q], &tf, &prim);
        if (prim != ORACLE_INVALID_ID) {
            tfar[q] = tf;
            prim_id[q] = prim;
            geom_id[q] = 0;
        }
    }
}

/* rtcOccluded1M stand-in (shape.py:409-421): tfar <- -inf when anything is hit */
void oracle_occluded1M(const oracle_scene *s, size_t n, const float *org, const float *dir,
                       const float *tnear, float *tfar, int use_bvh, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (long long q = 0; q < (long long)n; ++q) {
        float tf = tfar[q];
        uint32_t prim;
        closest_hit(s, use_bvh, org + 3 * q, dir + 3 * q, tnear[q], &tf, &prim);
        if (prim != ORACLE_INVALID_ID) tfar[q] = -INFINITY;
    }
}

/* ------------------------------------------------------------------------ */
/* Ray set-up of EmbreeTrimeshShapeModel._get_visibility (shape.py:349-385),  */
/* restated for one (i, j) pair in the shape model's dtype.                   */
/* Returns 0 when the pair is masked out (norm <= eps: "vis by default").     */
/* ------------------------------------------------------------------------ */

static inline int setup_ray_f32(const float *Pi, const float *Pj, float org[3], float dir[3]) {
    const float eps = ray_eps_f32();
    float d[3];
    for (int k = 0; k < 3; ++k) d[k] = Pj[k] - Pi[k];          /* shape.py:359 */
    const float nrm = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]); /* :360 */
    if (!(nrm > eps)) return 0;                                  /* :361 */
    for (int k = 0; k < 3; ++k) {
        dir[k] = d[k] / nrm;                                     /* :362 */
        org[k] = Pi[k] + eps * dir[k];                           /* :380 */
    }
    return 1;
}

static inline int setup_ray_f64(const double *Pi, const double *Pj, float org[3], float dir[3]) {
    const double eps = (double)ray_eps_f32(); /* float32 scalar promoted by NumPy */
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = Pj[k] - Pi[k];
    const double nrm = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    if (!(nrm > eps)) return 0;
    for (int k = 0; k < 3; ++k) {
        const double D = d[k] / nrm;
        dir[k] = (float)D;                  /* rayhit.dir[:] = D  (float32 buffer) */
        org[k] = (float)(Pi[k] + eps * D);  /* rayhit.org[:] = P + eps*D */
    }
    return 1;
}

/* exposed so tests can compare the ray set-up with the reference's arrays */
int oracle_setup_ray_f32(const float *Pi, const float *Pj, float *org, float *dir) {
    return setup_ray_f32(Pi, Pj, org, dir);
}
int oracle_setup_ray_f64(const double *Pi, const double *Pj, float *org, float *dir) {
    return setup_ray_f64(Pi, Pj, org, dir);
}

/* vis[p*n+q] for I[p], J[q]  (shape.py:349-398): masked pairs are visible by
 * default, traced pairs are visible iff the closest hit is triangle J[q]. */
#define DEFINE_VISIBILITY(SUFFIX, REAL, SETUP)                                                   \
    void oracle_visibility_##SUFFIX(const oracle_scene *s, const REAL *P, const uint64_t *I,    \
                                    size_t m, const uint64_t *J, size_t n, uint8_t *vis,        \
                                    int use_bvh, int nthreads) {                                 \
        (void)nthreads;                                                                          \
        _Pragma("omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())") \
        for (long long p = 0; p < (long long)m; ++p) {                                           \
            const REAL *Pi = P + 3 * I[p];                                                       \
            for (size_t q = 0; q < n; ++q) {                                                     \
                float org[3], dir[3];                                                            \
                if (!SETUP(Pi, P + 3 * J[q], org, dir)) {                                        \
                    vis[(size_t)p * n + q] = 1;                                                  \
                    continue;                                                                    \
                }                                                                                \
                float tf = INFINITY;                                                             \
                uint32_t prim;                                                                   \
                closest_hit(s, use_bvh, org, dir, 0.0f, &tf, &prim);                             \
                vis[(size_t)p * n + q] = (prim != ORACLE_INVALID_ID) && (prim == (uint32_t)J[q]); \
            }                                                                                    \
        }                                                                                        \
    }
DEFINE_VISIBILITY(f32, float, setup_ray_f32)
DEFINE_VISIBILITY(f64, double, setup_ray_f64)

/* _is_occluded (shape.py:400-421): origin P[I] + eps*N[I], direction D (one
 * shared vector when nd == 1 and shared != 0, else row p of D for face p when
 * per_face != 0, else the full m x nd product used by the CGAL 2-D variant,
 * src/flux/cgal/aabb.pyx:77-87).  occ[p*nd + k]. */
#define DEFINE_OCCLUDED(SUFFIX, REAL)                                                            \
    void oracle_is_occluded_##SUFFIX(const oracle_scene *s, const REAL *P, const REAL *N,       \
                                     const uint64_t *I, size_t m, const REAL *D, size_t nd,     \
                                     int per_face, uint8_t *occ, int use_bvh, int nthreads) {   \
        (void)nthreads;                                                                          \
        const REAL eps = (REAL)ray_eps_f32();                                                    \
        _Pragma("omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())") \
        for (long long p = 0; p < (long long)m; ++p) {                                           \
            const size_t i = I[p];                                                               \
            float org[3];                                                                        \
            for (int k = 0; k < 3; ++k) org[k] = (float)(P[3 * i + k] + eps * N[3 * i + k]);     \
            const size_t cols = per_face ? 1 : nd;                                               \
            for (size_t c = 0; c < cols; ++c) {                                                  \
                const REAL *d = per_face ? D + 3 * (size_t)p : D + 3 * c;                        \
                float dir[3] = {(float)d[0], (float)d[1], (float)d[2]};                          \
                float tf = INFINITY;                                                             \
                uint32_t prim;                                                                   \
                closest_hit(s, use_bvh, org, dir, 0.0f, &tf, &prim);                             \
                occ[(size_t)p * cols + c] = prim != ORACLE_INVALID_ID;                           \
            }                                                                                    \
        }                                                                                        \
    }
DEFINE_OCCLUDED(f32, float)
DEFINE_OCCLUDED(f64, double)

/* ------------------------------------------------------------------------ */
/* get_form_factor_matrix (form_factors.py:11-72), "direct" evaluation        */
/*                                                                          */
/* Deviation from the reference's round-off, on purpose: the two  */
/* dot products are evaluated in double from the shape model's own P, N, A    */
/* values as n_i.(p_j-p_i) and n_j.(p_i-p_j) -- not as P[i]@N[J].T - NJ_PJ,   */
/* which cancels -- and rounded once to the shape model's dtype.  The cull    */
/* (form_factors.py:52) is applied to that rounded numerator against eps      */
/* converted to the same dtype, as NumPy does for a Python-float eps.         */
/* ------------------------------------------------------------------------ */

typedef struct oracle_ff_result {
    size_t m;
    int64_t *indptr;     /* m + 1 */
    uint64_t **row_idx;  /* per row column positions (into J), ascending */
    void **row_val;      /* per row values, REAL */
    int is_f64;
    int64_t tested;      /* pairs that survived the cull (one ray each) */
} oracle_ff_result;

#define PI_D 3.141592653589793

static inline double dot3d(const double a[3], const double b[3]) {
    return fma(a[0], b[0], fma(a[1], b[1], a[2] * b[2]));
}

#define DEFINE_FF(SUFFIX, REAL, SETUP, ISF64)                                                    \
    oracle_ff_result *oracle_ff_assemble_##SUFFIX(                                               \
        const oracle_scene *s, const REAL *P, const REAL *N, const REAL *A, const uint64_t *I,  \
        size_t m, const uint64_t *J, size_t n, double eps_in, int use_bvh, int nthreads) {      \
        oracle_ff_result *r = (oracle_ff_result *)calloc(1, sizeof(*r));                         \
        r->m = m;                                                                                \
        r->is_f64 = ISF64;                                                                       \
        r->indptr = (int64_t *)calloc(m + 1, sizeof(int64_t));                                   \
        r->row_idx = (uint64_t **)calloc(m ? m : 1, sizeof(uint64_t *));                         \
        r->row_val = (void **)calloc(m ? m : 1, sizeof(void *));                                 \
        const REAL eps = (REAL)eps_in;                                                           \
        int64_t tested = 0;                                                                      \
        (void)nthreads;                                                                          \
        _Pragma("omp parallel for schedule(dynamic, 1) reduction(+ : tested) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())") \
        for (long long p = 0; p < (long long)m; ++p) {                                           \
            const size_t i = I[p];                                                               \
            const REAL *Pi = P + 3 * i;                                                          \
            const double pi_[3] = {Pi[0], Pi[1], Pi[2]};                                         \
            const double ni[3] = {N[3 * i], N[3 * i + 1], N[3 * i + 2]};                         \
            uint64_t *idx = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));                  \
            REAL *val = (REAL *)malloc(sizeof(REAL) * (n ? n : 1));                              \
            size_t cnt = 0;                                                                      \
            for (size_t q = 0; q < n; ++q) {                                                     \
                const size_t j = J[q];                                                           \
                const REAL *Pj = P + 3 * j;                                                      \
                const double d[3] = {(double)Pj[0] - pi_[0], (double)Pj[1] - pi_[1],             \
                                     (double)Pj[2] - pi_[2]};                                    \
                const double nj[3] = {N[3 * j], N[3 * j + 1], N[3 * j + 2]};                     \
                double a = dot3d(ni, d);           /* form_factors.py:46 */                      \
                double b = -dot3d(nj, d);          /* :47, evaluated directly */                 \
                a = a > 0.0 ? a : 0.0;                                                           \
                b = b > 0.0 ? b : 0.0;                                                           \
                double num = a * b;                                                              \
                if (i == j) num = 0.0;             /* :50 */                                     \
                const REAL num_r = (REAL)num;                                                    \
                if (!(num_r > eps || -num_r > eps)) continue; /* :52 abs(row) > eps */           \
                ++tested;                                                                        \
                float org[3], dir[3];                                                            \
                int visible = 1;                   /* shape.py:392 "vis by default" */           \
                if (SETUP(Pi, Pj, org, dir)) {                                                   \
                    float tf = INFINITY;                                                         \
                    uint32_t prim;                                                               \
                    closest_hit(s, use_bvh, org, dir, 0.0f, &tf, &prim);                         \
                    visible = (prim != ORACLE_INVALID_ID) && (prim == (uint32_t)j);              \
                }                                                                                \
                if (!visible) continue;            /* :60 */                                     \
                const double r2 = dot3d(d, d);                                                   \
                const double sden = PI_D * (r2 * r2); /* :62 */                                  \
                const double v = (sden == 0.0) ? 0.0 : (num * (double)A[j]) / sden; /* :63-64 */ \
                idx[cnt] = q;                                                                    \
                val[cnt] = (REAL)v;                                                              \
                ++cnt;                                                                           \
            }                                                                                    \
            r->row_idx[p] = (uint64_t *)realloc(idx, sizeof(uint64_t) * (cnt ? cnt : 1));        \
            r->row_val[p] = realloc(val, sizeof(REAL) * (cnt ? cnt : 1));                        \
            r->indptr[p + 1] = (int64_t)cnt;                                                     \
        }                                                                                        \
        for (size_t p = 0; p < m; ++p) r->indptr[p + 1] += r->indptr[p];                         \
        r->tested = tested;                                                                      \
        return r;                                                                                \
    }
DEFINE_FF(f32, float, setup_ray_f32, 0)
DEFINE_FF(f64, double, setup_ray_f64, 1)

int64_t oracle_ff_nnz(const oracle_ff_result *r) { return r->indptr[r->m]; }
int64_t oracle_ff_tested(const oracle_ff_result *r) { return r->tested; }

void oracle_ff_copy(const oracle_ff_result *r, int64_t *indptr, int64_t *indices, void *data) {
    memcpy(indptr, r->indptr, sizeof(int64_t) * (r->m + 1));
    const size_t w = r->is_f64 ? 8 : 4;
    for (size_t p = 0; p < r->m; ++p) {
        const int64_t b = r->indptr[p], c = r->indptr[p + 1] - b;
        for (int64_t k = 0; k < c; ++k) indices[b + k] = (int64_t)r->row_idx[p][k];
        memcpy((char *)data + w * (size_t)b, r->row_val[p], w * (size_t)c);
    }
}

void oracle_ff_free(oracle_ff_result *r) {
    if (!r) return;
    for (size_t p = 0; p < r->m; ++p) {
        free(r->row_idx[p]);
        free(r->row_val[p]);
    }
    free(r->row_idx);
    free(r->row_val);
    free(r->indptr);
    free(r);
}

/* ------------------------------------------------------------------------ */
/* Face geometry (shape.py:16-45) in the array dtype, NumPy operation order   */
/* ------------------------------------------------------------------------ */
#define DEFINE_GEOM(SUFFIX, REAL, SQRT)                                                          \
    void oracle_face_geometry_##SUFFIX(const REAL *V, const int64_t *F, size_t nf, REAL *P,     \
                                       REAL *N, REAL *A) {                                       \
        for (size_t f = 0; f < nf; ++f) {                                                        \
            const REAL *v0 = V + 3 * F[3 * f], *v1 = V + 3 * F[3 * f + 1], *v2 = V + 3 * F[3 * f + 2]; \
            REAL a[3], b[3], c[3];                                                               \
            for (int k = 0; k < 3; ++k) {                                                        \
                /* V[F].mean(axis=1): add.reduce over the 3 vertices, then / 3 */                \
                P[3 * f + k] = ((v0[k] + v1[k]) + v2[k]) / (REAL)3;                              \
                a[k] = v1[k] - v0[k];                                                            \
                b[k] = v2[k] - v0[k];                                                            \
            }                                                                                    \
            c[0] = a[1] * b[2] - a[2] * b[1]; /* np.cross */                                     \
            c[1] = a[2] * b[0] - a[0] * b[2];                                                    \
            c[2] = a[0] * b[1] - a[1] * b[0];                                                    \
            const REAL nrm = SQRT((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);                    \
            for (int k = 0; k < 3; ++k) N[3 * f + k] = c[k] / nrm;                               \
            A[f] = nrm / (REAL)2;                                                                \
        }                                                                                        \
    }
DEFINE_GEOM(f32, float, sqrtf)
DEFINE_GEOM(f64, double, sqrt)
