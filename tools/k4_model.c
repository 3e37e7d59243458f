/* k4_model.c -- CPU model of the trace kernel's traversal (K4, fluxpy_b200/csrc/assemble.cuh), for
 * counting work, not for results.  It rebuilds the LBVH the way lbvh.cuh does (Morton codes of the box
 * centres with one scale for all axes, Karras hierarchy, padded boxes, fitted slabs), then walks a sample
 * of (row, chunk) work units exactly as the kernel does -- per-unit record list, shaft filter, batches of
 * 32 surviving columns, phase A / B / C -- and reports the warp-level iteration counts per batch that
 * decide the kernel's instruction count (it is issue-bound: profiles/r01b_kernels_ncu.md).  Design variants
 * (chunk size, a second, per-batch shaft filter) are switches, so their effect on the counts can be read
 * off without GPU time.  Not part of the product or of the oracle; nothing imports it.
 *
 *   gcc -O2 -shared -fPIC -o /tmp/libk4model.so tools/k4_model.c -lm     (tools/k4_model.py does this)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;                 /* faces */
    const float *V;        /* vertices */
    const int *F;          /* faces */
    const double *P, *N;   /* centroids, normals (per face) */
    int *left, *right, *parent, *first, *last; /* Karras nodes: internal 0..n-2, leaf k -> n-1+k */
    float *box;            /* 6 per node */
    float *sdir, *smin, *smax; /* fitted slab per node */
    int *leaf_face, *face_leaf;
    uint64_t *keys;
    float scale;
} Model;

static uint64_t spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static const uint64_t *g_keys;
static int cmp_key(const void *a, const void *b) {
    const int x = *(const int *)a, y = *(const int *)b;
    if (g_keys[x] != g_keys[y]) return g_keys[x] < g_keys[y] ? -1 : 1;
    return x - y; /* stable, as the LSD radix sort */
}

static int delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __builtin_clz((unsigned)(i ^ j));
    return __builtin_clzll(a ^ b);
}

Model *k4_build(int nv, const float *V, int nf, const int *F, const double *P, const double *N) {
    (void)nv;
    Model *M = calloc(1, sizeof(Model));
    const int n = nf, nn = 2 * n - 1;
    M->n = n; M->V = V; M->F = F; M->P = P; M->N = N;
    M->left = malloc(sizeof(int) * n); M->right = malloc(sizeof(int) * n);
    M->parent = malloc(sizeof(int) * nn); M->first = malloc(sizeof(int) * n); M->last = malloc(sizeof(int) * n);
    M->box = malloc(sizeof(float) * 6 * nn); M->sdir = malloc(sizeof(float) * 3 * nn);
    M->smin = malloc(sizeof(float) * nn); M->smax = malloc(sizeof(float) * nn);
    M->leaf_face = malloc(sizeof(int) * n); M->face_leaf = malloc(sizeof(int) * n);
    uint64_t *code = malloc(sizeof(uint64_t) * n);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, mx = 0.f;
    for (int f = 0; f < n; ++f)
        for (int k = 0; k < 3; ++k) {
            const float a = V[3 * F[3 * f] + k], b = V[3 * F[3 * f + 1] + k], c = V[3 * F[3 * f + 2] + k];
            const float l = fminf(fminf(a, b), c), h = fmaxf(fmaxf(a, b), c), ce = 0.5f * (l + h);
            lo[k] = fminf(lo[k], ce); hi[k] = fmaxf(hi[k], ce);
            mx = fmaxf(mx, fmaxf(fabsf(l), fabsf(h)));
        }
    M->scale = mx;
    const double ext = fmax(fmax((double)hi[0] - lo[0], (double)hi[1] - lo[1]), (double)hi[2] - lo[2]);
    for (int f = 0; f < n; ++f) {
        uint64_t c64 = 0;
        for (int k = 0; k < 3; ++k) {
            const float a = V[3 * F[3 * f] + k], b = V[3 * F[3 * f + 1] + k], c = V[3 * F[3 * f + 2] + k];
            const double ce = 0.5f * (fminf(fminf(a, b), c) + fmaxf(fmaxf(a, b), c));
            double u = ext > 0 ? (ce - lo[k]) / ext : 0.0;
            u = fmin(fmax(u, 0.0), 1.0);
            c64 |= spread21((uint64_t)fmin(u * 2097152.0, 2097151.0)) << (2 - k);
        }
        code[f] = c64;
    }
    for (int f = 0; f < n; ++f) M->leaf_face[f] = f;
    g_keys = code;
    qsort(M->leaf_face, n, sizeof(int), cmp_key);
    M->keys = malloc(sizeof(uint64_t) * n);
    for (int k = 0; k < n; ++k) { M->keys[k] = code[M->leaf_face[k]]; M->face_leaf[M->leaf_face[k]] = k; }
    free(code);
    const uint64_t *keys = M->keys;
    for (int x = 0; x < nn; ++x) M->parent[x] = -1;
    for (int i = 0; i < n - 1; ++i) { /* karras_kernel */
        const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta(keys, n, i, i - d);
        int lmax = 2;
        while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d, dnode = delta(keys, n, i, j);
        int s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const int gamma = i + s * d + (d < 0 ? d : 0);
        const int lo_ = i < j ? i : j, hi_ = i < j ? j : i;
        const int lc = (lo_ == gamma) ? (n - 1) + gamma : gamma;
        const int rc = (hi_ == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
        M->left[i] = lc; M->right[i] = rc; M->parent[lc] = i; M->parent[rc] = i;
        M->first[i] = lo_; M->last[i] = hi_;
    }
    M->parent[0] = -1;
    /* leaf boxes + area normals, then bottom-up by decreasing depth: process internal nodes in an order
     * where children come first = sort by subtree size is overkill; use a post-order stack walk */
    float *an = malloc(sizeof(float) * 3 * nn);
    const float pad = 1.0e-6f * mx + 1e-30f;
    for (int k = 0; k < n; ++k) {
        const int f = M->leaf_face[k], x = n - 1 + k;
        const float *v0 = V + 3 * F[3 * f], *v1 = V + 3 * F[3 * f + 1], *v2 = V + 3 * F[3 * f + 2];
        for (int c = 0; c < 3; ++c) {
            M->box[6 * x + c] = fminf(fminf(v0[c], v1[c]), v2[c]) - pad;
            M->box[6 * x + 3 + c] = fmaxf(fmaxf(v0[c], v1[c]), v2[c]) + pad;
        }
        const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
        const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
        an[3 * x] = e1y * e2z - e1z * e2y; an[3 * x + 1] = e1z * e2x - e1x * e2z; an[3 * x + 2] = e1x * e2y - e1y * e2x;
    }
    if (n > 1) {
        int *stack = malloc(sizeof(int) * 2 * nn), sp = 0;
        char *seen = calloc(nn, 1);
        stack[sp++] = 0;
        while (sp) {
            const int x = stack[sp - 1];
            if (x >= n - 1) { --sp; continue; }
            if (!seen[x]) { seen[x] = 1; stack[sp++] = M->left[x]; stack[sp++] = M->right[x]; continue; }
            --sp;
            const int lc = M->left[x], rc = M->right[x];
            for (int c = 0; c < 3; ++c) {
                M->box[6 * x + c] = fminf(M->box[6 * lc + c], M->box[6 * rc + c]);
                M->box[6 * x + 3 + c] = fmaxf(M->box[6 * lc + 3 + c], M->box[6 * rc + 3 + c]);
                an[3 * x + c] = an[3 * lc + c] + an[3 * rc + c];
            }
        }
        free(stack); free(seen);
    }
    for (int x = 0; x < nn; ++x) {
        const float ax = an[3 * x], ay = an[3 * x + 1], az = an[3 * x + 2], l = sqrtf(ax * ax + ay * ay + az * az);
        if (l > 1e-30f && isfinite(l)) { M->sdir[3 * x] = ax / l; M->sdir[3 * x + 1] = ay / l; M->sdir[3 * x + 2] = az / l; }
        else { M->sdir[3 * x] = 0; M->sdir[3 * x + 1] = 0; M->sdir[3 * x + 2] = 1; }
        M->smin[x] = INFINITY; M->smax[x] = -INFINITY;
    }
    free(an);
    for (int k = 0; k < n; ++k) { /* slab_extent_kernel, every ancestor (slab_limit = inf) */
        const int f = M->leaf_face[k];
        int x = n - 1 + k;
        while (x >= 0) {
            const float *d = M->sdir + 3 * x;
            for (int v = 0; v < 3; ++v) {
                const float *p = V + 3 * F[3 * f + v];
                const float t = d[0] * p[0] + d[1] * p[1] + d[2] * p[2];
                M->smin[x] = fminf(M->smin[x], t); M->smax[x] = fmaxf(M->smax[x], t);
            }
            x = M->parent[x];
        }
    }
    const float spad = 8.0e-6f * mx + 1e-30f;
    for (int x = 0; x < nn; ++x) { M->smin[x] -= spad; M->smax[x] += spad; }
    return M;
}

void k4_free(Model *M) {
    free(M->left); free(M->right); free(M->parent); free(M->first); free(M->last); free(M->box); free(M->sdir);
    free(M->smin); free(M->smax); free(M->leaf_face); free(M->face_leaf); free(M->keys); free(M);
}

typedef struct { float ox, oy, oz, dx, dy, dz, ix, iy, iz, qx, qy, qz, tmax; } Ray;

static int node_hit(const Model *M, int x, const Ray *r) { /* child_hit of trace.cuh */
    const float *b = M->box + 6 * x, *d = M->sdir + 3 * x;
    const float no = d[0] * r->ox + d[1] * r->oy + d[2] * r->oz, nd = d[0] * r->dx + d[1] * r->dy + d[2] * r->dz;
    const float rn = 1.0f / nd;
    const float s0 = (M->smin[x] - no) * rn, s1 = (M->smax[x] - no) * rn;
    float tn = fmaxf(fminf(s0, s1), 0.0f), tf = fminf(fmaxf(s0, s1), r->tmax);
    const float x0 = b[0] * r->ix - r->qx, x1 = b[3] * r->ix - r->qx;
    const float y0 = b[1] * r->iy - r->qy, y1 = b[4] * r->iy - r->qy;
    const float z0 = b[2] * r->iz - r->qz, z1 = b[5] * r->iz - r->qz;
    tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tn));
    tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tf));
    return tn <= tf * 1.000002f;
}

static void node_range(const Model *M, int x, int *lo, int *hi) {
    if (x >= M->n - 1) { *lo = *hi = x - (M->n - 1); }
    else { *lo = M->first[x]; *hi = M->last[x]; }
}
static double g_margin = 1e-4; /* sine margin of the horizon comparison */
void k4_set_margin(double m) { g_margin = m; }

static int zone_of(const Model *M, int leafnode, int max_leaves) { /* largest ancestor (or the leaf) with <= max_leaves */
    int x = leafnode;
    while (M->parent[x] >= 0) {
        const int p = M->parent[x];
        if (M->last[p] - M->first[p] + 1 > max_leaves) break;
        x = p;
    }
    return x;
}

/* Variant "horizon": hor[f] = upper bound of sin(elevation), measured from the centroid of face f against its
 * normal, of every point of the OTHER triangles in f's zone (the ancestor of f with at most zone_leaves
 * leaves).  A ray that leaves f's centroid (or arrives at it) with a larger sine cannot meet any of them.
 * Sampled densely here (model accuracy); +inf when a zone triangle comes closer than 1e-6 of the scale. */
void k4_horizons(const Model *M, int zone_leaves, float *hor) {
    const int n = M->n;
    for (int f = 0; f < n; ++f) {
        const int z = zone_of(M, n - 1 + M->face_leaf[f], zone_leaves);
        int lo, hi;
        if (z >= n - 1) { lo = hi = z - (n - 1); } else { lo = M->first[z]; hi = M->last[z]; }
        const double *P = M->P + 3 * f, *N = M->N + 3 * f;
        double best = -1.0;
        for (int k = lo; k <= hi; ++k) {
            const int g = M->leaf_face[k];
            if (g == f) continue;
            const float *v[3] = {M->V + 3 * M->F[3 * g], M->V + 3 * M->F[3 * g + 1], M->V + 3 * M->F[3 * g + 2]};
            for (int a = 0; a <= 8; ++a)
                for (int b = 0; a + b <= 8; ++b) {
                    const double wa = a / 8.0, wb = b / 8.0, wc = 1.0 - wa - wb;
                    double d[3], l2 = 0, h = 0;
                    for (int c = 0; c < 3; ++c) { d[c] = wa * v[0][c] + wb * v[1][c] + wc * v[2][c] - P[c]; l2 += d[c] * d[c]; h += d[c] * N[c]; }
                    const double l = sqrt(l2);
                    if (l < 1e-6 * M->scale) { best = INFINITY; continue; }
                    if (h / l > best) best = h / l;
                }
        }
        hor[f] = (float)best;
    }
}

/* Variant "multi-level source horizon": further, larger source-side zones (leaf counts g_src_zone[k] with
 * horizons g_src_hor[k], valid for the sampled rows only); a batch skips the records inside the largest zone
 * whose horizon all its rays clear. */
double g_chist[20];
double *k4_chist(void) { return g_chist; }
static int g_src_levels = 0;
static int g_src_zone[4];
static const float *g_src_hor[4];
void k4_set_source_levels(int nlev, const int *zones, const float **hors) {
    g_src_levels = nlev;
    for (int k = 0; k < nlev && k < 4; ++k) { g_src_zone[k] = zones[k]; g_src_hor[k] = hors[k]; }
}
/* horizons of the given faces only (vertices + edge midpoints + centroid sampled; model accuracy) */
void k4_horizons_rows(const Model *M, int zone_leaves, int nrows, const int *rows, float *hor) {
    const int n = M->n;
    for (int r = 0; r < nrows; ++r) {
        const int f = rows[r];
        const int z = zone_of(M, n - 1 + M->face_leaf[f], zone_leaves);
        int lo, hi;
        if (z >= n - 1) { lo = hi = z - (n - 1); } else { lo = M->first[z]; hi = M->last[z]; }
        const double *P = M->P + 3 * f, *N = M->N + 3 * f;
        double best = -1.0;
        for (int k = lo; k <= hi; ++k) {
            const int g = M->leaf_face[k];
            if (g == f) continue;
            const float *v[3] = {M->V + 3 * M->F[3 * g], M->V + 3 * M->F[3 * g + 1], M->V + 3 * M->F[3 * g + 2]};
            for (int a = 0; a <= 2; ++a)
                for (int b = 0; a + b <= 2; ++b) {
                    const double wa = a / 2.0, wb = b / 2.0, wc = 1.0 - wa - wb;
                    double d[3], l2 = 0, h = 0;
                    for (int c = 0; c < 3; ++c) { d[c] = wa * v[0][c] + wb * v[1][c] + wc * v[2][c] - P[c]; l2 += d[c] * d[c]; h += d[c] * N[c]; }
                    const double l = sqrt(l2);
                    if (l < 1e-6 * M->scale) { best = INFINITY; continue; }
                    if (h / l > best) best = h / l;
                }
        }
        hor[f] = (float)best;
    }
}

static int zone64(const Model *M, int leafnode) { /* largest ancestor (or the leaf) holding at most 64 leaves */
    int x = leafnode;
    while (M->parent[x] >= 0) {
        const int p = M->parent[x];
        if (M->last[p] - M->first[p] + 1 > 64) break;
        x = p;
    }
    return x;
}
static int in_zone(const Model *M, int zone, int leafpos) {
    int lo, hi;
    if (zone >= M->n - 1) { lo = hi = zone - (M->n - 1); } else { lo = M->first[zone]; hi = M->last[zone]; }
    return leafpos >= lo && leafpos <= hi;
}
static int sibling(const Model *M, int x) {
    const int p = M->parent[x];
    return M->left[p] == x ? M->right[p] : M->left[p];
}

/* out[]: 0 batches, 1 rays, 2 A warp-iterations, 3 B warp-iterations, 4 C warp-iterations, 5 A lane hits,
 * 6 B lane-iterations, 7 C lane-iterations, 8 leaf candidates, 9 flush warp-iterations, 10 units,
 * 11 list length before the filter (sum over units), 12 after the chunk filter, 13 A warp-iterations with
 * the per-batch filter, 14 candidate pairs, 15 units with a usable common ancestor, 16 / 17 phase-C lane visits
 * rooted in a phase-A / phase-B hit, 18 deferred candidates that are the source triangle itself, 19 / 20
 * candidates under the 64-leaf ancestor of the source / of the target, 21 phase-A lane hits on records of
 * at most 64 leaves, 22 phase-B lane hits */
/* projection of an AABB (lo, hi) on axis a: [*mn, *mx] */
static void proj_box(const float *lo, const float *hi, const float *a, float *mn, float *mx) {
    *mn = *mx = 0.f;
    for (int k = 0; k < 3; ++k) {
        const float p = a[k] * lo[k], q = a[k] * hi[k];
        *mn += fminf(p, q); *mx += fmaxf(p, q);
    }
}

void k4_count(const Model *M, int nrows, const int *rows, int chunk, double eps, int batch_filter, int axes,
              int expand_leaves, int zone_leaves, const float *hor, double *out) {
    const int n = M->n, nchunks = (n + chunk - 1) / chunk;
    int *list = malloc(sizeof(int) * 256), *surv = malloc(sizeof(int) * chunk);
    int *stack = malloc(sizeof(int) * 32 * 128);
    for (int ri = 0; ri < nrows; ++ri) {
        const int i = rows[ri], ileaf = M->face_leaf[i];
        const int szone = zone64(M, M->n - 1 + ileaf);
        const double *Pi = M->P + 3 * i, *Ni = M->N + 3 * i;
        for (int c = 0; c < nchunks; ++c) {
            const int s0 = c * chunk, s1 = (s0 + chunk < n ? s0 + chunk : n);
            /* list: own record, siblings up to the root */
            int npath = 0, x = n - 1 + ileaf;
            list[npath++] = x;
            while (M->parent[x] >= 0) { list[npath++] = sibling(M, x); x = M->parent[x]; }
            const int leaf_lo = s0, leaf_hi = s1 - 1;
            int xe = -1, cref = -1;
            for (int e = 0; e < npath; ++e) { int lo, hi; node_range(M, list[e], &lo, &hi); if (lo <= leaf_lo && leaf_hi <= hi) xe = e; }
            if (xe >= 0 && leaf_lo != leaf_hi) {
                int cn = M->parent[n - 1 + leaf_lo];
                while (!(M->first[cn] <= leaf_lo && leaf_hi <= M->last[cn])) cn = M->parent[cn];
                int cur = cn, ok = 1;
                while (cur != list[xe]) { if (M->parent[cur] < 0 || npath >= 250) { ok = 0; break; } list[npath++] = sibling(M, cur); cur = M->parent[cur]; }
                if (ok) cref = cn; /* (an incomplete walk leaves extra records in this model's list: rare) */
            }
            /* cull + chunk bbox */
            float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
            int ns = 0;
            for (int s = s0; s < s1; ++s) {
                const int j = M->leaf_face[s];
                const double *Pj = M->P + 3 * j, *Nj = M->N + 3 * j;
                for (int k = 0; k < 3; ++k) { bl[k] = fminf(bl[k], (float)Pj[k]); bh[k] = fmaxf(bh[k], (float)Pj[k]); }
                const double dx = Pj[0] - Pi[0], dy = Pj[1] - Pi[1], dz = Pj[2] - Pi[2];
                double a = Ni[0] * dx + Ni[1] * dy + Ni[2] * dz, b = -(Nj[0] * dx + Nj[1] * dy + Nj[2] * dz);
                a = a > 0 ? a : 0; b = b > 0 ? b : 0;
                if (j != i && (float)(a * b) > (float)eps) surv[ns++] = s;
            }
            out[14] += s1 - s0;
            out[10] += 1; out[11] += npath; if (cref >= 0) out[15] += 1;
            /* shaft filter -> sel (X removed when cref valid) */
            const float px = (float)Pi[0], py = (float)Pi[1], pz = (float)Pi[2];
            int sel[256], nsel = 0;
            {
                const float pad = 3e-5f * M->scale + 1e-4f * fmaxf(fmaxf(bh[0] - bl[0], bh[1] - bl[1]), bh[2] - bl[2]) + 2e-3f;
                const float hl[3] = {fminf(px, bl[0]) - pad, fminf(py, bl[1]) - pad, fminf(pz, bl[2]) - pad};
                const float hh[3] = {fmaxf(px, bh[0]) + pad, fmaxf(py, bh[1]) + pad, fmaxf(pz, bh[2]) + pad};
                int work[512], nwork = 0;
                for (int e = npath - 1; e >= 0; --e) {
                    if (cref >= 0 && e == xe) continue;
                    work[nwork++] = list[e];
                }
                while (nwork > 0) {
                    const int y = work[--nwork]; int lo, hi; node_range(M, y, &lo, &hi);
                    int keep = 1;
                    if (hi < leaf_lo || lo > leaf_hi) {
                        const float *b = M->box + 6 * y, *d = M->sdir + 3 * y;
                        if (b[0] > hh[0] || b[3] < hl[0] || b[1] > hh[1] || b[4] < hl[1] || b[2] > hh[2] || b[5] < hl[2]) keep = 0;
                        const float sp = d[0] * px + d[1] * py + d[2] * pz;
                        const float lo_s = fminf(d[0] * bl[0], d[0] * bh[0]) + fminf(d[1] * bl[1], d[1] * bh[1]) + fminf(d[2] * bl[2], d[2] * bh[2]);
                        const float hi_s = fmaxf(d[0] * bl[0], d[0] * bh[0]) + fmaxf(d[1] * bl[1], d[1] * bh[1]) + fmaxf(d[2] * bl[2], d[2] * bh[2]);
                        const float spad = pad * (fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]));
                        if (fminf(sp, lo_s) - spad > M->smax[y] || fmaxf(sp, hi_s) + spad < M->smin[y]) keep = 0;
                    }
                    if (keep && axes && (hi < leaf_lo || lo > leaf_hi)) {
                        /* variant: two more separating axes, perpendicular to the source -> chunk direction:
                         * the axis-aligned hull of a diagonal shaft is mostly empty */
                        const float cx = 0.5f * (bl[0] + bh[0]) - px, cy = 0.5f * (bl[1] + bh[1]) - py, cz = 0.5f * (bl[2] + bh[2]) - pz;
                        const float cl = sqrtf(cx * cx + cy * cy + cz * cz);
                        if (cl > 1e-20f) {
                            const float d0[3] = {cx / cl, cy / cl, cz / cl};
                            float u[3] = {-d0[1], d0[0], 0.f};
                            float ul = sqrtf(u[0] * u[0] + u[1] * u[1]);
                            if (ul < 1e-6f) { u[0] = 1.f; u[1] = 0.f; ul = 1.f; }
                            u[0] /= ul; u[1] /= ul;
                            const float w[3] = {d0[1] * u[2] - d0[2] * u[1], d0[2] * u[0] - d0[0] * u[2], d0[0] * u[1] - d0[1] * u[0]};
                            const float *ax[2] = {u, w};
                            const float *b = M->box + 6 * y;
                            for (int t = 0; t < 2 && keep; ++t) {
                                float hmn, hmx, rmn, rmx;
                                proj_box(bl, bh, ax[t], &hmn, &hmx);
                                const float sp2 = ax[t][0] * px + ax[t][1] * py + ax[t][2] * pz;
                                hmn = fminf(hmn, sp2) - 2.f * pad; hmx = fmaxf(hmx, sp2) + 2.f * pad;
                                proj_box(b, b + 3, ax[t], &rmn, &rmx);
                                if (rmn > hmx || rmx < hmn) keep = 0;
                            }
                        }
                    }
                    if (!keep) continue;
                    /* variant: a large record that survives is replaced by its two children, which go through
                     * the filter themselves (the uniform loop is cheaper per record than a phase-C visit and the
                     * filter works on the smaller boxes) */
                    if (expand_leaves > 0 && y < n - 1 && hi - lo + 1 > expand_leaves && (hi < leaf_lo || lo > leaf_hi) && nwork < 500 && nsel < 200) {
                        work[nwork++] = M->right[y];
                        work[nwork++] = M->left[y];
                        continue;
                    }
                    sel[nsel++] = y;
                }
            }
            out[12] += nsel;
            /* batches */
            for (int b0 = 0; b0 < ns; b0 += 32) {
                const int nb = ns - b0 < 32 ? ns - b0 : 32;
                Ray ray[32]; int tleaf[32];
                float tb_lo[3] = {INFINITY, INFINITY, INFINITY}, tb_hi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (int l = 0; l < nb; ++l) {
                    const int s = surv[b0 + l], j = M->leaf_face[s];
                    const double *Pj = M->P + 3 * j;
                    float dx = (float)(Pj[0] - Pi[0]), dy = (float)(Pj[1] - Pi[1]), dz = (float)(Pj[2] - Pi[2]);
                    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                    dx /= nrm; dy /= nrm; dz /= nrm;
                    Ray *r = &ray[l];
                    r->dx = dx; r->dy = dy; r->dz = dz;
                    r->ox = px + 1e-3f * dx; r->oy = py + 1e-3f * dy; r->oz = pz + 1e-3f * dz;
                    r->ix = 1.0f / (fabsf(dx) < 1e-30f ? copysignf(1e-30f, dx) : dx);
                    r->iy = 1.0f / (fabsf(dy) < 1e-30f ? copysignf(1e-30f, dy) : dy);
                    r->iz = 1.0f / (fabsf(dz) < 1e-30f ? copysignf(1e-30f, dz) : dz);
                    r->qx = r->ox * r->ix; r->qy = r->oy * r->iy; r->qz = r->oz * r->iz;
                    r->tmax = (nrm - 1e-3f) * 1.000002f; /* the ray meets its target at the centroid */
                    tleaf[l] = s;
                    for (int k = 0; k < 3; ++k) { tb_lo[k] = fminf(tb_lo[k], (float)Pj[k]); tb_hi[k] = fmaxf(tb_hi[k], (float)Pj[k]); }
                }
                out[0] += 1; out[1] += nb;
                /* variant "horizon": the source's near zone is skipped for the whole batch when every ray leaves
                 * above its horizon; a lane starts the upward walk at the target's zone ancestor when its ray
                 * arrives above the target's horizon */
                int src_skip = zone_leaves > 0, tgt_skip[32] = {0};
                int szlo = 0, szhi = -1;
                if (zone_leaves > 0) {
                    const int sz = zone_of(M, n - 1 + ileaf, zone_leaves);
                    node_range(M, sz, &szlo, &szhi);
                    for (int l = 0; l < nb; ++l) {
                        const int j = M->leaf_face[tleaf[l]];
                        const double *Nj = M->N + 3 * j;
                        const double es = Ni[0] * ray[l].dx + Ni[1] * ray[l].dy + Ni[2] * ray[l].dz;
                        const double et = -(Nj[0] * ray[l].dx + Nj[1] * ray[l].dy + Nj[2] * ray[l].dz);
                        if (!(es > hor[i] + g_margin)) src_skip = 0;
                        tgt_skip[l] = et > hor[j] + g_margin;
                        out[23] += tgt_skip[l];
                    }
                    out[24] += src_skip;
                    for (int k = 0; k < g_src_levels; ++k) { /* larger source zones, smallest first */
                        int ok = src_skip;
                        for (int l = 0; l < nb && ok; ++l) {
                            const double es = Ni[0] * ray[l].dx + Ni[1] * ray[l].dy + Ni[2] * ray[l].dz;
                            if (!(es > g_src_hor[k][i] + g_margin)) ok = 0;
                        }
                        if (!ok) break;
                        const int sz2 = zone_of(M, n - 1 + ileaf, g_src_zone[k]);
                        node_range(M, sz2, &szlo, &szhi);
                        out[25 + k] += 1;
                    }
                }
                /* phase A */
                int sp[32] = {0}, cand[32] = {0}, xref[32];
                int a_iter = 0, a_iter_bf = 0;
                const float bpad = 3e-5f * M->scale + 1e-4f * fmaxf(fmaxf(tb_hi[0] - tb_lo[0], tb_hi[1] - tb_lo[1]), tb_hi[2] - tb_lo[2]) + 2e-3f;
                for (int l = 0; l < nb; ++l) xref[l] = -2;
                for (int e = 0; e < nsel; ++e) {
                    const int y = sel[e]; int lo, hi; node_range(M, y, &lo, &hi);
                    if (src_skip && lo >= szlo && hi <= szhi) continue; /* in the source's zone: below every ray */
                    ++a_iter;
                    /* per-batch filter (variant): drop the record for this batch if it is separated from the hull
                     * of the source and the batch's targets along x, y, z or its slab direction */
                    int bkeep = 1;
                    if (hi < leaf_lo || lo > leaf_hi) {
                        const float *b = M->box + 6 * y, *d = M->sdir + 3 * y;
                        const float hl[3] = {fminf(px, tb_lo[0]) - bpad, fminf(py, tb_lo[1]) - bpad, fminf(pz, tb_lo[2]) - bpad};
                        const float hh[3] = {fmaxf(px, tb_hi[0]) + bpad, fmaxf(py, tb_hi[1]) + bpad, fmaxf(pz, tb_hi[2]) + bpad};
                        if (b[0] > hh[0] || b[3] < hl[0] || b[1] > hh[1] || b[4] < hl[1] || b[2] > hh[2] || b[5] < hl[2]) bkeep = 0;
                        const float sp_ = d[0] * px + d[1] * py + d[2] * pz;
                        const float lo_s = fminf(d[0] * tb_lo[0], d[0] * tb_hi[0]) + fminf(d[1] * tb_lo[1], d[1] * tb_hi[1]) + fminf(d[2] * tb_lo[2], d[2] * tb_hi[2]);
                        const float hi_s = fmaxf(d[0] * tb_lo[0], d[0] * tb_hi[0]) + fmaxf(d[1] * tb_lo[1], d[1] * tb_hi[1]) + fmaxf(d[2] * tb_lo[2], d[2] * tb_hi[2]);
                        const float spad = bpad * (fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]));
                        if (fminf(sp_, lo_s) - spad > M->smax[y] || fmaxf(sp_, hi_s) + spad < M->smin[y]) bkeep = 0;
                    }
                    if (bkeep) ++a_iter_bf;
                    if (batch_filter && !bkeep) continue;
                    for (int l = 0; l < nb; ++l) {
                        if (cref < 0 && tleaf[l] >= lo && tleaf[l] <= hi) { xref[l] = y; continue; }
                        if (node_hit(M, y, &ray[l])) {
                            out[5] += 1;
                            { int lo2, hi2; node_range(M, y, &lo2, &hi2); if (hi2 - lo2 + 1 <= 64) out[21] += 1; }
                            if (y >= n - 1) { cand[l]++; const int lp = y - (n - 1); if (lp == ileaf) out[18] += 1; if (in_zone(M, szone, lp)) out[19] += 1; }
                            else stack[128 * l + sp[l]++] = y;
                        }
                    }
                }
                out[2] += a_iter; out[13] += a_iter_bf;
                /* phase B */
                int b_max = 0;
                for (int l = 0; l < nb; ++l) {
                    const int stop = cref >= 0 ? cref : (xref[l] >= 0 ? xref[l] : -1);
                    int cur = n - 1 + tleaf[l], it = 0;
                    if (tgt_skip[l] && stop >= 0) { /* start above the target's zone if that is still below `stop` */
                        const int tz = zone_of(M, cur, zone_leaves);
                        int zlo, zhi, slo, shi;
                        node_range(M, tz, &zlo, &zhi); node_range(M, stop, &slo, &shi);
                        if (tz != stop && slo <= zlo && zhi <= shi) cur = tz;
                    }
                    while (cur != stop && M->parent[cur] >= 0) {
                        const int y = sibling(M, cur);
                        cur = M->parent[cur];
                        ++it;
                        if (node_hit(M, y, &ray[l])) {
                            out[22] += 1;
                            if (y >= n - 1) { cand[l]++; const int lp = y - (n - 1); if (in_zone(M, szone, lp)) out[19] += 1; if (in_zone(M, zone64(M, n - 1 + tleaf[l]), lp)) out[20] += 1; }
                            else stack[128 * l + sp[l]++] = y | (1 << 30);
                        }
                    }
                    out[6] += it;
                    if (it > b_max) b_max = it;
                }
                out[3] += b_max;
                /* phase C */
                int c_max = 0, f_max = 0;
                int useen[4096], nseen = 0; /* distinct nodes any ray of the batch visits (a packet traversal's count) */
                for (int l = 0; l < nb; ++l) {
                    int it = 0;
                    while (sp[l] > 0) {
                        int node = stack[128 * l + --sp[l]];
                        const int side = (node >> 30) & 1;
                        node &= ~(1 << 30);
                        const int tz = zone64(M, n - 1 + tleaf[l]);
                        for (;;) {
                            ++it; out[16 + side] += 1;
                            { int k = 0; while (k < nseen && useen[k] != node) ++k;
                              if (k == nseen && nseen < 4096) { useen[nseen++] = node; } }
                            { int sz = M->last[node] - M->first[node] + 1, lg = 0; while ((1 << (lg + 1)) <= sz) ++lg; g_chist[lg < 20 ? lg : 19] += 1; }
                            const int lc = M->left[node], rc = M->right[node];
                            const int h0 = node_hit(M, lc, &ray[l]), h1 = node_hit(M, rc, &ray[l]);
                            if (h0 && lc >= n - 1 && lc - (n - 1) != tleaf[l]) { cand[l]++; if (in_zone(M, szone, lc - (n - 1))) out[19] += 1; if (in_zone(M, tz, lc - (n - 1))) out[20] += 1; }
                            if (h1 && rc >= n - 1 && rc - (n - 1) != tleaf[l]) { cand[l]++; if (in_zone(M, szone, rc - (n - 1))) out[19] += 1; if (in_zone(M, tz, rc - (n - 1))) out[20] += 1; }
                            const int i0 = h0 && lc < n - 1, i1 = h1 && rc < n - 1;
                            if (i0 && i1) stack[128 * l + sp[l]++] = rc | (side << 30);
                            if (i0) node = lc; else if (i1) node = rc; else break;
                        }
                    }
                    out[7] += it; out[8] += cand[l];
                    if (it > c_max) c_max = it;
                    if (cand[l] > f_max) f_max = cand[l];
                }
                out[4] += c_max; out[9] += f_max; out[29] += nseen;
            }
        }
    }
    free(list); free(surv); free(stack);
}

/* ---------------------------------------------------------------------------------------------
 * Exact horizons and a brute-force check of the "horizon skip" (profiles/r01b_k4_model.md): for every
 * sampled ray that clears the horizon of its source (target) zone, every triangle of that zone is put
 * through the same float32 Pluecker test the kernel uses; a hit in [0, t_target] would be an occluder the
 * skip loses.  The ray is set up in float32 exactly as trace.cuh does (compile without FMA contraction).
 * ------------------------------------------------------------------------------------------- */
typedef struct { float ox, oy, oz, dx, dy, dz; } FRay;

static int pluecker(const FRay *r, float tnear, float tfar, const float *p0, const float *p1, const float *p2, float *t_out) {
    const float v0x = p0[0] - r->ox, v0y = p0[1] - r->oy, v0z = p0[2] - r->oz;
    const float v1x = p1[0] - r->ox, v1y = p1[1] - r->oy, v1z = p1[2] - r->oz;
    const float v2x = p2[0] - r->ox, v2y = p2[1] - r->oy, v2z = p2[2] - r->oz;
    const float e0x = v2x - v0x, e0y = v2y - v0y, e0z = v2z - v0z;
    const float e1x = v0x - v1x, e1y = v0y - v1y, e1z = v0z - v1z;
    const float e2x = v1x - v2x, e2y = v1y - v2y, e2z = v1z - v2z;
#define MSUB(a, b, c) fmaf((a), (b), -(c))
#define DOT3(ax, ay, az, bx, by, bz) fmaf((ax), (bx), fmaf((ay), (by), (az) * (bz)))
    float sx, sy, sz, cx, cy, cz;
    sx = v2x + v0x; sy = v2y + v0y; sz = v2z + v0z;
    cx = MSUB(e0y, sz, e0z * sy); cy = MSUB(e0z, sx, e0x * sz); cz = MSUB(e0x, sy, e0y * sx);
    const float U = DOT3(cx, cy, cz, r->dx, r->dy, r->dz);
    sx = v0x + v1x; sy = v0y + v1y; sz = v0z + v1z;
    cx = MSUB(e1y, sz, e1z * sy); cy = MSUB(e1z, sx, e1x * sz); cz = MSUB(e1x, sy, e1y * sx);
    const float V = DOT3(cx, cy, cz, r->dx, r->dy, r->dz);
    sx = v1x + v2x; sy = v1y + v2y; sz = v1z + v2z;
    cx = MSUB(e2y, sz, e2z * sy); cy = MSUB(e2z, sx, e2x * sz); cz = MSUB(e2x, sy, e2y * sx);
    const float W = DOT3(cx, cy, cz, r->dx, r->dy, r->dz);
    const float UVW = (U + V) + W;
    const float eps = 1.1920929e-7f * fabsf(UVW);
    const float mn = fminf(fminf(U, V), W), mx = fmaxf(fmaxf(U, V), W);
    if (!(mn >= -eps || mx <= eps)) return 0;
    const float ab_x = e0z * e1y, ab_y = e0x * e1z, ab_z = e0y * e1x;
    const float bc_x = e1z * e2y, bc_y = e1x * e2z, bc_z = e1y * e2x;
    const float cabx = MSUB(e0y, e1z, ab_x), caby = MSUB(e0z, e1x, ab_y), cabz = MSUB(e0x, e1y, ab_z);
    const float cbcx = MSUB(e1y, e2z, bc_x), cbcy = MSUB(e1z, e2x, bc_y), cbcz = MSUB(e1x, e2y, bc_z);
    const float Ngx = fabsf(ab_x) < fabsf(bc_x) ? cabx : cbcx;
    const float Ngy = fabsf(ab_y) < fabsf(bc_y) ? caby : cbcy;
    const float Ngz = fabsf(ab_z) < fabsf(bc_z) ? cabz : cbcz;
    const float dn = DOT3(Ngx, Ngy, Ngz, r->dx, r->dy, r->dz), den = dn + dn;
    const float Tn = DOT3(v0x, v0y, v0z, Ngx, Ngy, Ngz), T = Tn + Tn;
    if (den == 0.0f) return 0;
    const float t = T / den;
    if (!(tnear <= t && t <= tfar)) return 0;
    *t_out = t;
    return 1;
#undef MSUB
#undef DOT3
}

/* sup over the triangle (a, b, c) of n.(x - o)/|x - o|, and the smallest |x - o| */
static void tri_elevation(const double *o, const double *n, const float *a, const float *b, const float *c,
                          double *sup, double *rmin) {
    const float *v[3] = {a, b, c};
    double best = -INFINITY, rm = INFINITY;
    double q[3][3];
    for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 3; ++d) q[k][d] = (double)v[k][d] - o[d];
    for (int k = 0; k < 3; ++k) { /* vertices and edges */
        const double *A = q[k], *B = q[(k + 1) % 3];
        const double la = sqrt(A[0] * A[0] + A[1] * A[1] + A[2] * A[2]);
        if (la < rm) rm = la;
        if (la > 0) { const double e = (n[0] * A[0] + n[1] * A[1] + n[2] * A[2]) / la; if (e > best) best = e; }
        double E[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
        /* f(s) = (alpha + beta s) / sqrt(gamma + 2 delta s + eps s^2), s in [0, 1]; stationary point is linear */
        const double alpha = n[0] * A[0] + n[1] * A[1] + n[2] * A[2], beta = n[0] * E[0] + n[1] * E[1] + n[2] * E[2];
        const double gamma = la * la, delta = A[0] * E[0] + A[1] * E[1] + A[2] * E[2], eps = E[0] * E[0] + E[1] * E[1] + E[2] * E[2];
        const double den = beta * delta - alpha * eps;
        if (den != 0) {
            const double s = (alpha * delta - beta * gamma) / den;
            if (s > 0 && s < 1) {
                const double l2 = gamma + 2 * delta * s + eps * s * s;
                if (l2 > 0) { const double e = (alpha + beta * s) / sqrt(l2); if (e > best) best = e; }
            }
        }
        /* closest point of the edge to o */
        if (eps > 0) {
            double s = -delta / eps; s = s < 0 ? 0 : (s > 1 ? 1 : s);
            const double l2 = gamma + 2 * delta * s + eps * s * s;
            if (l2 >= 0 && sqrt(l2) < rm) rm = sqrt(l2);
        }
    }
    /* interior: the ray o + t n (t > 0) pierces the triangle -> elevation 1 there; also the foot of o on the plane */
    {
        const double *A = q[0];
        const double E1[3] = {q[1][0] - A[0], q[1][1] - A[1], q[1][2] - A[2]}, E2[3] = {q[2][0] - A[0], q[2][1] - A[1], q[2][2] - A[2]};
        const double m[3] = {E1[1] * E2[2] - E1[2] * E2[1], E1[2] * E2[0] - E1[0] * E2[2], E1[0] * E2[1] - E1[1] * E2[0]};
        const double mm = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
        if (mm > 0) {
            const double mA = m[0] * A[0] + m[1] * A[1] + m[2] * A[2];
            for (int pass = 0; pass < 2; ++pass) { /* pass 0: along n; pass 1: along the plane normal (closest point) */
                const double *dir = pass ? m : n;
                const double md = m[0] * dir[0] + m[1] * dir[1] + m[2] * dir[2];
                if (md == 0) continue;
                const double t = mA / md;
                if (!pass && t <= 0) continue;
                const double X[3] = {t * dir[0] - A[0], t * dir[1] - A[1], t * dir[2] - A[2]};
                /* barycentric coordinates of X in (E1, E2) */
                const double d11 = E1[0] * E1[0] + E1[1] * E1[1] + E1[2] * E1[2], d12 = E1[0] * E2[0] + E1[1] * E2[1] + E1[2] * E2[2];
                const double d22 = E2[0] * E2[0] + E2[1] * E2[1] + E2[2] * E2[2];
                const double x1 = X[0] * E1[0] + X[1] * E1[1] + X[2] * E1[2], x2 = X[0] * E2[0] + X[1] * E2[1] + X[2] * E2[2];
                const double det = d11 * d22 - d12 * d12;
                if (det <= 0) continue;
                const double u = (x1 * d22 - x2 * d12) / det, w = (x2 * d11 - x1 * d12) / det;
                if (u >= -1e-9 && w >= -1e-9 && u + w <= 1 + 1e-9) {
                    if (!pass) best = 1.0;
                    else { const double l = fabs(t) * sqrt(mm); if (l < rm) rm = l; }
                }
            }
        }
    }
    *sup = best; *rmin = rm;
}

/* hor[f] with the perturbation term: the traced ray is the float32 image of the ideal one (origin and
 * direction off by a few ulps of the largest coordinate); a point at distance r is displaced by at most
 * pert / r in the sine.  c_pert in ulps. */
static int g_formula = 0; /* 0: r = rmin - 1e-3 (first version); 1: r = rmin, twice the perturbation */
void k4_set_formula(int f) { g_formula = f; }
void k4_horizons_exact(const Model *M, int zone_leaves, double c_pert, float *hor) {
    const int n = M->n;
    const double pert = c_pert * 1.1920929e-7 * M->scale;
    for (int f = 0; f < n; ++f) {
        const int z = zone_of(M, n - 1 + M->face_leaf[f], zone_leaves);
        int lo, hi;
        if (z >= n - 1) { lo = hi = z - (n - 1); } else { lo = M->first[z]; hi = M->last[z]; }
        double best = -1.0;
        for (int k = lo; k <= hi; ++k) {
            const int g = M->leaf_face[k];
            if (g == f) continue;
            double sup, rmin;
            tri_elevation(M->P + 3 * f, M->N + 3 * f, M->V + 3 * M->F[3 * g], M->V + 3 * M->F[3 * g + 1], M->V + 3 * M->F[3 * g + 2], &sup, &rmin);
            double r, e;
            if (g_formula == 0) {
                /* the ray starts 1e-3 along itself: a zone point can be that much closer to the origin than to p_f */
                r = rmin - 1.0e-3 * 1.001;
                e = r > pert ? sup + pert / r : INFINITY;
            } else {
                /* a point x of the triangle within delta of a point y of the ray: the unit directions from p_f
                 * differ by at most 2 delta / max(|x - p_f|, |y - p_f|) <= 2 delta / rmin (either end of the ray) */
                r = rmin;
                e = r > 0 ? sup + 2.0 * pert / r : INFINITY;
            }
            if (e > best) best = e;
        }
        hor[f] = (float)best;
        if ((double)hor[f] < best) hor[f] = nextafterf(hor[f], INFINITY);
    }
}

/* out: 0 rays, 1 rays clearing the source horizon, 2 clearing the target horizon, 3 VIOLATIONS at the source
 * end (a zone triangle hit in [0, t_target]), 4 violations at the target end, 5 Pluecker tests done,
 * 6 rays that miss their own target */
void k4_check_horizon(const Model *M, int nrows, const int *rows, int zone_leaves, const float *hor, double eps, double *out) {
    const int n = M->n;
    for (int ri = 0; ri < nrows; ++ri) {
        const int i = rows[ri];
        const double *Pi = M->P + 3 * i, *Ni = M->N + 3 * i;
        const int sz = zone_of(M, n - 1 + M->face_leaf[i], zone_leaves);
        int slo, shi; node_range(M, sz, &slo, &shi);
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double *Pj = M->P + 3 * j, *Nj = M->N + 3 * j;
            const double ddx = Pj[0] - Pi[0], ddy = Pj[1] - Pi[1], ddz = Pj[2] - Pi[2];
            double a = Ni[0] * ddx + Ni[1] * ddy + Ni[2] * ddz, b = -(Nj[0] * ddx + Nj[1] * ddy + Nj[2] * ddz);
            a = a > 0 ? a : 0; b = b > 0 ? b : 0;
            if (!((float)(a * b) > (float)eps)) continue;
            /* setup_ray<float> of trace.cuh */
            const float pix = (float)Pi[0], piy = (float)Pi[1], piz = (float)Pi[2];
            const float fdx = (float)Pj[0] - pix, fdy = (float)Pj[1] - piy, fdz = (float)Pj[2] - piz;
            const float nrm = sqrtf((fdx * fdx + fdy * fdy) + fdz * fdz);
            const float feps = 1e-3f;
            if (!(nrm > feps)) continue;
            FRay r;
            r.dx = fdx / nrm; r.dy = fdy / nrm; r.dz = fdz / nrm;
            r.ox = pix + feps * r.dx; r.oy = piy + feps * r.dy; r.oz = piz + feps * r.dz;
            float tj;
            const float *t0 = M->V + 3 * M->F[3 * j], *t1 = M->V + 3 * M->F[3 * j + 1], *t2 = M->V + 3 * M->F[3 * j + 2];
            out[0] += 1;
            if (!pluecker(&r, 0.f, INFINITY, t0, t1, t2, &tj)) { out[6] += 1; continue; }
            const double es = Ni[0] * r.dx + Ni[1] * r.dy + Ni[2] * r.dz, et = -(Nj[0] * r.dx + Nj[1] * r.dy + Nj[2] * r.dz);
            if (es > hor[i]) {
                out[1] += 1;
                for (int k = slo; k <= shi; ++k) {
                    const int g = M->leaf_face[k];
                    if (g == i || g == j) continue;
                    float t;
                    out[5] += 1;
                    if (pluecker(&r, 0.f, tj, M->V + 3 * M->F[3 * g], M->V + 3 * M->F[3 * g + 1], M->V + 3 * M->F[3 * g + 2], &t)) out[3] += 1;
                }
            }
            if (et > hor[j]) {
                out[2] += 1;
                const int tz = zone_of(M, n - 1 + M->face_leaf[j], zone_leaves);
                int tlo, thi; node_range(M, tz, &tlo, &thi);
                for (int k = tlo; k <= thi; ++k) {
                    const int g = M->leaf_face[k];
                    if (g == j || g == i) continue;
                    float t;
                    out[5] += 1;
                    if (pluecker(&r, 0.f, tj, M->V + 3 * M->F[3 * g], M->V + 3 * M->F[3 * g + 1], M->V + 3 * M->F[3 * g + 2], &t)) out[4] += 1;
                }
            }
        }
    }
}
