/*
 * fluxb200.h -- C ABI of libfluxb200.so, the B200 (sm_100a) implementation of
 * fluxpy's form-factor assembly hot path.
 *
 * Plain C: opaque handle, raw pointers and sizes, int status returns, no
 * exceptions and no torch / CUDA types cross this boundary.  Every entry point
 * names the reference interface (relative to the fluxpy source tree) it
 * replaces.  The shape of the interface mirrors the reference's one existing
 * native ABI, src/flux/cgal/aabb_wrapper.h:8-16 (opaque struct, alloc / init /
 * dealloc, batch queries over index arrays).
 *
 * Conventions
 *   - status: 0 = ok, non-zero = failure; fluxb200_last_error() gives the text
 *     (thread-local, valid until the next call on the same thread).
 *   - all array arguments are C-contiguous HOST buffers owned by the caller
 *     unless a parameter is documented as a device pointer.
 *   - dtype_code: FLUXB200_F32 / FLUXB200_F64 = dtype of V, P, N, A and of the
 *     returned CSR `data` (shape_model.dtype, src/flux/form_factors.py:32-37).
 *   - face index arrays I, J are int64, arbitrary order, may repeat; NULL means
 *     arange(num_faces) (form_factors.py:18-21).
 *   - calls on one handle are serialised on the handle's own CUDA stream; use
 *     one handle per device / per host thread.
 */
#ifndef FLUXB200_H
#define FLUXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUXB200_F32 0
#define FLUXB200_F64 1

#define FLUXB200_ABI_VERSION 7

#define FLUXB200_OK 0
#define FLUXB200_ERROR 1
#define FLUXB200_OVERFLOW 2 /* fluxb200_ff_assemble: `capacity` too small, see stats.nnz */

typedef struct fluxb200_mesh fluxb200_mesh; /* cf. struct cgal_aabb, aabb_wrapper.h:6 */
typedef struct fluxb200_csr fluxb200_csr;   /* a CSR slab resident in device memory */

/* Counters and device timings of the last assembly on a handle. */
typedef struct fluxb200_ff_stats {
    int64_t pairs_all;    /* m * n */
    int64_t pairs_tested; /* pairs surviving the cull form_factors.py:52 (one ray each) */
    int64_t nnz;          /* stored entries */
    float ms_prepare;     /* index-set upload, sort of J by leaf order, gathers */
    float ms_trace;       /* fused cull + occlusion kernel (K4) */
    float ms_scan;        /* row counts -> indptr (K5) */
    float ms_fill;        /* order-preserving CSR fill (K6) */
    float ms_d2h;         /* device -> host copies of the CSR arrays */
    int32_t trace_launches;
    int32_t kernel_launches; /* all kernels launched by the last count+fill */
    int64_t h2d_bytes;    /* bytes copied host -> device by the call (index sets) */
    int64_t d2h_bytes;    /* bytes copied device -> host by the call (CSR arrays or visibility words, counts) */
} fluxb200_ff_stats;

typedef struct fluxb200_bvh_info {
    int64_t num_faces;
    int64_t num_nodes;     /* internal (two-child) nodes: num_faces - 1 */
    int32_t num_top_nodes; /* nodes staged in shared memory by the trace kernel */
    int32_t max_depth;
    float ms_build;        /* device time of the last LBVH build */
    float scene_lo[3], scene_hi[3];
} fluxb200_bvh_info;

const char *fluxb200_last_error(void);
int fluxb200_abi_version(void);
int fluxb200_device_count(int *count);

/* ---- scene / shape model -------------------------------------------------- */

/* Replaces EmbreeTrimeshShapeModel._make_scene (src/flux/shape.py:296-344) and
 * cgal_aabb_alloc + cgal_aabb_init_from_trimesh (aabb_wrapper.cpp:36-58):
 * uploads V (nv x 3, dtype_code) and F (nf x 3 int64), computes centroids,
 * unit normals and areas on the device in the array dtype with NumPy's
 * operation order (shape.py:16-45), converts the vertices to the float32
 * buffer the ray tracer uses (shape.py:319-325) and builds the LBVH. */
int fluxb200_mesh_create(const void *V, size_t nv, const int64_t *F, size_t nf, int dtype_code,
                         int device, fluxb200_mesh **out);
/* cgal_aabb_dealloc (aabb_wrapper.cpp:60-63) */
int fluxb200_mesh_destroy(fluxb200_mesh *mesh);

/* TrimeshShapeModel keeps P, N, A as mutable public attributes
 * (shape.py:104-106; the reference tests flip N in place,
 * tests/test_form_factors.py:33-34,48): the host mirror re-sends them before
 * every query.  Each pointer may be NULL to keep the device copy.
 * P, N: nf x 3; A: nf. */
int fluxb200_mesh_set_face_data(fluxb200_mesh *mesh, const void *P, const void *N, const void *A);
/* device-computed get_centroids / get_surface_normals_and_face_areas
 * (shape.py:16-45) back to the host; any pointer may be NULL */
int fluxb200_mesh_get_face_data(fluxb200_mesh *mesh, void *P, void *N, void *A);

/* Rebuild the LBVH (Morton codes, radix sort, hierarchy, refit, layout).  Done
 * once by fluxb200_mesh_create; exported so it can be timed. */
int fluxb200_bvh_build(fluxb200_mesh *mesh);
int fluxb200_bvh_info_get(fluxb200_mesh *mesh, fluxb200_bvh_info *info);
/* Debug/test export of the flattened tree: nodes = num_nodes x 24 floats, two
 * children of 12 floats each: (lo.xyz, ref-as-int-bits, hi.xyz, slab_min,
 * slab_dir.xyz, slab_max); ref >= 0 internal node, ref < 0 triangle ~ref in
 * leaf order.  leaf_face = nf int32 (face id of every leaf position). */
int fluxb200_bvh_export(fluxb200_mesh *mesh, float *nodes, int32_t *leaf_face);

/* ---- get_form_factor_matrix (src/flux/form_factors.py:11-72) -------------- */

/* Pass 1: per row of I, cull (form_factors.py:46-52), trace the survivors
 * (shape.py:349-398 semantics) and count the stored entries.
 * row_counts: int64[m] (host, may be NULL).  The visibility bits stay on the
 * device for fluxb200_ff_fill; a second count call discards them. */
int fluxb200_ff_count(fluxb200_mesh *mesh, const int64_t *I, size_t m, const int64_t *J, size_t n,
                      double eps, int64_t *row_counts, fluxb200_ff_stats *stats);
/* Pass 2: write the CSR arrays of the last fluxb200_ff_count.
 * index_width: 4 (int32) or 8 (int64) bytes for indices and indptr.
 * destination: 0 = host buffers (indptr m+1, indices nnz, data nnz);
 *              1 = caller-supplied DEVICE buffers of the same sizes;
 *              2 = library-owned device buffers (see fluxb200_ff_device_csr),
 *                  the three pointers are ignored.
 * Columns are positions into J, ascending within each row (form_factors.py:52,69). */
int fluxb200_ff_fill(fluxb200_mesh *mesh, int index_width, int destination, void *indptr,
                     void *indices, void *data, fluxb200_ff_stats *stats);
/* One-call streaming form of the two passes above (what get_form_factor_matrix
 * uses): rows are processed in sub-slabs; while sub-slab k+1 is traced on the
 * handle's stream, sub-slab k is filled and copied out on a second stream.
 * destination 0: host buffers holding `capacity` entries (indices, data) and
 *                m+1 indptr entries; page-locked buffers (fluxb200_host_alloc)
 *                make the copies asynchronous.  Returns FLUXB200_OVERFLOW with
 *                stats->nnz = entries needed when capacity is too small
 *                (nothing usable was written; call again with more room).
 * destination 2: library-owned device buffers (grown as needed; capacity <= 0
 *                uses the previous size as the first guess).
 * destination 3: as 0, for ORDINARY (pageable) host buffers: the values pass through page-locked staging slots
 *                inside the library and host threads move them on (the column indices are written by host
 *                threads in either case).  Page-locking a multi-gigabyte result costs several times what
 *                assembling it does (about 0.5 s per GB measured); this mode pays a memcpy instead.
 * row_counts: optional int64[m] (host). */
int fluxb200_ff_assemble(fluxb200_mesh *mesh, const int64_t *I, size_t m, const int64_t *J, size_t n,
                         double eps, int index_width, int destination, void *indptr, void *indices,
                         void *data, int64_t capacity, int64_t *row_counts, fluxb200_ff_stats *stats);
/* Page-locked host memory for the CSR outputs (cudaHostAlloc / cudaFreeHost). */
int fluxb200_host_alloc(size_t bytes, void **ptr);
int fluxb200_host_free(void *ptr);
/* Device pointers of the library-owned CSR of the last destination==2 fill / assemble. */
int fluxb200_ff_device_csr(fluxb200_mesh *mesh, void **indptr, void **indices, void **data,
                           int64_t *nnz);

/* ---- device-resident slab: products for the radiosity iteration (next row N2) -- */

/* Take ownership of the library-owned device CSR of the last destination==2
 * assembly (rows = the I of that call).  The mesh handle can assemble again. */
int fluxb200_ff_detach_csr(fluxb200_mesh *mesh, fluxb200_csr **out);
int fluxb200_csr_destroy(fluxb200_csr *csr);
int fluxb200_csr_info(fluxb200_csr *csr, int64_t *m, int64_t *n, int64_t *nnz, int *dtype_code,
                      int *index_width, float *last_ms);
/* Download (what scipy.sparse.save_npz would store); indptr/indices in the slab's index width. */
int fluxb200_csr_to_host(fluxb200_csr *csr, void *indptr, void *indices, void *data);
/* y = E + FF @ (rho * x) on the slab: one step of _solve_radiosity_jacobi_right
 * (src/flux/solve.py:36-45); with E = NULL, rho = 1 it is the plain product FF @ x
 * of src/flux/model.py:17.  All vectors are DEVICE pointers to float64: E[m] or
 * NULL, rho[n] or NULL (then rho_scalar is used), x[n], y[m].  When diffmax_host
 * is not NULL it receives max_r |y[r] - x[row_offset + r]| (solve.py:41). */
int fluxb200_csr_jacobi_step(fluxb200_csr *csr, const double *E_dev, const double *rho_dev,
                             double rho_scalar, const double *x_dev, double *y_dev,
                             double *diffmax_host, int64_t row_offset);

/* ---- feeding the hierarchical compression from the resident slab (next row N3) -- */

/* out = src[rows, :][:, cols]: the CSR slicing FormFactor2dTreeBlock does on the
 * parent's matrix (src/flux/compressed_form_factors.py:562).  rows: local row
 * numbers of the slab (any order, may repeat); cols: column positions of src,
 * without repeats; the new column ids are positions into `cols`, entries keep
 * the source order (ascending when cols is ascending). */
int fluxb200_csr_extract(fluxb200_csr *src, const int64_t *rows, size_t mr, const int64_t *cols,
                         size_t nc, fluxb200_csr **out);
/* Thin dense products, DEVICE pointers to row-major float64, 1 <= k <= 32:
 * transpose == 0:  Y[m x k] = A @ X[n x k];   transpose != 0:  Y[n x k] = A^T @ X[m x k].
 * The two products of a randomised range finder for the SVD leaves
 * (src/flux/compressed_form_factors.py:388-405, src/flux/linalg.py:8-50). */
int fluxb200_csr_matmat(fluxb200_csr *csr, const double *X_dev, int k, double *Y_dev, int transpose);

/* ---- TrimeshShapeModel hooks (src/flux/shape.py:129-188, 349-421) ---------- */

/* _get_visibility(I, J) -> bool[m, n]: Embree semantics (masked pairs closer
 * than 1e-3 are visible, otherwise visible iff the closest hit of the ray
 * p_i -> p_j is triangle j); replaces cgal_aabb_test_face_to_face_vis batched
 * by AABB.test_face_to_face_vis_MN (src/flux/cgal/aabb.pyx:58-70). */
int fluxb200_visibility(fluxb200_mesh *mesh, const int64_t *I, size_t m, const int64_t *J,
                        size_t n, uint8_t *vis);
/* _is_occluded(I, D): origin P[I] + 1e-3*N[I] (shape.py:400-421).
 * mode 0: D is one vector (3), out[m];
 * mode 1: D is m x 3, row p belongs to I[p], out[m]        (Embree backend);
 * mode 2: D is nd x 3, all directions for every face, out[m x nd]
 *         (ray_from_centroid_is_occluded_2d, aabb.pyx:77-87). */
int fluxb200_is_occluded(fluxb200_mesh *mesh, const int64_t *I, size_t m, const void *D, size_t nd,
                         int mode, uint8_t *occluded);
/* _intersect1(x, d): closest hit of one ray (double in, as cgal_aabb_intersect1,
 * aabb_wrapper.cpp:109-129).  *hit = 1 and face / t / xt filled when something is hit. */
int fluxb200_intersect1(fluxb200_mesh *mesh, const double x[3], const double d[3], int *hit,
                        int64_t *face, double *t, double xt[3]);

/* ---- test hooks ------------------------------------------------------------ */

/* Same as fluxb200_visibility but every ray is tested against every triangle
 * (no BVH): validates the tree and the conservative box test on the device. */
int fluxb200_visibility_bruteforce(fluxb200_mesh *mesh, const int64_t *I, size_t m,
                                   const int64_t *J, size_t n, uint8_t *vis);

/* ---- multi-GPU helpers ----------------------------------------------------- */

/* Contiguous row slabs for `nranks` devices (rows shard):
 * starts[nranks+1].  weights (int64[m], e.g. row counts of a previous pass) may
 * be NULL for an equal split. */
int fluxb200_slab_plan(size_t m, int nranks, const int64_t *weights, int64_t *starts);

/* Host helper of the copy-out of fluxb200_ff_assemble (destination 0): the column
 * indices of a CSR row (form_factors.py:52, 69) are the positions of the set bits
 * of the row's visibility words in the caller's column order; the words cross
 * PCIe, this writes the positions (ascending; int32 or int64) and their number.
 * `out` must hold popcount(words) entries.  Exported for the host-logic tests. */
int fluxb200_expand_words(const uint32_t *words, size_t nwords, int index_width, void *out,
                          int64_t *count);

/* The same for `mr` rows at once on `nthreads` host threads (0 = automatic): words is mr x nwords,
 * row r goes to indices[offs[r] .. offs[r+1]).  A row whose set bits do not match its length is
 * reported as an error and left unwritten.  This is the worker pool fluxb200_ff_assemble uses. */
int fluxb200_expand_rows(const uint32_t *words, size_t nwords, size_t mr, const int64_t *offs,
                         int index_width, void *indices, int nthreads);

/* The CUDA stream (cudaStream_t) all work of this handle is enqueued on, so a
 * caller can bracket calls with its own events. */
int fluxb200_mesh_stream(fluxb200_mesh *mesh, void **stream);
/* Tunables: "top_nodes" (BVH nodes staged in shared memory), "slab_limit"
 * (largest subtree, in faces, that gets a fitted slab), "blocks_per_sm" (persistent CTAs per SM of
 * the trace kernel), "shaft_filter" (0 disables the per-unit record filter: A/B checks),
 * "sub_rows" (rows per sub-slab of fluxb200_ff_assemble), "fill_rows" (rows per CTA of the
 * CSR fill's un-permute kernel: 0 = as many as fit in shared memory, -1 = no shared memory),
 * "host_expand" (1: fluxb200_ff_assemble ships visibility words to the host and host threads
 * write the column indices; 0: the indices themselves are copied), "host_threads" (0 = automatic),
 * "horizon_skip" (1: the trace kernel skips a face's near zone for rays that clear its horizon --
 * exact, see csrc/horizon.cuh; default 1 since it was measured on a B200: trace kernel -30 %), "horizon_zone"
 * (leaves per near zone, default 1023; the target end of the skip needs zones below the 1024-column chunk),
 * "colset_cache" (prepared column sets kept per handle, default 8, 0 = none: the sort of J by BVH position and
 * the gathers are reused by later calls with the same J while P, N, A and the tree are unchanged -- the 16 / 64
 * root-block calls of CompressedFormFactorMatrix share 4 / 8 column parts),
 * "trace_variant" (2: second-generation trace kernel, csrc/trace2.cuh, the default;
 * 1: the first-generation kernel with per-lane stacks -- identical results, kept as the A/B reference),
 * "pipeline_ramp" (host output of fluxb200_ff_assemble; 1, the default: short first sub-slabs that double up to
 * "sub_rows", and the fill of a sub-slab completes before the next one is traced, so that the copy-out -- the
 * longest of the three overlapped stages -- starts ~2 ms into the call; 0: equal sub-slabs, fill and next trace
 * left to the hardware scheduler). */
int fluxb200_set_option(fluxb200_mesh *mesh, const char *name, int64_t value);
/* Counters of the last assembly's trace launches: out[0] rays traced (= stats.pairs_tested), and with
 * "horizon_skip" on: out[1] 32-ray batches, out[2] batches walked without the records of the source
 * face's near zone, out[3] rays whose upward walk started at the target's zone node; "trace_variant" 2:
 * out[4], out[5] unused (counters of a dropped work-queue variant),
 * out[6] rays handed to the follow-up kernel because a shared-memory stack / list was full; out[7] calls on this
 * handle (cumulative) that found their column set J already prepared ("colset_cache"). */
int fluxb200_trace_counters(fluxb200_mesh *mesh, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* FLUXB200_H */
