/* iglb200 — C ABI of the B200-native point-vs-mesh query path.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, no C++/torch
 * types.  libigl itself has no FFI for this path — its interface is the set of
 * C++ function templates cited beside each entry point below (paths relative to
 * /root/reference/include/igl/).  include/iglb200/ *.h re-creates those
 * templates on top of this ABI; INTEGRATION.md shows the few lines a libigl
 * maintainer would add to route igl::signed_distance through it.
 *
 * Conventions
 *  - Matrices are passed as (pointer, is_f64, rows, row_stride, col_stride),
 *    strides in ELEMENTS, which covers Eigen column-major (rs=1, cs=rows),
 *    row-major (rs=cols, cs=1) and Map<..., Stride<>> views.
 *  - All pointers in the non-_dev entry points are HOST pointers (pageable or
 *    pinned); the library copies to/from the device(s) holding the handle.  Batches are
 *    cut into chunks and streamed by worker threads, so uploads, kernels and downloads
 *    overlap; pageable arrays go through a page-locked staging ring.
 *  - The caller owns every array (AABB.h:20-21: "The mesh (V,Ele) is stored and
 *    managed by the caller").  A mesh handle COPIES V and F at creation, so the
 *    inputs may be freed afterwards (unlike UT_SolidAngle, which keeps raw
 *    pointers: FastWindingNumberForSoups.h:6232-6234).
 *  - Return value 0 = ok, negative = error; iglb200_last_error() gives the
 *    message (thread local).  There is NO CPU fallback: without a CUDA device
 *    every compute entry point returns IGLB200_ERR_CUDA.
 *  - Mesh and cloud handles are immutable after creation and may be queried
 *    from several host threads at once (the reference structures are
 *    const-queried from igl::parallel_for threads: AABB.h:569-574).  A grid
 *    handle carries state (its values S): use it from one thread at a time.
 */
#ifndef IGLB200_H
#define IGLB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IGLB200_OK 0
#define IGLB200_ERR_ARG (-1)
#define IGLB200_ERR_CUDA (-2)
#define IGLB200_ERR_OOM (-3)
#define IGLB200_ERR_UNSUPPORTED (-4)
#define IGLB200_ERR_INTERNAL (-5)

/* igl::SignedDistanceType, signed_distance.h:20-34.  All five run on the device: UNSIGNED,
 * FAST_WINDING_NUMBER, PSEUDONORMAL (mesh created with IGLB200_MESH_PSEUDONORMALS) and
 * WINDING_NUMBER = DEFAULT in 3-D (mesh created with IGLB200_MESH_WINDING_NUMBER); a mesh that
 * lacks the tables a sign type needs returns IGLB200_ERR_ARG (no silent fallback). */
#define IGLB200_SIGN_PSEUDONORMAL 0
#define IGLB200_SIGN_WINDING_NUMBER 1
#define IGLB200_SIGN_DEFAULT 2
#define IGLB200_SIGN_UNSIGNED 3
#define IGLB200_SIGN_FAST_WINDING_NUMBER 4

typedef struct iglb200_mesh iglb200_mesh;   /* AABB<MatrixXd,3> + FastWindingNumberBVH in one */
typedef struct iglb200_cloud iglb200_cloud; /* octree + CM/R/EC of the point-cloud FWN */

const char *iglb200_last_error(void);
const char *iglb200_version(void);
int iglb200_device_count(void);          /* <0 on error, 0 when no GPU */
int iglb200_set_device(int device);      /* device used by the calling thread */
/* One process, several GPUs.  iglb200_init(n) makes devices 0..n-1 the process's GROUP: every mesh created
 * afterwards is built once (on the calling thread's device) and its arena is replicated to all devices of the
 * group with ncclBroadcast over NVLink (NCCL is loaded at run time; peer-to-peer DMA when it cannot be), and every
 * host-pointer query on such a mesh is dealt — chunk by chunk, one atomic counter — to all of those devices, each
 * of which DMA-writes its chunks straight into the caller's output arrays.  This is the loop igl::signed_distance
 * runs over igl::parallel_for threads (signed_distance.cpp:90-138; AABB.cpp:1241-1265;
 * fast_winding_number.cpp:429-436): queries are independent, the structures read-only.
 * iglb200_init(1) returns to single-device operation.  Optional: without it the group is the current device. */
int iglb200_init(int n_gpus);
/* info = {devices in the group, 1 if NCCL replicates, 1 if peer access is on, microseconds of the last replication};
 * transport (may be NULL) names what the last replication used */
int iglb200_group_info(int64_t info[4], char *transport, uint64_t cap);
void iglb200_shutdown(void);             /* optional: waits for the calling thread's device to go idle */

/* pinned host memory for fast transfers (optional; any host pointer works) */
int iglb200_host_alloc(void **out, uint64_t bytes);
int iglb200_host_free(void *p);
/* page-lock / release memory the caller owns (a shared-memory mapping, an Eigen matrix that is reused many times) */
int iglb200_host_register(void *p, uint64_t bytes);
int iglb200_host_unregister(void *p);

/* ---- mesh: replaces AABB::init (AABB.h:494-496, AABB.cpp:877-883) and
 *      fast_winding_number(V,F,order,bvh) (fast_winding_number.h:160-170,
 *      fast_winding_number.cpp:383-416).  Both trees are built on the device. */
int iglb200_mesh_create(const void *V, int v_is_f64, int64_t nV, int64_t v_row_stride, int64_t v_col_stride,
                        const int32_t *F, int64_t nF, int64_t f_row_stride, int64_t f_col_stride,
                        int fwn_order, iglb200_mesh **out);
/* flags: IGLB200_MESH_PSEUDONORMALS also precomputes face / angle-weighted vertex / uniform edge
 * normals (signed_distance.cpp:68-76: per_face_normals, per_vertex_normals, per_edge_normals). */
#define IGLB200_MESH_PSEUDONORMALS 1
/* IGLB200_MESH_WINDING_NUMBER also builds the boundary chains of the exact hierarchical winding number
 * (WindingNumberAABB::set_mesh + grow, signed_distance.cpp:60-64), needed by sign types WINDING_NUMBER / DEFAULT
 * and by iglb200_winding_number. */
#define IGLB200_MESH_WINDING_NUMBER 2
int iglb200_mesh_create_ex(const void *V, int v_is_f64, int64_t nV, int64_t v_row_stride, int64_t v_col_stride,
                           const int32_t *F, int64_t nF, int64_t f_row_stride, int64_t f_col_stride,
                           int fwn_order, int flags, iglb200_mesh **out);
int iglb200_mesh_destroy(iglb200_mesh *m);
/* (re)replicate a mesh onto the devices of the current group (meshes created after iglb200_init are already) */
int iglb200_mesh_replicate(iglb200_mesh *m);
/* info[0]=#faces, [1]=#tree nodes, [2]=#levels, [3]=device bytes, [4]=#faces excluded (NaN/Inf boxes),
 * [5]=fwn order, [6]=has pseudonormals, [7]=#devices holding a copy */
int iglb200_mesh_info(const iglb200_mesh *m, int64_t info[8]);
/* device build time of the last create on this handle, milliseconds (CUDA events) */
double iglb200_mesh_build_ms(const iglb200_mesh *m);

/* Flat, position-independent image of the built trees, for replication to
 * other GPUs (ncclBroadcast of one byte range) or caching.  The arena is one
 * device allocation owned by the handle. */
int iglb200_mesh_arena(const iglb200_mesh *m, void **device_ptr, uint64_t *bytes);
/* Copy the arena to host memory (serialisation; AABB::serialize is the
 * reference's analogue, AABB.cpp:1089-1116). `bytes` must equal the arena size. */
int iglb200_mesh_arena_download(const iglb200_mesh *m, void *host_dst, uint64_t bytes);
/* Adopt an arena that already sits in device memory of the current device.
 * copy!=0: the handle makes its own copy; copy==0: the caller keeps `device_ptr`
 * alive for the lifetime of the handle. */
int iglb200_mesh_from_arena(const void *device_ptr, uint64_t bytes, int copy, iglb200_mesh **out);

/* ---- igl::signed_distance(P,V,F,sign_type,lower,upper,S,I,C,N)
 *      signed_distance.h:61-89, signed_distance.cpp:28-139, 372-416.
 *      N (optional, may be NULL) is written only for PSEUDONORMAL, as in the reference
 *      (the normal the sign was taken from, signed_distance.cpp:125-131).
 *      W (optional, may be NULL) receives the winding number the sign was made
 *      from — an extension (NaN for the other sign types).  Out-of-bounds rows: S=NaN, I=#F+1,
 *      C=0, W=NaN, N=0 (signed_distance.cpp:99-105).  #F==0: S=NaN, I=-1, C=NaN, N=NaN (:398-405). */
int iglb200_signed_distance(const iglb200_mesh *m,
                            const void *P, int p_is_f64, int64_t nP, int64_t p_row_stride, int64_t p_col_stride,
                            int sign_type, double lower_bound, double upper_bound,
                            void *S, int s_is_f64, int32_t *I,
                            void *C, int c_is_f64, int64_t c_row_stride, int64_t c_col_stride,
                            void *W, int w_is_f64,
                            void *N, int n_is_f64, int64_t n_row_stride, int64_t n_col_stride);

/* The same call made COOPERATIVELY by several processes (one per GPU, e.g. under torchrun) on ONE query set: all
 * of them map the same host arrays (shared memory) and pass the same arguments; each takes chunks of `chunk_rows`
 * rows off `shared_chunk_counter` (an int64 in that shared memory, zeroed before the first call starts) until the
 * rows run out, and writes only the rows of its chunks.  Matrices must be column-major or dense row-major. */
int iglb200_signed_distance_coop(const iglb200_mesh *m,
                                 const void *P, int p_is_f64, int64_t nP, int64_t p_row_stride, int64_t p_col_stride,
                                 int sign_type, double lower_bound, double upper_bound,
                                 void *S, int s_is_f64, int32_t *I,
                                 void *C, int c_is_f64, int64_t c_row_stride, int64_t c_col_stride,
                                 void *W, int w_is_f64,
                                 void *N, int n_is_f64, int64_t n_row_stride, int64_t n_col_stride,
                                 int64_t *shared_chunk_counter, int64_t chunk_rows);
/* the chunk size the library itself would use for nP rows on n_devices devices */
int64_t iglb200_host_chunk_rows(int64_t nP, int n_devices);

/* ---- AABB::squared_distance(V,Ele,P,sqrD,I,C) batch, AABB.h:722-728,
 *      AABB.cpp:1241-1265; also point_mesh_squared_distance.h:54-60. */
int iglb200_squared_distance(const iglb200_mesh *m,
                             const void *P, int p_is_f64, int64_t nP, int64_t p_row_stride, int64_t p_col_stride,
                             void *sqrD, int d_is_f64, int32_t *I,
                             void *C, int c_is_f64, int64_t c_row_stride, int64_t c_col_stride);

/* AABB::squared_distance(V,Ele,p,low_sqr_d,up_sqr_d,i,c) for a batch (AABB.h:575-614, AABB.cpp:1135-1218): only primitives
 * closer than up_sqr_d count — a row without any gets sqrD = up_sqr_d, I = -1, C = 0 (the reference leaves i and c
 * untouched) — and low_sqr_d > up_sqr_d returns low_sqr_d.  low_sqr_d is otherwise the reference's licence to stop early
 * at ANY primitive below it; this search always returns the closest one. */
int iglb200_squared_distance_bounded(const iglb200_mesh *m,
                                     const void *P, int p_is_f64, int64_t nP, int64_t p_row_stride, int64_t p_col_stride,
                                     double low_sqr_d, double up_sqr_d,
                                     void *sqrD, int d_is_f64, int32_t *I,
                                     void *C, int c_is_f64, int64_t c_row_stride, int64_t c_col_stride);

/* ---- fast_winding_number(fwn_bvh, accuracy_scale, Q, W)
 *      fast_winding_number.h:172-187, fast_winding_number.cpp:421-437. */
int iglb200_fwn_eval(const iglb200_mesh *m,
                     const void *Q, int q_is_f64, int64_t nQ, int64_t q_row_stride, int64_t q_col_stride,
                     float accuracy_scale, void *W, int w_is_f64);

/* ---- device-resident variants: every pointer is a DEVICE pointer on the
 *      current device, canonical layout (P/C column-major doubles, rs=1,
 *      cs=nP; S,W doubles; I int32), `stream` is a cudaStream_t (NULL = default
 *      stream).  Asynchronous: the call returns after enqueueing.  `scratch`
 *      must hold iglb200_query_scratch_bytes(nP) bytes.  sort_queries!=0 groups
 *      queries in Morton order before traversal. */
uint64_t iglb200_query_scratch_bytes(int64_t nP);
int iglb200_signed_distance_dev(const iglb200_mesh *m, const double *P, int64_t nP,
                                int sign_type, double lower_bound, double upper_bound,
                                double *S, int32_t *I, double *C, double *W, double *N,
                                void *scratch, int sort_queries, void *stream);
int iglb200_squared_distance_dev(const iglb200_mesh *m, const double *P, int64_t nP,
                                 double *sqrD, int32_t *I, double *C,
                                 void *scratch, int sort_queries, void *stream);
int iglb200_fwn_eval_dev(const iglb200_mesh *m, const double *Q, int64_t nQ, float accuracy_scale,
                         double *W, void *scratch, int sort_queries, void *stream);

/* ---- SEVERAL GPUs on ONE device-resident query set (the loop of signed_distance.cpp:90-138 sharded).
 * The query set P and the outputs S, I, C, W live on one device, the OWNER; every GPU holds a copy of the mesh
 * (iglb200_init in one process; iglb200_mesh_from_arena after an NCCL broadcast between processes).
 *   begin  (owner)   Morton-sorts the queries once: keys and permutation go to `shard_scratch`
 *                    (iglb200_shard_scratch_bytes(nP) bytes of the owner's memory), the work counter in it is zeroed and
 *                    the hand-out order of the super-chunks is fixed (Morton order).  `super_rows` >= 8192, the same in
 *                    begin and run.
 *   run    (each GPU, after begin) runs the same kernels as the single-GPU call over the super-chunks of `super_rows`
 *                    consecutive sorted queries it takes off that counter — an atomic over NVLink when the owner is a
 *                    peer — until they run out.  P, the permutation and S/I/C/W are read from / written to the OWNER's
 *                    memory directly (peer loads and stores): no separate scatter, gather or un-sort pass.  All
 *                    pointers are the owner's buffers as seen from the calling device (peer-enabled pointers in one
 *                    process, iglb200_ipc_open between processes).  max_supers > 0 caps what this GPU takes.
 *                    The super-chunks form n_shards (<= 64) contiguous slices of the sorted order with a counter each:
 *                    the caller with shard_index k works through slice k first — one region of space, hence one part of
 *                    the tree in its L2 — and then helps with whichever slice has the most left, so the slices need
 *                    not cost the same.  Every participant passes the same n_shards and its own shard_index; a single
 *                    caller passes (0, 1).
 *                    The winding kernel then runs over the rows this GPU took, from a stash the search left in its
 *                    own memory (no scattered peer reads).
 *   finish (owner, after every run) writes the sentinel rows of a degenerate mesh (#F == 0); otherwise nothing.
 * Ordering between the phases is the caller's: events between streams in one process, a tiny collective on the
 * streams between processes.  mode 0: signed distance (sign_type FAST_WINDING_NUMBER or UNSIGNED; S,I,C[,W]),
 * 1: squared distance (S=sqrD,I,C), 2: fast winding number (W; rows are claimed 2048 at a time, super_rows unused).
 * All asynchronous on `stream` of the calling thread's current device. */
uint64_t iglb200_shard_scratch_bytes(int64_t nP);
int64_t iglb200_shard_super_rows(int64_t nP, int n_shards);   /* the library's choice; a multiple of 1024 */
int iglb200_shard_begin_dev(const iglb200_mesh *m, const double *P, int64_t nP, void *shard_scratch, int sort_queries,
                            int64_t super_rows, void *stream);
int iglb200_shard_run_dev(const iglb200_mesh *m, const double *P, int64_t nP, void *shard_scratch, int mode, int sign_type,
                          double lower_bound, double upper_bound, float accuracy_scale, int sort_queries,
                          double *S, int32_t *I, double *C, double *W,
                          int64_t super_rows, int64_t max_supers, int shard_index, int n_shards, void *stream);
int iglb200_shard_finish_dev(const iglb200_mesh *m, int64_t nP, void *shard_scratch, int mode, int sign_type,
                             double lower_bound, double upper_bound, int sort_queries,
                             double *S, int32_t *I, double *C, double *W, void *stream);
/* device memory that other processes on the node can map (cudaIpcGetMemHandle / cudaIpcOpenMemHandle): the
 * owner allocates and publishes the 64-byte handle (e.g. torch.distributed broadcast), the others open it. */
int iglb200_ipc_alloc(uint64_t bytes, void **device_ptr, unsigned char handle[64]);
int iglb200_ipc_open(const unsigned char handle[64], void **device_ptr);
int iglb200_ipc_close(void *device_ptr);
int iglb200_ipc_free(void *device_ptr);

/* igl::winding_number(V,F,O,W) for triangle meshes (winding_number.h:35-43, winding_number.cpp:17-52): the exact
 * generalised winding number, double precision, evaluated hierarchically (WindingNumberTree.h:262-335).
 * The mesh must have been created with IGLB200_MESH_WINDING_NUMBER. */
int iglb200_winding_number(const iglb200_mesh *m,
                           const void *O, int o_is_f64, int64_t nO, int64_t o_row_stride, int64_t o_col_stride,
                           void *W, int w_is_f64);
int iglb200_winding_number_dev(const iglb200_mesh *m, const double *O, int64_t nO, double *W, void *scratch,
                               int sort_queries, void *stream);
/* igl::iterative_closest_point(VX,FX,VY,FY,num_samples,max_iters,R,t) (iterative_closest_point.h, iterative_closest_point.cpp:40-130):
 * `meshY` is the handle of (VY,FY); per iteration num_samples area-weighted random points of X*R+t, their closest points
 * on Y (the hot path's search), Y's face normals there and the point-to-plane fit of rigid_alignment.cpp:24-108 run on the
 * device; only the 6x6 normal equations are solved on the host.  R is 3x3 ROW-major with VX*R + t ~ Y (row vectors), t 1x3.
 * The reference seeds its sampler from std::rand(); here the samples are a counter-based hash of `seed`. */
int iglb200_iterative_closest_point(const iglb200_mesh *meshY,
                                    const void *VX, int vx_is_f64, int64_t nVX, int64_t vx_row_stride, int64_t vx_col_stride,
                                    const int32_t *FX, int64_t nFX, int64_t fx_row_stride, int64_t fx_col_stride,
                                    int num_samples, int max_iters, uint64_t seed, double R[9], double t[3]);

/* After synchronising a stream that ran *_dev calls: 0, or an error if a device-side
 * condition (traversal stack overflow) invalidated their results.  The host-pointer
 * entry points check this themselves. */
int iglb200_poll_device_error(void);
/* number of kernels the last *_dev / host call on this thread enqueued */
int iglb200_last_launch_count(void);

/* Mean work per query of both traversals on this mesh, for the algorithmic-flop
 * figure: out = {closest: nodes, faces PER WARP of 32 Morton-adjacent queries (they
 * share one walk); winding: nodes, faces per query}.
 * P is a HOST pointer, column-major doubles; keep n below ~1e6. */
int iglb200_debug_visit_counts(const iglb200_mesh *m, const double *P_host_colmajor, int64_t n, double out[4]);

/* Measured FP32 FMA peak of the current device in TFLOP/s (a dependent-chain FMA micro-benchmark);
 * the denominator of bench.py's fp32 roofline fraction.  <0 on error. */
double iglb200_debug_fp32_peak_tflops(void);

/* Optional per-kernel timing (CUDA events on the launching stream) for the
 * calling thread.  collect() synchronises, writes "name=ms;name=ms;..." summed
 * per kernel group since the previous collect, and returns the string length. */
void iglb200_profile_enable(int on);
int iglb200_profile_collect(char *out, uint64_t cap);

/* ---- point cloud: fast_winding_number(P,N,A,Q,order,beta,WN)
 *      fast_winding_number.h:44-53,88-131; octree.h; fast_winding_number.cpp:17-344 */
int iglb200_cloud_create(const void *P, const void *N, const void *A, int is_f64, int64_t nP,
                         int expansion_order, iglb200_cloud **out);
int iglb200_cloud_destroy(iglb200_cloud *c);
int iglb200_cloud_info(const iglb200_cloud *c, int64_t info[8]);
int iglb200_cloud_eval(const iglb200_cloud *c,
                       const void *Q, int q_is_f64, int64_t nQ, int64_t q_row_stride, int64_t q_col_stride,
                       double beta, void *WN, int w_is_f64);
int iglb200_cloud_eval_dev(const iglb200_cloud *c, const double *Q, int64_t nQ, double beta, double *WN,
                           void *stream);

/* ---- voxel grids around the path, device resident
 *      voxel_grid.h:30-37 (voxel_grid.cpp:14-88), flood_fill.h (flood_fill.cpp:11-98),
 *      marching_cubes.h:39-66 (marching_cubes.cpp:19-112, march_cube.cpp:24-142),
 *      offset_surface.h (offset_surface.cpp:18-48: the host wrappers compose it from these calls).
 * A grid owns its vertex positions GV (n x 3) and one scalar S per vertex, both in device memory;
 * n = side[0]*side[1]*side[2] <= 2^31-1, x fastest (grid.cpp:22-44). */
typedef struct iglb200_grid iglb200_grid;
typedef struct iglb200_surface iglb200_surface;

/* igl::voxel_grid(box, s, pad_count, GV, side).  The (V, offset, ...) overload is the box
 * [min(V) - offset, max(V) + offset] (voxel_grid.cpp:80-86). */
int iglb200_grid_create(const double box_min[3], const double box_max[3], int s, int pad_count, iglb200_grid **out);
/* wraps caller-supplied vertex positions (host, strided n x 3 doubles) of an nx*ny*nz grid */
int iglb200_grid_create_from(const double *GV, int64_t gv_row_stride, int64_t gv_col_stride, int nx, int ny, int nz,
                             iglb200_grid **out);
int iglb200_grid_destroy(iglb200_grid *g);
/* side, vertex count, and the first and last vertex (corners[0..2], corners[3..5]); any may be NULL */
int iglb200_grid_info(const iglb200_grid *g, int side[3], int64_t *n, double corners[6]);
/* device pointers: GV column-major (x[n], y[n], z[n]) and S[n]; valid until destroy */
int iglb200_grid_device(iglb200_grid *g, const double **GV_colmajor, double **S);
int iglb200_grid_get_gv(const iglb200_grid *g, double *GV, int64_t gv_row_stride, int64_t gv_col_stride);
int iglb200_grid_set_s(iglb200_grid *g, const double *S);
int iglb200_grid_get_s(const iglb200_grid *g, double *S);
/* S = igl::signed_distance(GV, mesh, sign_type, lower, upper) (signed_distance.cpp:28-139); I, C, N are not kept */
int iglb200_grid_signed_distance(iglb200_grid *g, const iglb200_mesh *m, int sign_type, double lower, double upper);
/* igl::flood_fill(side, S) */
int iglb200_grid_flood_fill(iglb200_grid *g);
/* igl::marching_cubes(S - shift, GV, side..., isovalue, V, F, E2V); offset_surface.cpp:46-47 is
 * shift = isolevel, isovalue = 0 */
int iglb200_grid_marching_cubes(const iglb200_grid *g, double shift, double isovalue, iglb200_surface **out);
int iglb200_surface_info(const iglb200_surface *s, int64_t *nV, int64_t *nF);
/* V: nV x 3 row-major doubles; F: nF x 3 row-major; E: nV x 2 = the grid edge (i < j) under each vertex,
 * i.e. the E2V map of marching_cubes.h:58-66 listed by vertex.  Any may be NULL. */
int iglb200_surface_get(const iglb200_surface *s, double *V, int32_t *F, int32_t *E);
int iglb200_surface_device(const iglb200_surface *s, const double **V, const int32_t **F, const int32_t **E);
int iglb200_surface_destroy(iglb200_surface *s);
/* host-pointer one-shots: igl::flood_fill(res, S) in place; igl::marching_cubes(S, GV, nx, ny, nz, isovalue, ...) */
int iglb200_flood_fill(const int res[3], double *S);
int iglb200_marching_cubes(const double *S, const double *GV, int64_t gv_row_stride, int64_t gv_col_stride,
                           int nx, int ny, int nz, double isovalue, iglb200_surface **out);
/* igl::marching_cubes(S, GV, GI, isovalue, V, F, E2V) (marching_cubes.h:68-92, marching_cubes.cpp:114-176): cube c has
 * the corners GI(c, 0..7), vertex indices into the n rows of GV / entries of S (e.g. a sparse voxel grid). */
int iglb200_marching_cubes_sparse(const double *S, const double *GV, int64_t n, int64_t gv_row_stride, int64_t gv_col_stride,
                                  const int32_t *GI, int64_t n_cubes, int64_t gi_row_stride, int64_t gi_col_stride,
                                  double isovalue, iglb200_surface **out);

#ifdef __cplusplus
}
#endif
#endif
