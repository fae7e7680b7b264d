/*
 * shm3d_grid.h -- C ABI of the B200-native signed-heat grid solver (libshm3d_grid.so).
 *
 * This is the drop-in boundary for ONE hot path of nzfeng/signed-heat-3d:
 * SignedHeatGridSolver::computeDistance (reference: include/signed_heat_grid_solver.h:16-20,
 * src/signed_heat_grid_solver.cpp:5-222).  Everything behind these entry points runs as
 * hand-written sm_100a CUDA; there is no CPU fallback -- calls fail with SHM3D_ERR_CUDA when
 * no device is usable.
 *
 * What the reference's FFI for this path would bind (file:line = interface replaced):
 *   shm3d_ctx_create / destroy      <- SignedHeatGridSolver() ctor + unique_ptr lifetime
 *                                      (src/signed_heat_grid_solver.cpp:3, src/main.cpp:52,290)
 *   shm3d_solve                     <- computeDistance(...), Steps 1-3 + shift
 *                                      (src/signed_heat_grid_solver.cpp:38-113 and :146-221)
 *   shm3d_step12 / shm3d_rhs        <- the Step 1-2 loop (:48-65, :157-174) and D^T Y (:70-74,:179-180);
 *                                      exposed so parity tests can check intermediate fields
 *   shm3d_last_error                <- the C++ exceptions geometry-central raises
 *                                      (deps/geometry-central/src/numerical/square_solvers.cpp:123-125,164-169)
 * Host-side work the caller (adapter) keeps: grid setup a4, lambda a5, face areas/normals/
 * barycentres a6 -- see include/shm3d_host.hpp for the dependency-free
 * C++ mirror and INTEGRATION.md for the geometry-central adapter.
 *
 * Layouts: all host arrays are caller-owned, plain row-major.  pos/nrm are [M][3] doubles,
 * area is [M].  Source ORDER matters: it defines which source pins each grid cell
 * (src/signed_heat_grid_solver.cpp:86-98 "first face per cell wins").
 * phi_out is double[nx*ny*(k1-k0)], node index i + j*nx + (k-k0)*nx*ny  (x fastest,
 * src/signed_heat_grid_solver.cpp:505-508); k0..k1 is the calling rank's z-slab
 * (the whole grid when the context is single-GPU).
 */
#ifndef SHM3D_GRID_H
#define SHM3D_GRID_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHM3D_OK 0
#define SHM3D_ERR_INVALID_ARG 1
#define SHM3D_ERR_CUDA 2            /* no device / CUDA runtime failure */
#define SHM3D_ERR_NONFINITE 3       /* non-finite source data or rhs (reference: checkFinite throws) */
#define SHM3D_ERR_FACTORIZATION 4   /* constraint system A A^T not positive definite */
#define SHM3D_ERR_NO_CONVERGENCE 5  /* constrained PCG hit cg_max_iters */
#define SHM3D_ERR_NCCL 6

#define SHM3D_FLAG_FAST 1u             /* SignedHeat3DOptions.fastIntegration (include/signed_heat_3d.h:27): greedy BFS integration
                                          (src/signed_heat_grid_solver.cpp:224-275) instead of the constrained solve */
#define SHM3D_FLAG_SCRUB_NONFINITE 2u  /* mesh overload zeroes non-finite rhs entries (:72-74); point overload does not */
#define SHM3D_FLAG_VERBOSE 4u          /* SignedHeatGridSolver::VERBOSE */
#define SHM3D_FLAG_NO_MG 8u            /* diagnostics: plain projected CG (no multigrid preconditioner) */
#define SHM3D_FLAG_PROFILE 32u         /* time selected kernels with CUDA events on the solver's stream (fills the ms_pcg_* stats) */
#define SHM3D_FLAG_FP64_UNDERFLOW 64u  /* follow a floating-point artefact of the reference: where the summed field X is so small
                                        * (lambda * distance-to-the-surface >~ 355: far corners of the box for finely
                                        * triangulated inputs such as data/SprayBottle.obj) that its X /= X.norm()
                                        * (src/signed_heat_grid_solver.cpp:61, plain double) loses precision (max|X| < 2^-511)
                                        * or squares to zero (max|X| < 2^-537.5), Y is of the wrong length or non-finite; the
                                        * mesh overload then zeroes the right-hand-side entries that touch a non-finite node
                                        * (:72-74), the point overload throws.  Steps 1-2 here are range-shifted and finite
                                        * everywhere; with this flag Step 2 evaluates the reference's expression in IEEE double
                                        * on the true values at those nodes, so phi follows the reference (2e-2 relative L2
                                        * on SprayBottle otherwise).  Set by shm3d_prepare_mesh / shm3d_prepare_points, the
                                        * adapter and the class mirrors (a drop-in returns what the reference returns); a
                                        * caller filling shm3d_params by hand opts in. */
#define SHM3D_FLAG_NO_TMA 128u          /* diagnostics: row-streaming stencil kernels instead of the TMA-staged marching ones */
#define SHM3D_FLAG_TAIL_PROGRAM 256u    /* experiment (off by default: measured 1 % slower than graph-replayed launches): the coarsest
                                          multigrid levels (<= 16^3) as ONE launch of a recorded op program (csrc/mg_tail.cuh) */
#define SHM3D_FLAG_NO_CYCLIC_SUM 2048u  /* diagnostics (slab contexts): every rank sums its own slab instead of the round-robin
                                          z-chunks that balance Steps 1-2 across the ranks */
#define SHM3D_FLAG_NO_PDL 1024u         /* diagnostics: the projector's sweep kernels launched fully serialised instead of with
                                          programmatic dependent launch */
#define SHM3D_FLAG_NO_GRAPH 512u        /* diagnostics: launch every PCG iteration kernel by kernel instead of replaying the
                                          captured CUDA graph */
#define SHM3D_FLAG_PLAIN_MG 16u        /* unconstrained Poisson V-cycle as preconditioner, projector on the fine level only */

typedef struct shm3d_ctx shm3d_ctx;

typedef struct shm3d_params {
    int32_t nx, ny, nz;   /* grid nodes per axis (reference: nx=ny=nz=16*2^hCoef, :24) */
    double bbox_min[3];   /* position of node (0,0,0) (:17-18) */
    double cell;          /* node spacing 2s/(nx-1) (:26) */
    double lambda;        /* 1/sqrt(tCoef h^2) (:42-44) */
    uint32_t flags;
    double cull_tau;      /* far-field cut: drop sources with lambda*(r - r_min) > tau.  <=0: default 10.
                             +inf: brute force (every source at every node, like the reference) */
    double cg_rel_tol;    /* <=0: default 1e-5 (relative preconditioned residual; csrc/solver.cu run_pcg says why) */
    int32_t cg_max_iters; /* <=0: default 2000 */
    int32_t mg_smooth;    /* Jacobi sweeps per multigrid leg; <=0: default 2 */
    int32_t mg_constrained_from; /* first multigrid level (0 = finest) whose smoothers are projected onto that level's
                             constraints; finer levels run the plain Poisson smoother.  0: default (2); <0: every level */
    int32_t reserved;
} shm3d_params;

typedef struct shm3d_stats {
    double ms_total;        /* wall time of the call */
    double ms_h2d;          /* source upload + clustering */
    double ms_sum;          /* Steps 1-2 kernel(s) (device time) */
    double ms_rhs;          /* D^T Y */
    double ms_constraints;  /* constraint rows + A A^T nested-dissection factorisation (host) + upload */
    double ms_pcg;          /* constrained multigrid-PCG (device time) */
    double ms_shift;        /* source average + shift */
    double ms_d2h;          /* phi download */
    int64_t pairs_evaluated;  /* (node,source) pairs the summation kernel actually evaluated */
    int64_t pairs_bruteforce; /* N*M */
    int32_t n_clusters;
    int32_t m_constraints;
    int32_t cg_iters;
    double cg_rel_residual;
    double shift;
    int64_t kernel_launches;  /* CUDA kernels launched by this call */
    double ms_pcg_stencil;    /* device time inside the fused stencil-apply+dot kernel, summed */
    int64_t pcg_stencil_launches;
    double ms_pcg_vcycle;     /* device time inside multigrid V-cycles, summed */
    double ms_pcg_projector;  /* device time inside fine-level projector applications, summed */
    int64_t pcg_projector_applies;
    double ms_pcg_update;     /* device time inside the fused x/r update kernel, summed */
    int64_t pcg_vcycles;      /* V-cycles the ms_pcg_vcycle sum covers (SHM3D_FLAG_PROFILE times the first iterations only;
                                 the rest of the solve replays a CUDA graph) */
    int32_t tail_ops;         /* ops of the V-cycle tail's cluster program (0: every level runs as separate launches) */
    int32_t graph_replays;    /* PCG iterations executed as CUDA-graph replays */
} shm3d_stats;

/* Context: one per GPU.  device = CUDA ordinal.  Returns SHM3D_ERR_CUDA when no usable device. */
int shm3d_ctx_create(shm3d_ctx** out, int device);
/* One rank of a z-slab-partitioned solve over `world` GPUs (NCCL over NVLink).  nccl_id is the 128-byte
 * ncclUniqueId from shm3d_nccl_unique_id() on rank 0, distributed by the caller (e.g. torch.distributed). */
int shm3d_ctx_create_dist(shm3d_ctx** out, int device, int rank, int world, const void* nccl_id);
int shm3d_nccl_unique_id(void* out128);
void shm3d_ctx_destroy(shm3d_ctx* ctx);
/* Message of the last failing call on this context (or of a failed create when ctx == NULL). */
const char* shm3d_last_error(const shm3d_ctx* ctx);
/* z-slab [k0,k1) of rank `rank` of `world` for a grid of nz planes (pure function; no context needed). */
int shm3d_slab_range(int32_t rank, int32_t world, int32_t nz, int32_t* k0, int32_t* k1);
/* z-slab [k0,k1) this context owns for a grid of nz planes. */
int shm3d_slab(const shm3d_ctx* ctx, int32_t nz, int32_t* k0, int32_t* k1);

/* The CUDA stream (cudaStream_t, returned as void*) every kernel of this context is launched on -- for callers that
 * time the path with their own CUDA events or order their own device work against it. */
void* shm3d_ctx_stream(const shm3d_ctx* ctx);
/* Page-locked host buffers (cudaHostAlloc) for the phi_out / source arrays of shm3d_solve: the D2H copy of a 512^3
 * double field is ~10x faster into pinned memory.  NULL on failure. */
void* shm3d_host_alloc(size_t bytes);
void shm3d_host_free(void* p);

/* computeDistance: Steps 1-3 + shift.  phi_out: double[nx*ny*(k1-k0)]. stats may be NULL. */
int shm3d_solve(shm3d_ctx* ctx, const shm3d_params* p, int64_t n_sources, const double* pos, const double* nrm,
                const double* area, double* phi_out, shm3d_stats* stats);

/* Same solve with the source arrays already resident on the device (device pointers, same layout) and phi
 * left on the device as float[local N] (phi_dev).  Used to time the path without host<->device copies. */
int shm3d_solve_device(shm3d_ctx* ctx, const shm3d_params* p, int64_t n_sources, const double* d_pos,
                       const double* d_nrm, const double* d_area, float* phi_dev, shm3d_stats* stats);

/* Steps 1-2 only: Y_out float[3][local N] (component-major: Yx, then Yy, then Yz). */
int shm3d_step12(shm3d_ctx* ctx, const shm3d_params* p, int64_t n_sources, const double* pos, const double* nrm,
                 const double* area, float* Y_out, shm3d_stats* stats);

/* Steps 1-2 at arbitrary query points instead of grid nodes -- what the reference's tet solver evaluates at tet
 * barycentres (src/signed_heat_tet_solver.cpp:54-72, :131-147).  query: double[n_query][3];
 * Y_out: float[n_query][3] (interleaved unit vectors).  Every source is evaluated at every point (no culling). */
int shm3d_step12_points(shm3d_ctx* ctx, double lambda, int64_t n_sources, const double* pos, const double* nrm,
                        const double* area, int64_t n_query, const double* query, float* Y_out);

/* b = cell^2 * D^T Y for a given Y (component-major float[3][local N]) -> b_out float[local N]. */
int shm3d_rhs(shm3d_ctx* ctx, const shm3d_params* p, const float* Y, float* b_out);

/* Step 3 only, from a given right-hand side b (= cell^2 D^T Y, float[local N]): constrained solve + shift. */
int shm3d_step3(shm3d_ctx* ctx, const shm3d_params* p, int64_t n_sources, const double* pos, const double* area,
                const float* b, double* phi_out, shm3d_stats* stats);

/* Host half of the reference interface (rows a4-a6 of SURVEY.md section 8a), dependency-free:
 * fills `out` (grid, lambda, flags) and, when the three arrays are non-NULL, the per-face barycentre / unit
 * normal / area arrays ([nF][3], [nF][3], [nF]) that shm3d_solve consumes.  Faces are polygons:
 * face f uses vertices face_vertices[face_offsets[f] .. face_offsets[f+1]).
 * Replaces centroid/radius (src/signed_heat_3d.cpp:3-22), the grid set-up (src/signed_heat_grid_solver.cpp:13-26),
 * meanEdgeLength (src/signed_heat_3d.cpp:51-60), setFaceVectorAreas (:62-89), barycenter (grid_solver.cpp:498-503). */
int shm3d_prepare_mesh(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets,
                       int64_t nF, double tCoef, double hCoef, double scale, shm3d_params* out, double* pos_out,
                       double* nrm_out, double* area_out, double* h_out);
/* Point-cloud overload set-up (src/signed_heat_grid_solver.cpp:124-137, :151-153): h is the mean edge length of
 * the caller's tufted triangulation; the per-point areas are passed to shm3d_solve as `area`. */
int shm3d_prepare_points(const double* P, int64_t nP, double h, double tCoef, double hCoef, double scale,
                         shm3d_params* out);

/* Source weights for the point-cloud overload: what the reference reads from
 * geometry-central's tufted triangulation -- per-point vertex dual areas and the mean edge length h
 * (src/signed_heat_grid_solver.cpp:149-151,165) -- computed from positions + normals by restating that pipeline:
 * kNN(k_neighbors <= 0: 30), tangent-plane local Delaunay 1-rings (deps/geometry-central/src/pointcloud/
 * local_triangulation.cpp:10-210), the triangle soup of all local triangles, intrinsic mollification, the tufted cover
 * (src/surface/tufted_laplacian.cpp:39-121) and intrinsic Delaunay flips (src/surface/simple_idt.cpp).  Equal to
 * geometry-central's own code (its sources compiled for the tests, oracle/_ref/libshm_gc_ref.so) to <= 1e-12 relative on
 * areas and h, degenerate inputs included: the kNN restates nanoflann's kd-tree (tie-breaking by visiting order), the
 * in-circle test Eigen 3.3's determinant expression (tests/test_point_weights.py, tests/test_cli.py).
 * Optional diagnostics: number of flips, smallest edge cotan weight after the flips (>= -1e-6 = intrinsically
 * Delaunay), total cover area before the flips (= sum of areas_out).  Host only; no device work. */
int shm3d_point_weights(const double* P, const double* N, int64_t nP, int32_t k_neighbors, double* areas_out,
                        double* h_out, int64_t* n_triangles_out, int64_t* n_flips_out, double* min_cotan_out,
                        double* area_before_out);
/* probe for the tests: tufted cover + intrinsic Delaunay flips of a given triangle soup (tris: int64[T][3]) */
int shm3d_debug_tufted_weights(const double* P, int64_t nP, const int64_t* tris, int64_t T, double* areas_out, double* h_out,
                               int64_t* n_flips_out, double* min_cotan_out, double* area_before_out);
/* probe for the tests: the local Delaunay 1-ring of the origin among n tangent-plane points (returns the ring size) */
int shm3d_debug_local_ring(const double* coords2d, int32_t n, int32_t* ring_out, int32_t* tri_after_out);
/* test probe (host logic, no GPU): the exchange plan of the load-balanced Steps 1-2 on slab contexts -- the 8-plane z-chunks
 * rank `rank` computes for slab `peer` (to_out) and the chunks of its own slab that `peer` computes (from_out), in message
 * order.  Returns -1 when nz is not a multiple of 8 * world (such grids keep slab-local summation). */
int shm3d_debug_cyclic_plan(int32_t nz, int32_t world, int32_t rank, int32_t peer, int32_t* to_out, int32_t* n_to,
                            int32_t* from_out, int32_t* n_from);
/* test probe: which k-nearest-neighbour search shm3d_point_weights uses on the calling process.  0 (default): the
 * restatement of nanoflann's kd-tree (ties among exactly equidistant points in its visiting order, like geometry-central);
 * 1: an independent cell-list search with ties broken by point index -- identical wherever no distances tie, used by the
 * tests as a cross-check of the former.  Not thread-safe; not part of the product path. */
void shm3d_debug_knn_mode(int32_t mode);

/* ---- Row N3: the consumer of phi, on the device -------------------------------------------------
 * The reference hands N doubles to polyscope, which narrows them to float32
 * (deps/polyscope/include/polyscope/volume_grid.ipp:103-106) and, when the user contours (src/main.cpp:116-128), runs
 * registerIsosurfaceAsMesh (deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228) = MC::marching_cube of
 * deps/polyscope/deps/MarchingCubeCpp/include/MarchingCube/MC.h:242-315 + a swizzle/scale/translate of the vertices.
 * shm3d_isosurface produces the same indexed mesh on the GPU -- same float32 coordinates, same vertex numbering, same
 * triangle order -- from a field that can stay where shm3d_solve_device left it.  Single-GPU contexts only. */
#define SHM3D_FIELD_DEVICE_F32 0 /* float[nx*ny*nz] in device memory (phi_dev of shm3d_solve_device) */
#define SHM3D_FIELD_HOST_F64 1   /* double[nx*ny*nz] on the host (phi_out of shm3d_solve); narrowed to float32 on the device */
#define SHM3D_FIELD_HOST_F32 2   /* float[nx*ny*nz] on the host */
#define SHM3D_ISO_LATTICE 1u     /* flag: vertices in the marching-cubes library's own lattice coordinates (k, j, i) */
typedef struct shm3d_iso_stats {
    int64_t n_vertices, n_triangles;
    double ms_device;     /* CUDA-event time of the extraction kernels */
    int64_t gpu_launches; /* kernels launched by the call (incl. the narrowing of a host field) */
} shm3d_iso_stats;
/* Uses p->nx,ny,nz (and bbox_min, cell when the bounds are NULL).  bound_min / bound_max: the float[3] bounds the volume
 * grid was registered with (src/signed_heat_grid_solver.cpp:20-24,35); NULL = (float)bbox_min and
 * (float)(bbox_min + cell*(n-1)).  The mesh stays in buffers owned by the context until the next call.
 * z-slab contexts (shm3d_ctx_create_dist): a collective call -- every rank passes ITS slab as SHM3D_FIELD_DEVICE_F32 (the
 * phi_dev of shm3d_solve_device); the slabs are gathered over NVLink on rank 0, which extracts the mesh; the other ranks
 * report 0 vertices / triangles.  shm3d_slice works the same way (rank 0's `out` is filled). */
int shm3d_isosurface(shm3d_ctx* ctx, const shm3d_params* p, const void* phi, int32_t field_kind, float isoval,
                     const float* bound_min, const float* bound_max, uint32_t iso_flags, shm3d_iso_stats* out);
/* Copies the last mesh to the host: vertices_out float[n_vertices][3], triangles_out uint32[n_triangles][3]
 * (either may be NULL). */
int shm3d_isosurface_fetch(shm3d_ctx* ctx, float* vertices_out, uint32_t* triangles_out);
/* Device pointers of the last mesh (valid until the next shm3d_isosurface on this context). */
int shm3d_isosurface_device(shm3d_ctx* ctx, const float** d_vertices, const uint32_t** d_triangles);
/* Plane slice through the field (what the reference's slice plane displays, src/main.cpp:101-103): out[a + b*nu] =
 * trilinear interpolant of the node values (src/signed_heat_grid_solver.cpp:405-431) at origin + a*du + b*dv,
 * a < nu, b < nv; NaN where the point is outside the grid.  out: float[nu*nv] on the host. */
int shm3d_slice(shm3d_ctx* ctx, const shm3d_params* p, const void* phi, int32_t field_kind, const double* origin,
                const double* du, const double* dv, int32_t nu, int32_t nv, float* out);

/* Host-logic probes for the CPU test-suite (no device code runs; not part of the product path):
 * constraint rows (src/signed_heat_grid_solver.cpp:80-100, :433-464) and the nested-dissection factor of
 * A D^-1 A^T applied on the host with the same block layout the GPU kernels consume. */
int shm3d_debug_constraints(const shm3d_params* p, int64_t n_sources, const double* pos, int32_t* m_out,
                            int64_t* node_out /*[m*8]*/, double* w_out /*[m*8]*/, int64_t* src_out /*[m]*/,
                            int64_t capacity_rows);
int shm3d_debug_factor_solve(const shm3d_params* p, int64_t n_sources, const double* pos, int32_t uniform,
                             double* v /*[m] in/out*/, int32_t m_expected, double* factor_megabytes,
                             int32_t* tree_height);

/* GPU-test probe: one stencil operation of the PCG / V-cycle (op: 0 p-update + K'p + p.q, 1 Jacobi sweep, 2 sweep + the
 * PCG dots, 3 residual, 4 the two fused first sweeps) on padded vectors of (k1-k0)+2 planes, through the row-streaming
 * kernels (use_tma = 0) or the TMA-staged marching kernels (use_tma = 1); tests/test_gpu_march.py requires identical
 * results.  scal: {mean, beta | omega, omega2}.  reps > 0: also the device time per launch over `reps` launches. */
int shm3d_debug_stencil_op(shm3d_ctx* ctx, int32_t op, int32_t nx, int32_t ny, int32_t nz, int32_t k0, int32_t k1,
                           const float* in0, const float* in1, const float* pw, const double* scal, int32_t use_tma,
                           float* out0, float* out1, double* red /*[2]*/, int32_t reps, double* ms_per_launch);

/* Library / build identification ("shm3d-b200 <version> sm_100a"). */
const char* shm3d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SHM3D_GRID_H */
