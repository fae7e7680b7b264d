/*
 * oracle/csrc/shm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU fp64 restatement ("port") of the O(N*M) loops of the reference grid solver.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (signed-heat-3d_b200/) never does.
 *
 * Parity status: checked against the reference's own translation units compiled against a
 * shim (oracle/ref_shim, oracle/_ref/libshm_ref.so, tests/test_reference_build.py); the
 * reference ships no tests or golden vectors for this path.  See oracle/shm_oracle.py.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* yukawaPotential: src/signed_heat_3d.cpp:45-49  exp(-lambda r)/r */
static inline double yukawa(double dx, double dy, double dz, double lambda) {
    double r = sqrt(dx * dx + dy * dy + dz * dz);
    return exp(-lambda * r) / r;
}

/*
 * Steps 1-2 on a z-range [k0,k1) of the grid.
 * src/signed_heat_grid_solver.cpp:48-65 (mesh) and :157-174 (points): for every node,
 * X = sum_s n_s A_s yukawa(x, y_s); Y = X/|X|.  Node index i + j*nx + k*nx*ny (:505-508),
 * node position bboxMin + cell*(i,j,k) (:510-514).  Sources are visited in input order,
 * exactly as the reference's inner loop does; r == 0 gives Inf/NaN like the reference.
 * Y is interleaved [3*idx+p] like the reference's Eigen vector.
 * threads <= 1: single-threaded (what the reference is); otherwise OpenMP over k planes.
 */
void oracle_step12(int nx, int ny, int nz, int k0, int k1, const double* bmin, double cell, double lambda,
                   int64_t M, const double* pos, const double* nrm, const double* area, double* Y, int threads) {
    (void)nz;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int k = k0; k < k1; k++) {
        for (int j = 0; j < ny; j++) {
            for (int i = 0; i < nx; i++) {
                size_t idx = (size_t)i + (size_t)j * nx + (size_t)k * nx * ny;
                double px = bmin[0] + i * cell, py = bmin[1] + j * cell, pz = bmin[2] + k * cell;
                double X0 = 0, X1 = 0, X2 = 0;
                for (int64_t s = 0; s < M; s++) {
                    double w = area[s] * yukawa(px - pos[3 * s], py - pos[3 * s + 1], pz - pos[3 * s + 2], lambda);
                    X0 += nrm[3 * s] * w;
                    X1 += nrm[3 * s + 1] * w;
                    X2 += nrm[3 * s + 2] * w;
                }
                double n = sqrt(X0 * X0 + X1 * X1 + X2 * X2);
                Y[3 * idx] = X0 / n;
                Y[3 * idx + 1] = X1 / n;
                Y[3 * idx + 2] = X2 / n;
            }
        }
    }
}

/*
 * Same loop restricted to rows [j0,j1) of planes [k0,k1): the bounded sample bench.py's cpu_baseline / --impl
 * reference legs time (every node costs the same M kernel evaluations, so any sub-box is representative).
 * Y_box is packed [(k-k0)][(j-j0)][i][3].  OpenMP over (k,j) rows.
 */
void oracle_step12_box(int nx, int ny, int j0, int j1, int k0, int k1, const double* bmin, double cell,
                       double lambda, int64_t M, const double* pos, const double* nrm, const double* area,
                       double* Y_box, int threads) {
    (void)ny;
    const int nj = j1 - j0, nk = k1 - k0;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int row = 0; row < nj * nk; row++) {
        const int k = k0 + row / nj, j = j0 + row % nj;
        for (int i = 0; i < nx; i++) {
            double px = bmin[0] + i * cell, py = bmin[1] + j * cell, pz = bmin[2] + k * cell;
            double X0 = 0, X1 = 0, X2 = 0;
            for (int64_t s = 0; s < M; s++) {
                double w = area[s] * yukawa(px - pos[3 * s], py - pos[3 * s + 1], pz - pos[3 * s + 2], lambda);
                X0 += nrm[3 * s] * w;
                X1 += nrm[3 * s + 1] * w;
                X2 += nrm[3 * s + 2] * w;
            }
            double n = sqrt(X0 * X0 + X1 * X1 + X2 * X2);
            double* o = Y_box + 3 * ((size_t)row * nx + i);
            o[0] = X0 / n;
            o[1] = X1 / n;
            o[2] = X2 / n;
        }
    }
}

/*
 * "As written" variant for the CPU baseline: the reference recomputes each face's
 * barycentre inside the inner loop (src/signed_heat_grid_solver.cpp:55 -> :498-503, a
 * halfedge walk).  Approximated by an indexed 3-vertex gather + average per pair.
 */
void oracle_step12_aswritten(int nx, int ny, int nz, int k0, int k1, const double* bmin, double cell, double lambda,
                             int64_t M, const double* verts, const int32_t* tris, const double* nrm,
                             const double* area, double* Y, int threads) {
    (void)nz;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int k = k0; k < k1; k++) {
        for (int j = 0; j < ny; j++) {
            for (int i = 0; i < nx; i++) {
                size_t idx = (size_t)i + (size_t)j * nx + (size_t)k * nx * ny;
                double px = bmin[0] + i * cell, py = bmin[1] + j * cell, pz = bmin[2] + k * cell;
                double X0 = 0, X1 = 0, X2 = 0;
                for (int64_t s = 0; s < M; s++) {
                    const double* a = verts + 3 * (size_t)tris[3 * s];
                    const double* b = verts + 3 * (size_t)tris[3 * s + 1];
                    const double* c = verts + 3 * (size_t)tris[3 * s + 2];
                    double bx = (a[0] + b[0] + c[0]) / 3., by = (a[1] + b[1] + c[1]) / 3.,
                           bz = (a[2] + b[2] + c[2]) / 3.;
                    double w = area[s] * yukawa(px - bx, py - by, pz - bz, lambda);
                    X0 += nrm[3 * s] * w;
                    X1 += nrm[3 * s + 1] * w;
                    X2 += nrm[3 * s + 2] * w;
                }
                double n = sqrt(X0 * X0 + X1 * X1 + X2 * X2);
                Y[3 * idx] = X0 / n;
                Y[3 * idx + 1] = X1 / n;
                Y[3 * idx + 2] = X2 / n;
            }
        }
    }
}

/*
 * K u = -L u, L from laplacian() src/signed_heat_grid_solver.cpp:278-334:
 * (K u)[idx] = sum over in-range axis neighbours (u[idx] - u[nbr]) / cell^2.
 */
void oracle_apply_K(int nx, int ny, int nz, double cell, const double* u, double* out, int threads) {
    double ic2 = 1.0 / (cell * cell);
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                size_t idx = (size_t)i + (size_t)j * nx + (size_t)k * nx * ny;
                double c = u[idx], s = 0;
                if (i > 0) s += c - u[idx - 1];
                if (i < nx - 1) s += c - u[idx + 1];
                if (j > 0) s += c - u[idx - nx];
                if (j < ny - 1) s += c - u[idx + nx];
                if (k > 0) s += c - u[idx - (size_t)nx * ny];
                if (k < nz - 1) s += c - u[idx + (size_t)nx * ny];
                out[idx] = s * ic2;
            }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
