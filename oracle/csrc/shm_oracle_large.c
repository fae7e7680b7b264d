/*
 * oracle/csrc/shm_oracle_large.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64 helpers that make the oracle affordable at the BASELINE.json grid sizes (256^3, 512^3) on this container's
 * host cores.  Used ONLY by tests/golden/make_golden_baseline.py (run here, results committed as fixtures) and by
 * tests/test_oracle_large.py (which pins these helpers to the plain restatement in shm_oracle.c / shm_oracle.py).
 * Nothing under signed-heat-3d_b200/ loads this file.
 *
 *  oracle_step12_bricks   Steps 1-2 (reference src/signed_heat_grid_solver.cpp:48-65 / :157-174 with yukawaPotential,
 *                         src/signed_heat_3d.cpp:45-49) in double precision, AVX-512 over the nodes of an 8^3 brick,
 *                         with far clusters of sources skipped under an A-POSTERIORI bound: for every node the summed
 *                         magnitude of everything that was skipped is at most `eps` times the largest component of the
 *                         X that was kept -- otherwise the brick is redone with every source.  eps = 1e-13 leaves the
 *                         result equal to the plain loop to rounding (checked on subsamples by the tests); it is the
 *                         same sum, evaluated in another order with provably negligible terms left out.
 *                         The division X / |X| is the reference's: sqrt of the sum of squares in double, so the
 *                         underflow artefact of `X /= X.norm()` (:61) at far nodes of finely triangulated inputs is
 *                         reproduced (no flush-to-zero: this file is compiled without -ffast-math).
 *  oracle_cg_*            fused OpenMP kernels of the fp64 projected CG on the matrix-free stencil K = -L
 *                         (laplacian(), :278-334): the same iteration as shm_oracle.solve_projected_cg.
 */
#include <immintrin.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* glibc libmvec: 8 x exp(double), <= 4 ulp (irrelevant at the 1e-13 level this file is checked to: see the tests) */
__m512d _ZGVeN8v_exp(__m512d x);

#define BR 8
#define BN (BR * BR * BR)

typedef struct {
    int nx, ny, nz;
    const double* bmin;
    double cell, lambda;
    const double *sx, *sy, *sz;    /* sources, cluster-sorted, structure of arrays */
    const double *wx, *wy, *wz;    /* n * A */
    int nc;
    const int64_t* cbeg;           /* nc + 1 cluster offsets */
    const double* ccen;            /* 3 * nc */
    const double* crad;            /* nc */
    const double* cmass;           /* nc: sum of |A| */
} Job;

static void brick_sum(const Job* J, const double* px, const double* py, const double* pz, const int* keep, int nkeep,
                      double* X0, double* X1, double* X2) {
    const __m512d nlam = _mm512_set1_pd(-J->lambda);
    for (int v = 0; v < BN; v += 8) {
        _mm512_storeu_pd(X0 + v, _mm512_setzero_pd());
        _mm512_storeu_pd(X1 + v, _mm512_setzero_pd());
        _mm512_storeu_pd(X2 + v, _mm512_setzero_pd());
    }
    for (int t = 0; t < nkeep; t++) {
        const int c = keep[t];
        for (int64_t s = J->cbeg[c]; s < J->cbeg[c + 1]; s++) {
            const __m512d sx = _mm512_set1_pd(J->sx[s]), sy = _mm512_set1_pd(J->sy[s]), sz = _mm512_set1_pd(J->sz[s]);
            const __m512d wx = _mm512_set1_pd(J->wx[s]), wy = _mm512_set1_pd(J->wy[s]), wz = _mm512_set1_pd(J->wz[s]);
            for (int v = 0; v < BN; v += 8) {
                const __m512d dx = _mm512_sub_pd(_mm512_loadu_pd(px + v), sx);
                const __m512d dy = _mm512_sub_pd(_mm512_loadu_pd(py + v), sy);
                const __m512d dz = _mm512_sub_pd(_mm512_loadu_pd(pz + v), sz);
                const __m512d r2 = _mm512_fmadd_pd(dz, dz, _mm512_fmadd_pd(dy, dy, _mm512_mul_pd(dx, dx)));
                const __m512d r = _mm512_sqrt_pd(r2);
                const __m512d w = _mm512_div_pd(_ZGVeN8v_exp(_mm512_mul_pd(nlam, r)), r);
                _mm512_storeu_pd(X0 + v, _mm512_fmadd_pd(wx, w, _mm512_loadu_pd(X0 + v)));
                _mm512_storeu_pd(X1 + v, _mm512_fmadd_pd(wy, w, _mm512_loadu_pd(X1 + v)));
                _mm512_storeu_pd(X2 + v, _mm512_fmadd_pd(wz, w, _mm512_loadu_pd(X2 + v)));
            }
        }
    }
}

/*
 * Y (interleaved [3*idx + a], idx = i + j*nx + (k - k0)*nx*ny) for planes [k0, k1).
 * stats[0] = pairs evaluated, stats[1] = bricks redone in full, stats[2] = max over nodes of (skipped bound / max|X|).
 */
void oracle_step12_bricks(int nx, int ny, int nz, int k0, int k1, const double* bmin, double cell, double lambda,
                          const double* sx, const double* sy, const double* sz, const double* wx, const double* wy,
                          const double* wz, int nc, const int64_t* cbeg, const double* ccen, const double* crad,
                          const double* cmass, double tau, double eps, double* Y, double* stats, int threads) {
    Job J = {nx, ny, nz, bmin, cell, lambda, sx, sy, sz, wx, wy, wz, nc, cbeg, ccen, crad, cmass};
    const int bx = (nx + BR - 1) / BR, by = (ny + BR - 1) / BR, bz = (k1 - k0 + BR - 1) / BR;
    const long nbricks = (long)bx * by * bz;
    double pairs = 0, redone = 0, worst = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : pairs, redone) reduction(max : worst)
    {
        double *px = aligned_alloc(64, BN * 8), *py = aligned_alloc(64, BN * 8), *pz = aligned_alloc(64, BN * 8);
        double *X0 = aligned_alloc(64, BN * 8), *X1 = aligned_alloc(64, BN * 8), *X2 = aligned_alloc(64, BN * 8);
        int* keep = malloc(sizeof(int) * (size_t)nc);
        double* rlo = malloc(sizeof(double) * (size_t)nc);
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < nbricks; b++) {
            const int bi = (int)(b % bx), bj = (int)((b / bx) % by), bk = (int)(b / ((long)bx * by));
            const int i0 = bi * BR, j0 = bj * BR, kk0 = k0 + bk * BR;
            /* node positions: bboxMin + cell*i (:510-514); nodes outside the grid repeat the last valid one */
            for (int c = 0; c < BN; c++) {
                int i = i0 + (c % BR), j = j0 + ((c / BR) % BR), k = kk0 + c / (BR * BR);
                if (i > nx - 1) i = nx - 1;
                if (j > ny - 1) j = ny - 1;
                if (k > k1 - 1) k = k1 - 1;
                px[c] = bmin[0] + i * cell;
                py[c] = bmin[1] + j * cell;
                pz[c] = bmin[2] + k * cell;
            }
            const double lo[3] = {px[0], py[0], pz[0]}, hi[3] = {px[BN - 1], py[BN - 1], pz[BN - 1]};
            const double ce[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
            const double hd = 0.5 * sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) +
                                         (hi[2] - lo[2]) * (hi[2] - lo[2]));
            /* lower bound of the distance from any node of the brick to any source of cluster c; and the best upper
             * bound of the distance from any node of the brick to its nearest source */
            double rbest = 1e300;
            for (int c = 0; c < nc; c++) {
                const double* q = ccen + 3 * c;
                double d2 = 0;
                for (int a = 0; a < 3; a++) {
                    double d = q[a] < lo[a] ? lo[a] - q[a] : (q[a] > hi[a] ? q[a] - hi[a] : 0.0);
                    d2 += d * d;
                }
                double l = sqrt(d2) - crad[c];
                rlo[c] = l > 0 ? l : 0;
                double dc = sqrt((q[0] - ce[0]) * (q[0] - ce[0]) + (q[1] - ce[1]) * (q[1] - ce[1]) +
                                 (q[2] - ce[2]) * (q[2] - ce[2])) + hd + crad[c];
                if (dc < rbest) rbest = dc;
            }
            int nkeep = 0;
            double skipped = 0; /* sum over skipped clusters of mass * exp(-lambda rlo)/rlo >= what they could add */
            for (int c = 0; c < nc; c++) {
                if (lambda * (rlo[c] - rbest) <= tau) keep[nkeep++] = c;
                else skipped += cmass[c] * exp(-lambda * rlo[c]) / rlo[c];
            }
            brick_sum(&J, px, py, pz, keep, nkeep, X0, X1, X2);
            double ratio = 0;
            if (nkeep < nc) {
                for (int c = 0; c < BN; c++) {
                    double m = fmax(fmax(fabs(X0[c]), fabs(X1[c])), fabs(X2[c]));
                    double q = m > 0 ? skipped / m : (skipped > 0 ? 1e300 : 0);
                    if (q > ratio) ratio = q;
                }
                if (!(ratio <= eps)) { /* not provably negligible: every source */
                    for (int c = 0; c < nc; c++) keep[c] = c;
                    nkeep = nc;
                    brick_sum(&J, px, py, pz, keep, nkeep, X0, X1, X2);
                    redone += 1;
                    ratio = 0;
                }
            }
            if (ratio > worst) worst = ratio;
            for (int t = 0; t < nkeep; t++) pairs += (double)(cbeg[keep[t] + 1] - cbeg[keep[t]]) * BN;
            for (int c = 0; c < BN; c++) {
                const int i = i0 + (c % BR), j = j0 + ((c / BR) % BR), k = kk0 + c / (BR * BR);
                if (i >= nx || j >= ny || k >= k1) continue;
                const size_t idx = (size_t)i + (size_t)j * nx + (size_t)(k - k0) * nx * ny;
                /* X /= X.norm() as the reference evaluates it (:61): plain double, squares may underflow */
                const double n = sqrt(X0[c] * X0[c] + X1[c] * X1[c] + X2[c] * X2[c]);
                Y[3 * idx] = X0[c] / n;
                Y[3 * idx + 1] = X1[c] / n;
                Y[3 * idx + 2] = X2[c] / n;
            }
        }
        free(px); free(py); free(pz); free(X0); free(X1); free(X2); free(keep); free(rlo);
    }
    if (stats) {
        stats[0] = pairs;
        stats[1] = redone;
        stats[2] = worst;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * projected CG, fused dense parts.  K' = cell^2 K (integer stencil), as in oracle_apply_K.
 * ------------------------------------------------------------------------------------------------------------------ */

/* q = K p (1/cell^2 included), returns p.q */
double oracle_cg_apply_dot(int nx, int ny, int nz, double cell, const double* p, double* q, int threads) {
    const double ic2 = 1.0 / (cell * cell);
    const size_t pl = (size_t)nx * ny;
    double acc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : acc)
    for (int k = 0; k < nz; k++) {
        double a = 0;
        for (int j = 0; j < ny; j++) {
            const size_t row = (size_t)j * nx + (size_t)k * pl;
            const int cyz = (j > 0) + (j < ny - 1) + (k > 0) + (k < nz - 1);
            const double* c = p + row;
            const double* ym = j > 0 ? c - nx : NULL;
            const double* yp = j < ny - 1 ? c + nx : NULL;
            const double* zm = k > 0 ? c - pl : NULL;
            const double* zp = k < nz - 1 ? c + pl : NULL;
            for (int i = 0; i < nx; i++) {
                double s = (cyz + (i > 0) + (i < nx - 1)) * c[i];
                if (i > 0) s -= c[i - 1];
                if (i < nx - 1) s -= c[i + 1];
                if (ym) s -= ym[i];
                if (yp) s -= yp[i];
                if (zm) s -= zm[i];
                if (zp) s -= zp[i];
                s *= ic2;
                q[row + i] = s;
                a += c[i] * s;
            }
        }
        acc += a;
    }
    return acc;
}

/* x += alpha p ; r -= alpha q ; returns r.r */
double oracle_cg_update(size_t n, double alpha, const double* p, const double* q, double* x, double* r, int threads) {
    double acc = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : acc)
    for (size_t i = 0; i < n; i++) {
        x[i] += alpha * p[i];
        const double v = r[i] - alpha * q[i];
        r[i] = v;
        acc += v * v;
    }
    return acc;
}

/* p = r + beta p */
void oracle_cg_direction(size_t n, double beta, const double* r, double* p, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i];
}

/* b = D^T Y on the interior form of gradient() (:336-402): row 3*idx+a of D is u[next_a] - u[idx] (forward) or, on the
 * far face, u[idx] - u[prev_a]; everything / cell.  Non-finite results are zeroed when scrub != 0 (:72-74).
 * Returns the number of scrubbed entries. */
long oracle_div_rhs(int nx, int ny, int nz, double cell, const double* Y, int scrub, double* b, int threads) {
    const size_t pl = (size_t)nx * ny;
    const double ic = 1.0 / cell;
    long bad = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : bad)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const size_t idx = (size_t)i + (size_t)j * nx + (size_t)k * pl;
                /* column idx of D collects: from its own rows (a = x,y,z): -1 (forward row) or +1 (backward row at the
                 * far face); from the row of the previous node along a (if that row is a forward row, i.e. always: the
                 * previous node is never on the far face): +1; from the row of the next node along a if that node is
                 * on the far face (backward row: -1 on its predecessor = this node). */
                double s = 0;
                const int n3[3] = {nx, ny, nz}, id3[3] = {i, j, k};
                const size_t st[3] = {1, (size_t)nx, pl};
                for (int a = 0; a < 3; a++) {
                    const int t = id3[a], n = n3[a];
                    s += (t == n - 1 ? 1.0 : -1.0) * Y[3 * idx + a];
                    if (t > 0) s += Y[3 * (idx - st[a]) + a];
                    if (t == n - 2) s -= Y[3 * (idx + st[a]) + a];
                }
                s *= ic;
                if (scrub && !isfinite(s)) {
                    s = 0;
                    bad++;
                }
                b[idx] = s;
            }
    return bad;
}
