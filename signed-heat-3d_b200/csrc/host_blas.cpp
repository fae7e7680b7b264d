// host_blas.cpp -- the two dense inner kernels of the host-side multifrontal factorisation (projector.cu), written
// with AVX2/FMA intrinsics and compiled by the host compiler only.
//
//  * rank-nb update of the trailing lower triangle of a front (the O(f^2 nb) bulk of the partial Cholesky):
//        F[i][k1 + j] -= sum_t F[i][k0 + t] * F[k1 + j][k0 + t],   j <= i - k1
//    done as C -= A * Bt with Bt the TRANSPOSED panel, so the output row is contiguous: 4 rows x 8 columns of C live in
//    eight ymm accumulators, one broadcast + two loads feed eight FMAs per step (the first version was a dot product
//    per output element: two loads per FMA);
//  * W = L11^-1 by rows (forward substitution in axpy form, unit stride) instead of by columns.
#include <immintrin.h>

#include <algorithm>
#include <cstddef>

namespace shm3d {

// Bt[t * ldb + j] = F[(k1 + j) * f + k0 + t],  j in [0, f - k1), zero-padded up to ldb (a multiple of 8)
void host_transpose_panel(const double* F, int f, int k0, int nb, int k1, double* Bt, int ldb) {
    const int rows = f - k1;
    for (int t = 0; t < nb; t++) {
        double* bt = Bt + (size_t)t * ldb;
        for (int j = 0; j < rows; j++) bt[j] = F[(size_t)(k1 + j) * f + k0 + t];
        for (int j = rows; j < ldb; j++) bt[j] = 0.0;
    }
}

// rows [i0, i1) of the trailing update (lower triangle only; entries right of the diagonal inside an 8-block may be
// overwritten with meaningless values -- the upper triangle of a front is never read)
void host_syrk_rows(double* F, int f, int k0, int nb, int k1, int i0, int i1, const double* Bt, int ldb) {
    const int ncol = f - k1;  // columns of the trailing block
    int i = i0;
    for (; i + 4 <= i1; i += 4) {
        const double* a0 = F + (size_t)i * f + k0;
        const double* a1 = a0 + f;
        const double* a2 = a1 + f;
        const double* a3 = a2 + f;
        double* c0 = F + (size_t)i * f + k1;
        double* c1 = c0 + f;
        double* c2 = c1 + f;
        double* c3 = c2 + f;
        const int jn = std::min(ncol, i + 3 - k1 + 1);  // columns needed by the last of the four rows
        int j0 = 0;
        for (; j0 + 8 <= ncol && j0 < jn; j0 += 8) {
            __m256d s00 = _mm256_setzero_pd(), s01 = s00, s10 = s00, s11 = s00, s20 = s00, s21 = s00, s30 = s00, s31 = s00;
            const double* b = Bt + j0;
            for (int t = 0; t < nb; t++, b += ldb) {
                const __m256d b0 = _mm256_loadu_pd(b), b1 = _mm256_loadu_pd(b + 4);
                __m256d a = _mm256_broadcast_sd(a0 + t);
                s00 = _mm256_fmadd_pd(a, b0, s00);
                s01 = _mm256_fmadd_pd(a, b1, s01);
                a = _mm256_broadcast_sd(a1 + t);
                s10 = _mm256_fmadd_pd(a, b0, s10);
                s11 = _mm256_fmadd_pd(a, b1, s11);
                a = _mm256_broadcast_sd(a2 + t);
                s20 = _mm256_fmadd_pd(a, b0, s20);
                s21 = _mm256_fmadd_pd(a, b1, s21);
                a = _mm256_broadcast_sd(a3 + t);
                s30 = _mm256_fmadd_pd(a, b0, s30);
                s31 = _mm256_fmadd_pd(a, b1, s31);
            }
            _mm256_storeu_pd(c0 + j0, _mm256_sub_pd(_mm256_loadu_pd(c0 + j0), s00));
            _mm256_storeu_pd(c0 + j0 + 4, _mm256_sub_pd(_mm256_loadu_pd(c0 + j0 + 4), s01));
            _mm256_storeu_pd(c1 + j0, _mm256_sub_pd(_mm256_loadu_pd(c1 + j0), s10));
            _mm256_storeu_pd(c1 + j0 + 4, _mm256_sub_pd(_mm256_loadu_pd(c1 + j0 + 4), s11));
            _mm256_storeu_pd(c2 + j0, _mm256_sub_pd(_mm256_loadu_pd(c2 + j0), s20));
            _mm256_storeu_pd(c2 + j0 + 4, _mm256_sub_pd(_mm256_loadu_pd(c2 + j0 + 4), s21));
            _mm256_storeu_pd(c3 + j0, _mm256_sub_pd(_mm256_loadu_pd(c3 + j0), s30));
            _mm256_storeu_pd(c3 + j0 + 4, _mm256_sub_pd(_mm256_loadu_pd(c3 + j0 + 4), s31));
        }
        // tail columns (fewer than 8 left in the row): scalar, lower triangle only
        for (int r = 0; r < 4; r++) {
            const double* a = F + (size_t)(i + r) * f + k0;
            double* c = F + (size_t)(i + r) * f + k1;
            const int jr = std::min(ncol, i + r - k1 + 1);
            for (int j = j0; j < jr; j++) {
                double acc = 0;
                for (int t = 0; t < nb; t++) acc += a[t] * Bt[(size_t)t * ldb + j];
                c[j] -= acc;
            }
        }
    }
    for (; i < i1; i++) {  // remaining rows: 1 x 8
        const double* a = F + (size_t)i * f + k0;
        double* c = F + (size_t)i * f + k1;
        const int jn = std::min(ncol, i - k1 + 1);
        int j0 = 0;
        for (; j0 + 8 <= ncol && j0 < jn; j0 += 8) {
            __m256d s0 = _mm256_setzero_pd(), s1 = s0;
            const double* b = Bt + j0;
            for (int t = 0; t < nb; t++, b += ldb) {
                const __m256d av = _mm256_broadcast_sd(a + t);
                s0 = _mm256_fmadd_pd(av, _mm256_loadu_pd(b), s0);
                s1 = _mm256_fmadd_pd(av, _mm256_loadu_pd(b + 4), s1);
            }
            _mm256_storeu_pd(c + j0, _mm256_sub_pd(_mm256_loadu_pd(c + j0), s0));
            _mm256_storeu_pd(c + j0 + 4, _mm256_sub_pd(_mm256_loadu_pd(c + j0 + 4), s1));
        }
        for (int j = j0; j < jn; j++) {
            double acc = 0;
            for (int t = 0; t < nb; t++) acc += a[t] * Bt[(size_t)t * ldb + j];
            c[j] -= acc;
        }
    }
}

// W = L^-1 for the s x s lower-triangular L stored in the leading block of F (row stride f); W row-major s x s,
// columns [c0, c1) only (independent column ranges can run on different threads); W must be zero on entry.
void host_tri_inverse_rows(const double* F, int f, int s, double* W, int c0, int c1) {
    for (int i = c0; i < s; i++) {
        double* wi = W + (size_t)i * s;
        const double* li = F + (size_t)i * f;
        const int ce = std::min(i, c1);  // columns [c0, ce) get contributions from rows k < i
        for (int k = c0; k < i; k++) {
            const double l = li[k];
            if (l == 0.0) continue;
            const double* wk = W + (size_t)k * s;
            const int ck = std::min(k + 1, ce);  // W[k][c] != 0 only for c <= k
            int c = c0;
            const __m256d lv = _mm256_set1_pd(l);
            for (; c + 4 <= ck; c += 4)
                _mm256_storeu_pd(wi + c, _mm256_fnmadd_pd(lv, _mm256_loadu_pd(wk + c), _mm256_loadu_pd(wi + c)));
            for (; c < ck; c++) wi[c] -= l * wk[c];
        }
        if (i >= c0 && i < c1) wi[i] += 1.0;
        const double inv = 1.0 / li[i];
        for (int c = c0; c < std::min(i + 1, c1); c++) wi[c] *= inv;
    }
}

}  // namespace shm3d
