// farfield_probe.c -- EXPERIMENT (round-2 preparation, CPU only): how accurate are Steps 1-2 if source clusters that are
// far from a node in units of the kernel's footprint are replaced by a few equivalent sources?
// X(x) = sum_s w_s exp(-lam r)/r.  Clusters: contiguous ranges of <= 32 Morton-sorted sources with centre c, radius rho.
// Per node: d0 = min over clusters of (R + rho) >= the nearest-source distance; clusters with (R - rho) - d0 > tau/lam are
// culled (k_sum's safe rule without its brick slack: dropped terms are < e^-tau of the leading one);
// a kept cluster is "admissible" if lam*rho^2 / max(R - rho, 1e-30) < eps, and is then evaluated through its
// n_eq equivalent sources instead of its members.  With cl_gadm != NULL the rule is g >= cl_gadm[c] instead (a per-cluster
// admissible distance tabulated beforehand from the cluster's own equivalent-source error).
//   gcc -O3 -march=native -fopenmp -shared -fPIC -o libfarfield.so farfield_probe.c -lm
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void farfield_eval(int64_t n_nodes, const double* nodes, int64_t n_cl, const int64_t* cl_first, const int64_t* cl_count,
                   const double* cl_centre, const double* cl_rho, const double* cl_rho_gate, const double* src_pos, const double* src_w,
                   int n_eq, const double* eq_pos, const double* eq_w,  // [n_cl][n_eq][3]
                   double lam, double tau, double eps, const double* cl_gadm,
                   int ng, const double* gaps, int na, const double* cl_normal, const double* cl_err /*[n_cl][ng][na]*/, double tol,
                   const double* node_gamma /* per-node tolerance scale or NULL */, double* S_out /* sum of magnitudes or NULL */,
                   double* X_out, int64_t* pairs_exact, int64_t* pairs_equiv) {
    int64_t pe = 0, pq = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pe, pq)
    for (int64_t i = 0; i < n_nodes; i++) {
        const double x = nodes[3 * i], y = nodes[3 * i + 1], z = nodes[3 * i + 2];
        double d0 = 1e300;
        for (int64_t c = 0; c < n_cl; c++) {
            double dx = x - cl_centre[3 * c], dy = y - cl_centre[3 * c + 1], dz = z - cl_centre[3 * c + 2];
            double g = sqrt(dx * dx + dy * dy + dz * dz) + cl_rho[c];  // upper bound of the nearest-source distance
            if (g < d0) d0 = g;
        }
        double ax = 0, ay = 0, az = 0, S = 0;
        for (int64_t c = 0; c < n_cl; c++) {
            double dx = x - cl_centre[3 * c], dy = y - cl_centre[3 * c + 1], dz = z - cl_centre[3 * c + 2];
            double R = sqrt(dx * dx + dy * dy + dz * dz);
            double g = R - cl_rho[c];
            if (g < 0) g = 0;
            if ((g - d0) * lam > tau) continue;
            int adm;
            if (n_eq <= 0) adm = 0;
            else if (cl_err) {
                // tabulated relative error of the cluster's equivalent sources at (gap, angle from the cluster normal),
                // weighted by the largest share this cluster can have of the node's sum
                adm = 0;
                if (g >= gaps[0]) {
                    int gi = 0;
                    while (gi + 1 < ng && gaps[gi + 1] <= g) gi++;
                    int gj = gi + 1 < ng ? gi + 1 : gi;
                    double cs = fabs(dx * cl_normal[3 * c] + dy * cl_normal[3 * c + 1] + dz * cl_normal[3 * c + 2]) / R;
                    if (cs > 1) cs = 1;
                    double th = acos(cs) / (0.5 * 3.14159265358979323846) * (na - 1);  // 0 .. na-1
                    int ai = (int)th, aj = ai + 1 < na ? ai + 1 : ai;
                    const double* E = cl_err + (size_t)c * ng * na;
                    double e = fmax(fmax(E[gi * na + ai], E[gi * na + aj]), fmax(E[gj * na + ai], E[gj * na + aj]));
                    adm = e * exp(-lam * (g - d0)) < tol * (node_gamma ? node_gamma[i] : 1.0);
                }
            } else if (cl_gadm) adm = g >= cl_gadm[c];
            else adm = lam * cl_rho_gate[c] * cl_rho_gate[c] < eps * (g > 1e-30 ? g : 1e-30);
            if (adm) {
                for (int e = 0; e < n_eq; e++) {
                    const double* p = eq_pos + 3 * (c * n_eq + e);
                    const double* w = eq_w + 3 * (c * n_eq + e);
                    double ex = x - p[0], ey = y - p[1], ez = z - p[2];
                    double r = sqrt(ex * ex + ey * ey + ez * ez);
                    double k = exp(-lam * (r - d0)) / r;  // common factor exp(-lam d0) cancels in the normalisation
                    ax += w[0] * k; ay += w[1] * k; az += w[2] * k;
                    S += sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) * k;
                }
                pq += n_eq;
            } else {
                for (int64_t s = cl_first[c]; s < cl_first[c] + cl_count[c]; s++) {
                    double ex = x - src_pos[3 * s], ey = y - src_pos[3 * s + 1], ez = z - src_pos[3 * s + 2];
                    double r = sqrt(ex * ex + ey * ey + ez * ez);
                    double k = exp(-lam * (r - d0)) / r;
                    ax += src_w[3 * s] * k; ay += src_w[3 * s + 1] * k; az += src_w[3 * s + 2] * k;
                    S += sqrt(src_w[3 * s] * src_w[3 * s] + src_w[3 * s + 1] * src_w[3 * s + 1] + src_w[3 * s + 2] * src_w[3 * s + 2]) * k;
                }
                pe += cl_count[c];
            }
        }
        if (S_out) S_out[i] = S;
        X_out[3 * i] = ax; X_out[3 * i + 1] = ay; X_out[3 * i + 2] = az;
    }
    *pairs_exact = pe;
    *pairs_equiv = pq;
}
