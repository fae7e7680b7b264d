// grid_rows.cuh -- row-oriented stencil kernels for grids with nx % 4 == 0 (every grid the reference can produce:
// nx = 16 * 2^h).  Included by grid_ops.cu inside its anonymous namespace.
//
// ncu on the first, element-indexed versions showed the stencil kernels were instruction-bound (63-75 % SM throughput
// at 45 % of HBM peak): per-quad integer divisions to decode (i,j,k), per-element fp32 divisions by the diagonal, two
// extra scalar loads per quad for the x-neighbours.  Here a WARP owns a grid row: (j,k) and the four boundary
// predicates are warp-uniform and computed once per row, the reciprocal diagonals once per row, the x-neighbours come
// from warp shuffles (only lanes 0/31 touch memory for them), and each lane streams float4 quads of the row.
#pragma once

struct RowInfo {
    int j, k;         // global y / z index of the row
    unsigned base;    // element index of node (0, j, k - k0)
    bool ym, yp, zm, zp;
    float cyz;        // number of in-range y/z neighbours
};

__device__ __forceinline__ RowInfo row_info(const LevelDims& L, unsigned row) {
    RowInfo R;
    const unsigned kl = row / (unsigned)L.ny;
    R.j = (int)(row - kl * (unsigned)L.ny);
    R.k = L.k0 + (int)kl;
    R.base = row * (unsigned)L.nx;
    R.ym = R.j > 0;
    R.yp = R.j < L.ny - 1;
    R.zm = R.k > 0;
    R.zp = R.k < L.nz - 1;
    R.cyz = (float)((int)R.ym + (int)R.yp + (int)R.zm + (int)R.zp);
    return R;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

#define ROWS_BEGIN(L_)                                                                             \
    const int lane = threadIdx.x & 31;                                                             \
    const unsigned _gw = (blockIdx.x * kT + threadIdx.x) >> 5, _nw = (gridDim.x * kT) >> 5;        \
    const unsigned _nrows = (unsigned)(L_).ny * (unsigned)(L_).nzl();                              \
    const int nq = (L_).nx >> 2;                                                                   \
    const unsigned pl = (unsigned)(L_).nx * (unsigned)(L_).ny;                                     \
    (void)pl;                                                                                      \
    for (unsigned _row = _gw; _row < _nrows; _row += _nw) {                                        \
        const RowInfo R = row_info((L_), _row);                                                    \
        for (int _q0 = 0; _q0 < nq; _q0 += 32) {                                                   \
            const int q = _q0 + lane;                                                              \
            const bool act = q < nq;                                                               \
            const unsigned e = R.base + 4u * (unsigned)(act ? q : 0);
#define ROWS_END() \
        }          \
    }

// x-neighbours of a quad whose (possibly derived) values are c: shuffles, lanes 0 / 31 use the supplied scalars
__device__ __forceinline__ void x_neighbours(const float4& c, int lane, int q, int nq, float lane0_left, float lane31_right,
                                             float& l, float& r) {
    l = __shfl_up_sync(0xffffffffu, c.w, 1);
    r = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0) l = lane0_left;
    if (lane == 31) r = lane31_right;
    if (q == 0) l = 0.f;
    if (q >= nq - 1) r = 0.f;
}

// K'u of a quad from its centre / y / z neighbour quads and x end scalars; cnt = per-element diagonal
__device__ __forceinline__ float4 stencil_quad(const float4& c, const float4& a, const float4& b, const float4& d,
                                               const float4& f, float l, float r, float cnt0, float cnt12, float cnt3) {
    float4 K;
    K.x = cnt0 * c.x - (l + c.y + a.x + b.x + d.x + f.x);
    K.y = cnt12 * c.y - (c.x + c.z + a.y + b.y + d.y + f.y);
    K.z = cnt12 * c.z - (c.y + c.w + a.z + b.z + d.z + f.z);
    K.w = cnt3 * c.w - (c.z + r + a.w + b.w + d.w + f.w);
    return K;
}

// ---------------------------------------------------------------- q = K'p, out = sum p q
__global__ void __launch_bounds__(kT) k_row_stencil_dot(LevelDims L, const float* __restrict__ p, float* __restrict__ qo,
                                                        RedScratch rs, double* out) {
    double acc[1] = {0.0};
    ROWS_BEGIN(L)
        const float4 c = ld4(p + e);
        const float4 a = R.ym ? ld4(p + e - L.nx) : zero4();
        const float4 b = R.yp ? ld4(p + e + L.nx) : zero4();
        const float4 d = R.zm ? ld4(p + (ptrdiff_t)e - (ptrdiff_t)pl) : zero4();
        const float4 f = R.zp ? ld4(p + e + pl) : zero4();
        const float sl = (lane == 0 && q > 0) ? p[e - 1] : 0.f;
        const float sr = (lane == 31 && q < nq - 1) ? p[e + 4] : 0.f;
        float l, r;
        x_neighbours(c, lane, q, nq, sl, sr, l, r);
        const float cin = R.cyz + 2.f;
        const float4 K = stencil_quad(c, a, b, d, f, l, r, q == 0 ? cin - 1.f : cin, cin, q == nq - 1 ? cin - 1.f : cin);
        if (act) {
            st4(qo + e, K);
            acc[0] += (double)(c.x * K.x + c.y * K.y + c.z * K.z + c.w * K.w);
        }
    ROWS_END()
    block_reduce_commit<1>(acc, rs, out);
}

// ---------------------------------------------------------------- damped Jacobi sweep (optionally with the PCG dots)
template <bool DOT>
__global__ void __launch_bounds__(kT) k_row_smooth(LevelDims L, float* __restrict__ xo, const float* __restrict__ x,
                                                   const float* __restrict__ bb, const double* sum_b, double n_global,
                                                   float omega, RedScratch rs, double* out) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    double acc[2] = {0.0, 0.0};
    ROWS_BEGIN(L)
        const float4 c = ld4(x + e);
        const float4 a = R.ym ? ld4(x + e - L.nx) : zero4();
        const float4 b = R.yp ? ld4(x + e + L.nx) : zero4();
        const float4 d = R.zm ? ld4(x + (ptrdiff_t)e - (ptrdiff_t)pl) : zero4();
        const float4 f = R.zp ? ld4(x + e + pl) : zero4();
        const float4 rhs = ld4(bb + e);
        const float sl = (lane == 0 && q > 0) ? x[e - 1] : 0.f;
        const float sr = (lane == 31 && q < nq - 1) ? x[e + 4] : 0.f;
        float l, r;
        x_neighbours(c, lane, q, nq, sl, sr, l, r);
        const float cin = R.cyz + 2.f, ced = R.cyz + 1.f;
        const float win = omega / cin, wed = omega / ced;  // warp-uniform: two divisions per row
        const bool e0 = q == 0, e3 = q == nq - 1;
        const float4 K = stencil_quad(c, a, b, d, f, l, r, e0 ? ced : cin, cin, e3 ? ced : cin);
        float4 o;
        o.x = fmaf(e0 ? wed : win, (rhs.x - shift) - K.x, c.x);
        o.y = fmaf(win, (rhs.y - shift) - K.y, c.y);
        o.z = fmaf(win, (rhs.z - shift) - K.z, c.z);
        o.w = fmaf(e3 ? wed : win, (rhs.w - shift) - K.w, c.w);
        if (act) {
            st4(xo + e, o);
            if (DOT) {
                acc[0] += (double)(rhs.x * o.x + rhs.y * o.y + rhs.z * o.z + rhs.w * o.w);
                acc[1] += (double)((o.x + o.y) + (o.z + o.w));
            }
        }
    ROWS_END()
    if (DOT) block_reduce_commit<2>(acc, rs, out);
}

// ---------------------------------------------------------------- r = (b - shift) - K'x
__global__ void __launch_bounds__(kT) k_row_residual(LevelDims L, const float* __restrict__ x, const float* __restrict__ bb,
                                                     const double* sum_b, double n_global, float* __restrict__ ro) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    ROWS_BEGIN(L)
        const float4 c = ld4(x + e);
        const float4 a = R.ym ? ld4(x + e - L.nx) : zero4();
        const float4 b = R.yp ? ld4(x + e + L.nx) : zero4();
        const float4 d = R.zm ? ld4(x + (ptrdiff_t)e - (ptrdiff_t)pl) : zero4();
        const float4 f = R.zp ? ld4(x + e + pl) : zero4();
        const float4 rhs = ld4(bb + e);
        const float sl = (lane == 0 && q > 0) ? x[e - 1] : 0.f;
        const float sr = (lane == 31 && q < nq - 1) ? x[e + 4] : 0.f;
        float l, r;
        x_neighbours(c, lane, q, nq, sl, sr, l, r);
        const float cin = R.cyz + 2.f;
        const float4 K = stencil_quad(c, a, b, d, f, l, r, q == 0 ? cin - 1.f : cin, cin, q == nq - 1 ? cin - 1.f : cin);
        if (act) st4(ro + e, make_float4((rhs.x - shift) - K.x, (rhs.y - shift) - K.y, (rhs.z - shift) - K.z, (rhs.w - shift) - K.w));
    ROWS_END()
}

// ---------------------------------------------------------------- first two Jacobi sweeps from a zero guess, one pass over b
//   x1 = omega (b - shift)/d ;  x2 = x1 + omega ((b - shift) - K'x1)/d, with x1 of the six neighbours recomputed from b
__global__ void __launch_bounds__(kT) k_row_smooth01(LevelDims L, float* __restrict__ xo, const float* __restrict__ bb,
                                                     const double* sum_b, double n_global, float omega, float omega2) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    ROWS_BEGIN(L)
        // diagonals of this row and of the four adjacent rows (warp-uniform)
        const int cy = (int)R.ym + (int)R.yp, cz = (int)R.zm + (int)R.zp;
        const int cy_m = (int)(R.j - 1 > 0) + 1, cy_p = 1 + (int)(R.j + 1 < L.ny - 1);
        const int cz_m = (int)(R.k - 1 > 0) + 1, cz_p = 1 + (int)(R.k + 1 < L.nz - 1);
        const float cin = R.cyz + 2.f, ced = R.cyz + 1.f;
        const float wc_in = omega / cin, wc_ed = omega / ced;
        const float wa_in = omega / (float)(cy_m + cz + 2), wa_ed = omega / (float)(cy_m + cz + 1);
        const float wb_in = omega / (float)(cy_p + cz + 2), wb_ed = omega / (float)(cy_p + cz + 1);
        const float wd_in = omega / (float)(cy + cz_m + 2), wd_ed = omega / (float)(cy + cz_m + 1);
        const float wf_in = omega / (float)(cy + cz_p + 2), wf_ed = omega / (float)(cy + cz_p + 1);
        const bool e0 = q == 0, e3 = q == nq - 1;
        auto x1 = [&](const float4& v, float win, float wed) {
            return make_float4((e0 ? wed : win) * (v.x - shift), win * (v.y - shift), win * (v.z - shift),
                               (e3 ? wed : win) * (v.w - shift));
        };
        const float4 rhs = ld4(bb + e);
        const float4 c = x1(rhs, wc_in, wc_ed);
        const float4 a = R.ym ? x1(ld4(bb + e - L.nx), wa_in, wa_ed) : zero4();
        const float4 b = R.yp ? x1(ld4(bb + e + L.nx), wb_in, wb_ed) : zero4();
        const float4 d = R.zm ? x1(ld4(bb + (ptrdiff_t)e - (ptrdiff_t)pl), wd_in, wd_ed) : zero4();
        const float4 f = R.zp ? x1(ld4(bb + e + pl), wf_in, wf_ed) : zero4();
        // x-end neighbours that live in another warp iteration: always x-interior nodes of this row
        const float sl = (lane == 0 && q > 0) ? wc_in * (bb[e - 1] - shift) : 0.f;
        const float sr = (lane == 31 && q < nq - 1) ? wc_in * (bb[e + 4] - shift) : 0.f;
        float l, r;
        x_neighbours(c, lane, q, nq, sl, sr, l, r);
        const float4 K = stencil_quad(c, a, b, d, f, l, r, e0 ? ced : cin, cin, e3 ? ced : cin);
        float4 o;
        const float w2_in = omega2 / cin, w2_ed = omega2 / ced;  // the second sweep may use another damping (Chebyshev)
        o.x = fmaf(e0 ? w2_ed : w2_in, (rhs.x - shift) - K.x, c.x);
        o.y = fmaf(w2_in, (rhs.y - shift) - K.y, c.y);
        o.z = fmaf(w2_in, (rhs.z - shift) - K.z, c.z);
        o.w = fmaf(e3 ? w2_ed : w2_in, (rhs.w - shift) - K.w, c.w);
        if (act) st4(xo + e, o);
    ROWS_END()
}

// ---------------------------------------------------------------- p <- (z - mean) + beta p ; q = K'p ; out = sum p q
// The new p at the six neighbours is recomputed from z and the old p there, so p_new must not alias p_old.
__global__ void __launch_bounds__(kT) k_row_update_p_stencil(LevelDims L, float* __restrict__ pn, const float* __restrict__ po,
                                                             const float* __restrict__ z, float* __restrict__ qo,
                                                             const double* sum_z, double n_global, const double* rho_new,
                                                             const double* rho_old, int first, RedScratch rs, double* out) {
    const float mean = (float)(*sum_z / n_global);
    const float beta = first ? 0.f : (float)(*rho_new / *rho_old);
    double acc[1] = {0.0};
    auto comb = [&](const float4& zv, const float4& pv) {
        return make_float4(fmaf(beta, pv.x, zv.x - mean), fmaf(beta, pv.y, zv.y - mean), fmaf(beta, pv.z, zv.z - mean),
                           fmaf(beta, pv.w, zv.w - mean));
    };
    ROWS_BEGIN(L)
        const float4 c = comb(ld4(z + e), ld4(po + e));
        const float4 a = R.ym ? comb(ld4(z + e - L.nx), ld4(po + e - L.nx)) : zero4();
        const float4 b = R.yp ? comb(ld4(z + e + L.nx), ld4(po + e + L.nx)) : zero4();
        const float4 d = R.zm ? comb(ld4(z + (ptrdiff_t)e - (ptrdiff_t)pl), ld4(po + (ptrdiff_t)e - (ptrdiff_t)pl)) : zero4();
        const float4 f = R.zp ? comb(ld4(z + e + pl), ld4(po + e + pl)) : zero4();
        const float sl = (lane == 0 && q > 0) ? fmaf(beta, po[e - 1], z[e - 1] - mean) : 0.f;
        const float sr = (lane == 31 && q < nq - 1) ? fmaf(beta, po[e + 4], z[e + 4] - mean) : 0.f;
        float l, r;
        x_neighbours(c, lane, q, nq, sl, sr, l, r);
        const float cin = R.cyz + 2.f;
        const float4 K = stencil_quad(c, a, b, d, f, l, r, q == 0 ? cin - 1.f : cin, cin, q == nq - 1 ? cin - 1.f : cin);
        if (act) {
            st4(pn + e, c);
            st4(qo + e, K);
            acc[0] += (double)(c.x * K.x + c.y * K.y + c.z * K.z + c.w * K.w);
        }
    ROWS_END()
    block_reduce_commit<1>(acc, rs, out);
}

// ---------------------------------------------------------------- restriction bc = 0.5 P^T r   (Lc.nx % 4 == 0)
// A warp owns a COARSE row (J,K): lanes stream 8 fine values (two float4) of each of the <= 16 contributing fine rows,
// reduce them over y/z in registers, exchange the two x-end values by shuffle, then apply the x weights.
__device__ __forceinline__ void rweights4(int I, int nc, float (&wt)[4]) {
    wt[0] = (I > 0) ? 0.25f : 0.f;
    wt[1] = (I > 0) ? 0.75f : 1.0f;
    wt[2] = (I < nc - 1) ? 0.75f : 1.0f;
    wt[3] = (I < nc - 1) ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(kT) k_row_restrict(LevelDims Lf, LevelDims Lc, const float* __restrict__ r,
                                                     float* __restrict__ bc) {
    const ptrdiff_t plf = (ptrdiff_t)Lf.plane();
    ROWS_BEGIN(Lc)
        float wy[4], wz[4];
        rweights4(R.j, Lc.ny, wy);
        rweights4(R.k, Lc.nz, wz);
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float sR = 0.f;  // fine value right of this lane's 8 (lane 31 only)
        float sL = 0.f;  // fine value left of this lane's 8 (lane 0 only)
        const bool needR = lane == 31 && q < nq - 1, needL = lane == 0 && q > 0;
#pragma unroll
        for (int cz = 0; cz < 4; cz++) {
            if (wz[cz] == 0.f) continue;
            const int kf = 2 * R.k - 1 + cz - Lf.k0;  // local fine plane (may be a ghost plane: -1 or nzl)
#pragma unroll
            for (int cy = 0; cy < 4; cy++) {
                if (wy[cy] == 0.f) continue;
                const int jf = 2 * R.j - 1 + cy;
                const float w = wy[cy] * wz[cz];
                const float* row = r + (ptrdiff_t)kf * plf + (ptrdiff_t)jf * Lf.nx + 8 * (act ? q : 0);
                const float4 u = ld4(row), v = ld4(row + 4);
                s[0] = fmaf(w, u.x, s[0]);
                s[1] = fmaf(w, u.y, s[1]);
                s[2] = fmaf(w, u.z, s[2]);
                s[3] = fmaf(w, u.w, s[3]);
                s[4] = fmaf(w, v.x, s[4]);
                s[5] = fmaf(w, v.y, s[5]);
                s[6] = fmaf(w, v.z, s[6]);
                s[7] = fmaf(w, v.w, s[7]);
                if (needR) sR = fmaf(w, row[8], sR);
                if (needL) sL = fmaf(w, row[-1], sL);
            }
        }
        float left = __shfl_up_sync(0xffffffffu, s[7], 1);
        float right = __shfl_down_sync(0xffffffffu, s[0], 1);
        if (lane == 0) left = sL;
        if (lane == 31) right = sR;
        // coarse node t gathers fine 2t-1 .. 2t+2 of this lane's window (index -1 = left, 8 = right)
        const bool first = q == 0, last = q == nq - 1;
        float4 o;
        o.x = first ? (s[0] + 0.75f * s[1] + 0.25f * s[2]) : (0.25f * left + 0.75f * s[0] + 0.75f * s[1] + 0.25f * s[2]);
        o.y = 0.25f * s[1] + 0.75f * s[2] + 0.75f * s[3] + 0.25f * s[4];
        o.z = 0.25f * s[3] + 0.75f * s[4] + 0.75f * s[5] + 0.25f * s[6];
        o.w = last ? (0.25f * s[5] + 0.75f * s[6] + s[7]) : (0.25f * s[5] + 0.75f * s[6] + 0.75f * s[7] + 0.25f * right);
        if (act) st4(bc + e, make_float4(0.5f * o.x, 0.5f * o.y, 0.5f * o.z, 0.5f * o.w));
    ROWS_END()
}

// ---------------------------------------------------------------- x += P ec  (clamped cell-centred trilinear prolongation)
// A warp owns a FINE row: the four contributing coarse rows are read as float2 (coarse nodes 2q, 2q+1 of fine quad q),
// interpolated in y/z first, the two x-end coarse values come from the neighbouring lanes.
__global__ void __launch_bounds__(kT) k_row_prolong_add(LevelDims Lf, LevelDims Lc, float* __restrict__ x,
                                                        const float* __restrict__ ec) {
    const ptrdiff_t plc = (ptrdiff_t)Lc.plane();
    ROWS_BEGIN(Lf)
        const int J0 = R.j >> 1, K0 = R.k >> 1;
        const int J1 = min(max((R.j & 1) ? J0 + 1 : J0 - 1, 0), Lc.ny - 1);
        const int K1 = min(max((R.k & 1) ? K0 + 1 : K0 - 1, 0), Lc.nz - 1);
        const float* r00 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;
        const float* r01 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        const float* r10 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;  // may be a ghost plane
        const float* r11 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        const int I = 2 * (act ? q : 0);
        auto yz = [&](int off) {  // y/z-interpolated coarse value at coarse column I + off
            return 0.75f * (0.75f * r00[I + off] + 0.25f * r01[I + off]) + 0.25f * (0.75f * r10[I + off] + 0.25f * r11[I + off]);
        };
        const float2 a00 = *reinterpret_cast<const float2*>(r00 + I), a01 = *reinterpret_cast<const float2*>(r01 + I);
        const float2 a10 = *reinterpret_cast<const float2*>(r10 + I), a11 = *reinterpret_cast<const float2*>(r11 + I);
        const float c0 = 0.75f * (0.75f * a00.x + 0.25f * a01.x) + 0.25f * (0.75f * a10.x + 0.25f * a11.x);
        const float c1 = 0.75f * (0.75f * a00.y + 0.25f * a01.y) + 0.25f * (0.75f * a10.y + 0.25f * a11.y);
        float cl = __shfl_up_sync(0xffffffffu, c1, 1);
        float cr = __shfl_down_sync(0xffffffffu, c0, 1);
        if (lane == 0 && q > 0) cl = yz(-1);
        if (lane == 31 && q < nq - 1) cr = yz(2);
        if (q == 0) cl = c0;        // clamp at the domain boundary
        if (q >= nq - 1) cr = c1;
        if (act) {
            float4 v = ld4(x + e);
            v.x += 0.75f * c0 + 0.25f * cl;
            v.y += 0.75f * c0 + 0.25f * c1;
            v.z += 0.75f * c1 + 0.25f * c0;
            v.w += 0.75f * c1 + 0.25f * cr;
            st4(x + e, v);
        }
    ROWS_END()
}
