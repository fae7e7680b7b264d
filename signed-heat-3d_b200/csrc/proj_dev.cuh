// proj_dev.cuh -- device-side view of one Projector and the per-row bodies of its four phases (gather A v, forward
// sweep, backward sweep, scatter D^-1 A^T sol).  Shared by the per-level launches of projector.cu and by the cluster
// program of mg_tail.cuh, which runs a whole application (2*height + 2 phases) inside one launch.
#pragma once
#include "projector.cuh"

namespace shm3d {

struct ProjDev {
    int m, n_touched, n_levels;
    const int64_t* rnode;
    const double* rw;
    const int* rperm;
    const ProjLevelInfo* levels;
    const int* rowmaps;
    const ProjNodeDesc* nodes;
    const double* mat;
    const int* bidx;
    const int64_t* tnode;
    const int* tptr;
    const int* trow;
    const double* tw;
    double *rhs, *y, *sol;
};

// rhs[perm r] = A_r (v - w - shift)
__device__ __forceinline__ void proj_gather_row(const ProjDev& A, int r, const float* v, const float* w, double shift) {
    double acc = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int64_t n = A.rnode[(size_t)r * 8 + c];
        if (n >= 0) {
            double val = (double)v[n] - shift;
            if (w) val -= (double)w[n];
            acc += A.rw[(size_t)r * 8 + c] * val;
        }
    }
    A.rhs[A.rperm[r]] = acc;
}

// forward: row R of a supernode at this height: val = <FWD[R, 0:s], rhs[s0:s0+s]>;
// R < s -> y[s0+R] = val ; else rhs[B[R-s]] -= val.
// G lanes (a power of two <= 32, aligned sub-group of a warp) share one row; `valid` = this sub-group has a row (all 32
// lanes of the warp must make the call: the reduction shuffles name the full warp).
template <int G>
__device__ __forceinline__ void proj_fwd_row(const ProjDev& A, const int* row_node, const int* row_local, int R, int gl,
                                             bool valid = true) {
    double acc = 0;
    ProjNodeDesc nd = {};
    int rl = 0;
    if (valid) {
        nd = A.nodes[row_node[R]];
        rl = row_local[R];
        const double* row = A.mat + nd.fwd + (long long)rl * nd.s;
        const double* x = A.rhs + nd.s0;
        const int cend = rl < nd.s ? rl + 1 : nd.s;  // W is lower triangular
        for (int c = gl; c < cend; c += G) acc += row[c] * x[c];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (valid && gl == 0) {
        if (rl < nd.s)
            A.y[nd.s0 + rl] = acc;
        else
            atomicAdd(&A.rhs[A.bidx[nd.bidx + rl - nd.s]], -acc);
    }
}

// backward: sol[s0+R] = <BWD[R, 0:s], y[s0:]> + <BWD[R, s:s+b], sol[B]>
template <int G>
__device__ __forceinline__ void proj_bwd_row(const ProjDev& A, const int* row_node, const int* row_local, int R, int gl,
                                             bool valid = true) {
    double acc = 0;
    ProjNodeDesc nd = {};
    int rl = 0;
    if (valid) {
        nd = A.nodes[row_node[R]];
        rl = row_local[R];
        const int f = nd.s + nd.b;
        const double* row = A.mat + nd.bwd + (long long)rl * f;
        for (int c = rl + gl; c < nd.s; c += G) acc += row[c] * A.y[nd.s0 + c];  // W^T is upper triangular
        const int* bi = A.bidx + nd.bidx;
        for (int c = gl; c < nd.b; c += G) acc += row[nd.s + c] * A.sol[bi[c]];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (valid && gl == 0) A.sol[nd.s0 + rl] = acc;
}

// v[node t] -= (D^-1 A^T sol)_t
__device__ __forceinline__ void proj_scatter_node(const ProjDev& A, int t, float* v) {
    double acc = 0;
    for (int e = A.tptr[t]; e < A.tptr[t + 1]; e++) acc += A.tw[e] * A.sol[A.trow[e]];
    const int64_t n = A.tnode[t];
    v[n] = (float)((double)v[n] - acc);
}

}  // namespace shm3d
