// mg_tail.cuh -- the latency-bound tail of a V-cycle as ONE launch.  Included by grid_ops.cu inside its anonymous
// namespace (after the element-indexed kernels, whose helpers it reuses).
//
// Below ~32^3 every multigrid operation (and every level of the constraint projector's elimination tree) is a kernel of
// a few microseconds whose cost is the launch boundary itself (~3.4 us per node even inside a CUDA graph): at 512^3 the
// projected levels 32^3 .. 4^3 are ~90 launches per V-cycle.  Here that sequence is a PROGRAM -- an array of TailOp in
// device memory, recorded once per solve by the host (Solver::record_tail) -- interpreted by ONE CTA of 1024 threads:
// ops are separated by __syncthreads() (tens of cycles; the level's vectors and factor blocks stay in this SM's L1 / L2)
// instead of a kernel boundary.
// Measured on B200 (profiles/experiments/r02_pcg_probe_cluster16_vs_graph.jsonl): the first version ran the levels
// <= 64^3 on a 16-CTA thread-block cluster with the hardware cluster barrier between ops; every barrier carries a
// GPU-scope MEMBAR plus an L1 invalidate (SASS: MEMBAR.ALL.GPU, UCGABAR_ARV/WAIT, CCTL.IVALL), ~6 us per op all told --
// slower than graph-replayed launches (4.78 vs 4.07 ms per PCG iteration).  The kernel still accepts a multi-CTA cluster
// (launch_cluster_program's `ctas`), but the solver uses one CTA.
#pragma once
// (grid_ops.cu includes <cooperative_groups.h> at file scope before this header)

constexpr int kTailThreads = 1024;          // 32 warps per CTA, <= 64 registers per thread
constexpr int kTailSmemBytes = 120 * 1024;  // multi-CTA launches only: unused, asks for one CTA per SM

__device__ __forceinline__ const float* tail_in(const float* p, const float* v, const float* w) {
    return p == kTailSlotV ? v : (p == kTailSlotW ? w : p);
}

// The ops are LATENCY-bound (a level's vectors and factor blocks sit in L2; the L1 is invalidated by every barrier), so
// their bodies are written for memory-level parallelism, not for instruction count: one float4 quad per thread with all
// seven neighbour loads independent (the element-indexed stencil<4> of grid_ops.cu, not the warp-per-row form whose rows
// would queue up behind each other in a 16-SM cluster), and sub-warp groups per matrix row in the projector sweeps.
template <int MODE>  // 0: o = x + omega (rhs - K'x) / d ; 1: o = rhs - K'x
__device__ __forceinline__ void tail_stencil(const LevelDims& L, const float* x, const float* rhs, float* o, float omega,
                                             unsigned gt, unsigned nt) {
    const unsigned ng = (unsigned)(L.n() >> 2);
    for (unsigned g = gt; g < ng; g += nt) {
        const unsigned e = g * 4;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        Vec<4> c = Vec<4>::ld(x + e), bv = Vec<4>::ld(rhs + e), Ku, dg, out;
        stencil<4>(x, e, c, i0, j, L.k0 + kl, L, Ku, dg);
#pragma unroll
        for (int t = 0; t < 4; t++)
            out.v[t] = MODE == 0 ? c.v[t] + omega * (bv.v[t] - Ku.v[t]) / dg.v[t] : bv.v[t] - Ku.v[t];
        out.st(o + e);
    }
}

__device__ __forceinline__ void tail_smooth0(const LevelDims& L, const float* rhs, float* o, float omega, unsigned gt,
                                             unsigned nt) {
    const unsigned ng = (unsigned)(L.n() >> 2);
    for (unsigned g = gt; g < ng; g += nt) {
        const unsigned e = g * 4;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        const int k = L.k0 + kl;
        const int cyz = (j > 0) + (j < L.ny - 1) + (k > 0) + (k < L.nz - 1);
        Vec<4> bv = Vec<4>::ld(rhs + e), xv;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = i0 + t;
            xv.v[t] = omega * bv.v[t] / (float)(cyz + (i > 0) + (i < L.nx - 1));
        }
        xv.st(o + e);
    }
}

// bc = 0.5 P^T r: one coarse node per thread (k_mg_restrict's body)
__device__ __forceinline__ void tail_restrict(const LevelDims& Lf, const LevelDims& Lc, const float* r, float* bc, unsigned gt,
                                              unsigned nt) {
    const unsigned nc = (unsigned)Lc.n();
    const ptrdiff_t plf = (ptrdiff_t)Lf.plane();
    for (unsigned e = gt; e < nc; e += nt) {
        int I, J, Kl;
        decode(e, Lc.nx, Lc.ny, I, J, Kl);
        const int K = Lc.k0 + Kl;
        float wx[4], wy[4], wz[4];
        rweights(I, Lc.nx, wx);
        rweights(J, Lc.ny, wy);
        rweights(K, Lc.nz, wz);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (wz[c] == 0.f) continue;
            const int kf = 2 * K - 1 + c - Lf.k0;
#pragma unroll
            for (int bq = 0; bq < 4; bq++) {
                if (wy[bq] == 0.f) continue;
                const int jf = 2 * J - 1 + bq;
                const float* row = r + (ptrdiff_t)kf * plf + (ptrdiff_t)jf * Lf.nx;
                float s = 0.f;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (wx[a] == 0.f) continue;
                    s = fmaf(wx[a], row[2 * I - 1 + a], s);
                }
                acc = fmaf(wy[bq] * wz[c], s, acc);
            }
        }
        bc[e] = 0.5f * acc;
    }
}

// x += P ec: one fine quad per thread (k_mg_prolong_add<4>'s body)
__device__ __forceinline__ void tail_prolong_add(const LevelDims& Lf, const LevelDims& Lc, float* x, const float* ec,
                                                 unsigned gt, unsigned nt) {
    const unsigned ng = (unsigned)(Lf.n() >> 2);
    const ptrdiff_t plc = (ptrdiff_t)Lc.plane();
    for (unsigned g = gt; g < ng; g += nt) {
        const unsigned e = g * 4;
        int i0, j, kl;
        decode(e, Lf.nx, Lf.ny, i0, j, kl);
        const int k = Lf.k0 + kl;
        const int J0 = j >> 1, K0 = k >> 1;
        const int J1 = min(max((j & 1) ? J0 + 1 : J0 - 1, 0), Lc.ny - 1);
        const int K1 = min(max((k & 1) ? K0 + 1 : K0 - 1, 0), Lc.nz - 1);
        const float* r00 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;
        const float* r01 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        const float* r10 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;
        const float* r11 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        Vec<4> xv = Vec<4>::ld(x + e);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = i0 + t;
            const int I0 = i >> 1;
            const int I1 = min(max((i & 1) ? I0 + 1 : I0 - 1, 0), Lc.nx - 1);
            const float c0 = 0.75f * (0.75f * r00[I0] + 0.25f * r01[I0]) + 0.25f * (0.75f * r10[I0] + 0.25f * r11[I0]);
            const float c1 = 0.75f * (0.75f * r00[I1] + 0.25f * r01[I1]) + 0.25f * (0.75f * r10[I1] + 0.25f * r11[I1]);
            xv.v[t] += 0.75f * c0 + 0.25f * c1;
        }
        xv.st(x + e);
    }
}

// projector sweep of one tree height: G lanes per matrix row, every sub-group of the cluster takes rows round-robin
template <int G, bool FWD>
__device__ __forceinline__ void tail_sweep(const ProjDev& A, const int* row_node, const int* row_local, int n_rows, unsigned gt,
                                           unsigned nt) {
    const int ngroups = (int)(nt / G), grp = (int)(gt / G), gl = (int)(gt % G);
    for (int R0 = 0; R0 < n_rows; R0 += ngroups) {  // (uniform trip count: the row reductions shuffle over full warps)
        const int R = R0 + grp;
        if (FWD) proj_fwd_row<G>(A, row_node, row_local, R, gl, R < n_rows);
        else proj_bwd_row<G>(A, row_node, row_local, R, gl, R < n_rows);
    }
}

__global__ void __launch_bounds__(kTailThreads, 1)
    k_cluster_program(const TailOp* __restrict__ ops, int n_ops, float* v, const float* w, const double* shift_num,
                      double shift_den) {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const bool single = gridDim.x == 1;  // one CTA: ops are separated by __syncthreads()
    const unsigned wpb = kTailThreads / 32;
    const unsigned rank = single ? 0u : cl.block_rank(), nblk = single ? 1u : cl.num_blocks();
    const unsigned gw = rank * wpb + (threadIdx.x >> 5), nw = nblk * wpb;
    const unsigned gt = rank * kTailThreads + threadIdx.x, nt = nblk * kTailThreads;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < n_ops; i++) {
        const TailOp op = ops[i];
        const float* a = tail_in(op.a, v, w);
        const float* b = tail_in(op.b, v, w);
        float* o = const_cast<float*>(tail_in(op.o, v, w));
        switch (op.code) {
            case kTSmooth0:  // o = omega a / d
                tail_smooth0(op.L, a, o, op.omega, gt, nt);
                break;
            case kTSmooth:  // o = b + omega (a - K'b) / d
                tail_stencil<0>(op.L, b, a, o, op.omega, gt, nt);
                break;
            case kTResidual:  // o = a - K'b
                tail_stencil<1>(op.L, b, a, o, 0.f, gt, nt);
                break;
            case kTRestrict:  // o (level Lc) = 0.5 P^T a (level L)
                tail_restrict(op.L, op.Lc, a, o, gt, nt);
                break;
            case kTProlong:  // o (level L) += P a (level Lc)
                tail_prolong_add(op.L, op.Lc, o, a, gt, nt);
                break;
            case kTCoarse: {  // o = pinv(a) b, dense n3 x n3, one warp per row
                const int n3 = op.h;
                for (int t = (int)gw; t < n3; t += (int)nw) {
                    float acc = 0.f;
                    for (int c = lane; c < n3; c += 32) acc = fmaf(a[(size_t)t * n3 + c], b[c], acc);
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                    if (lane == 0) o[t] = acc;
                }
                break;
            }
            case kTCopy: {
                const unsigned n4 = (unsigned)(op.L.n() >> 2);
                for (unsigned e = gt; e < n4; e += nt) Vec<4>::ld(a + 4 * e).st(o + 4 * e);
                break;
            }
            case kTGather: {  // rhs = A (a - b - shift); h != 0: shift = *shift_num / shift_den of the launch
                const ProjDev A = *op.proj;
                const double shift = (op.h && shift_num) ? *shift_num / shift_den : 0.0;
                for (int r = (int)gt; r < A.m; r += (int)nt) proj_gather_row(A, r, a, b, shift);
                break;
            }
            case kTFwd: {
                const ProjDev A = *op.proj;
                const ProjLevelInfo li = A.levels[op.h];
                // many short rows (leaves): 8 lanes per row, 4x the rows in flight; few long rows (tree top): a warp per row
                if (li.n_fwd > (int)nw) tail_sweep<8, true>(A, A.rowmaps + li.fwd_node, A.rowmaps + li.fwd_local, li.n_fwd, gt, nt);
                else tail_sweep<32, true>(A, A.rowmaps + li.fwd_node, A.rowmaps + li.fwd_local, li.n_fwd, gt, nt);
                break;
            }
            case kTBwd: {
                const ProjDev A = *op.proj;
                const ProjLevelInfo li = A.levels[op.h];
                if (li.n_bwd > (int)nw) tail_sweep<8, false>(A, A.rowmaps + li.bwd_node, A.rowmaps + li.bwd_local, li.n_bwd, gt, nt);
                else tail_sweep<32, false>(A, A.rowmaps + li.bwd_node, A.rowmaps + li.bwd_local, li.n_bwd, gt, nt);
                break;
            }
            case kTScatter: {  // o -= D^-1 A^T sol
                const ProjDev A = *op.proj;
                for (int t = (int)gt; t < A.n_touched; t += (int)nt) proj_scatter_node(A, t, o);
                break;
            }
            default:
                break;
        }
        if (single) __syncthreads();
        else cl.sync();
    }
}

// largest cluster (<= want) this device can co-schedule for the program kernel (16 needs the non-portable opt-in)
inline int tail_cluster_size(int want) {
    SHM3D_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_program, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes));
    SHM3D_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_program, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {16, 8, 4, 2}) {
        if (cs > want) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(kTailThreads);
        cfg.dynamicSmemBytes = kTailSmemBytes;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = cs;
        at.val.clusterDim.y = at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k_cluster_program, &cfg) == cudaSuccess && n >= 1) return cs;
        cudaGetLastError();
    }
    return 1;
}
