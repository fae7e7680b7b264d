// mg_tail.cuh -- the latency-bound tail of a V-cycle as ONE launch.  Included by grid_ops.cu inside its anonymous
// namespace (after grid_rows.cuh, whose row bodies it reuses).
//
// Below ~64^3 every multigrid operation (and every level of the constraint projector's elimination tree) is a kernel of
// a few microseconds whose cost is the launch itself: at 512^3 the projected levels 128^3 .. 8^3 were ~210 launches and
// 1.3 ms of a 4.7 ms PCG iteration (profiles/r02_launches_sphere512_tma_first3000.csv).  Here such a sequence is a
// PROGRAM -- an array of TailOp in device memory, recorded once per solve by the host -- interpreted by one thread-block
// cluster (up to 16 CTAs on one GPC): every op is spread over all warps of the cluster, ops are separated by the
// hardware cluster barrier (barrier.cluster arrive.release / wait.acquire: ~0.2 us, and ptxas emits the L1 invalidate
// with it, so plain loads see what other CTAs of the cluster wrote in the previous op) instead of a kernel boundary
// (~5 us).  Two kinds of program: the whole V-cycle from the first level with <= 64^3 nodes down to the dense coarsest
// solve and back up (Solver::record_tail), and one application of a small projector (Projector::build).
#pragma once
// (grid_ops.cu includes <cooperative_groups.h> at file scope before this header)

constexpr int kTailThreads = 512;           // 16 warps per CTA, <= 128 registers per thread
constexpr int kTailSmemBytes = 120 * 1024;  // unused; asks for one CTA per SM so the cluster spreads over 16 SMs

__device__ __forceinline__ const float* tail_in(const float* p, const float* v, const float* w) {
    return p == kTailSlotV ? v : (p == kTailSlotW ? w : p);
}

__global__ void __launch_bounds__(kTailThreads, 1)
    k_cluster_program(const TailOp* __restrict__ ops, int n_ops, float* v, const float* w, const double* shift_num,
                      double shift_den) {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned wpb = kTailThreads / 32;
    const unsigned gw = cl.block_rank() * wpb + (threadIdx.x >> 5), nw = cl.num_blocks() * wpb;
    const unsigned gt = cl.block_rank() * kTailThreads + threadIdx.x, nt = cl.num_blocks() * kTailThreads;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < n_ops; i++) {
        const TailOp op = ops[i];
        const float* a = tail_in(op.a, v, w);
        const float* b = tail_in(op.b, v, w);
        float* o = const_cast<float*>(tail_in(op.o, v, w));
        switch (op.code) {
            case kTSmooth0:  // o = omega a / d
                rows_smooth0(op.L, o, a, 0.f, op.omega, gw, nw);
                break;
            case kTSmooth: {  // o = b + omega (a - K'b) / d
                double acc[2];
                rows_smooth<false>(op.L, o, b, a, 0.f, op.omega, acc, gw, nw);
                break;
            }
            case kTResidual:  // o = a - K'b
                rows_residual(op.L, b, a, 0.f, o, gw, nw);
                break;
            case kTRestrict:  // o (level Lc) = 0.5 P^T a (level L)
                rows_restrict(op.L, op.Lc, a, o, gw, nw);
                break;
            case kTProlong:  // o (level L) += P a (level Lc)
                rows_prolong_add(op.L, op.Lc, o, a, gw, nw);
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
                for (unsigned e = gt; e < n4; e += nt) st4(o + 4 * e, ld4(a + 4 * e));
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
                for (int R = (int)gw; R < li.n_fwd; R += (int)nw)
                    proj_fwd_row(A, A.rowmaps + li.fwd_node, A.rowmaps + li.fwd_local, R, lane);
                break;
            }
            case kTBwd: {
                const ProjDev A = *op.proj;
                const ProjLevelInfo li = A.levels[op.h];
                for (int R = (int)gw; R < li.n_bwd; R += (int)nw)
                    proj_bwd_row(A, A.rowmaps + li.bwd_node, A.rowmaps + li.bwd_local, R, lane);
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
        cl.sync();
    }
}

// cluster size this device can co-schedule for the program kernel (16 = one GPC's worth, non-portable; else 8)
inline int tail_cluster_size() {
    static thread_local int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    SHM3D_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_program, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes));
    SHM3D_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_program, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int best = 0;
    for (int cs : {16, 8, 4, 2, 1}) {
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
        if (cudaOccupancyMaxActiveClusters(&n, k_cluster_program, &cfg) == cudaSuccess && n >= 1) {
            best = cs;
            break;
        }
        cudaGetLastError();
    }
    if (!best) throw Error(SHM3D_ERR_CUDA, "thread-block clusters are not available on this device");
    if (dev >= 0 && dev < 64) cached[dev] = best;
    return best;
}
