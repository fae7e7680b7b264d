// k_sum.cu -- Steps 1-2 of the signed heat method on the grid: heat-kernel (Yukawa) summation of the source
// normals at every node, then per-node normalisation.
//
// Replaces the O(N*M) double loop of the reference, src/signed_heat_grid_solver.cpp:48-65 (mesh) and
// :157-174 (points), with yukawaPotential = exp(-lambda r)/r from src/signed_heat_3d.cpp:45-49:
//     X(x) = sum_s n_s A_s exp(-lambda |x - y_s|) / |x - y_s|,   Y = X / |X|.
//
// B200 design (compute/SFU-bound; no tensor cores -- this is not a contraction):
//  * one CTA per 8x8x8 tile of nodes, one warp per 4x4x4 brick, two nodes per lane (same x,y; z and z+2);
//  * sources arrive Morton-clustered (<=32 per cluster, sources.cu).  A tile first filters ALL clusters with a
//    conservative sphere/box test into a shared-memory candidate list (ordered compaction -> deterministic
//    summation order), then each warp filters the candidates again for its own brick and evaluates the kept
//    clusters: the cluster's sources are staged into a per-warp shared-memory buffer by one coalesced float4
//    load per lane and read back as LDS.128 broadcasts in the pair loop;
//  * far-field culling: a cluster is dropped for a brick when lambda*(r_lo - r_min_hi) > tau, where r_lo is a
//    lower bound of the distance from any brick node to any source of the cluster and r_min_hi an upper
//    bound on the distance from any brick node to its nearest source (exact nearest-source distance at the
//    brick centre + brick half-diagonal).  Dropped terms are < e^-tau relative to the leading term of that
//    node; the normalisation in Step 2 makes only relative size matter;
//  * range: exp(-lambda r) underflows fp32 for lambda*r > 87.  Every node keeps a running
//    reference distance m <= every r it has seen (a lower bound from the cluster spheres) and accumulates
//    sum w exp(-lambda (r - m)); when m decreases the accumulator is rescaled.  Arguments of ex2 are <= 0
//    and the nearest source's term is >= e^{-2 lambda rad_cluster} (>= e^-8 by construction);
//  * per pair: 2 MUFU (rsqrt, ex2) + ~10 FP32 instructions -> MUFU-bound at 16/clk/SM.
#include "kernels.cuh"

namespace shm3d {

namespace {

constexpr int kTile = 8;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kListCap = 2048;

__device__ __forceinline__ float fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of work).  The two nodes a lane owns
// (same x,y; z and z+2) travel as the low / high half of one 64-bit register; ptxas folds pack2(v, v) into the
// instruction's scalar-broadcast operand form, so per-source constants need no extra moves.  Each half is computed by
// the same IEEE operation (fma.rn / mul.rn / sub.rn) as the scalar form: results are bit-identical to it.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// (Measured and dropped, profiles/experiments/r02_ksum_poly_ex2_mix.jsonl: evaluating a fraction of the exponentials on the
// FMA pipe -- round-to-nearest split by the 1.5*2^23 trick, degree-5 polynomial in FFMA2 with immediate coefficients, 2^n by
// an integer add into the exponent -- to relieve the XU pipe.  0 / 2 / 3 / 4 / 5 of every 8 sources: 509 / 520 / 539 / 565 /
// 599 ms.  The extra 12 issue slots per offloaded pair cost more than the 16 XU cycles they free.)

// one source against the lane's two nodes (packed: low half z0, high half z1)
__device__ __forceinline__ void pair_step(const float4& p, const float4& n, float px, float py, f32x2 pz, f32x2 nl, f32x2 cm,
                                          f32x2& Xx, f32x2& Xy, f32x2& Xz) {
    const float dx = px - p.x, dy = py - p.y;
    const float dxy2 = fmaf(dy, dy, dx * dx);
    const f32x2 dz = sub2(pz, pack2(p.z, p.z));
    const f32x2 r2 = fma2(dz, dz, pack2(dxy2, dxy2));
    float r20, r21;
    unpack2(r2, r20, r21);
    const f32x2 ri = pack2(fast_rsqrt(r20), fast_rsqrt(r21));
    const f32x2 arg = fma2(nl, mul2(r2, ri), cm);
    float a0, a1;
    unpack2(arg, a0, a1);
    const f32x2 wgt = mul2(pack2(fast_ex2(a0), fast_ex2(a1)), ri);
    Xx = fma2(wgt, pack2(n.x, n.x), Xx);
    Xy = fma2(wgt, pack2(n.y, n.y), Xy);
    Xz = fma2(wgt, pack2(n.z, n.z), Xz);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float dist3(float ax, float ay, float az, const float4& b) {
    float dx = ax - b.x, dy = ay - b.y, dz = az - b.z;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

// running-reference update for one node when a new cluster (sphere b) is about to be accumulated
__device__ __forceinline__ void update_ref(float px, float py, float pz, const float4& b, float lam2, float& m,
                                           float& cm, float& X0, float& X1, float& X2) {
    float lb = fmaxf(0.f, dist3(px, py, pz, b) - b.w);
    if (lb < m) {
        float f = fast_ex2(lam2 * (lb - m));  // m = 3e38 initially -> f = 0, X = 0
        X0 *= f;
        X1 *= f;
        X2 *= f;
        m = lb;
        cm = lam2 * lb;
    }
}

// Step 2 for one node: Y = X / |X|.  The kernel holds X~ = wscale * 2^cm * X (cm = lam2 * running reference distance).
// Default: scale by the largest component, normalise in fp32 -- finite wherever X~ is.
// SHM3D_FLAG_FP64_UNDERFLOW: where the TRUE X is small enough for the reference's `X /= X.norm()`
// (src/signed_heat_grid_solver.cpp:61, plain sqrt(x*x + y*y + z*z) in double) to lose precision -- squares below the
// normal range, i.e. max|X| < 2^-511 -- or to underflow to zero altogether (max|X| < 2^-537.5: Y non-finite), that very
// expression is evaluated in IEEE double (subnormals included, no contraction) on the true values, so Y equals the
// reference's there: non-finite in the core, of wrong length in the gradual-underflow shell around it.
__device__ __forceinline__ void normalise_node(const SumParams& P, float x, float y, float z, float cm, float& a, float& b,
                                               float& c) {
    const float s = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
    if (P.uf_enable && log2f(s) - cm < P.uf_thr + 28.f) {
        const double sc = exp2((double)P.uf_log2_unscale - (double)cm);  // X = X~ * sc; >= 2^-1022 wherever X~ is normal
        const double xd = __dmul_rn((double)x, sc), yd = __dmul_rn((double)y, sc), zd = __dmul_rn((double)z, sc);
        const double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)), __dmul_rn(zd, zd)));
        a = (float)(xd / n);
        b = (float)(yd / n);
        c = (float)(zd / n);
        return;
    }
    const float u = x / s, v = y / s, w = z / s;
    const float nrm = sqrtf(u * u + v * v + w * w);
    a = u / nrm;
    b = v / nrm;
    c = w / nrm;
}


__global__ void __launch_bounds__(kThreads, 3)
k_sum(SumParams P, const float4* __restrict__ cl_bounds, const int2* __restrict__ cl_range,
      const float4* __restrict__ src_pos, const float4* __restrict__ src_wn, float* __restrict__ Y, size_t ystride,
      unsigned long long* __restrict__ pair_counter) {
    __shared__ int s_list[kListCap];
    __shared__ float4 s_src[kWarps][2][32];
    __shared__ float s_redf[kWarps];
    __shared__ int s_wcount[kWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ntx = (P.nx + kTile - 1) / kTile, nty = (P.ny + kTile - 1) / kTile;
    int t = blockIdx.x;
    const int tx = t % ntx;
    t /= ntx;
    const int ty = t % nty;
    const int tz = t / nty;
    const int nzl = P.k1 - P.k0;

    // tile centre / half diagonal (always the full 8^3 box: conservative for partial tiles)
    const float hdT = P.cell * 3.5f * 1.7320508f * 1.0001f;
    const float cTx = P.ox + P.cell * (tx * kTile + 3.5f);
    const float cTy = P.oy + P.cell * (ty * kTile + 3.5f);
    const float cTz = P.oz + P.cell * (P.k0 + tz * kTile + 3.5f);

    // ---- A1: upper bound of the nearest-source distance over the tile
    float best = 3e38f;
    for (int c = tid; c < P.n_clusters; c += kThreads) {
        float4 b = cl_bounds[c];
        best = fminf(best, dist3(cTx, cTy, cTz, b) + b.w);
    }
    best = warp_min(best);
    if (lane == 0) s_redf[w] = best;
    __syncthreads();
    float UT = s_redf[0];
#pragma unroll
    for (int i = 1; i < kWarps; i++) UT = fminf(UT, s_redf[i]);
    UT += hdT;
    __syncthreads();

    // ---- brick / node geometry
    const int bx0 = tx * kTile + 4 * (w & 1), by0 = ty * kTile + 4 * ((w >> 1) & 1), bz0 = tz * kTile + 4 * (w >> 2);
    const float hdB = P.cell * 1.5f * 1.7320508f * 1.0001f;
    const float cBx = P.ox + P.cell * (bx0 + 1.5f), cBy = P.oy + P.cell * (by0 + 1.5f),
                cBz = P.oz + P.cell * (P.k0 + bz0 + 1.5f);
    const int ix = bx0 + (lane & 3), iy = by0 + ((lane >> 2) & 3), iz0 = bz0 + (lane >> 4), iz1 = iz0 + 2;
    const float px = P.ox + P.cell * ix, py = P.oy + P.cell * iy;
    const float pz0 = P.oz + P.cell * (P.k0 + iz0), pz1 = P.oz + P.cell * (P.k0 + iz1);

    float m0 = 3e38f, m1 = 3e38f, cm0 = 0.f, cm1 = 0.f;
    float X00 = 0.f, X01 = 0.f, X02 = 0.f, X10 = 0.f, X11 = 0.f, X12 = 0.f;
    float best2 = 3e38f;  // squared distance brick centre -> nearest source seen so far
    unsigned long long npairs = 0;
    const float lam2 = P.lam2, nlam2 = -P.lam2, tol = P.tol;

    for (int cbase = 0; cbase < P.n_clusters;) {
        // ---- A2: ordered compaction of the clusters that can matter anywhere in the tile
        int nl = 0;
        while (cbase < P.n_clusters && nl + kThreads <= kListCap) {
            int c = cbase + tid;
            bool keep = false;
            if (c < P.n_clusters) {
                float4 b = cl_bounds[c];
                float lb = fmaxf(0.f, dist3(cTx, cTy, cTz, b) - b.w - hdT);
                keep = (lb - UT) <= tol;
            }
            unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wcount[w] = __popc(mask);
            __syncthreads();
            int off = nl, tot = 0;
#pragma unroll
            for (int i = 0; i < kWarps; i++) {
                int cnt = s_wcount[i];
                if (i < w) off += cnt;
                tot += cnt;
            }
            if (keep) s_list[off + __popc(mask & ((1u << lane) - 1u))] = c;
            nl += tot;
            cbase += kThreads;
            __syncthreads();
        }
        const int nlist = nl;

        // ---- B0: exact nearest-source distance at the brick centre (culling only; never enters the sum)
        {
            float ub = 3e38f;
            for (int i = lane; i < nlist; i += 32) {
                float4 b = cl_bounds[s_list[i]];
                ub = fminf(ub, dist3(cBx, cBy, cBz, b) + b.w);
            }
            ub = warp_min(ub);
            best2 = fminf(best2, ub * ub);
            for (int base = 0; base < nlist; base += 32) {
                int i = base + lane;
                bool need = false;
                if (i < nlist) {
                    float4 b = cl_bounds[s_list[i]];
                    float lb = dist3(cBx, cBy, cBz, b) - b.w;
                    need = lb * fabsf(lb) < best2;
                }
                unsigned mask = __ballot_sync(0xffffffffu, need);
                while (mask) {
                    int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    int2 rg = cl_range[s_list[base + j]];
                    float d2 = 3e38f;
                    if (lane < rg.y) {
                        float4 p = src_pos[rg.x + lane];
                        float dx = p.x - cBx, dy = p.y - cBy, dz = p.z - cBz;
                        d2 = dx * dx + dy * dy + dz * dz;
                    }
                    best2 = fminf(best2, warp_min(d2));
                }
            }
        }
        const float UB = sqrtf(best2) * 1.00001f + hdB;

        // ---- B1: evaluate the clusters kept for this brick
        for (int base = 0; base < nlist; base += 32) {
            int i = base + lane;
            bool keep = false;
            int c = 0;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nlist) {
                c = s_list[i];
                b = cl_bounds[c];
                float lb = fmaxf(0.f, dist3(cBx, cBy, cBz, b) - b.w - hdB);
                keep = (lb - UB) <= tol;
            }
            unsigned mask = __ballot_sync(0xffffffffu, keep);
            while (mask) {
                int j = __ffs(mask) - 1;
                mask &= mask - 1;
                int cj = __shfl_sync(0xffffffffu, c, j);
                float4 bj;
                bj.x = __shfl_sync(0xffffffffu, b.x, j);
                bj.y = __shfl_sync(0xffffffffu, b.y, j);
                bj.z = __shfl_sync(0xffffffffu, b.z, j);
                bj.w = __shfl_sync(0xffffffffu, b.w, j);
                int2 rg = cl_range[cj];
                __syncwarp();
                if (lane < rg.y) {
                    s_src[w][0][lane] = src_pos[rg.x + lane];
                    s_src[w][1][lane] = src_wn[rg.x + lane];
                }
                __syncwarp();
                update_ref(px, py, pz0, bj, lam2, m0, cm0, X00, X01, X02);
                update_ref(px, py, pz1, bj, lam2, m1, cm1, X10, X11, X12);
                // the pair loop, both nodes of the lane at once in packed fp32 (low half: z0, high half: z1): per source
                // 2 LDS.128 + 4 scalar (dx, dy, dx^2 + dy^2) + 9 packed + 4 MUFU = 19 issue slots for two pairs (the
                // scalar form took 31: instruction issue was the co-limiter next to the MUFU pipe)
                {
                    const f32x2 pz = pack2(pz0, pz1), cm = pack2(cm0, cm1), nl = pack2(nlam2, nlam2);
                    f32x2 Xx = pack2(X00, X10), Xy = pack2(X01, X11), Xz = pack2(X02, X12);
                    int s = 0;
                    for (; s + 8 <= rg.y; s += 8) {  // (blocks of 8 + remainder: 509 ms against 517 with `#pragma unroll 4`)
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            pair_step(s_src[w][0][s + u], s_src[w][1][s + u], px, py, pz, nl, cm, Xx, Xy, Xz);
                    }
#pragma unroll 4
                    for (; s < rg.y; s++) pair_step(s_src[w][0][s], s_src[w][1][s], px, py, pz, nl, cm, Xx, Xy, Xz);
                    unpack2(Xx, X00, X10);
                    unpack2(Xy, X01, X11);
                    unpack2(Xz, X02, X12);
                }
                npairs += 2ull * (unsigned)rg.y;
            }
        }
        __syncthreads();  // s_list is rebuilt by the next chunk
    }

    // ---- Step 2: normalise and store (component-major Y)
    const size_t nloc = ystride;
    if (ix < P.nx && iy < P.ny) {
        if (iz0 < nzl) {
            float a, b, c;
            normalise_node(P, X00, X01, X02, cm0, a, b, c);
            size_t idx = (size_t)ix + (size_t)iy * P.nx + (size_t)iz0 * P.nx * P.ny;
            Y[idx] = a;
            Y[idx + nloc] = b;
            Y[idx + 2 * nloc] = c;
        }
        if (iz1 < nzl) {
            float a, b, c;
            normalise_node(P, X10, X11, X12, cm1, a, b, c);
            size_t idx = (size_t)ix + (size_t)iy * P.nx + (size_t)iz1 * P.nx * P.ny;
            Y[idx] = a;
            Y[idx + nloc] = b;
            Y[idx + 2 * nloc] = c;
        }
    }
    if (pair_counter) {
        // one atomic per warp: every lane evaluated the same number of (node,source) pairs
        if (lane == 0) atomicAdd(pair_counter, npairs * 32ull);
    }
}

// ------------------------------------------------------------------------------------------------
// Steps 1-2 at ARBITRARY query points (the tet solver evaluates the same sum at tet barycentres,
// reference src/signed_heat_tet_solver.cpp:54-72, :131-147 -- SURVEY.md section 8f row N4).  One thread per query
// point, sources tiled through shared memory, two passes: (1) exact nearest-source distance r0, (2) the sum with
// exp(-lambda (r - r0)) so the exponent is <= 0 whatever lambda*r is.  No culling: every source at every point.
constexpr int kPtsThreads = 256;
__global__ void __launch_bounds__(kPtsThreads) k_sum_points(int n_src, const float4* __restrict__ src_pos,
                                                            const float4* __restrict__ src_wn, float lam2, long long n_q,
                                                            const float4* __restrict__ qpts, float* __restrict__ Y) {
    __shared__ float4 s_p[kPtsThreads], s_n[kPtsThreads];
    const long long qi = (long long)blockIdx.x * kPtsThreads + threadIdx.x;
    const float4 qp = qi < n_q ? qpts[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
    float best = 3e38f;
    for (int base = 0; base < n_src; base += kPtsThreads) {
        const int cnt = min(kPtsThreads, n_src - base);
        __syncthreads();
        if ((int)threadIdx.x < cnt) s_p[threadIdx.x] = src_pos[base + threadIdx.x];
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < cnt; t++) {
            const float dx = qp.x - s_p[t].x, dy = qp.y - s_p[t].y, dz = qp.z - s_p[t].z;
            best = fminf(best, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        }
    }
    const float cm = lam2 * sqrtf(best);
    float X0 = 0.f, X1 = 0.f, X2 = 0.f;
    for (int base = 0; base < n_src; base += kPtsThreads) {
        const int cnt = min(kPtsThreads, n_src - base);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            s_p[threadIdx.x] = src_pos[base + threadIdx.x];
            s_n[threadIdx.x] = src_wn[base + threadIdx.x];
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < cnt; t++) {
            const float4 p = s_p[t], n = s_n[t];
            const float dx = qp.x - p.x, dy = qp.y - p.y, dz = qp.z - p.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float ri = fast_rsqrt(r2);
            const float w = fast_ex2(fmaf(-lam2, r2 * ri, cm)) * ri;
            X0 = fmaf(w, n.x, X0);
            X1 = fmaf(w, n.y, X1);
            X2 = fmaf(w, n.z, X2);
        }
    }
    if (qi < n_q) {
        const float s = fmaxf(fmaxf(fabsf(X0), fabsf(X1)), fabsf(X2));
        const float a = X0 / s, b = X1 / s, c = X2 / s;
        const float nrm = sqrtf(a * a + b * b + c * c);
        Y[3 * qi] = a / nrm;
        Y[3 * qi + 1] = b / nrm;
        Y[3 * qi + 2] = c / nrm;
    }
}

}  // namespace

void launch_heat_sum_points(int n_src, const float4* src_pos, const float4* src_wn, float lam2, long long n_q,
                            const float4* qpts, float* Y, cudaStream_t stream) {
    if (n_q <= 0) return;
    k_sum_points<<<(unsigned)((n_q + kPtsThreads - 1) / kPtsThreads), kPtsThreads, 0, stream>>>(n_src, src_pos, src_wn, lam2,
                                                                                                  n_q, qpts, Y);
    SHM3D_LAUNCHED();
    SHM3D_CUDA_CHECK(cudaGetLastError());
}

void launch_heat_sum(const SumParams& P, const float4* cl_bounds, const int2* cl_range, const float4* src_pos,
                     const float4* src_wn, float* Y, size_t ystride, unsigned long long* pair_counter,
                     cudaStream_t stream) {
    int ntx = (P.nx + kTile - 1) / kTile, nty = (P.ny + kTile - 1) / kTile, ntz = (P.k1 - P.k0 + kTile - 1) / kTile;
    size_t nblocks = (size_t)ntx * nty * ntz;
    k_sum<<<(unsigned)nblocks, kThreads, 0, stream>>>(P, cl_bounds, cl_range, src_pos, src_wn, Y, ystride, pair_counter);
    SHM3D_LAUNCHED();
    SHM3D_CUDA_CHECK(cudaGetLastError());
}

}  // namespace shm3d
