// grid_march.cuh -- TMA-staged, z-marching versions of the 7-point-stencil kernels of the PCG / V-cycle fine levels.
// Included by grid_ops.cu inside its anonymous namespace (after grid_rows.cuh, whose arithmetic they reproduce bit for
// bit: same expression order in stencil_quad, same fused multiply-adds).
//
// Why: the row kernels (grid_rows.cuh) fetch the y+-1 and z+-1 rows of every quad through L1/L2 again and sit at
// 0.62-0.66 of the measured HBM peak with ~50 % of the warps stalled on those loads.  Here a persistent CTA (one per SM)
// owns tiles of TY rows x ZC planes x the full x extent and MARCHES along z:
//   * every input plane of the tile (TY + 2 rows: one halo row on each side, zero-filled by the TMA unit outside the
//     grid) is brought into a shared-memory ring by cp.async.bulk.tensor (3-D tensor map over the padded vector, one
//     box of <= 256 x (TY+2) x 1 floats per 1 KB of row), completion signalled on an mbarrier per ring stage -- the
//     loads of the next STAGES-1 planes are in flight while a plane is processed;
//   * each plane is read from shared memory ONCE: a thread keeps its own quads of the planes k-1, k, k+1 in
//     registers (the z neighbours are its own data), the y neighbours and the two x end values come from the tile;
//   * the stream of planes is flat over the CTA's tiles, so the pipeline never drains between tiles.
// DRAM traffic = algorithmic bytes x (1 + 2/ZC) on the stencil inputs (the halo rows are L2 hits: the neighbouring strip
// is processed at the same time by the next CTA).
#pragma once
#include <cuda.h>  // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through the runtime (no -lcuda)

constexpr int kMT = 512;          // threads per marching CTA
constexpr int kMW = kMT / 32;     // warps
constexpr int kMaxStages = 6;

struct MarchGeom {
    LevelDims L;
    int ty, zc;            // tile: rows, planes
    int nstrips, nchunks;  // ny / ty, ceil(nzl / zc)
    int bx, nbx;           // TMA box width in floats, boxes per row
    int segs;              // 32-quad segments per row (nx / 128)
    int stages;
    unsigned halo_bytes;   // one stencil input plane tile: (ty + 2) * nx * 4
    unsigned pw_bytes;     // one pointwise input plane tile: ty * nx * 4
    unsigned stage_bytes;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

template <int K>
__device__ __forceinline__ void march_reduce_commit(double (&v)[K], RedScratch rs, double* out) {
    __shared__ double s_w[K][kMW + 1];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_w[k][w] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = 0;
#pragma unroll
            for (int i = 0; i < kMW; i++) x += s_w[k][i];  // (the producer warp holds no partial sum)
            rs.partials[(size_t)blockIdx.x * K + k] = x;
        }
        __threadfence();
        unsigned int ticket = atomicAdd(rs.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {  // <= 148 partials per sum: one warp folds them in a fixed order
        __threadfence();
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = 0;
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) x += rs.partials[(size_t)b * K + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (threadIdx.x == 0) out[k] = x;
        }
        if (threadIdx.x == 0) *rs.counter = 0;
    }
}

// ---------------------------------------------------------------- the operations
// An Op defines u (the field the stencil is applied to) from NS stencil inputs, and what is written from u, K'u and NP
// pointwise inputs.  cin = number of in-range neighbours (the diagonal of K') of the x-interior nodes of the row, e0 / e3 =
// this quad holds the first / last node of the row (diagonal cin - 1).  Damping weights omega / diagonal are looked up in
// a small shared-memory table (wt[0..1][diagonal], filled once per CTA with the very divisions the row kernels evaluate
// per row): a division per quad and plane made the sweeps instruction-bound.
struct OpUpdateP {  // u = (z - mean) + beta p ; out0 = u ; out1 = K'u ; red0 = sum u K'u      (k_row_update_p_stencil)
    static constexpr int NS = 2, NP = 0, NOUT = 2, NRED = 1;
    struct Params { const double *sum_z, *rho_new, *rho_old; double n_global; int first; };
    struct Ctx { float mean, beta; };
    __device__ static void table(const Params&, float*) {}
    __device__ static Ctx ctx(const Params& p, const float*) {
        return Ctx{(float)(*p.sum_z / p.n_global), p.first ? 0.f : (float)(*p.rho_new / *p.rho_old)};
    }
    __device__ static float u1(const Ctx& c, float z, float p, int) { return fmaf(c.beta, p, z - c.mean); }
    __device__ static float4 u(const Ctx& c, const float4& z, const float4& p, int, bool, bool) {
        return make_float4(fmaf(c.beta, p.x, z.x - c.mean), fmaf(c.beta, p.y, z.y - c.mean), fmaf(c.beta, p.z, z.z - c.mean),
                           fmaf(c.beta, p.w, z.w - c.mean));
    }
    __device__ static void emit(const Ctx&, const float4& c, const float4& K, const float4&, int, bool, bool, float4& o0,
                                float4& o1, double* acc) {
        o0 = c;
        o1 = K;
        acc[0] += (double)(c.x * K.x + c.y * K.y + c.z * K.z + c.w * K.w);
    }
};

template <bool DOT>
struct OpSmooth {  // u = x ; out0 = x + omega ((b - shift) - K'x) / cnt ; [red0 = sum b out0, red1 = sum out0]   (k_row_smooth)
    static constexpr int NS = 1, NP = 1, NOUT = 1, NRED = DOT ? 2 : 0;
    struct Params { const double* sum_b; double n_global; float omega; };
    struct Ctx { float shift; const float* wt; };
    __device__ static void table(const Params& p, float* wt) {
        for (int c = 1; c < 8; c++) wt[c] = p.omega / (float)c;
    }
    __device__ static Ctx ctx(const Params& p, const float* wt) { return Ctx{p.sum_b ? (float)(*p.sum_b / p.n_global) : 0.f, wt}; }
    __device__ static float u1(const Ctx&, float x, float, int) { return x; }
    __device__ static float4 u(const Ctx&, const float4& x, const float4&, int, bool, bool) { return x; }
    __device__ static void emit(const Ctx& c, const float4& x, const float4& K, const float4& rhs, int cin, bool e0, bool e3,
                                float4& o, float4&, double* acc) {
        const float win = c.wt[cin], wed = c.wt[cin - 1];
        o.x = fmaf(e0 ? wed : win, (rhs.x - c.shift) - K.x, x.x);
        o.y = fmaf(win, (rhs.y - c.shift) - K.y, x.y);
        o.z = fmaf(win, (rhs.z - c.shift) - K.z, x.z);
        o.w = fmaf(e3 ? wed : win, (rhs.w - c.shift) - K.w, x.w);
        if (DOT) {
            acc[0] += (double)(rhs.x * o.x + rhs.y * o.y + rhs.z * o.z + rhs.w * o.w);
            acc[1] += (double)((o.x + o.y) + (o.z + o.w));
        }
    }
};

struct OpResidual {  // u = x ; out0 = (b - shift) - K'x        (k_row_residual)
    static constexpr int NS = 1, NP = 1, NOUT = 1, NRED = 0;
    struct Params { const double* sum_b; double n_global; };
    struct Ctx { float shift; };
    __device__ static void table(const Params&, float*) {}
    __device__ static Ctx ctx(const Params& p, const float*) { return Ctx{p.sum_b ? (float)(*p.sum_b / p.n_global) : 0.f}; }
    __device__ static float u1(const Ctx&, float x, float, int) { return x; }
    __device__ static float4 u(const Ctx&, const float4& x, const float4&, int, bool, bool) { return x; }
    __device__ static void emit(const Ctx& c, const float4&, const float4& K, const float4& rhs, int, bool, bool, float4& o,
                                float4&, double*) {
        o = make_float4((rhs.x - c.shift) - K.x, (rhs.y - c.shift) - K.y, (rhs.z - c.shift) - K.z, (rhs.w - c.shift) - K.w);
    }
};

struct OpSmooth01 {  // u = x1 = omega (b - shift) / cnt ; out0 = x1 + omega2 ((b - shift) - K'x1) / cnt     (k_row_smooth01)
    static constexpr int NS = 1, NP = 0, NOUT = 1, NRED = 0;
    struct Params { const double* sum_b; double n_global; float omega, omega2; };
    struct Ctx { float shift; const float* wt; };
    __device__ static void table(const Params& p, float* wt) {
        for (int c = 1; c < 8; c++) {
            wt[c] = p.omega / (float)c;
            wt[8 + c] = p.omega2 / (float)c;
        }
    }
    __device__ static Ctx ctx(const Params& p, const float* wt) { return Ctx{p.sum_b ? (float)(*p.sum_b / p.n_global) : 0.f, wt}; }
    __device__ static float u1(const Ctx& c, float b, float, int cin) { return c.wt[cin] * (b - c.shift); }
    __device__ static float4 u(const Ctx& c, const float4& b, const float4&, int cin, bool e0, bool e3) {
        const float win = c.wt[cin], wed = c.wt[cin - 1];
        return make_float4((e0 ? wed : win) * (b.x - c.shift), win * (b.y - c.shift), win * (b.z - c.shift),
                           (e3 ? wed : win) * (b.w - c.shift));
    }
    // rhs = the tile's raw value at the centre
    __device__ static void emit(const Ctx& c, const float4& x1, const float4& K, const float4& rhs, int cin, bool e0, bool e3,
                                float4& o, float4&, double*) {
        const float win = c.wt[8 + cin], wed = c.wt[8 + cin - 1];
        o.x = fmaf(e0 ? wed : win, (rhs.x - c.shift) - K.x, x1.x);
        o.y = fmaf(win, (rhs.y - c.shift) - K.y, x1.y);
        o.z = fmaf(win, (rhs.z - c.shift) - K.z, x1.z);
        o.w = fmaf(e3 ? wed : win, (rhs.w - c.shift) - K.w, x1.w);
    }
};

// ---------------------------------------------------------------- the kernel
// Warp-specialised: warp kMW (one elected lane) is the PRODUCER -- it walks the flat stream of (tile, plane) arrivals and
// issues the TMA loads of arrival g into ring stage g % stages as soon as the consumers have released it (empty[s]);
// warps 0..kMW-1 are CONSUMERS -- they wait for full[s], never for each other (no CTA-wide barrier in the loop), so the
// warps drift apart and the shared-memory reads, the arithmetic and the global stores of different warps overlap.
// A consumer touches the stage of arrival a twice: when it arrives (its own quad of the new plane -> registers: the
// z-neighbour of the plane below), and one step later (y neighbours, x end values, pointwise input of the plane that is
// now written); then it releases it.  Q = quads per thread and plane.
// Dynamic shared memory: [stages][stage_bytes] ring, then full[stages], empty[stages].
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class Op, int Q>
__global__ void __launch_bounds__(kMT + 32, 1)
    k_march(const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1,
            const __grid_constant__ CUtensorMap tmP, const MarchGeom G, const typename Op::Params prm,
            float* __restrict__ out0, float* __restrict__ out1, RedScratch rs, double* red_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)G.stages * G.stage_bytes);
    uint64_t* empty = full + kMaxStages;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const LevelDims& L = G.L;
    const int nx = L.nx, nzl = L.nzl();
    const int nq = nx >> 2;
    const size_t pl = (size_t)nx * L.ny;
    const int ntiles = G.nstrips * G.nchunks;

    __shared__ float s_wt[16];
    if (threadIdx.x == 0) {
        Op::table(prm, s_wt);
        for (int s = 0; s < G.stages; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kMW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    double acc[Op::NRED > 0 ? Op::NRED : 1] = {0.0};
    if (warp == kMW) {
        // ================================================================ producer
        if (lane == 0) {
            const unsigned box_h = (unsigned)(G.ty + 2) * G.bx * 4u, box_p = (unsigned)G.ty * G.bx * 4u;
            int g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int strip = tile % G.nstrips, ch = tile / G.nstrips;
                const int j0 = strip * G.ty;
                const int zcur = min(G.zc, nzl - ch * G.zc);
                for (int a = 0; a < zcur + 2; a++, g++) {
                    const int s = g % G.stages;
                    mbar_wait(empty + s, (uint32_t)(((g / G.stages) & 1) ^ 1));
                    unsigned char* base = smem + (size_t)s * G.stage_bytes;
                    const int kp = ch * G.zc + a;  // padded plane index of local plane ch*zc - 1 + a
                    const bool ghost = (a == 0) || (a == zcur + 1);
                    mbar_expect_tx(full + s, Op::NS * G.halo_bytes + ((Op::NP && !ghost) ? G.pw_bytes : 0u));
                    for (int b = 0; b < G.nbx; b++) {
                        tma_load_3d(base + (size_t)b * box_h, &tmS0, full + s, b * G.bx, j0 - 1, kp);
                        if (Op::NS == 2)
                            tma_load_3d(base + G.halo_bytes + (size_t)b * box_h, &tmS1, full + s, b * G.bx, j0 - 1, kp);
                        if (Op::NP && !ghost)
                            tma_load_3d(base + Op::NS * G.halo_bytes + (size_t)b * box_p, &tmP, full + s, b * G.bx, j0, kp);
                    }
                }
            }
        }
    } else {
        // ================================================================ consumers
        const typename Op::Ctx cx = Op::ctx(prm, s_wt);
        // per-slot constants (do not depend on the tile): row within the strip, quad, shared-memory offsets (in floats)
        int s_row[Q], s_q[Q];
        unsigned o_c[Q], o_l[Q], o_r[Q], o_p[Q];
        bool s_act[Q];
#pragma unroll
        for (int t = 0; t < Q; t++) {
            const int sg = warp + t * kMW;
            s_act[t] = sg < G.ty * G.segs;
            const int row = s_act[t] ? sg / G.segs : 0, seg = s_act[t] ? sg % G.segs : 0;
            const int q = seg * 32 + lane;
            s_row[t] = row;
            s_q[t] = q;
            auto hoff = [&](int x) { return (unsigned)((x / G.bx) * (G.ty + 2) * G.bx + (row + 1) * G.bx + (x % G.bx)); };
            o_c[t] = hoff(4 * q);
            o_l[t] = hoff(max(4 * q - 1, 0));
            o_r[t] = hoff(min(4 * q + 4, nx - 1));
            o_p[t] = (unsigned)(((4 * q) / G.bx) * G.ty * G.bx + row * G.bx + ((4 * q) % G.bx));
        }
        int g = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int strip = tile % G.nstrips, ch = tile / G.nstrips;
            const int j0 = strip * G.ty, kz0 = ch * G.zc;
            const int zcur = min(G.zc, nzl - kz0);
            float4 c_prev[Q], c_cur[Q];
#pragma unroll
            for (int t = 0; t < Q; t++) c_prev[t] = c_cur[t] = zero4();
            for (int a = 0; a < zcur + 2; a++, g++) {
                const int s = g % G.stages;
                const float* S0 = reinterpret_cast<const float*>(smem + (size_t)s * G.stage_bytes);
                const float* S1 = S0 + (G.halo_bytes >> 2);
                mbar_wait(full + s, (uint32_t)((g / G.stages) & 1));
                const int kl = kz0 - 1 + a;  // local plane that just arrived
                const int k = L.k0 + kl;     // global
                const bool ghost = (a == 0) || (a == zcur + 1);
                // ---- own quads of the arrived plane
                float4 n_c[Q];
                {
                    const int cz = (int)(k > 0) + (int)(k < L.nz - 1);
#pragma unroll
                    for (int t = 0; t < Q; t++) {
                        const int j = j0 + s_row[t];
                        const int cin = (int)(j > 0) + (int)(j < L.ny - 1) + cz + 2;
                        n_c[t] = Op::u(cx, ld4(S0 + o_c[t]), Op::NS == 2 ? ld4(S1 + o_c[t]) : zero4(), cin, s_q[t] == 0,
                                       s_q[t] == nq - 1);
                    }
                }
                if (ghost) {  // a ghost plane is only a z-neighbour: done with its stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + s);
                }
                // ---- write plane kl - 1 (arrived one step ago, still in its stage): d = two steps ago, f = just arrived
                if (a >= 2) {
                    const int sp = (g - 1) % G.stages;
                    const float* T0 = reinterpret_cast<const float*>(smem + (size_t)sp * G.stage_bytes);
                    const float* T1 = T0 + (G.halo_bytes >> 2);
                    const float* TP = T0 + Op::NS * (G.halo_bytes >> 2);
                    const int ko = k - 1;  // global index of the output plane
                    const bool zm = ko > 0, zp = ko < L.nz - 1;
                    const int czo = (int)zm + (int)zp;
                    float4 ya[Q], yb[Q], pw[Q];
                    float xl[Q], xr[Q];
#pragma unroll
                    for (int t = 0; t < Q; t++) {
                        const int j = j0 + s_row[t];
                        const bool e0 = s_q[t] == 0, e3 = s_q[t] == nq - 1;
                        const int cin = (int)(j > 0) + (int)(j < L.ny - 1) + czo + 2;
                        ya[t] = yb[t] = pw[t] = zero4();
                        xl[t] = xr[t] = 0.f;
                        if (j > 0) {
                            const int cin_m = (int)(j - 1 > 0) + 1 + czo + 2;
                            ya[t] = Op::u(cx, ld4(T0 + o_c[t] - G.bx), Op::NS == 2 ? ld4(T1 + o_c[t] - G.bx) : zero4(), cin_m, e0, e3);
                        }
                        if (j < L.ny - 1) {
                            const int cin_p = 1 + (int)(j + 1 < L.ny - 1) + czo + 2;
                            yb[t] = Op::u(cx, ld4(T0 + o_c[t] + G.bx), Op::NS == 2 ? ld4(T1 + o_c[t] + G.bx) : zero4(), cin_p, e0, e3);
                        }
                        // x end values held by another warp / slot: always x-interior nodes of this row
                        if (lane == 0 && s_q[t] > 0) xl[t] = Op::u1(cx, T0[o_l[t]], Op::NS == 2 ? T1[o_l[t]] : 0.f, cin);
                        if (lane == 31 && s_q[t] < nq - 1) xr[t] = Op::u1(cx, T0[o_r[t]], Op::NS == 2 ? T1[o_r[t]] : 0.f, cin);
                        if (Op::NP) pw[t] = ld4(TP + o_p[t]);
                        else if (Op::NS == 1) pw[t] = ld4(T0 + o_c[t]);  // (OpSmooth01: the raw right-hand side at the centre)
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + sp);  // this warp is done with the stage of the output plane
#pragma unroll
                    for (int t = 0; t < Q; t++) {
                        if (!s_act[t]) continue;  // (warp-uniform)
                        const int j = j0 + s_row[t];
                        const int q = s_q[t];
                        const bool e0 = q == 0, e3 = q == nq - 1;
                        const int cin = (int)(j > 0) + (int)(j < L.ny - 1) + czo + 2;
                        const float4 d = zm ? c_prev[t] : zero4();
                        const float4 f = zp ? n_c[t] : zero4();
                        float l, r;
                        x_neighbours(c_cur[t], lane, q, nq, xl[t], xr[t], l, r);
                        const float fc = (float)cin;
                        const float4 K = stencil_quad(c_cur[t], ya[t], yb[t], d, f, l, r, e0 ? fc - 1.f : fc, fc, e3 ? fc - 1.f : fc);
                        float4 o0, o1;
                        Op::emit(cx, c_cur[t], K, pw[t], cin, e0, e3, o0, o1, acc);
                        const size_t e = (size_t)(kl - 1) * pl + (size_t)j * nx + 4u * (unsigned)q;
                        st4(out0 + e, o0);
                        if (Op::NOUT == 2) st4(out1 + e, o1);
                    }
                }
#pragma unroll
                for (int t = 0; t < Q; t++) {
                    c_prev[t] = c_cur[t];
                    c_cur[t] = n_c[t];
                }
            }
        }
    }
    if (Op::NRED > 0) march_reduce_commit<(Op::NRED > 0 ? Op::NRED : 1)>(acc, rs, red_out);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled encode_tiled_fn() {
    static PFN_encodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p)
            throw Error(SHM3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        return reinterpret_cast<PFN_encodeTiled>(p);
    }();
    return fn;
}

// tensor map over a padded vector (nx, ny, nzl + 2 planes) with boxes of bx x rows x 1
inline CUtensorMap make_tmap(const float* padded_base, const LevelDims& L, int bx, int rows) {
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)L.nx, (cuuint64_t)L.ny, (cuuint64_t)(L.nzl() + 2)};
    const cuuint64_t strides[2] = {(cuuint64_t)L.nx * 4, (cuuint64_t)L.nx * L.ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(padded_base), dims, strides, box,
                                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(SHM3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return tm;
}

// Is the marching path usable / worthwhile for this level?  (grids the row kernels handle at launch-latency cost stay there)
inline bool march_ok(const LevelDims& L) {
    if (g_march_disabled) return false;
    return L.nx % 128 == 0 && L.nx <= 1024 && L.ny % 4 == 0 && L.nzl() >= 8 && L.n() >= (size_t)96 * 96 * 96;
}

template <class Op>
inline MarchGeom march_geom(const LevelDims& L) {
    MarchGeom G;
    G.L = L;
    G.bx = L.nx % 256 == 0 ? 256 : 128;
    G.nbx = L.nx / G.bx;
    G.segs = L.nx / 128;
    // rows per tile: as many as keep a ring stage <= 48 KB (y-halo re-reads from L2 shrink with ty), dividing ny
    const int per_row = (Op::NS + Op::NP) * L.nx * 4;
    int ty = 16;
    // ... and at most two quads per thread and plane (register budget of 128 at 512 threads)
    while (ty > 2 && ((Op::NS * (ty + 2) + Op::NP * ty) * L.nx * 4 > 48 * 1024 || L.ny % ty != 0 || ty * (L.nx / 128) > 2 * kMW))
        ty >>= 1;
    (void)per_row;
    G.ty = ty;
    G.nstrips = L.ny / ty;
    G.halo_bytes = (unsigned)(ty + 2) * L.nx * 4u;
    G.pw_bytes = (unsigned)ty * L.nx * 4u;
    G.stage_bytes = Op::NS * G.halo_bytes + Op::NP * G.pw_bytes;
    G.stages = (int)std::min<size_t>(kMaxStages, (200 * 1024) / G.stage_bytes);
    // planes per tile: balance the z-halo overhead (2/zc) against the tail of the last wave over the SMs
    const int nzl = L.nzl();
    double best = 1e30;
    int bestzc = nzl;
    for (int zc : {8, 16, 32, 64, 128}) {
        if (zc > nzl && zc != 8) continue;
        const int z = std::min(zc, nzl);
        const int tiles = G.nstrips * ((nzl + z - 1) / z);
        const int waves = (tiles + g_march_sms - 1) / g_march_sms;
        const double cost = (double)waves * (z + 2 + G.stages * 0.5);  // planes streamed by the busiest CTA (+ pipeline fill)
        if (cost < best) {
            best = cost;
            bestzc = z;
        }
    }
    G.zc = bestzc;
    G.nchunks = (nzl + G.zc - 1) / G.zc;
    return G;
}

template <class Op>
inline void march_launch(const LevelDims& L, const float* s0, const float* s1, const float* pw, float* out0, float* out1,
                         const typename Op::Params& prm, RedScratch rs, double* red_out, cudaStream_t s) {
    const MarchGeom G = march_geom<Op>(L);
    const size_t pl = L.plane();
    const CUtensorMap t0 = make_tmap(s0 - pl, L, G.bx, G.ty + 2);
    const CUtensorMap t1 = Op::NS == 2 ? make_tmap(s1 - pl, L, G.bx, G.ty + 2) : t0;
    const CUtensorMap tp = Op::NP ? make_tmap(pw - pl, L, G.bx, G.ty) : t0;
    const size_t shmem = (size_t)G.stages * G.stage_bytes + 2 * kMaxStages * sizeof(uint64_t);
    const int slots = G.ty * G.segs;
    const int Q = (slots + kMW - 1) / kMW;
    const int grid = std::min(g_march_sms, G.nstrips * G.nchunks);
    static bool configured[64][3] = {};  // per device and Q: the opt-in for > 48 KB of dynamic shared memory
    int dev = 0;
    SHM3D_CUDA_CHECK(cudaGetDevice(&dev));
    auto go = [&](auto kern) {
        if (dev >= 64 || !configured[dev][Q]) {
            SHM3D_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
            if (dev < 64) configured[dev][Q] = true;
        }
        kern<<<grid, kMT + 32, shmem, s>>>(t0, t1, tp, G, prm, out0, out1, rs, red_out);
    };
    switch (Q) {
        case 1: go(k_march<Op, 1>); break;
        case 2: go(k_march<Op, 2>); break;
        default: throw Error(SHM3D_ERR_INVALID_ARG, "internal: marching tile too large");
    }
}
