// isosurface_core.h -- per-column marching-cubes logic shared by the three kernels of isosurface.cu (row N3).
//
// Replaces, on the device, what the reference's consumer does on the host with the solver's output: polyscope narrows
// phi to float32 (deps/polyscope/include/polyscope/volume_grid.ipp:103-106) and registerIsosurfaceAsMesh
// (deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228) runs MC::marching_cube
// (deps/polyscope/deps/MarchingCubeCpp/include/MarchingCube/MC.h:242-315) over it, then swizzles / scales / translates the
// vertices.  The goal is the SAME indexed mesh: same float32 vertex coordinates, same vertex numbering, same triangle
// order -- so everything below is phrased in that library's lattice:
//
//   lattice X = grid k (slowest in memory), Y = grid j, Z = grid i (fastest)       [MC.h:74, volume_grid_scalar_quantity.cpp:222]
//   cell corner b (0..7) sits at (X + (b&1), Y + (b>>1&1), Z + (b>>2&1));  case = sum of (value_b < 0) << b
//   cell edge e: 0-3 along X at (Y,Z)+{00,10,01,11}; 4-7 along Y at (X,Z)+{00,10,01,11}; 8-11 along Z at (X,Y)+{00,10,01,11}
//   cells are visited Z outermost, Y, X innermost; a vertex is created by the FIRST visited cell that sees its edge
//   in the slot order 0..11, which is the cell for which the edge is a "far" edge (3, 7, 11) except on the low faces.
//
// The sequential library keeps two slabs of edge->vertex indices.  Here a *column* (fixed Y,Z; marching along X) is the
// unit of work: consecutive columns c = Z*(SY-1)+Y and, inside a column, increasing X reproduce the visiting order, so
//   vertex id   = (vertices created by earlier columns) + (created earlier in this column)
//   triangle id = likewise
// come out of one exclusive scan over per-column counts; an edge's vertex id is found by locating its creating cell
// (creator_of) and searching that column's short, sorted list of (X*16 + slot) keys.
//
// No CUDA-only constructs here apart from the rounding intrinsics: tests/csrc/mc_emulate.cpp compiles this same header
// with g++ and runs the columns in a loop to check the logic against the reference's library without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace shm3d {
namespace mc {

// IEEE single operations that must not be contracted or approximated (vertex coordinates are compared bit for bit)
#if defined(__CUDA_ARCH__)
MC_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
MC_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
MC_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
MC_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
MC_HD int popcount(unsigned m) { return __popc(m); }
MC_HD int lowest_bit(unsigned m) { return __ffs((int)m) - 1; }
#else  // host builds of this header use -ffp-contract=off
MC_HD float fadd(float a, float b) { return a + b; }
MC_HD float fsub(float a, float b) { return a - b; }
MC_HD float fmul(float a, float b) { return a * b; }
MC_HD float fdiv(float a, float b) { return a / b; }
MC_HD int popcount(unsigned m) { return __builtin_popcount(m); }
MC_HD int lowest_bit(unsigned m) { return __builtin_ctz(m); }
#endif

struct Lattice {
    int SX, SY, SZ;            // nodes along lattice X, Y, Z ( = grid nz, ny, nx )
    long long strideX;         // memory stride of one X step ( = nx*ny ); Y stride = SZ, Z stride = 1
    float isoval;
    int world;                 // 1: emit world coordinates, 0: lattice coordinates (X,Y,Z)
    float scale[3], bmin[3];   // per WORLD axis x,y,z:  spacing and bound_min   (volume_grid.ipp:72-76)
    MC_HD int ncols() const { return (SY - 1) * (SZ - 1); }
};

MC_HD int column_id(const Lattice& L, int y, int z) { return z * (L.SY - 1) + y; }

// grid (nx,ny,nz; index i + j*nx + k*nx*ny) -> lattice.  bound_min / bound_max: the glm::vec3 bounds the volume grid was
// registered with (src/signed_heat_grid_solver.cpp:20-24,35); NULL for lattice coordinates.  All float, like glm.
inline Lattice make_lattice(int nx, int ny, int nz, float isoval, const float* bound_min, const float* bound_max) {
    Lattice L;
    L.SX = nz;
    L.SY = ny;
    L.SZ = nx;
    L.strideX = (long long)nx * ny;
    L.isoval = isoval;
    L.world = (bound_min && bound_max) ? 1 : 0;
    const int n[3] = {nx, ny, nz};
    for (int a = 0; a < 3; a++) {
        L.bmin[a] = L.world ? bound_min[a] : 0.f;
        float width = L.world ? bound_max[a] - bound_min[a] : (float)(n[a] - 1);
        L.scale[a] = width / (float)(unsigned)(n[a] - 1);  // gridSpacing(): width / vec3(gridCellDim)
    }
    return L;
}

// edges whose two corners differ in sign, as a 12-bit mask
MC_HD unsigned crossing_mask(unsigned cfg) {
    unsigned ax = cfg ^ (cfg >> 1);  // corner pairs (0,1) (2,3) (4,5) (6,7) at bits 0,2,4,6
    unsigned ay = cfg ^ (cfg >> 2);  // (0,2) (1,3) (4,6) (5,7) at bits 0,1,4,5
    unsigned az = cfg ^ (cfg >> 4);  // (0,4) (1,5) (2,6) (3,7) at bits 0..3
    unsigned ex = (ax & 1u) | ((ax >> 1) & 2u) | ((ax >> 2) & 4u) | ((ax >> 3) & 8u);
    unsigned ey = (ay & 3u) | ((ay >> 2) & 12u);
    unsigned ez = az & 15u;
    return ex | (ey << 4) | (ez << 8);
}

// edges whose vertex THIS cell creates if they cross (MC.h:279-304)
MC_HD unsigned creator_mask(int x, int y, int z) {
    unsigned m = 0x888u;  // 3, 7, 11: always
    if (z == 0) m |= (1u << 1) | (1u << 5);
    if (y == 0) m |= (1u << 2) | (1u << 9);
    if (x == 0) m |= (1u << 6) | (1u << 10);
    if (y == 0 && z == 0) m |= 1u << 0;
    if (x == 0 && z == 0) m |= 1u << 4;
    if (x == 0 && y == 0) m |= 1u << 8;
    return m;
}

// the cell that creates the vertex of edge e of cell (x,y,z), and the slot under which it does
MC_HD void creator_of(int e, int x, int y, int z, int& rx, int& ry, int& rz, int& slot) {
    int a = e & 1, b = (e >> 1) & 1;
    rx = x;
    ry = y;
    rz = z;
    if (e < 4) {  // along X at (y+a, z+b)
        int ey = y + a, ez = z + b;
        ry = ey > 0 ? ey - 1 : 0;
        rz = ez > 0 ? ez - 1 : 0;
        slot = (ey > 0) + 2 * (ez > 0);
    } else if (e < 8) {  // along Y at (x+a, z+b)
        int ex = x + a, ez = z + b;
        rx = ex > 0 ? ex - 1 : 0;
        rz = ez > 0 ? ez - 1 : 0;
        slot = 4 + (ex > 0) + 2 * (ez > 0);
    } else {  // along Z at (x+a, y+b)
        int ex = x + a, ey = y + b;
        rx = ex > 0 ? ex - 1 : 0;
        ry = ey > 0 ? ey - 1 : 0;
        slot = 8 + (ex > 0) + 2 * (ey > 0);
    }
}

// vertex of edge e of cell (x,y,z): lower corner + va/(va-vb) along the edge's axis (MC.h:183-193), then the
// consumer's swizzle * scale + bound_min (volume_grid_scalar_quantity.cpp:220-224)
MC_HD void edge_vertex(const Lattice& L, int e, int x, int y, int z, const float vs[8], float out[3]) {
    int a = e & 1, b = (e >> 1) & 1;
    int axis = e >> 2;
    int lo, hi;   // corner numbers of the edge's ends
    float X = (float)x, Y = (float)y, Z = (float)z;
    if (axis == 0) {
        lo = 2 * a + 4 * b;
        hi = lo + 1;
        Y = (float)(y + a);
        Z = (float)(z + b);
    } else if (axis == 1) {
        lo = a + 4 * b;
        hi = lo + 2;
        X = (float)(x + a);
        Z = (float)(z + b);
    } else {
        lo = a + 2 * b;
        hi = lo + 4;
        X = (float)(x + a);
        Y = (float)(y + b);
    }
    float va = vs[lo], vb = vs[hi];
    float t = fdiv(va, fsub(va, vb));
    if (axis == 0) X = fadd(X, t);
    else if (axis == 1) Y = fadd(Y, t);
    else Z = fadd(Z, t);
    if (L.world) {
        out[0] = fadd(fmul(Z, L.scale[0]), L.bmin[0]);
        out[1] = fadd(fmul(Y, L.scale[1]), L.bmin[1]);
        out[2] = fadd(fmul(X, L.scale[2]), L.bmin[2]);
    } else {
        out[0] = X;
        out[1] = Y;
        out[2] = Z;
    }
}

// March column (y,z) along X.  `visit(x, cfg, vs)` is called for every cell whose case is neither 0 nor 255, in
// increasing x.  field index of lattice node (X,Y,Z) = X*strideX + Y*SZ + Z.  Per step the four values of the next X
// plane are loaded; a warp whose lanes hold consecutive z reads four (nearly) contiguous 128-byte rows.
// [x_begin, x_end): the cells of the column to visit -- the whole column, or the range outside which the count pass
// found only trivial cells (cases 0 / 255 create and emit nothing, so skipping them changes no output).
template <typename Visit>
MC_HD void march_column(const Lattice& L, const float* __restrict__ field, int y, int z, Visit& visit, int x_begin = 0,
                        int x_end = -1) {
    if (x_end < 0) x_end = L.SX - 1;
    const float* p = field + (long long)x_begin * L.strideX + (long long)y * L.SZ + z;
    const float niso = -L.isoval;
    float a0 = fadd(niso, p[0]), a1 = fadd(niso, p[L.SZ]), a2 = fadd(niso, p[1]), a3 = fadd(niso, p[L.SZ + 1]);
    for (int x = x_begin; x < x_end; x++) {
        p += L.strideX;
        float b0 = fadd(niso, p[0]), b1 = fadd(niso, p[L.SZ]), b2 = fadd(niso, p[1]), b3 = fadd(niso, p[L.SZ + 1]);
        unsigned cfg = (unsigned)(a0 < 0.f) | ((unsigned)(b0 < 0.f) << 1) | ((unsigned)(a1 < 0.f) << 2) |
                       ((unsigned)(b1 < 0.f) << 3) | ((unsigned)(a2 < 0.f) << 4) | ((unsigned)(b2 < 0.f) << 5) |
                       ((unsigned)(a3 < 0.f) << 6) | ((unsigned)(b3 < 0.f) << 7);
        if (cfg != 0u && cfg != 255u) {
            float vs[8] = {a0, b0, a1, b1, a2, b2, a3, b3};
            visit(x, cfg, vs);
        }
        a0 = b0;
        a1 = b1;
        a2 = b2;
        a3 = b3;
    }
}

// ---- the count pass in bit form (k_mc_count): one sign nibble per column and lattice plane, two nibbles = one case ----
// nibble of plane X for column (y,z): bit0 = node (y,z), bit1 = (y+1,z), bit2 = (y,z+1), bit3 = (y+1,z+1) below the level
MC_HD unsigned sign_bit(float niso, float v) { return (unsigned)(fadd(niso, v) < 0.f); }
// case number of the cell between plane X (nibble n0) and plane X+1 (nibble n1): corner c of march_column's cfg is
// bit 2*(c/2) of n0 (even c) or of n1 (odd c)
MC_HD unsigned case_of_nibbles(unsigned n0, unsigned n1) {
    unsigned s0 = (n0 & 1u) | ((n0 & 2u) << 1) | ((n0 & 4u) << 2) | ((n0 & 8u) << 3);
    unsigned s1 = (n1 & 1u) | ((n1 & 2u) << 1) | ((n1 & 4u) << 2) | ((n1 & 8u) << 3);
    return s0 | (s1 << 1);
}
// per-column record of the count pass: vertices created, triangles emitted, and the cell range [x_lo, x_hi) outside which
// every cell of the column is trivial (x_lo = x_hi = 0 for a column the surface does not touch)
struct ColumnCount {
    unsigned nv, nt;
    int x_lo, x_hi;
};
MC_HD void count_cell(ColumnCount& cc, const unsigned long long* table, unsigned cfg, int x, int y, int z) {
    if (cfg == 0u || cfg == 255u) return;
    cc.nv += (unsigned)popcount(crossing_mask(cfg) & creator_mask(x, y, z));
    cc.nt += (unsigned)(table[cfg] & 0xFull);
    if (cc.x_hi == 0) cc.x_lo = x;
    cc.x_hi = x + 1;
}
MC_HD unsigned pack_range(int x_lo, int x_hi) { return (unsigned)x_lo | ((unsigned)x_hi << 16); }  // SX <= 65535
MC_HD void unpack_range(unsigned r, int& x_lo, int& x_hi) {
    x_lo = (int)(r & 0xFFFFu);
    x_hi = (int)(r >> 16);
}

// ---- launch geometry and the chunked scan, shared with the host emulation ------------------------------------------
constexpr int kLanesZ = 32;        // lattice Z (grid i, contiguous in memory) across the lanes of a warp
constexpr int kRowsY = 8;          // warps per CTA: consecutive lattice Y (grid j)
constexpr int kScanBlock = 256;    // threads per CTA of the two scan kernels
constexpr int kScanBlocks = 128;   // CTAs: 32768 scan threads, each owning a short contiguous chunk of the column counts
constexpr int kScanThreads = kScanBlock * kScanBlocks;

MC_HD void launch_grid(const Lattice& L, unsigned& gx, unsigned& gy) {
    gx = (unsigned)((L.SZ - 1 + kLanesZ - 1) / kLanesZ);
    gy = (unsigned)((L.SY - 1 + kRowsY - 1) / kRowsY);
}
// the column thread (tx,ty) of CTA (bx,by) works on; false = idle thread
MC_HD bool thread_column(const Lattice& L, unsigned bx, unsigned by, unsigned tx, unsigned ty, int& y, int& z) {
    z = (int)(bx * kLanesZ + tx);
    y = (int)(by * kRowsY + ty);
    return z < L.SZ - 1 && y < L.SY - 1;
}
// scan thread t owns the contiguous chunk [b,e) of the n column counts
MC_HD void scan_chunk(int n, int t, int& b, int& e) {
    const int chunk = (n + kScanThreads - 1) / kScanThreads;
    const long long b0 = (long long)t * chunk;
    b = (int)(b0 < n ? b0 : n);
    e = (int)(b0 + chunk < n ? b0 + chunk : n);
}
MC_HD void scan_chunk_sum(const unsigned* col_v, const unsigned* col_t, int b, int e, unsigned long long& sv,
                          unsigned long long& st) {
    sv = 0;
    st = 0;
    for (int i = b; i < e; i++) {
        sv += col_v[i];
        st += col_t[i];
    }
}
// pv / pt: sum of all chunks before this one
MC_HD void scan_chunk_write(const unsigned* col_v, const unsigned* col_t, int b, int e, unsigned long long pv,
                            unsigned long long pt, unsigned long long* voff, unsigned long long* toff) {
    for (int i = b; i < e; i++) {
        voff[i] = pv;
        toff[i] = pt;
        pv += col_v[i];
        pt += col_t[i];
    }
}

// ---- the three visitors ------------------------------------------------------------------------------------------
struct CountVisitor {  // pass 1: how many vertices this column creates, how many triangles it emits
    const unsigned long long* table;
    int y, z;
    unsigned nv, nt;
    MC_HD void operator()(int x, unsigned cfg, const float*) {
        nv += (unsigned)popcount(crossing_mask(cfg) & creator_mask(x, y, z));
        nt += (unsigned)(table[cfg] & 0xFull);
    }
};

struct VertexVisitor {  // pass 2: positions and search keys of the vertices this column creates
    const Lattice* L;
    int y, z;
    unsigned long long v;  // running vertex id
    float* vertices;       // [3*nV]
    uint32_t* vkey;        // [nV]: x*16 + slot, ascending within a column
    MC_HD void operator()(int x, unsigned cfg, const float* vs) {
        unsigned m = crossing_mask(cfg) & creator_mask(x, y, z);
        while (m) {
            int e = lowest_bit(m);
            m &= m - 1;
            float q[3];
            edge_vertex(*L, e, x, y, z, vs, q);
            vertices[3 * v] = q[0];
            vertices[3 * v + 1] = q[1];
            vertices[3 * v + 2] = q[2];
            vkey[v] = (uint32_t)x * 16u + (uint32_t)e;
            v++;
        }
    }
};

struct TriangleVisitor {  // pass 3: vertex ids of the triangles this column emits
    const Lattice* L;
    const unsigned long long* table;
    const unsigned long long* voff;  // [ncols+1] exclusive scan of per-column vertex counts
    const uint32_t* vkey;
    int y, z;
    unsigned long long t;  // running triangle id
    uint32_t* triangles;   // [3*nT]
    MC_HD uint32_t vertex_id(int e, int x) const {
        int rx, ry, rz, slot;
        creator_of(e, x, y, z, rx, ry, rz, slot);
        int c = column_id(*L, ry, rz);
        unsigned long long lo = voff[c], hi = voff[c + 1];
        uint32_t key = (uint32_t)rx * 16u + (uint32_t)slot;
        while (lo < hi) {  // first entry >= key
            unsigned long long mid = (lo + hi) >> 1;
            if (vkey[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        return (uint32_t)lo;
    }
    MC_HD void operator()(int x, unsigned cfg, const float*) {
        unsigned long long w = table[cfg];
        int n = (int)(w & 0xFull);
        w >>= 4;
        for (int i = 0; i < 3 * n; i++) {
            triangles[3 * t + i] = vertex_id((int)(w & 0xFull), x);
            w >>= 4;
        }
        t += (unsigned long long)n;
    }
};

}  // namespace mc
}  // namespace shm3d
