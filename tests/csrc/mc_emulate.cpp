// mc_emulate.cpp -- TEST INFRASTRUCTURE.  Runs the device logic of row N3 (signed-heat-3d_b200/csrc/isosurface_core.h
// and the product's case table, the very files the CUDA kernels are built from) on the host, one "thread" after the
// other: count per column -> exclusive scan -> vertices -> triangles, exactly the structure of isosurface.cu.  Lets
// the CPU test-suite compare the kernels' arithmetic and index logic with the reference's marching-cubes library
// bit for bit; the launch geometry and the device scan are what only the GPU tests cover.
// Built by tests/test_isosurface.py with g++ -ffp-contract=off.  Never loaded by the product.
#include <cstdint>
#include <vector>

#include "isosurface_core.h"

using namespace shm3d::mc;

static const unsigned long long kTable[256] = {
#include "mc_table.inc"
};

extern "C" int mc_emulate(const float* field, int nx, int ny, int nz, float isoval, const float* bound_min,
                          const float* bound_max, float* vertices, int64_t vertex_capacity, uint32_t* triangles,
                          int64_t triangle_capacity, int64_t* n_vertices, int64_t* n_triangles) {
    Lattice L = make_lattice(nx, ny, nz, isoval, bound_min, bound_max);
    const int nc = L.ncols();
    std::vector<unsigned long long> voff(nc + 1, 0), toff(nc + 1, 0);
    for (int z = 0; z < L.SZ - 1; z++)
        for (int y = 0; y < L.SY - 1; y++) {
            CountVisitor cv{kTable, y, z, 0u, 0u};
            march_column(L, field, y, z, cv);
            voff[column_id(L, y, z) + 1] = cv.nv;
            toff[column_id(L, y, z) + 1] = cv.nt;
        }
    for (int c = 0; c < nc; c++) {
        voff[c + 1] += voff[c];
        toff[c + 1] += toff[c];
    }
    *n_vertices = (int64_t)voff[nc];
    *n_triangles = (int64_t)toff[nc];
    if (!vertices) return 0;
    if (*n_vertices > vertex_capacity || *n_triangles > triangle_capacity) return 2;
    std::vector<uint32_t> vkey(voff[nc] + 1);
    // columns in a scrambled order: nothing may depend on the order in which the "threads" run
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < nc; i++) {
            int c = (int)(((long long)i * 7919 + 13) % nc);
            if (nc % 7919 == 0) c = i;
            int z = c / (L.SY - 1), y = c % (L.SY - 1);
            if (pass == 0) {
                VertexVisitor vv{&L, y, z, voff[c], vertices, vkey.data()};
                march_column(L, field, y, z, vv);
                if (vv.v != voff[c + 1]) return 3;
            } else {
                TriangleVisitor tv{&L, kTable, voff.data(), vkey.data(), y, z, toff[c], triangles};
                march_column(L, field, y, z, tv);
                if (tv.t != toff[c + 1]) return 4;
            }
        }
    return 0;
}
