// mc_emulate.cpp -- TEST INFRASTRUCTURE.  Runs the device logic of row N3 (signed-heat-3d_b200/csrc/isosurface_core.h
// and the product's case table, the very files the CUDA kernels are built from) on the host, one "thread" after the
// other, over the kernels' own launch geometry: count -> chunked scan -> vertices -> triangles, as in isosurface.cu.  Lets
// the CPU test-suite compare the kernels' arithmetic and index logic with the reference's marching-cubes library
// bit for bit, including the thread -> column mapping and the chunking of the scan; what remains for the GPU tests is
// the execution itself (barriers, memory staging, launches).
// Built by tests/test_row_n3_isosurface.py with g++ -ffp-contract=off.  Never loaded by the product.
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
    unsigned gx, gy;
    launch_grid(L, gx, gy);
    // every (CTA, thread) of the kernels' launch geometry, CTAs in a scrambled order: nothing may depend on the order
    // in which CTAs run
    auto for_each_thread = [&](auto&& body) {
        const unsigned nb = gx * gy;
        for (unsigned i = 0; i < nb; i++) {
            const unsigned blk = (unsigned)(((unsigned long long)i * 7919ull + 13ull) % nb);
            const unsigned b = (nb % 7919u == 0) ? i : blk, bx = b % gx, by = b / gx;
            for (unsigned ty = 0; ty < (unsigned)kRowsY; ty++)
                for (unsigned tx = 0; tx < (unsigned)kLanesZ; tx++) {
                    int y, z;
                    if (thread_column(L, bx, by, tx, ty, y, z)) body(y, z);
                }
        }
    };
    // k_mc_count
    std::vector<unsigned> col_v(nc, 0xdeadbeefu), col_t(nc, 0xdeadbeefu);
    std::vector<unsigned> col_x(nc, 0xdeadbeefu);
    int count_mismatch = 0;
    for_each_thread([&](int y, int z) {
        // the kernel's bit form: one sign nibble per plane (its z+1 half comes from the next lane there), two nibbles = a case
        const float niso = -L.isoval;
        auto nibble = [&](int x) {
            const float* q = field + (long long)x * L.strideX + (long long)y * L.SZ + z;
            return sign_bit(niso, q[0]) | (sign_bit(niso, q[L.SZ]) << 1) | (sign_bit(niso, q[1]) << 2) |
                   (sign_bit(niso, q[L.SZ + 1]) << 3);
        };
        ColumnCount cc{0u, 0u, 0, 0};
        unsigned n_prev = nibble(0);
        for (int x = 0; x < L.SX - 1; x++) {
            const unsigned n_cur = nibble(x + 1);
            count_cell(cc, kTable, case_of_nibbles(n_prev, n_cur), x, y, z);
            n_prev = n_cur;
        }
        // ... must equal the visitor form the emit passes use
        CountVisitor cv{kTable, y, z, 0u, 0u};
        march_column(L, field, y, z, cv);
        if (cv.nv != cc.nv || cv.nt != cc.nt) count_mismatch = 1;
        col_v[column_id(L, y, z)] = cc.nv;
        col_t[column_id(L, y, z)] = cc.nt;
        col_x[column_id(L, y, z)] = pack_range(cc.x_lo, cc.x_hi);
    });
    if (count_mismatch) return 6;
    // ... and over k_mc_count's OWN launch geometry: a warp holds 32 consecutive nodes (z0 .. z0+31; lanes past the last node
    // re-read it) and owns 31 columns; lane l takes the z+1 half of its nibble from lane l+1 (a shuffle on the device; lane
    // 31's result is never used).  Every column must get the counts and the cell range found above, exactly once.
    {
        const int cols = kLanesZ - 1;
        const unsigned cgx = (unsigned)((L.SZ - 1 + cols - 1) / cols);
        std::vector<unsigned> v2(nc, 0xdeadbeefu), t2(nc, 0xdeadbeefu), x2(nc, 0xdeadbeefu);
        const float niso = -L.isoval;
        for (unsigned by = 0; by < gy; by++)
            for (unsigned bx = 0; bx < cgx; bx++)
                for (unsigned ty = 0; ty < (unsigned)kRowsY; ty++) {
                    const int y = (int)(by * kRowsY + ty);
                    if (y >= L.SY - 1) continue;
                    ColumnCount cc[kLanesZ];
                    unsigned n_prev[kLanesZ], mine[kLanesZ];
                    auto plane = [&](int x, unsigned* nib) {
                        for (int l = 0; l < kLanesZ; l++) {
                            const int z = (int)(bx * cols) + l, zc = z < L.SZ - 1 ? z : L.SZ - 1;
                            const float* q = field + (long long)x * L.strideX + (long long)y * L.SZ + zc;
                            mine[l] = sign_bit(niso, q[0]) | (sign_bit(niso, q[L.SZ]) << 1);
                        }
                        for (int l = 0; l < kLanesZ; l++) nib[l] = mine[l] | (mine[l + 1 < kLanesZ ? l + 1 : l] << 2);
                    };
                    plane(0, n_prev);
                    for (int l = 0; l < kLanesZ; l++) cc[l] = ColumnCount{0u, 0u, 0, 0};
                    for (int x = 0; x < L.SX - 1; x++) {
                        unsigned n_cur[kLanesZ];
                        plane(x + 1, n_cur);
                        for (int l = 0; l < kLanesZ; l++) {
                            const int z = (int)(bx * cols) + l;
                            if (!((n_prev[l] | n_cur[l]) == 0u || (n_prev[l] & n_cur[l]) == 15u))
                                count_cell(cc[l], kTable, case_of_nibbles(n_prev[l], n_cur[l]), x, y, z);
                            n_prev[l] = n_cur[l];
                        }
                    }
                    for (int l = 0; l < cols; l++) {
                        const int z = (int)(bx * cols) + l;
                        if (z >= L.SZ - 1) continue;
                        const int c = column_id(L, y, z);
                        if (v2[c] != 0xdeadbeefu) return 7;  // written twice
                        v2[c] = cc[l].nv;
                        t2[c] = cc[l].nt;
                        x2[c] = pack_range(cc[l].x_lo, cc[l].x_hi);
                    }
                }
        for (int c = 0; c < nc; c++)
            if (v2[c] != col_v[c] || t2[c] != col_t[c] || x2[c] != col_x[c]) return 7;
    }
    // k_mc_scan_sums: one partial sum per CTA; k_mc_scan_write: offsets of the CTAs before + scan of the CTA's thread sums
    std::vector<unsigned long long> voff(nc + 1, ~0ull), toff(nc + 1, ~0ull), sv(kScanThreads), st(kScanThreads), a(kScanThreads),
        c(kScanThreads), bsum(2 * kScanBlocks, 0ull);
    for (int t = 0; t < kScanThreads; t++) {
        int b, e;
        scan_chunk(nc, t, b, e);
        scan_chunk_sum(col_v.data(), col_t.data(), b, e, a[t], c[t]);
        bsum[2 * (t / kScanBlock)] += a[t];
        bsum[2 * (t / kScanBlock) + 1] += c[t];
    }
    for (int blk = kScanBlocks - 1; blk >= 0; blk--) {  // (CTAs in any order)
        unsigned long long pa = 0, pc = 0;
        for (int i = 0; i < blk; i++) {
            pa += bsum[2 * i];
            pc += bsum[2 * i + 1];
        }
        unsigned long long ra = 0, rc = 0;
        for (int tt = 0; tt < kScanBlock; tt++) {
            const int t = blk * kScanBlock + tt;
            ra += a[t];
            rc += c[t];
            sv[t] = pa + ra;   // = offset of the CTA + inclusive scan of its threads' sums
            st[t] = pc + rc;
            int b, e;
            scan_chunk(nc, t, b, e);
            scan_chunk_write(col_v.data(), col_t.data(), b, e, sv[t] - a[t], st[t] - c[t], voff.data(), toff.data());
        }
    }
    voff[nc] = sv[kScanThreads - 1];
    toff[nc] = st[kScanThreads - 1];
    for (int i = 0; i < nc; i++)  // the scan the kernels rely on: exclusive, complete, monotone
        if (voff[i + 1] != voff[i] + col_v[i] || toff[i + 1] != toff[i] + col_t[i]) return 5;
    if (voff[0] != 0 || toff[0] != 0) return 5;
    *n_vertices = (int64_t)voff[nc];
    *n_triangles = (int64_t)toff[nc];
    if (!vertices) return 0;
    if (*n_vertices > vertex_capacity || *n_triangles > triangle_capacity) return 2;
    std::vector<uint32_t> vkey(voff[nc] + 1);
    int bad = 0;
    // k_mc_vertices
    for_each_thread([&](int y, int z) {
        const int col = column_id(L, y, z);
        if (voff[col + 1] == voff[col]) return;
        int x_lo, x_hi;
        unpack_range(col_x[col], x_lo, x_hi);
        VertexVisitor vv{&L, y, z, voff[col], vertices, vkey.data()};
        march_column(L, field, y, z, vv, x_lo, x_hi);
        if (vv.v != voff[col + 1]) bad = 3;
    });
    // k_mc_triangles
    for_each_thread([&](int y, int z) {
        const int col = column_id(L, y, z);
        if (toff[col + 1] == toff[col]) return;
        int x_lo, x_hi;
        unpack_range(col_x[col], x_lo, x_hi);
        TriangleVisitor tv{&L, kTable, voff.data(), vkey.data(), y, z, toff[col], triangles};
        march_column(L, field, y, z, tv, x_lo, x_hi);
        if (tv.t != toff[col + 1]) bad = 4;
    });
    if (bad) return bad;
    return 0;
}
