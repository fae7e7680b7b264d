// shm3d_cli -- headless driver of the grid solver (stands in for the reference's GUI-only src/main.cpp, which has no
// way to run a solve unattended: SURVEY.md section 0 D3/D4).
//
//   shm3d_cli INPUT.{obj,pc} [--grid|-g] [--h K] [--t TCOEF] [--fast|-f] [--verbose|-V] [--device D]
//             [-o OUT.{npy,raw}] [--isoval C --iso-out SURFACE.obj] [--dry-run]
//
// Flags follow the reference's (src/main.cpp:230-238: positional mesh, -g/--grid, -f/--fast, -V/--verbose, --help)
// plus the --h the README documents (README.md:70) but main.cpp never defined.  Input readers restate
// geometry-central's OBJ loader (deps/geometry-central/src/surface/simple_polygon_mesh.cpp:167-232 + meshio.cpp:22-29:
// v / f records, index before the first '/', unused vertices stripped, no vertex merging) and main.cpp's .pc reader
// (src/main.cpp:196-225: 'v x y z' and 'vn x y z' records).
// --isoval / --iso-out: what the GUI's "Contour" + "Export isosurface" buttons do (src/main.cpp:116-128, :160-190): the
// level set phi = C extracted on the GPU (row N3) and written as an OBJ (same vertices / faces polyscope would hold).
// --dry-run prints the grid / source scalars and exits without touching the GPU (used by the CPU tests).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../include/shm3d/signed_heat_grid_solver.hpp"

static bool read_obj(const std::string& path, shm3d::PolygonMesh& mesh) {
    std::ifstream in(path);
    if (!in) return false;
    std::vector<double> V;
    std::vector<std::vector<int64_t>> F;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tok;
        if (!(ss >> tok)) continue;
        if (tok == "v") {
            double x, y, z;
            ss >> x >> y >> z;
            V.push_back(x);
            V.push_back(y);
            V.push_back(z);
        } else if (tok == "f") {
            std::vector<int64_t> f;
            while (ss >> tok) {
                long long idx = std::atoll(tok.substr(0, tok.find('/')).c_str());
                f.push_back(idx > 0 ? idx - 1 : (long long)(V.size() / 3) + idx);
            }
            if (!f.empty()) F.push_back(f);
        }
    }
    // stripUnusedVertices
    const int64_t nV = (int64_t)V.size() / 3;
    std::vector<int64_t> remap(nV, -1);
    for (auto& f : F)
        for (int64_t v : f)
            if (v >= 0 && v < nV) remap[v] = 0;
    int64_t cnt = 0;
    for (int64_t v = 0; v < nV; v++)
        if (remap[v] == 0) {
            remap[v] = cnt++;
            for (int a = 0; a < 3; a++) mesh.vertexPositions.push_back(V[3 * v + a]);
        }
    mesh.faceOffsets.push_back(0);
    for (auto& f : F) {
        for (int64_t v : f) mesh.faceVertices.push_back((v >= 0 && v < nV) ? remap[v] : -1);
        mesh.faceOffsets.push_back((int64_t)mesh.faceVertices.size());
    }
    return true;
}

static bool read_pc(const std::string& path, shm3d::OrientedPointCloud& pc) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tok;
        if (!(ss >> tok)) continue;
        double x, y, z;
        if (tok == "v") {
            ss >> x >> y >> z;
            pc.positions.insert(pc.positions.end(), {x, y, z});
        } else if (tok == "vn") {
            ss >> x >> y >> z;
            pc.normals.insert(pc.normals.end(), {x, y, z});
        }
    }
    return true;
}

static bool write_npy(const std::string& path, const std::vector<double>& v, size_t nx, size_t ny, size_t nz) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    // phi[k][j][i] in C order == node index i + j*nx + k*nx*ny
    std::string hdr = "{'descr': '<f8', 'fortran_order': False, 'shape': (" + std::to_string(nz) + ", " + std::to_string(ny) +
                      ", " + std::to_string(nx) + "), }";
    while ((10 + hdr.size() + 1) % 64) hdr += ' ';
    hdr += '\n';
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    std::fwrite(magic, 1, 8, f);
    unsigned short hl = (unsigned short)hdr.size();
    std::fwrite(&hl, 2, 1, f);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
    return true;
}

static void usage() {
    std::fprintf(stderr,
                 "usage: shm3d_cli INPUT.{obj,pc} [-g|--grid] [--h K] [--t TCOEF] [-f|--fast] [-V|--verbose] [--device D]\n"
                 "                 [-o OUT.{npy,raw}] [--isoval C --iso-out SURFACE.obj] [--no-reference-underflow] [--dry-run]\n"
                 "  Generalized signed distance to INPUT on an nx^3 grid, nx = 16*2^K (B200 grid solver; %s)\n",
                 shm3d_version());
}

int main(int argc, char** argv) {
    std::string input, output, iso_output;
    float isoval = 0.f;
    shm3d::SignedHeat3DOptions opts;
    bool verbose = false, dry = false, ref_underflow = true;
    int device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char* what) -> const char* {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value for %s\n", what);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--help") { usage(); return 0; }
        else if (a == "-g" || a == "--grid") {}
        else if (a == "-f" || a == "--fast") opts.fastIntegration = true;
        else if (a == "-V" || a == "--verbose") verbose = true;
        else if (a == "--h") opts.hCoef = std::atof(need("--h"));
        else if (a == "--t") opts.tCoef = std::atof(need("--t"));
        else if (a == "--device") device = std::atoi(need("--device"));
        else if (a == "-o") output = need("-o");
        else if (a == "--isoval") isoval = (float)std::atof(need("--isoval"));
        else if (a == "--iso-out") iso_output = need("--iso-out");
        else if (a == "--reference-underflow") ref_underflow = true;  // (the default)
        else if (a == "--no-reference-underflow") ref_underflow = false;
        else if (a == "--dry-run") dry = true;
        else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); usage(); return 2; }
        else input = a;
    }
    if (input.empty()) {  // main.cpp:253-256
        std::fprintf(stderr, "Please specify a mesh file as argument.\n");
        usage();
        return 1;
    }
    const bool is_pc = input.size() > 3 && input.substr(input.size() - 3) == ".pc";
    try {
        shm3d::PolygonMesh mesh;
        shm3d::OrientedPointCloud pc;
        if (is_pc) {
            if (!read_pc(input, pc) || pc.nPoints() == 0) throw std::runtime_error("cannot read point cloud " + input);
            if ((int64_t)pc.normals.size() != 3 * pc.nPoints()) throw std::runtime_error("point cloud needs one 'vn' per 'v'");
            pc.computeWeights();  // row N1: see include/shm3d/signed_heat_grid_solver.hpp
            std::fprintf(stderr, "[shm3d_cli] point cloud: tufted-cover vertex dual areas, mean intrinsic edge length h = %g\n",
                         pc.meanEdgeLength);
        } else if (!read_obj(input, mesh) || mesh.nFaces() == 0) {
            throw std::runtime_error("cannot read mesh " + input);
        }
        if (dry) {
            shm3d_params p;
            double h = 0;
            int rc;
            int64_t n;
            if (is_pc) {
                h = pc.meanEdgeLength;
                n = pc.nPoints();
                rc = shm3d_prepare_points(pc.positions.data(), n, h, opts.tCoef, opts.hCoef, opts.scale, &p);
            } else {
                n = mesh.nFaces();
                rc = shm3d_prepare_mesh(mesh.vertexPositions.data(), mesh.nVertices(), mesh.faceVertices.data(),
                                        mesh.faceOffsets.data(), n, opts.tCoef, opts.hCoef, opts.scale, &p, nullptr, nullptr,
                                        nullptr, &h);
            }
            if (rc != SHM3D_OK) throw std::invalid_argument("invalid input geometry");
            std::printf("{\"nx\": %d, \"sources\": %lld, \"vertices\": %lld, \"h\": %.17g, \"lambda\": %.17g, \"cell\": %.17g, "
                        "\"bbox_min\": [%.17g, %.17g, %.17g]}\n",
                        p.nx, (long long)n, (long long)(is_pc ? n : mesh.nVertices()), h, p.lambda, p.cell, p.bbox_min[0],
                        p.bbox_min[1], p.bbox_min[2]);
            return 0;
        }
        shm3d::SignedHeatGridSolver solver(device);
        solver.VERBOSE = verbose;
        solver.referenceUnderflow = ref_underflow;
        std::vector<double> phi = is_pc ? solver.computeDistance(pc, opts) : solver.computeDistance(mesh, opts);
        double lo = 1e300, hi = -1e300;
        for (double v : phi) {
            lo = std::min(lo, v);
            hi = std::max(hi, v);
        }
        std::fprintf(stderr, "min: %g\tmax: %g\n", lo, hi);  // src/main.cpp:101
        const shm3d_stats& st = solver.lastStats();
        std::fprintf(stderr, "[shm3d_cli] %zu^3 nodes, %lld sources, m = %d constraints, %d PCG iterations, %.1f ms\n", solver.nx(),
                     (long long)(is_pc ? pc.nPoints() : mesh.nFaces()), st.m_constraints, st.cg_iters, st.ms_total);
        if (!output.empty()) {
            bool ok;
            if (output.size() > 4 && output.substr(output.size() - 4) == ".npy") {
                ok = write_npy(output, phi, solver.nx(), solver.ny(), solver.nz());
            } else {
                FILE* f = std::fopen(output.c_str(), "wb");
                ok = f && std::fwrite(phi.data(), sizeof(double), phi.size(), f) == phi.size();
                if (f) std::fclose(f);
            }
            if (!ok) throw std::runtime_error("cannot write " + output);
        }
        if (!iso_output.empty()) {
            shm3d::SignedHeatGridSolver::IsoMesh m = solver.isosurface(phi, isoval);
            FILE* f = std::fopen(iso_output.c_str(), "w");
            if (!f) throw std::runtime_error("cannot write " + iso_output);
            for (size_t v = 0; v < m.vertices.size() / 3; v++)
                std::fprintf(f, "v %.9g %.9g %.9g\n", m.vertices[3 * v], m.vertices[3 * v + 1], m.vertices[3 * v + 2]);
            for (size_t t = 0; t < m.triangles.size() / 3; t++)
                std::fprintf(f, "f %u %u %u\n", m.triangles[3 * t] + 1, m.triangles[3 * t + 1] + 1, m.triangles[3 * t + 2] + 1);
            std::fclose(f);
            std::fprintf(stderr, "[shm3d_cli] isosurface phi = %g: %lld vertices, %lld triangles, %.3f ms on the device -> %s\n",
                         (double)isoval, (long long)m.stats.n_vertices, (long long)m.stats.n_triangles, m.stats.ms_device,
                         iso_output.c_str());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "shm3d_cli: %s\n", e.what());
        return 3;
    }
    return 0;
}
