// mc_harness.cpp -- TEST INFRASTRUCTURE.  C entry points around the marching-cubes routine the reference's downstream
// consumer runs on the solver's output: polyscope's registerIsosurfaceAsMesh
// (deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228) calls MC::marching_cube of the vendored header-only
// library deps/polyscope/deps/MarchingCubeCpp/include/MarchingCube/MC.h.  That header and the vendored glm are compiled
// here from where they lie under /root/reference (recipe: oracle/Makefile -> oracle/_ref/libshm_mc_ref.so).
// polyscope itself cannot be built in this image, so the three lines around the call (the float narrowing of the
// values, volume_grid.ipp:103-106; gridSpacing(), volume_grid.ipp:72-76; the swizzle + scale + translate of the
// vertices, volume_grid_scalar_quantity.cpp:220-224) are restated below with the same glm expressions.
#include <cstdint>
#include <cstring>

#define GLM_ENABLE_EXPERIMENTAL
#define MC_IMPLEM_ENABLE
#include "MarchingCube/MC.h"

extern "C" {

// the 256-entry case table of the header (for tests that check the product's copy of the data)
const unsigned long long* ref_mc_table(void) { return MC::mc_internalMarching_cube_tris; }

// values: float[nx*ny*nz], index i + j*nx + k*nx*ny (what addNodeScalarQuantity stored after narrowing).
// node_dim = {nx,ny,nz} as registered; bound_min/max = the glm::vec3 the grid was registered with.
// world != 0 applies registerIsosurfaceAsMesh's transform; world == 0 returns MC's own lattice coordinates.
// The mesh is kept until the next call; ref_isosurface_copy hands it out.
static MC::mcMesh g_mesh;

int ref_isosurface(const float* values, float isoval, const uint32_t* node_dim, const float* bound_min,
                   const float* bound_max, int world, int64_t* n_vertices, int64_t* n_indices) {
    g_mesh = MC::mcMesh();
    MC::marching_cube(const_cast<float*>(values), isoval, node_dim[0], node_dim[1], node_dim[2], g_mesh);
    if (world) {
        glm::vec3 boundMin{bound_min[0], bound_min[1], bound_min[2]}, boundMax{bound_max[0], bound_max[1], bound_max[2]};
        glm::uvec3 gridNodeDim{node_dim[0], node_dim[1], node_dim[2]};
        glm::uvec3 gridCellDim = gridNodeDim - 1u;
        glm::vec3 width = boundMax - boundMin;
        glm::vec3 scale = width / (glm::vec3(gridCellDim));
        for (auto& p : g_mesh.vertices) p = glm::vec3{p.z, p.y, p.x} * scale + boundMin;
    }
    *n_vertices = (int64_t)g_mesh.vertices.size();
    *n_indices = (int64_t)g_mesh.indices.size();
    return 0;
}

int ref_isosurface_copy(float* vertices_out, int64_t vertex_capacity, uint32_t* indices_out, int64_t index_capacity) {
    if ((int64_t)g_mesh.vertices.size() > vertex_capacity || (int64_t)g_mesh.indices.size() > index_capacity) return 2;
    for (size_t i = 0; i < g_mesh.vertices.size(); i++)
        for (int a = 0; a < 3; a++) vertices_out[3 * i + a] = g_mesh.vertices[i][a];
    if (!g_mesh.indices.empty()) std::memcpy(indices_out, g_mesh.indices.data(), g_mesh.indices.size() * sizeof(uint32_t));
    g_mesh = MC::mcMesh();
    return 0;
}
}
