// polyscope/volume_grid.h -- TEST INFRASTRUCTURE.  Stand-in for the one polyscope entry point the grid solver touches
// (registerVolumeGrid, src/signed_heat_grid_solver.cpp:35,143), on top of the REAL glm vendored by the reference
// (deps/polyscope/deps/glm).  Used when the reference's headers are compiled against the real geometry-central headers
// (oracle/Makefile: _ref/libshm_adapter_gc.so); polyscope itself needs OpenGL / GLFW / imgui and cannot be built here.
#pragma once
#include <array>
#include <string>

#include <glm/glm.hpp>

namespace polyscope {
struct VolumeGrid {
    std::string name;
    glm::uvec3 dim{0u, 0u, 0u};
    glm::vec3 bmin{0.f, 0.f, 0.f}, bmax{0.f, 0.f, 0.f};
};
inline VolumeGrid& shim_last_grid() {
    static VolumeGrid g;
    return g;
}
inline VolumeGrid* registerVolumeGrid(const std::string& name, glm::uvec3 dim, glm::vec3 bmin, glm::vec3 bmax) {
    VolumeGrid& g = shim_last_grid();
    g.name = name;
    g.dim = dim;
    g.bmin = bmin;
    g.bmax = bmax;
    return &g;
}
}  // namespace polyscope
