// shim_defs.cpp -- TEST INFRASTRUCTURE: the two global objects of the shim (include/shm_ref_shim.h).
#include "shm_ref_shim.h"

namespace geometrycentral {
shim_solve_fn& shim_solver() {
    static shim_solve_fn f = nullptr;
    return f;
}
}  // namespace geometrycentral
namespace polyscope {
VolumeGrid& shim_last_grid() {
    static VolumeGrid g;
    return g;
}
}  // namespace polyscope
