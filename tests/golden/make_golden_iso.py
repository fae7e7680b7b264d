"""
Golden isosurfaces for row N3, made with the REFERENCE's own marching-cubes library (polyscope's vendored
MarchingCube/MC.h + glm compiled from /root/reference by oracle/Makefile -> oracle/_ref/libshm_mc_ref.so) plus
registerIsosurfaceAsMesh's vertex transform -- run here, on the CPU box; the GPU box has no /root/reference.

Input field: the committed golden phi of data/bunny_small.obj at hCoef=1 (32^3; tests/golden/bunny_small.npz, pinned to the
reference's source by tests/test_reference_build.py), narrowed to float32 as polyscope stores it.

    python tests/golden/make_golden_iso.py        # writes iso_bunny_small_h1.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import reference_build as rb  # noqa: E402
from oracle import shm_oracle as o  # noqa: E402


def main():
    assert rb.build() and rb.mc_available()
    z = np.load(os.path.join(HERE, "bunny_small.npz"))
    nx = int(z["h1_nx"])
    g = o.Grid(nx, nx, nx, z["h1_bmin"], float(z["h1_cell"]))
    bmin, bmax = o.grid_bounds_f32(g)
    out = dict(bound_min=bmin, bound_max=bmax, nx=nx)
    for tag, iso in (("iso0", 0.0), ("iso1", 0.25)):
        V, T = rb.isosurface(z["h1_phi"], iso, (nx, nx, nx), bmin, bmax)
        out[tag + "_isoval"] = np.float32(iso)
        out[tag + "_V"] = V
        out[tag + "_T"] = T
        print(tag, iso, V.shape, T.shape)
    np.savez_compressed(os.path.join(HERE, "iso_bunny_small_h1.npz"), **out)


if __name__ == "__main__":
    main()
