"""
Generates the committed golden fixtures under tests/golden/ (run here, on the CPU box; the GPU box has no
/root/reference).  Inputs are the reference's own sample data (data/bunny_small.obj, data/knot.obj, data/bunny.pc);
outputs come from the repo's fp64 oracle (oracle/shm_oracle.py) with Step 3 solved by the reference's own
formulation -- direct sparse LU of the KKT matrix (src/signed_heat_grid_solver.cpp:101-108).

    python tests/golden/make_golden.py            # writes *.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import shm_oracle as o  # noqa: E402

REF = "/root/reference/data"


def faces_to_arrays(F):
    off = np.zeros(len(F) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f) for f in F])
    return np.asarray([v for f in F for v in f], dtype=np.int32), off


def mesh_case(name, hcoefs, store_fields_upto=32):
    V, F = o.read_obj(os.path.join(REF, name + ".obj"))
    fv, fo = faces_to_arrays(F)
    s = o.mesh_sources(V, F)
    out = dict(V=V, face_vertices=fv, face_offsets=fo, h=s["h"], centroid=s["centroid"], radius=s["radius"])
    for hc in hcoefs:
        r = o.compute_distance(s["pos"], s["nrm"], s["area"], s["h"], s["centroid"], s["radius"], hCoef=hc,
                               return_all=True)
        g = r["grid"]
        phi = r["phi"]
        tag = f"h{hc}"
        out[tag + "_nx"] = g.nx
        out[tag + "_cell"] = g.cell
        out[tag + "_bmin"] = g.bmin
        out[tag + "_lambda"] = r["lam"]
        out[tag + "_m"] = r["m"]
        out[tag + "_phi_stats"] = np.array([phi.min(), phi.max(), np.linalg.norm(phi)])
        if g.nx <= store_fields_upto:
            out[tag + "_phi"] = phi
            out[tag + "_Y"] = r["Y"].astype(np.float32)  # interleaved [N,3]
            out[tag + "_b"] = r["b"]
        print(name, tag, g.nx, "m", r["m"], "phi min/max/L2", phi.min(), phi.max(), np.linalg.norm(phi))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    # BASELINE config[0]: bunny_small, hCoef 0 (16^3) and 1 (32^3, what BASELINE calls "32^3")
    mesh_case("bunny_small", [0, 1])
    # knot at 32^3 (SURVEY App. B known answer)
    mesh_case("knot", [1])
    # polygonal faces (degree 3-8): exercises the shoelace areas and polygon barycentres
    mesh_case("polygon-bear", [0])


if __name__ == "__main__":
    main()
