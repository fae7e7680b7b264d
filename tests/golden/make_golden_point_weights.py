"""
Golden point-cloud weights for row N1, made with geometry-central's OWN pipeline (its vendored sources compiled from
/root/reference/deps/geometry-central by oracle/Makefile -> oracle/_ref/libshm_gc_ref.so) exactly as the reference calls
it for the point-cloud overload (src/main.cpp:277-285, src/signed_heat_grid_solver.cpp:149-151,165) -- run here, on the
CPU box; the GPU box has no /root/reference.

Inputs: data/bunny.pc (tests/golden/bunny_pc.npz) and two synthetic clouds built by the formulas below.  Also stored: the
reference's end-to-end signed distance for bunny.pc at 32^3 (its own weights + its own grid solver).

    python tests/golden/make_golden_point_weights.py     # writes point_weights_gc.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import reference_build as rb  # noqa: E402


def clouds():
    d = np.load(os.path.join(HERE, "bunny_pc.npz"))
    yield "bunny_pc", d["P"], d["N"]
    rng = np.random.default_rng(3)
    P = rng.standard_normal((3000, 3))
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    yield "random_sphere_3000", P, P.copy()
    u, v = rng.uniform(0, 2 * np.pi, 4000), rng.uniform(0, 2 * np.pi, 4000)
    P = np.stack([(2 + 0.7 * np.cos(v)) * np.cos(u), (2 + 0.7 * np.cos(v)) * np.sin(u), 0.7 * np.sin(v)], axis=1)
    N = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], axis=1)
    yield "random_torus_4000", P, N


def main():
    assert rb.build() and rb.gc_available()
    out = {}
    for name, P, N in clouds():
        areas, h, nf, ne = rb.gc_point_weights(P, N)
        out[name + "_areas"] = areas
        out[name + "_h"] = h
        out[name + "_faces_edges"] = np.array([nf, ne])
        print(name, len(P), "h", h, "faces", nf, "edges", ne, "area", areas.sum())
    # BASELINE config[2] at 32^3, end to end through the reference: its own computeDistance(PointPositionNormalGeometry&)
    # linked with the real geometry-central (oracle/_ref/libshm_ref_gc.so: own tufted-cover weights, own grid solver)
    d = np.load(os.path.join(HERE, "bunny_pc.npz"))
    phi, info = rb.ref_gc_compute_distance_points(d["P"], d["N"], hCoef=1, return_info=True)
    out["bunny_pc_h1_phi"] = phi.astype(np.float64)
    out["bunny_pc_h1_bbox"] = info["bbox"]
    print("bunny_pc 32^3 reference phi", phi.min(), phi.max())
    np.savez_compressed(os.path.join(HERE, "point_weights_gc.npz"), **out)


if __name__ == "__main__":
    main()
