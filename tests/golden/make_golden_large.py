"""
CPU-only generator of the BASELINE-size fixtures (run here; the GPU box has no /root/reference):
  knot_h3.npz        -- data/knot.obj at hCoef 3 (128^3, BASELINE config[1]): the oracle's phi statistics and a fixed
                        strided subsample of the field.  Steps 1-2 by the fp64 C loop (6.4e10 pairs, minutes), Step 3 by
                        the oracle's fp64 projected CG (the reference's sparse LU is infeasible at this size; the two
                        are shown equal at <= 32^3 in tests/test_oracle.py).
  spraybottle_mesh.npz, bunny_pc.npz -- the raw inputs of BASELINE configs [3] and [2] (mesh / oriented points), so the
                        GPU tests can run them without the reference tree.
    python tests/golden/make_golden_large.py [knot] [inputs]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import shm_oracle as o  # noqa: E402

REF = "/root/reference/data"


def knot128():
    V, F = o.read_obj(os.path.join(REF, "knot.obj"))
    s = o.mesh_sources(V, F)
    g = o.make_grid(s["centroid"], s["radius"], 3)
    lam = o.lambda_from_h(s["h"])
    t = time.time()
    Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"])
    print("step12 %.1fs" % (time.time() - t), flush=True)
    b = o.div_rhs(g, Y)
    src, idx, w = o.constraints(g, s["pos"])
    t = time.time()
    phi, its = o.solve_projected_cg(g, b, idx, w, tol=1e-10,
                                    callback=lambda it, x, rel: print(it, rel, flush=True) if it % 100 == 0 else None)
    print("projected CG %d its %.1fs" % (its, time.time() - t), flush=True)
    phi = phi - o.source_average(g, phi, s["pos"], s["area"])
    sub = np.arange(0, g.N, 97)
    np.savez_compressed(os.path.join(HERE, "knot_h3.npz"), nx=g.nx, cell=g.cell, bmin=g.bmin, lam=lam, m=len(src),
                        phi_stats=np.array([phi.min(), phi.max(), np.linalg.norm(phi)]), sub_index=sub,
                        sub_phi=phi[sub], Y_sub=Y.reshape(-1, 3)[sub].astype(np.float32), its=its)
    print("knot 128^3: m", len(src), "phi min/max/L2", phi.min(), phi.max(), np.linalg.norm(phi))


def inputs():
    V, F = o.read_obj(os.path.join(REF, "SprayBottle.obj"))
    assert all(len(f) == 3 for f in F)
    np.savez_compressed(os.path.join(HERE, "spraybottle_mesh.npz"), V=V.astype(np.float64), F=np.asarray(F, dtype=np.int32))
    P, N = o.read_pc(os.path.join(REF, "bunny.pc"))
    np.savez_compressed(os.path.join(HERE, "bunny_pc.npz"), P=P, N=N)
    print("SprayBottle", V.shape, len(F), "bunny.pc", P.shape)


if __name__ == "__main__":
    what = sys.argv[1:] or ["inputs", "knot"]
    if "inputs" in what:
        inputs()
    if "knot" in what:
        knot128()
