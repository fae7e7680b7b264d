"""
CPU-only generator of fp64 oracle fixtures at the grid sizes BASELINE.json names (run in the build container; hours of
host-core time in total, results committed):

  spray_h3     data/SprayBottle.obj at 128^3  (config[3]'s input; the reference's X.norm() underflow artefact included)
  bunnypc_h4   data/bunny.pc at 256^3, point overload, geometry-central's own tufted-cover weights (config[2])
  sphere_h4    the bench workload (1e5-triangle sphere) at 256^3
  sphere_h5    the bench workload at 512^3 -- the headline configuration (config[4] on one GPU)
  spray_h5     data/SprayBottle.obj at 512^3 (config[3] at its own size)

Steps 1-2: oracle/shm_oracle_large.step12_bricks (fp64, far clusters skipped only under an a-posteriori bound of 1e-13 of
the kept sum; pinned to the plain loop by tests/test_oracle_large.py).  Step 3: fp64 projected CG on the KKT system (the
reference's sparse LU is infeasible beyond 64^3; the two are shown equal at <= 32^3 in tests/test_oracle.py), relative
residual 1e-9.  Stored: grid, lambda, m, phi statistics, and a strided subsample of phi, Y (float32) and the non-finite
mask -- a few hundred kB per fixture.

    python tests/golden/make_golden_baseline.py spray_h3 bunnypc_h4 sphere_h4 sphere_h5 spray_h5
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import shm_oracle as o  # noqa: E402
from oracle import shm_oracle_large as ol  # noqa: E402


def solve(name, pos, nrm, area, h, centroid, radius, hCoef, scrub, stride, tol=1e-9, extra=None):
    g = o.make_grid(centroid, radius, hCoef)
    lam = o.lambda_from_h(h)
    t = time.time()
    Y, st = ol.step12_bricks(g, lam, pos, nrm, area, tau=36.0, eps=1e-13)
    print("[%s] %d^3, M=%d, lambda=%.4f: steps 1-2 %.0f s, %.3g pairs (%.1f%% of brute force), worst skipped/kept %.1e, "
          "bricks redone %d" % (name, g.nx, len(area), lam, time.time() - t, st["pairs"],
                                100 * st["pairs"] / (g.N * len(area)), st["worst_skipped_ratio"], st["bricks_redone"]),
          flush=True)
    bad = ~np.isfinite(Y).all(axis=1)
    b, nscrub = ol.div_rhs(g, Y, scrub_nonfinite=scrub)
    assert scrub or np.isfinite(b).all(), "point overload: the reference's solve would throw on this input"
    src, idx, w = o.constraints(g, pos)
    t = time.time()
    phi, its = ol.solve_projected_cg(
        g, b, idx, w, tol=tol,
        callback=lambda it, x, rel: print("   cg", it, "%.3e" % np.sqrt(rel), flush=True) if it % 250 == 0 else None)
    print("[%s] projected CG: %d iterations, %.0f s, m = %d" % (name, its, time.time() - t, len(src)), flush=True)
    shift = o.source_average(g, phi, pos, area)
    phi -= shift
    sub = np.arange(stride // 2, g.N, stride)
    out = dict(nx=g.nx, cell=g.cell, bmin=g.bmin, lam=lam, h=h, m=len(src), its=its, cg_tol=tol, shift=shift,
               phi_stats=np.array([phi.min(), phi.max(), np.linalg.norm(phi)]), sub_index=sub, sub_phi=phi[sub],
               Y_sub=Y[sub].astype(np.float32), sub_nonfinite=bad[sub], n_nonfinite_nodes=int(bad.sum()),
               n_scrubbed_rhs=nscrub, step12_pairs=st["pairs"], step12_worst_skipped_ratio=st["worst_skipped_ratio"])
    # the gradual-underflow shell of the reference's X.norm() (|Y| != 1 although finite): how many nodes, how far off
    nrmY = np.linalg.norm(Y[~bad], axis=1)
    out["n_nodes_off_unit"] = int((np.abs(nrmY - 1) > 1e-6).sum())
    out["max_off_unit"] = float(np.abs(nrmY - 1).max())
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("[%s] phi min/max/L2 %.6g %.6g %.6g; non-finite Y nodes %d, scrubbed rhs %d, |Y| off unit at %d nodes (max %.2e)"
          % (name, phi.min(), phi.max(), np.linalg.norm(phi), bad.sum(), nscrub, out["n_nodes_off_unit"],
             out["max_off_unit"]), flush=True)


def spray(hCoef, stride):
    d = np.load(os.path.join(HERE, "spraybottle_mesh.npz"))
    s = o.mesh_sources(d["V"], d["F"].tolist())
    solve("spray_h%d" % hCoef, s["pos"], s["nrm"], s["area"], s["h"], s["centroid"], s["radius"], hCoef, True, stride)


def sphere(hCoef, stride):
    from synth import fibonacci_sphere
    V, F = fibonacci_sphere(100000)
    s = o.mesh_sources(V, F.tolist())
    solve("sphere_h%d" % hCoef, s["pos"], s["nrm"], s["area"], s["h"], s["centroid"], s["radius"], hCoef, True, stride)


def bunnypc(hCoef, stride):
    """Weights from geometry-central's own pipeline (oracle/_ref/libshm_gc_ref.so, fixture point_weights_gc.npz)."""
    d = np.load(os.path.join(HERE, "bunny_pc.npz"))
    w = np.load(os.path.join(HERE, "point_weights_gc.npz"))
    P, N, areas, h = d["P"], d["N"], w["bunny_pc_areas"], float(w["bunny_pc_h"])
    c = P.mean(axis=0)                       # centroid / radius of a point cloud: src/signed_heat_3d.cpp:14-43
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    solve("bunnypc_h%d" % hCoef, P, N, areas, h, c, r, hCoef, False, stride)


if __name__ == "__main__":
    for what in sys.argv[1:]:
        {"spray_h3": lambda: spray(3, 97), "bunnypc_h4": lambda: bunnypc(4, 251), "sphere_h4": lambda: sphere(4, 251),
         "sphere_h5": lambda: sphere(5, 2039), "spray_h5": lambda: spray(5, 2039), "spray_h4": lambda: spray(4, 251),
         "spray_h1": lambda: spray(1, 1)}[what]()
