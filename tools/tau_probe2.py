"""Far-field culling threshold tau vs accuracy on every parity input (after the constant-drift fix of round 2): max|dY| and
phi relative L2 against the fp64 fixtures (knot 128^3, SprayBottle 128^3, bunny.pc 256^3, sphere 256^3) and against
tau = inf on the small golden meshes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, shm3d
from conftest import load_golden, GOLDEN
from synth import fibonacci_sphere
ctx = shm3d.Context(0)
TAUS = (float("inf"), 12, 10, 9, 8, 7, 6)


def fixture_case(name, p, pos, nrm, area, gl):
    sub, ref, bad = gl["sub_index"], gl["sub_phi"], gl["sub_nonfinite"] if "sub_nonfinite" in gl.files else None
    for tau in TAUS[1:]:
        p.cull_tau = tau
        Y, _ = ctx.step12(p, pos, nrm, area)
        phi, st = ctx.solve(p, pos, nrm, area)
        Ys = Y[:, sub].T
        ok = np.isfinite(Ys).all(axis=1) if bad is None else ~bad
        dy = np.abs(Ys[ok] - gl["Y_sub"][ok]).max(axis=1)
        print(f"{name} tau {tau}: kept {st.pairs_evaluated/st.pairs_bruteforce:.3f} sum {st.ms_sum:.1f} ms |dY| max {dy.max():.2e} p99 {np.quantile(dy, 0.99):.2e}  phi rel-L2 vs fp64 oracle {np.linalg.norm(phi[sub]-ref)/np.linalg.norm(ref):.3e} its {st.cg_iters}", flush=True)


z, F = load_golden("knot")
p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=3)
fixture_case("knot128", p, pos, nrm, area, np.load(os.path.join(GOLDEN, "knot_h3.npz")))
d = np.load(os.path.join(GOLDEN, "spraybottle_mesh.npz"))
p, pos, nrm, area, _ = shm3d.prepare_mesh(d["V"], d["F"], hCoef=3)
fixture_case("spray128", p, pos, nrm, area, np.load(os.path.join(GOLDEN, "spray_h3.npz")))
d = np.load(os.path.join(GOLDEN, "bunny_pc.npz")); w = np.load(os.path.join(GOLDEN, "point_weights_gc.npz"))
p = shm3d.prepare_points(d["P"], float(w["bunny_pc_h"]), hCoef=4)
fixture_case("bunnypc256", p, d["P"], d["N"], w["bunny_pc_areas"], np.load(os.path.join(GOLDEN, "bunnypc_h4.npz")))
V, F = fibonacci_sphere(100000)
p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=4)
fixture_case("sphere256", p, pos, nrm, area, np.load(os.path.join(GOLDEN, "sphere_h4.npz")))
for name, hc in (("bunny_small", 2), ("polygon-bear", 2), ("bunny_small", 0), ("bunny_small", 1), ("knot", 1)):
    z, F = load_golden(name)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    p.cg_rel_tol = 1e-7
    base = None
    for tau in TAUS:
        p.cull_tau = tau
        Y, _ = ctx.step12(p, pos, nrm, area)
        phi, st = ctx.solve(p, pos, nrm, area)
        if base is None: base, Yb = phi.copy(), Y.copy()
        else: print(f"{name} {p.nx}^3 tau {tau}: kept {st.pairs_evaluated/st.pairs_bruteforce:.3f} max|dY| vs brute force {np.abs(Y-Yb).max():.2e} phi rel-L2 vs brute force {np.linalg.norm(phi-base)/np.linalg.norm(base):.3e}", flush=True)
