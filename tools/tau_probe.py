"""Development probe: far-field culling threshold tau vs accuracy (knot.obj @128^3 against the fp64 oracle fixture;
sphere @512^3 against tau = 16) and vs k_sum time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, shm3d
from conftest import load_golden
from synth import fibonacci_sphere
ctx = shm3d.Context(0)
z, F = load_golden("knot"); gl = np.load(os.path.join(ROOT, "tests/golden/knot_h3.npz"))
p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=3)
sub = gl["sub_index"]; ref = gl["sub_phi"]
for tau in (float("inf"), 12, 10, 9) if len(sys.argv) > 1 else (float("inf"), 14, 12, 10, 9, 8, 7):
    p.cull_tau = tau
    Y, s12 = ctx.step12(p, pos, nrm, area)
    phi, st = ctx.solve(p, pos, nrm, area)
    print(f"knot128 tau {tau}: kept {st.pairs_evaluated/st.pairs_bruteforce:.3f} sum {st.ms_sum:.1f} ms  max|dY| {np.abs(Y[:, sub].T - gl['Y_sub']).max():.2e}  phi rel-L2 vs fp64 oracle {np.linalg.norm(phi[sub]-ref)/np.linalg.norm(ref):.3e} its {st.cg_iters}", flush=True)
for name, hc in (("bunny_small", 2), ("polygon-bear", 2), ("bunny_small", 0)):
    z, F = load_golden(name)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    p.cg_rel_tol = 1e-7
    base = None
    for tau in (float("inf"), 12, 10, 9, 8, 7):
        p.cull_tau = tau
        Y, _ = ctx.step12(p, pos, nrm, area)
        phi, st = ctx.solve(p, pos, nrm, area)
        if base is None: base, Yb = phi.copy(), Y.copy()
        print(f"{name} {p.nx}^3 tau {tau}: kept {st.pairs_evaluated/st.pairs_bruteforce:.3f} max|dY| vs brute force {np.abs(Y-Yb).max():.2e} phi rel-L2 vs brute force {np.linalg.norm(phi-base)/np.linalg.norm(base):.3e}", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "small":
    sys.exit(0)
V, F = fibonacci_sphere(100000)
p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=5)
base = None
for tau in (16, 12, 10, 9, 8):
    p.cull_tau = tau
    phi, st = ctx.solve(p, pos, nrm, area)
    if base is None: base = phi.copy()
    print(f"sphere512 tau {tau}: kept {st.pairs_evaluated/st.pairs_bruteforce:.3f} sum {st.ms_sum:.1f} ms phi rel-L2 vs tau=16 {np.linalg.norm(phi-base)/np.linalg.norm(base):.3e} its {st.cg_iters}", flush=True)
