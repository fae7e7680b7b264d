"""PCG stopping tolerance vs distance to the fp64 oracle at the headline size (512^3 sphere, tests/golden/sphere_h5.npz), over
inputs that differ only in the last bits of Y (cull_tau nudged): how much of phi's error is where the iteration stops.
    python tools/tol_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import shm3d  # noqa: E402
import bench  # noqa: E402

p, pos, nrm, area, _ = bench.prepare("sphere512")
gl = np.load(os.path.join(ROOT, "tests", "golden", "sphere_h5.npz"))
sub, ref = gl["sub_index"], gl["sub_phi"]
ctx = shm3d.Context(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "tol"
if mode == "tol":
    for tol in ([float(a) for a in sys.argv[2:]] or [3e-6, 2e-6, 1.5e-6, 1e-6, 5e-7]):
        errs, its = [], []
        for tau in (10.0, 10.01, 10.02, 10.03, 10.05, 9.98):
            q = shm3d.Params.from_buffer_copy(p)
            q.cull_tau = tau
            q.cg_rel_tol = tol
            phi, st = ctx.solve(q, pos, nrm, area)
            errs.append(float(np.linalg.norm(phi[sub] - ref) / np.linalg.norm(ref)))
            its.append(int(st.cg_iters))
        print(json.dumps({"cg_rel_tol": tol, "phi_rel_l2": [round(e, 8) for e in errs], "worst": max(errs), "iters": its}), flush=True)
else:  # "tau": the far-field culling threshold vs error and Steps 1-2 time
    for tau0 in ([float(a) for a in sys.argv[2:]] or [10.0, 10.5, 11.0, 11.5, 12.0, 13.0, 14.0]):
        errs, ms, pairs = [], [], []
        for d in (0.0, 0.01, 0.02, 0.03, 0.05, -0.02):
            q = shm3d.Params.from_buffer_copy(p)
            q.cull_tau = tau0 + d
            phi, st = ctx.solve(q, pos, nrm, area)
            errs.append(float(np.linalg.norm(phi[sub] - ref) / np.linalg.norm(ref)))
            ms.append(st.ms_sum)
            pairs.append(st.pairs_evaluated / st.pairs_bruteforce)
        print(json.dumps({"cull_tau": tau0, "phi_rel_l2": [round(e, 8) for e in errs], "worst": max(errs),
                          "ms_sum": round(min(ms), 1), "pairs_frac": round(float(np.mean(pairs)), 4)}), flush=True)
ctx.close()
