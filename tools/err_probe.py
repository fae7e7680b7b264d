"""Where does phi's distance to the fp64 oracle sit at 512^3?  Error of two solves (cull_tau nudged) against
tests/golden/sphere_h5.npz: relative L2, mean (constant offset), what is left after removing the mean, and the error binned
by distance from the sphere's centre."""
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
n = p.nx
k, rem = np.divmod(sub, n * n)
j, i = np.divmod(rem, n)
xyz = np.stack([i, j, k], axis=1) * p.cell + np.array(p.bbox_min)
rad = np.linalg.norm(xyz - (np.array(p.bbox_min) + 0.5 * p.cell * (n - 1)), axis=1)
ctx = shm3d.Context(0)
prev = None
for tau in (10.0, 10.05, 10.02):
    q = shm3d.Params.from_buffer_copy(p)
    q.cull_tau = tau
    phi, st = ctx.solve(q, pos, nrm, area)
    e = phi[sub] - ref
    nr = np.linalg.norm(ref)
    bins = [0, 0.25, 0.5, 0.75, 0.95, 1.05, 1.5, 2.0, 2.5, 4.0]
    shells = []
    for a, b in zip(bins[:-1], bins[1:]):
        m = (rad >= a) & (rad < b)
        shells.append((b, int(m.sum()), float(np.sqrt((e[m] ** 2).mean())) if m.any() else 0.0, float(e[m].mean()) if m.any() else 0.0))
    out = {"cull_tau": tau, "rel_l2": float(np.linalg.norm(e) / nr), "mean_err": float(e.mean()),
           "rel_l2_without_mean": float(np.linalg.norm(e - e.mean()) / nr), "max_abs": float(np.abs(e).max()),
           "radius_of_max": float(rad[np.abs(e).argmax()]), "shift": st.shift, "oracle_shift": float(gl["shift"]),
           "shells(r_hi,count,rms,mean)": [(s[0], s[1], round(s[2], 8), round(s[3], 8)) for s in shells]}
    if prev is not None:
        d = phi[sub] - prev
        out["vs_previous_run"] = {"rel_l2": float(np.linalg.norm(d) / nr), "mean": float(d.mean()),
                                  "rel_l2_without_mean": float(np.linalg.norm(d - d.mean()) / nr)}
    prev = phi[sub].copy()
    print(json.dumps(out), flush=True)
ctx.close()
