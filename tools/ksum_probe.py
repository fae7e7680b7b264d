"""k_sum experiment probe: time Steps 1-2 of the bench workload (512^3, 1e5 triangles) with a given build of the library and
check the result against the fp64 fixture (tests/golden/sphere_h5.npz).
    python tools/ksum_probe.py [path/to/libshm3d_grid.so]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import shm3d  # noqa: E402

if len(sys.argv) > 1:
    shm3d.LIB_PATH = os.path.abspath(sys.argv[1])
import bench  # noqa: E402


def main():
    p, pos, nrm, area, _ = bench.prepare("sphere512")
    gl = np.load(os.path.join(ROOT, "tests", "golden", "sphere_h5.npz"))
    ctx = shm3d.Context(0)
    best = None
    for _ in range(3):
        Y, st = ctx.step12(p, pos, nrm, area)
        best = st.ms_sum if best is None else min(best, st.ms_sum)
    sub = gl["sub_index"]
    dy = np.abs(Y[:, sub].T - gl["Y_sub"]).max(axis=1)
    phi, st2 = ctx.solve(p, pos, nrm, area)
    e = np.linalg.norm(phi[sub] - gl["sub_phi"]) / np.linalg.norm(gl["sub_phi"])
    print(json.dumps({"lib": os.path.basename(shm3d.LIB_PATH), "ms_sum": round(best, 2), "pairs": int(st.pairs_evaluated),
                      "dY_max": float(dy.max()), "dY_p999": float(np.quantile(dy, 0.999)), "dY_p99": float(np.quantile(dy, 0.99)),
                      "phi_rel_l2": float(e), "pcg_its": int(st2.cg_iters), "ms_total": round(st2.ms_total, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
