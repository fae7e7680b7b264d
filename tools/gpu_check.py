"""Quick on-GPU parity walk-through against the golden fixtures (development aid; the real tests are tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))
import numpy as np
import shm3d

def rel(a, b):
    return np.linalg.norm(np.asarray(a, float) - b) / np.linalg.norm(b)

def faces(z):
    fo = z["face_offsets"]; fv = z["face_vertices"]
    return [fv[fo[i]:fo[i+1]].tolist() for i in range(len(fo)-1)]

ctx = shm3d.Context(0)
for name, hcs in [("bunny_small", [0, 1]), ("knot", [1]), ("polygon-bear", [0])]:
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    F = faces(z)
    for hc in hcs:
        tag = f"h{hc}"
        p, pos, nrm, area, h = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
        N = p.N
        Yg = z[tag + "_Y"].reshape(N, 3).T.astype(np.float64)
        for tau in (float("inf"), 12.0):
            p.cull_tau = tau
            Y, st = ctx.step12(p, pos, nrm, area)
            print(f"{name} {tag} tau={tau}: step12 max|dY| {np.abs(Y - Yg).max():.3e} pairs {st.pairs_evaluated}/{st.pairs_bruteforce} ms {st.ms_sum:.3f}")
        b = ctx.rhs(p, Yg.astype(np.float32))
        bg = z[tag + "_b"] * p.cell ** 2
        print(f"   rhs rel {rel(b, bg):.3e}")
        for flags in (shm3d.FLAG_SCRUB_NONFINITE | shm3d.FLAG_NO_MG, shm3d.FLAG_SCRUB_NONFINITE):
            p.flags = flags
            try:
                phi, st = ctx.step3(p, pos, area, bg.astype(np.float32))
                print(f"   step3 flags={flags}: rel {rel(phi, z[tag + '_phi']):.3e} its {st.cg_iters} res {st.cg_rel_residual:.2e} pcg ms {st.ms_pcg:.2f} constr ms {st.ms_constraints:.2f}")
            except shm3d.Shm3dError as e:
                print("   step3 FAILED", e)
        p.flags = shm3d.FLAG_SCRUB_NONFINITE
        phi, st = ctx.solve(p, pos, nrm, area)
        print(f"   solve: rel {rel(phi, z[tag + '_phi']):.3e} its {st.cg_iters} total ms {st.ms_total:.2f} launches {st.kernel_launches}")
        print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.asdict().items()})
