"""Development probe: solve synthetic spheres / golden meshes at growing grid sizes on the GPU and print stage times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import shm3d
from synth import fibonacci_sphere

ctx = shm3d.Context(0)
cases = sys.argv[1:] or ["sphere:3", "sphere:4", "sphere:5"]
t = time.time(); V, F = fibonacci_sphere(100000); print("sphere gen %.2fs" % (time.time() - t))
for c in cases:
    name, hc = c.split(":"); hc = int(hc)
    if name == "sphere":
        VV, FF = V, F
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")); fo = z["face_offsets"]; fv = z["face_vertices"]
        VV = z["V"]; FF = [fv[fo[i]:fo[i+1]].tolist() for i in range(len(fo) - 1)]
    t = time.time(); p, pos, nrm, area, h = shm3d.prepare_mesh(VV, FF, hCoef=hc); tp = time.time() - t
    p.flags |= int(os.environ.get("FLAGS", "0")); p.mg_smooth = int(os.environ.get("SMOOTH", "0"))
    p.cg_rel_tol = float(os.environ.get("TOL", "0")); p.cull_tau = float(os.environ.get("TAU", "0")); p.mg_constrained_from = int(os.environ.get("CMG", "0"))
    for rep in range(int(os.environ.get("REPS", "2"))):
        t = time.time(); phi, st = ctx.solve(p, pos, nrm, area); wall = time.time() - t
        d = st.asdict()
        print(f"{name} {p.nx}^3 M={len(area)} h={h:.4f} prep={tp*1e3:.0f}ms wall={wall*1e3:.0f}ms nodes/s={p.N/wall:.3e} "
              f"sum={d['ms_sum']:.1f} kept={d['pairs_evaluated']/d['pairs_bruteforce']:.3f} pairs/s={d['pairs_evaluated']/d['ms_sum']*1e3:.3e} "
              f"constr={d['ms_constraints']:.0f} m={d['m_constraints']} pcg={d['ms_pcg']:.1f} its={d['cg_iters']} "
              f"ms/it={d['ms_pcg']/max(1,d['cg_iters']):.2f} prof[vc={d['ms_pcg_vcycle']:.0f} pr={d['ms_pcg_projector']:.0f}/{d['pcg_projector_applies']} st={d['ms_pcg_stencil']:.0f} up={d['ms_pcg_update']:.0f}] h2d={d['ms_h2d']:.0f} d2h={d['ms_d2h']:.0f} launches={d['kernel_launches']}")
    if name == "sphere":
        g_cell = p.cell
        # accuracy vs analytic distance near the surface
        n = p.nx
        ax = np.array(p.bbox_min)[0] + p.cell * np.arange(n)
        k = n // 2
        sl = phi.reshape(n, n, n)[k]
        X, Yc = np.meshgrid(ax, ax, indexing="xy")
        d = np.sqrt(X**2 + Yc**2 + ax[k]**2) - 1
        band = np.abs(d) < 0.3
        print(f"   mid-slice |phi-(r-1)| max in band: {np.abs(sl - d)[band].max():.4f} (cell {p.cell:.4f})")
