"""Row N1 experiment (CPU; needs /root/reference/data): point-cloud weights with nanoflann's tie order (the product path)
vs an independent kNN with ties broken by point index (shm3d_debug_knn_mode 1), and what the difference does to phi through
the fp64 oracle at 32^3 / 64^3.  Result (profiles/experiments/r02_n1_knn_tiebreak_spraybottle.log): identical on bunny /
chair / knot / rocker; on SprayBottle.pc phi moves by 1.8e-4 relative L2 -> the kd-tree restatement stays.
    python tools/n1_tiebreak_probe.py [cloud ...]"""
import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/signed-heat-3d_b200')
import numpy as np, shm3d
from oracle import shm_oracle as o

def read_pc(path):
    P, N = [], []
    for ln in open(path):
        t = ln.split()
        if not t: continue
        if t[0] == 'v': P.append([float(x) for x in t[1:4]])
        elif t[0] == 'vn': N.append([float(x) for x in t[1:4]])
    return np.array(P), np.array(N)

L = shm3d.lib()
L.shm3d_debug_knn_mode.argtypes = [C.c_int32]
L.shm3d_debug_knn_mode.restype = None
for name in (sys.argv[1:] or ['bunny', 'chair', 'knot', 'rocker', 'SprayBottle']):
    P, N = read_pc(f'/root/reference/data/{name}.pc')
    L.shm3d_debug_knn_mode(0)
    t = time.time(); a0, h0, nt0 = shm3d.point_weights(P, N); t0 = time.time() - t
    L.shm3d_debug_knn_mode(1)
    t = time.time(); a1, h1, nt1 = shm3d.point_weights(P, N); t1 = time.time() - t
    L.shm3d_debug_knn_mode(0)
    da = np.abs(a1 - a0)
    line = f"{name:12s} nP={len(P):6d} h rel diff {abs(h1-h0)/h0:.2e}  areas: max rel {np.max(da/np.maximum(a0,1e-300)):.2e}, L1 rel {da.sum()/a0.sum():.2e}, changed {int((da>1e-12*a0.max()).sum())}, soup tris {nt0}/{nt1}, time {t0:.2f}/{t1:.2f}s"
    # effect on phi at 32^3 through the fp64 oracle (KKT LU)
    for hc in (1, 2):
        try:
            cen = P.mean(axis=0); rad = np.linalg.norm(P - cen, axis=1).max()
            kw = dict(hCoef=hc, scrub_nonfinite=False, step3="pcg", tol=1e-10)
            phi0 = o.compute_distance(P, N, a0, h0, cen, rad, **kw)
            phi1 = o.compute_distance(P, N, a1, h1, cen, rad, **kw)
            line += f"  phi@{16<<hc}^3 rel-L2 {np.linalg.norm(phi1-phi0)/np.linalg.norm(phi0):.2e}"
        except Exception as e:
            line += f"  phi: {e!r}"[:200]
    print(line, flush=True)
