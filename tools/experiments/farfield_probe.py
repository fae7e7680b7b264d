"""EXPERIMENT (round-2 preparation, CPU only; see farfield_probe.c): accuracy / work trade-off of replacing far source
clusters by equivalent sources in Steps 1-2.  Prints, per (eps, n_eq): pair-evaluation count relative to the culled exact
sum, max / rms angle between Y_exact and Y_approx, and (with --phi) the relative L2 change of phi after Step 3."""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import shm_oracle as o  # noqa: E402
from synth import fibonacci_sphere  # noqa: E402


def lib():
    import tempfile
    so = os.path.join(tempfile.gettempdir(), "libfarfield_probe.so")
    subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "farfield_probe.c"), "-lm"])
    L = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    L.farfield_eval.argtypes = [C.c_int64, dp, C.c_int64, ip, ip, dp, dp, dp, dp, dp, C.c_int, dp, dp, C.c_double, C.c_double,
                                C.c_double, dp, C.c_int, dp, C.c_int, dp, dp, C.c_double, dp, dp, dp, ip, ip]
    return L


def morton_order(pos, bits=10):
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    q = np.minimum(((pos - lo) / (hi - lo + 1e-300) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(len(pos), dtype=np.int64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(code, kind="stable")


def clusters(pos, w, area, size, rmax):
    """contiguous ranges of the Morton order: at most `size` sources, and every member within rmax of the first one
    (Morton order jumps across the surface now and then; an unbounded range would ruin the culling bound)"""
    M = len(pos)
    first = []
    f = 0
    while f < M:
        n = 1
        while n < size and f + n < M and np.linalg.norm(pos[f + n] - pos[f]) <= rmax:
            n += 1
        first.append(f)
        f += n
    first = np.asarray(first, dtype=np.int64)
    count = np.diff(np.append(first, M)).astype(np.int64)
    centre = np.zeros((len(first), 3))
    rho = np.zeros(len(first))
    for c, (f, n) in enumerate(zip(first, count)):
        p, a = pos[f:f + n], area[f:f + n]
        centre[c] = (a[:, None] * p).sum(axis=0) / a.sum()
        rho[c] = np.sqrt(((p - centre[c]) ** 2).sum(axis=1)).max()
    return first, count, centre, rho


def equivalents(pos, w, area, first, count, centre, n_eq):
    """n_eq = 1: the summed weight at the area-weighted centroid.  n_eq = 4: the cluster split in 4 along its two
    principal tangent directions (median cuts), each part as in n_eq = 1."""
    ncl = len(first)
    ep = np.zeros((ncl, n_eq, 3))
    ew = np.zeros((ncl, n_eq, 3))
    for c, (f, n) in enumerate(zip(first, count)):
        p, a, ww = pos[f:f + n], area[f:f + n], w[f:f + n]
        if n_eq == 1 or n < 4:
            parts = [np.arange(n)] + [np.arange(0)] * (n_eq - 1)
        else:
            u, s, vt = np.linalg.svd(p - centre[c], full_matrices=False)
            t0, t1 = (p - centre[c]) @ vt[0], (p - centre[c]) @ vt[1]
            h0 = t0 > np.median(t0)
            parts = []
            for side in (h0, ~h0):
                idx = np.nonzero(side)[0]
                m1 = t1[idx] > np.median(t1[idx])
                parts += [idx[m1], idx[~m1]]
        for e, idx in enumerate(parts):
            if len(idx) == 0:
                ep[c, e] = centre[c]
                continue
            ep[c, e] = (a[idx, None] * p[idx]).sum(axis=0) / a[idx].sum()
            ew[c, e] = ww[idx].sum(axis=0)
    return ep, ew


def equivalents_moment(pos, w, area, first, count, centre):
    """4 points c +- a_i e_i in the cluster's principal tangent directions that reproduce the cluster's moments up to
    order 2 (area measure) and the first moment of the vector weights: a_i = sqrt(2 mu_i) with mu_i the eigenvalues of
    sum A d d^T / A; weights W/4 +- B e_i / (2 a_i), B = sum w_s d_s^T."""
    ncl = len(first)
    ep = np.zeros((ncl, 4, 3))
    ew = np.zeros((ncl, 4, 3))
    for c, (f, n) in enumerate(zip(first, count)):
        d, a, ww = pos[f:f + n] - centre[c], area[f:f + n], w[f:f + n]
        W = ww.sum(axis=0)
        M2 = (a[:, None, None] * d[:, :, None] * d[:, None, :]).sum(axis=0) / a.sum()
        mu, E = np.linalg.eigh(M2)
        B = ww.T @ d                                        # 3x3: sum_s w_s d_s^T
        k = 0
        for i in (2, 1):                                    # the two largest eigenvalues
            e = E[:, i]
            ai = np.sqrt(2 * max(mu[i], 0.0))
            tilt = B @ e / (2 * ai) if ai > 1e-12 else np.zeros(3)
            ep[c, k], ew[c, k] = centre[c] + ai * e, W / 4 + tilt
            ep[c, k + 1], ew[c, k + 1] = centre[c] - ai * e, W / 4 - tilt
            k += 2
    return ep, ew


def fib_dirs(n):
    i = np.arange(n) + 0.5
    ph = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)], axis=1)


def admissible_distance(pos, w, area, first, count, centre, rho, ep, ew, lam, tol, ndir=42):
    """Per cluster: the smallest gap g (distance beyond the bounding sphere) from which on the equivalent sources
    reproduce the cluster's own contribution to within tol * (sum of magnitudes) in every sampled direction.
    Gaps are sampled geometrically from 0.25/lam to 64/lam; inf if never."""
    U = fib_dirs(ndir)
    gaps = (0.25 / lam) * 2.0 ** np.arange(0, 9.01, 0.5)
    out = np.full(len(first), np.inf)
    for c, (f, n) in enumerate(zip(first, count)):
        p, ww, a = pos[f:f + n], w[f:f + n], area[f:f + n]
        x = centre[c] + (rho[c] + gaps)[:, None, None] * U[None, :, :]            # [gap, dir, 3]
        r = np.linalg.norm(x[:, :, None, :] - p[None, None, :, :], axis=-1)       # [gap, dir, src]
        rmin = r.min(axis=-1, keepdims=True)
        k = np.exp(-lam * (r - rmin)) / r
        Xe = (k[..., None] * ww).sum(axis=2)
        mag = (k * a).sum(axis=2)
        re = np.linalg.norm(x[:, :, None, :] - ep[c][None, None, :, :], axis=-1)
        ke = np.exp(-lam * (re - rmin)) / re
        Xq = (ke[..., None] * ew[c]).sum(axis=2)
        err = (np.linalg.norm(Xq - Xe, axis=-1) / mag).max(axis=1)                # per gap: worst direction
        ok = err <= tol
        bad = np.nonzero(~ok)[0]
        if len(bad) == 0:
            out[c] = gaps[0]
        elif bad[-1] + 1 < len(gaps):
            out[c] = gaps[bad[-1] + 1]
    return out


def error_table(pos, w, area, first, count, centre, rho, ep, ew, lam, na=7, naz=8):
    """Per cluster: relative error (w.r.t. the cluster's own sum of magnitudes) of its equivalent sources at gaps
    0.25/lam .. 128/lam beyond the bounding sphere and angles 0..90 degrees from the cluster's mean normal (max over
    azimuth and over the two sides).  Returns (gaps[ng], normals[ncl,3], E[ncl,ng,na])."""
    gaps = (0.25 / lam) * 2.0 ** np.arange(0, 9.51, 0.5)
    th = np.linspace(0, np.pi / 2, na)
    az = np.arange(naz) * 2 * np.pi / naz
    ncl = len(first)
    E = np.zeros((ncl, len(gaps), na))
    normals = np.zeros((ncl, 3))
    for c, (f, n) in enumerate(zip(first, count)):
        p, ww, a = pos[f:f + n], w[f:f + n], area[f:f + n]
        nc = ww.sum(axis=0)
        nc = nc / max(np.linalg.norm(nc), 1e-300)
        normals[c] = nc
        t1 = np.cross(nc, [1.0, 0, 0] if abs(nc[0]) < 0.9 else [0, 1.0, 0])
        t1 /= np.linalg.norm(t1)
        t2 = np.cross(nc, t1)
        U = np.stack([sgn * np.cos(t) * nc + np.sin(t) * (np.cos(z) * t1 + np.sin(z) * t2)
                      for t in th for z in az for sgn in (1.0, -1.0)])                  # [na*naz*2, 3]
        x = centre[c] + (rho[c] + gaps)[:, None, None] * U[None, :, :]
        r = np.linalg.norm(x[:, :, None, :] - p[None, None, :, :], axis=-1)
        rmin = r.min(axis=-1, keepdims=True)
        k = np.exp(-lam * (r - rmin)) / r
        Xe = (k[..., None] * ww).sum(axis=2)
        mag = (k * a).sum(axis=2)
        re = np.linalg.norm(x[:, :, None, :] - ep[c][None, None, :, :], axis=-1)
        ke = np.exp(-lam * (re - rmin)) / re
        Xq = (ke[..., None] * ew[c]).sum(axis=2)
        err = np.linalg.norm(Xq - Xe, axis=-1) / mag                                    # [gap, dir]
        E[c] = err.reshape(len(gaps), na, naz * 2).max(axis=2)
    return np.ascontiguousarray(gaps), np.ascontiguousarray(normals), np.ascontiguousarray(E)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=100000)
    ap.add_argument("--hcoef", type=int, default=1)
    ap.add_argument("--cluster", type=int, default=32)
    ap.add_argument("--tau", type=float, default=10.0)
    ap.add_argument("--rmax-h", type=float, default=8.0, help="cluster extent limit in units of the mean edge length")
    ap.add_argument("--phi", action="store_true")
    ap.add_argument("--cone", type=float, default=0.0, help="flatness gate: clusters whose member normals deviate more than this (rad) from the mean normal are never approximated (0 = no gate)")
    ap.add_argument("--area-ratio", type=float, default=0.0)
    ap.add_argument("--min-count", type=int, default=0)
    ap.add_argument("--mesh", default=None, help="name of a tests/golden/*.npz mesh fixture instead of the sphere")
    a = ap.parse_args()
    if a.mesh:
        z = np.load(os.path.join(ROOT, "tests", "golden", a.mesh + ".npz"))
        if "F" in z:
            V, F = z["V"], z["F"].tolist()
        else:
            fo, fv = z["face_offsets"], z["face_vertices"]
            V, F = z["V"], [fv[fo[i]:fo[i + 1]].tolist() for i in range(len(fo) - 1)]
    else:
        V, F = fibonacci_sphere(a.tris)
    s = o.mesh_sources(V, F)
    order = morton_order(s["pos"])
    pos, nrm, area = s["pos"][order], s["nrm"][order], s["area"][order]
    w = nrm * area[:, None]
    lam = o.lambda_from_h(s["h"])
    g = o.make_grid(s["centroid"], s["radius"], hCoef=a.hcoef)
    ii, jj, kk = np.meshgrid(np.arange(g.nx), np.arange(g.ny), np.arange(g.nz), indexing="ij")
    nodes = np.ascontiguousarray((g.bmin + g.cell * np.stack([ii, jj, kk], axis=-1).transpose(2, 1, 0, 3).reshape(-1, 3)))
    print(f"M = {len(pos)}, h = {s['h']:.4f}, lambda = {lam:.2f}, grid {g.nx}^3, cell {g.cell:.4f}, cluster {a.cluster}")
    first, count, centre, rho = clusters(pos, w, area, a.cluster, a.rmax_h * s['h'])
    print(f"clusters {len(first)}, rho mean {rho.mean():.4f} max {rho.max():.4f}, lam*rho^2 mean {lam * (rho ** 2).mean():.3f}")
    L = lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    cone = np.zeros(len(first))
    for c, (f, n) in enumerate(zip(first, count)):
        nc = w[f:f + n].sum(axis=0)
        nc /= max(np.linalg.norm(nc), 1e-300)
        cone[c] = np.arccos(np.clip(nrm[f:f + n] @ nc, -1, 1)).max()
    print(f"normal cone per cluster: median {np.median(cone):.3f} rad, 90% {np.quantile(cone, 0.9):.3f}, max {cone.max():.3f}")
    rho_gate = rho.copy()
    if a.area_ratio > 0:
        ar = np.array([area[f:f + n].max() / area[f:f + n].min() for f, n in zip(first, count)])
        rho_gate[ar > a.area_ratio] = 1e30
        print(f"area-uniformity gate {a.area_ratio}: {np.mean(ar > a.area_ratio):.2f} of the clusters excluded (median ratio {np.median(ar):.1f})")
    if a.min_count > 0:
        rho_gate[count < a.min_count] = 1e30
    if a.cone > 0:
        rho_gate[cone > a.cone] = 1e30          # lam*rho^2/g < eps never holds
        print(f"flatness gate {a.cone}: {np.mean(cone > a.cone):.2f} of the clusters excluded")

    def run(n_eq, eps, ep=None, ew=None, gadm=None, table=None, tol=0.0, gamma=None, want_S=False):
        X = np.zeros((len(nodes), 3))
        S = np.zeros(len(nodes)) if want_S else None
        pe, pq = C.c_int64(), C.c_int64()
        z = np.zeros(3)
        t = time.time()
        L.farfield_eval(len(nodes), nodes.ctypes.data_as(dp), len(first), first.ctypes.data_as(ip), count.ctypes.data_as(ip),
                        centre.ctypes.data_as(dp), rho.ctypes.data_as(dp), rho_gate.ctypes.data_as(dp), pos.ctypes.data_as(dp), w.ctypes.data_as(dp),
                        n_eq, (ep if ep is not None else z).ctypes.data_as(dp), (ew if ew is not None else z).ctypes.data_as(dp),
                        lam, a.tau, eps, None if gadm is None else gadm.ctypes.data_as(dp),
                        0 if table is None else len(table[0]), None if table is None else table[0].ctypes.data_as(dp),
                        0 if table is None else table[2].shape[2], None if table is None else table[1].ctypes.data_as(dp),
                        None if table is None else table[2].ctypes.data_as(dp), tol,
                        None if gamma is None else gamma.ctypes.data_as(dp), None if S is None else S.ctypes.data_as(dp),
                        X.ctypes.data_as(dp), C.byref(pe),
                        C.byref(pq))
        if want_S:
            return X, pe.value, pq.value, time.time() - t, S
        return X, pe.value, pq.value, time.time() - t

    X0, pe0, _, dt = run(0, 0.0)
    Y0 = X0 / np.linalg.norm(X0, axis=1, keepdims=True)
    print(f"exact (tau = {a.tau}): {pe0 / len(nodes):.0f} pairs/node ({pe0 / len(nodes) / len(pos):.3f} of brute force), {dt:.1f} s")
    phi0 = None
    if a.phi:
        idx, wts = o.constraints(g, pos)[1:]
        b0 = o.div_rhs(g, Y0.reshape(-1), scrub_nonfinite=True)
        phi0 = o.solve_projected_cg(g, b0, idx, wts, tol=1e-10)[0]
    for n_eq, mode in ((1, "centroid"), (4, "median-split"), (4, "moments")):
        print(mode)
        ep, ew = (equivalents_moment(pos, w, area, first, count, centre) if mode == "moments" else
                  equivalents(pos, w, area, first, count, centre, n_eq))
        ep, ew = np.ascontiguousarray(ep), np.ascontiguousarray(ew)
        if mode == "moments":
            t = time.time()
            table = error_table(pos, w, area, first, count, centre, rho, ep, ew, lam)
            print(f"error table: {time.time() - t:.1f} s; median error face-on at gap 8/lam: "
                  f"{np.median(table[2][:, 10, 0]):.2e}, at 45 deg: {np.median(table[2][:, 10, 3]):.2e}, at 90 deg: {np.median(table[2][:, 10, 6]):.2e}")
            # cancellation ratio |X| / sum of magnitudes per node, estimated from the all-equivalent evaluation (cheap)
            Xc, pec, pqc, _, Sc = run(n_eq, 1e30, ep, ew, want_S=True)
            gamma = np.ascontiguousarray(np.linalg.norm(Xc, axis=1) / Sc)
            print(f"cancellation ratio |X|/S: min {gamma.min():.2e} median {np.median(gamma):.2e}; estimate pass work {(pec + pqc) / pe0:.3f}")
            for tol in (1e-2, 3e-3, 1e-3, 3e-4, 1e-4):
                X, pe, pq, dt = run(n_eq, 0.0, ep, ew, None, table, tol, gamma)
                Y = X / np.linalg.norm(X, axis=1, keepdims=True)
                ang = np.linalg.norm(Y - Y0, axis=1)
                msg = (f"node-level tol {tol:.0e}: work {(pe + pq) / pe0:.3f} (exact {pe / pe0:.3f} + equiv {pq / pe0:.3f})  "
                       f"|dY| max {ang.max():.2e} rms {np.sqrt((ang ** 2).mean()):.2e}")
                if a.phi:
                    b = o.div_rhs(g, Y.reshape(-1), scrub_nonfinite=True)
                    phi = o.solve_projected_cg(g, b, idx, wts, tol=1e-10)[0]
                    msg += f"  phi rel-L2 {np.linalg.norm(phi - phi0) / np.linalg.norm(phi0):.2e}"
                print(msg, flush=True)
            for tol in ():
                t = time.time()
                gadm = admissible_distance(pos, w, area, first, count, centre, rho, ep, ew, lam, tol)
                X, pe, pq, dt = run(n_eq, 0.0, ep, ew, gadm)
                Y = X / np.linalg.norm(X, axis=1, keepdims=True)
                ang = np.linalg.norm(Y - Y0, axis=1)
                msg = (f"tabulated tol {tol:.0e}: never-admissible clusters {np.isinf(gadm).mean():.2f}, median g_adm*lam "
                       f"{np.median(gadm[np.isfinite(gadm)]) * lam if np.isfinite(gadm).any() else np.inf:.1f}; work {(pe + pq) / pe0:.3f}  |dY| max {ang.max():.2e} rms "
                       f"{np.sqrt((ang ** 2).mean()):.2e}  (table {time.time() - t - dt:.1f} s)")
                if a.phi:
                    b = o.div_rhs(g, Y.reshape(-1), scrub_nonfinite=True)
                    phi = o.solve_projected_cg(g, b, idx, wts, tol=1e-10)[0]
                    msg += f"  phi rel-L2 {np.linalg.norm(phi - phi0) / np.linalg.norm(phi0):.2e}"
                print(msg, flush=True)
        for eps in ((0.3, 1.0, 3.0) if mode == "moments" else (1.0,)):
            X, pe, pq, dt = run(n_eq, eps, ep, ew)
            Y = X / np.linalg.norm(X, axis=1, keepdims=True)
            ang = np.linalg.norm(Y - Y0, axis=1)
            msg = (f"n_eq {n_eq} eps {eps:4.1f}: work {(pe + pq) / pe0:.3f} (exact {pe / pe0:.3f} + equiv {pq / pe0:.3f})  "
                   f"|dY| max {ang.max():.2e} rms {np.sqrt((ang ** 2).mean()):.2e}")
            if a.phi:
                b = o.div_rhs(g, Y.reshape(-1), scrub_nonfinite=True)
                phi = o.solve_projected_cg(g, b, idx, wts, tol=1e-10)[0]
                msg += f"  phi rel-L2 {np.linalg.norm(phi - phi0) / np.linalg.norm(phi0):.2e}"
            print(msg, flush=True)


if __name__ == "__main__":
    main()
