"""
oracle/shm_oracle_large.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The fp64 oracle made affordable at the BASELINE.json grid sizes (256^3 / 512^3) on a handful of host cores: a thin
driver over oracle/csrc/shm_oracle_large.c.  Same algorithm as oracle/shm_oracle.py (which restates the reference,
src/signed_heat_grid_solver.cpp) -- tests/test_oracle_large.py pins every function here to its plain counterpart there:

  step12_bricks     == shm_oracle.step12        (Steps 1-2, :48-65 / :157-174; provably negligible far terms skipped)
  div_rhs           == shm_oracle.div_rhs       (b = D^T Y, scrub, :70-74 / :336-402)
  solve_projected_cg == shm_oracle.solve_projected_cg (the KKT system of :101-108 solved in the null space of A)

Only tests/golden/make_golden_baseline.py (fixture generator, run in the build container) and the tests import it.
Needs AVX-512 (the container's Xeon); `available()` says whether this machine can run it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import shm_oracle as o

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


def _lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(_HERE, "_build", "libshm_oracle_large.so")
        src = os.path.join(_HERE, "csrc", "shm_oracle_large.c")
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            # no -ffast-math: subnormals must behave like the reference's doubles (the X.norm() underflow artefact)
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", out, src,
                                   "-lmvec", "-lm"])
        L = ctypes.CDLL(out)
        dp, i64p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
        ci, cd = ctypes.c_int, ctypes.c_double
        L.oracle_step12_bricks.argtypes = [ci] * 5 + [dp, cd, cd] + [dp] * 6 + [ci, i64p, dp, dp, dp, cd, cd, dp, dp, ci]
        L.oracle_cg_apply_dot.argtypes = [ci, ci, ci, cd, dp, dp, ci]
        L.oracle_cg_apply_dot.restype = cd
        L.oracle_cg_update.argtypes = [ctypes.c_size_t, cd, dp, dp, dp, dp, ci]
        L.oracle_cg_update.restype = cd
        L.oracle_cg_direction.argtypes = [ctypes.c_size_t, cd, dp, dp, ci]
        L.oracle_div_rhs.argtypes = [ci, ci, ci, cd, dp, ci, dp, ci]
        L.oracle_div_rhs.restype = ctypes.c_long
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _morton(q):
    """30-bit Morton code of integer triples in [0, 1024)."""
    def spread(v):
        v = v.astype(np.uint64) & 0x3FF
        v = (v | (v << 16)) & 0x30000FF
        v = (v | (v << 8)) & 0x300F00F
        v = (v | (v << 4)) & 0x30C30C3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def cluster_sources(pos, nrm, area, size=48):
    """Morton-sorted sources cut into runs of `size`: structure-of-arrays + bounding spheres + total |area| per run."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    q = np.floor((pos - lo) / np.maximum(hi - lo, 1e-300) * 1023.999).astype(np.int64)
    order = np.argsort(_morton(q), kind="stable")
    P = pos[order]
    W = (np.asarray(nrm, dtype=np.float64) * np.asarray(area, dtype=np.float64)[:, None])[order]
    A = np.abs(np.asarray(area, dtype=np.float64))[order]
    M = len(P)
    beg = np.arange(0, M + size, size, dtype=np.int64)
    beg[-1] = M
    if len(beg) >= 2 and beg[-2] >= M:
        beg = beg[:-1]
    nc = len(beg) - 1
    cen = np.add.reduceat(P, beg[:-1], axis=0) / np.diff(beg)[:, None]
    cid = np.repeat(np.arange(nc), np.diff(beg))
    d = np.linalg.norm(P - cen[cid], axis=1)
    rad = np.maximum.reduceat(d, beg[:-1]) * (1 + 1e-12) + 1e-300
    mass = np.add.reduceat(A, beg[:-1])
    cols = [np.ascontiguousarray(P[:, a]) for a in range(3)] + [np.ascontiguousarray(W[:, a]) for a in range(3)]
    return dict(cols=cols, beg=beg, cen=np.ascontiguousarray(cen), rad=np.ascontiguousarray(rad),
                mass=np.ascontiguousarray(mass), nc=nc)


def step12_bricks(g: o.Grid, lam, pos, nrm, area, tau=44.0, eps=1e-13, threads=None, k0=0, k1=None, cluster=48):
    """Y (N x 3, double) like shm_oracle.step12, and stats = dict(pairs, bricks_redone, worst_skipped_ratio)."""
    k1 = g.nz if k1 is None else k1
    cs = cluster_sources(pos, nrm, area, cluster)
    Y = np.empty(3 * g.nx * g.ny * (k1 - k0))
    st = np.zeros(3)
    bmin = np.ascontiguousarray(g.bmin, dtype=np.float64)
    _lib().oracle_step12_bricks(g.nx, g.ny, g.nz, k0, k1, _dp(bmin), g.cell, lam, *[_dp(c) for c in cs["cols"]],
                                cs["nc"], cs["beg"].ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(cs["cen"]),
                                _dp(cs["rad"]), _dp(cs["mass"]), tau, eps, _dp(Y), _dp(st),
                                threads or o.max_threads())
    return Y.reshape(-1, 3), dict(pairs=st[0], bricks_redone=int(st[1]), worst_skipped_ratio=st[2])


def div_rhs(g: o.Grid, Y, scrub_nonfinite=True, threads=None):
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1)
    b = np.empty(g.N)
    bad = _lib().oracle_div_rhs(g.nx, g.ny, g.nz, g.cell, _dp(Y), 1 if scrub_nonfinite else 0, _dp(b),
                                threads or o.max_threads())
    return b, int(bad)


def solve_projected_cg(g: o.Grid, b, idx, w, tol=1e-10, maxit=40000, threads=None, callback=None):
    """Same iteration as shm_oracle.solve_projected_cg (unpreconditioned CG in null A, fp64); the dense parts run in
    the fused OpenMP kernels, the projector correction touches only the 8m constrained nodes."""
    import scipy.sparse.linalg as spla
    L = _lib()
    T = threads or o.max_threads()
    A = o.constraint_matrix(g, idx, w)
    m = A.shape[0]
    fac = spla.splu((A @ A.T).tocsc()) if m else None
    nodes = np.unique(np.asarray(idx).reshape(-1)) if m else np.zeros(0, dtype=np.int64)
    As = A[:, nodes].tocsr() if m else None   # columns that carry entries
    AsT = As.T.tocsr() if m else None

    def project_inplace(v):
        if m:
            v[nodes] -= AsT @ fac.solve(As @ v[nodes])

    n = g.N
    x = np.zeros(n)
    r = np.array(b, dtype=np.float64)
    project_inplace(r)
    p = r.copy()
    q = np.empty(n)
    rho = float(r @ r)
    rho0 = rho
    its = 0
    while its < maxit and rho > tol * tol * rho0 and rho > 0:
        pq = L.oracle_cg_apply_dot(g.nx, g.ny, g.nz, g.cell, _dp(p), _dp(q), T)
        # r <- r - alpha P q: project q first (sparse), then the fused dense update; p.q is unchanged (p in null A)
        project_inplace(q)
        alpha = rho / pq
        rho_new = L.oracle_cg_update(n, alpha, _dp(p), _dp(q), _dp(x), _dp(r), T)
        L.oracle_cg_direction(n, rho_new / rho, _dp(r), _dp(p), T)
        rho = rho_new
        its += 1
        if callback:
            callback(its, x, rho / rho0)
    return x, its
