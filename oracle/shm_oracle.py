"""
oracle/shm_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU fp64 restatement ("port") of the reference's grid-solver hot path
(SignedHeatGridSolver::computeDistance).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The product
path (signed-heat-3d_b200/) never does and fails loudly without its CUDA library.

PARITY STATUS: pinned against the reference's OWN SOURCE, not against reference-published vectors (the reference
has no tests or golden vectors for this path).  oracle/_ref/libshm_ref.so is src/signed_heat_grid_solver.cpp +
src/signed_heat_3d.cpp of the reference compiled unmodified against oracle/ref_shim (a stand-in for the slices of
geometry-central / Eigen / polyscope they use: the real Eigen is fetched at configure time and absent here, SURVEY.md
section 0 D7); tests/test_reference_build.py shows this restatement equal to it to ~1e-13 on phi (mesh, polygon,
point-cloud and fastIntegration paths), with the identical KKT matrix and right-hand side.  A second build links the
same two files with geometry-central's REAL sources (oracle/_ref/libshm_ref_gc.so; only Eigen and polyscope stubbed) and
gives the same fields, including the point-cloud overload end to end with geometry-central's own tufted-cover weights.
NOT pinned: Eigen itself (its SparseLU: both builds hand the assembled system to scipy SuperLU, the same solver used
here; its sparse assembly and vector arithmetic are restated in the stubs).  Step 3 is also solved by an fp64 projected CG, which must
agree with the LU (tests/test_oracle.py).

Nothing here reads /root/reference at run time except the helper readers when a test
explicitly passes such a path (CPU-only fixture generation, tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------------------
# C helper (Steps 1-2 O(N*M) loop) -- built by oracle/Makefile or on first use
# --------------------------------------------------------------------------------------
def build_clib(force: bool = False) -> str:
    out = os.path.join(_HERE, "_build", "libshm_oracle.so")
    src = os.path.join(_HERE, "csrc", "shm_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # -O3 like the reference's Release build (CMakeLists.txt:46); x86-64-v3 instead of -march=native because
        # the built .so travels to the GPU box, whose host CPU differs from this container's
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"]
        subprocess.check_call(cmd)
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_clib())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        _LIB.oracle_step12.argtypes = [ctypes.c_int] * 5 + [dp, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                                            dp, dp, dp, dp, ctypes.c_int]
        _LIB.oracle_step12_aswritten.argtypes = [ctypes.c_int] * 5 + [dp, ctypes.c_double, ctypes.c_double,
                                                                      ctypes.c_int64, dp, ip, dp, dp, dp, ctypes.c_int]
        _LIB.oracle_step12_box.argtypes = [ctypes.c_int] * 6 + [dp, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                                                dp, dp, dp, dp, ctypes.c_int]
        _LIB.oracle_apply_K.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double, dp, dp, ctypes.c_int]
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def max_threads() -> int:
    return int(_lib().oracle_max_threads())


# --------------------------------------------------------------------------------------
# Input readers (restating the reference's loaders)
# --------------------------------------------------------------------------------------
def read_obj(path):
    """OBJ reader following deps/geometry-central/src/surface/simple_polygon_mesh.cpp:167-232
    (v / f tokens, index before the first '/', 1-based) and meshio.cpp:22-29
    (stripUnusedVertices; OBJ vertices at identical positions are NOT merged).
    Returns (V float64[nV,3], faces list[list[int]])."""
    verts, faces = [], []
    with open(path, "rb") as fh:
        for raw in fh:
            line = raw.decode("ascii", "replace").split()
            if not line:
                continue
            if line[0] == "v":
                verts.append([float(line[1]), float(line[2]), float(line[3])])
            elif line[0] == "f":
                faces.append([int(tok.split("/")[0]) - 1 for tok in line[1:]])
    V = np.asarray(verts, dtype=np.float64)
    used = np.zeros(len(V), dtype=bool)
    for f in faces:
        used[f] = True
    remap = np.cumsum(used) - 1
    V = V[used]
    faces = [[int(remap[i]) for i in f] for f in faces]
    return V, faces


def read_pc(path):
    """.pc reader following src/main.cpp:196-225 ('v x y z' / 'vn x y z' lines)."""
    P, Nn = [], []
    with open(path, "r") as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                P.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "vn":
                Nn.append([float(t[1]), float(t[2]), float(t[3])])
    return np.asarray(P, dtype=np.float64), np.asarray(Nn, dtype=np.float64)


# --------------------------------------------------------------------------------------
# Source quantities
# --------------------------------------------------------------------------------------
def mesh_sources(V, faces):
    """Per-face area / unit normal / barycentre and the mesh scalars the grid solver uses.

    setFaceVectorAreas (src/signed_heat_3d.cpp:62-89: shoelace vector area, always taken
    because the triangular fast path has no return), barycenter
    (src/signed_heat_grid_solver.cpp:498-503), meanEdgeLength (src/signed_heat_3d.cpp:51-60;
    edges = unique unordered vertex pairs, surface_mesh.cpp:145,162-166), centroid / radius
    (src/signed_heat_3d.cpp:3-22)."""
    V = np.asarray(V, dtype=np.float64)
    tri = all(len(f) == 3 for f in faces)
    M = len(faces)
    if tri:
        F = np.asarray(faces, dtype=np.int64).reshape(M, 3)
        p = V[F]  # M,3,3
        Nvec = 0.5 * (np.cross(p[:, 0], p[:, 1]) + np.cross(p[:, 1], p[:, 2]) + np.cross(p[:, 2], p[:, 0]))
        bary = (p[:, 0] + p[:, 1] + p[:, 2]) / 3.0
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    else:
        Nvec = np.zeros((M, 3))
        bary = np.zeros((M, 3))
        es = []
        for fi, f in enumerate(faces):
            pf = V[f]
            n = np.zeros(3)
            for a in range(len(f)):
                n += np.cross(pf[a], pf[(a + 1) % len(f)])
                es.append((f[a], f[(a + 1) % len(f)]))
            Nvec[fi] = 0.5 * n
            c = np.zeros(3)
            for a in range(len(f)):
                c += pf[a]
            bary[fi] = c / len(f)
        e = np.asarray(es, dtype=np.int64)
    area = np.linalg.norm(Nvec, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = Nvec / area[:, None]
    e = np.sort(e, axis=1)
    e = np.unique(e, axis=0)
    h = np.linalg.norm(V[e[:, 0]] - V[e[:, 1]], axis=1).sum() / len(e)
    c = V.sum(axis=0) / len(V)
    r = np.sqrt(((V - c) ** 2).sum(axis=1)).max()
    return dict(pos=np.ascontiguousarray(bary), nrm=np.ascontiguousarray(nrm), area=np.ascontiguousarray(area),
                h=float(h), centroid=c, radius=float(r))


@dataclass
class Grid:
    nx: int
    ny: int
    nz: int
    bmin: np.ndarray
    cell: float

    @property
    def N(self):
        return self.nx * self.ny * self.nz


def make_grid(centroid, radius, hCoef=0.0, scale=2.0) -> Grid:
    """src/signed_heat_grid_solver.cpp:13-26: cube c +- scale*r, nx = (size_t)(2*2^(hCoef+3)),
    cell = 2s/(nx-1)."""
    s = radius * scale
    nx = int(2 * 2.0 ** (hCoef + 3))
    cell = 2.0 * s / (nx - 1)
    return Grid(nx, nx, nx, np.asarray(centroid, dtype=np.float64) - s, float(cell))


def lambda_from_h(h, tCoef=1.0):
    """src/signed_heat_grid_solver.cpp:42-44: shortTime = tCoef h^2, lambda = sqrt(1/shortTime)."""
    return float(np.sqrt(1.0 / (tCoef * h * h)))


# --------------------------------------------------------------------------------------
# Steps 1-2
# --------------------------------------------------------------------------------------
def step12(g: Grid, lam, pos, nrm, area, threads=None, k0=0, k1=None):
    """Y[3N] interleaved (src/signed_heat_grid_solver.cpp:48-65).  C loop, fp64."""
    if k1 is None:
        k1 = g.nz
    if threads is None:
        threads = max_threads()
    Y = np.zeros(3 * g.N, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    nrm = np.ascontiguousarray(nrm, dtype=np.float64)
    area = np.ascontiguousarray(area, dtype=np.float64)
    bmin = np.ascontiguousarray(g.bmin, dtype=np.float64)
    _lib().oracle_step12(g.nx, g.ny, g.nz, k0, k1, _dp(bmin), g.cell, lam, len(area), _dp(pos), _dp(nrm), _dp(area),
                         _dp(Y), int(threads))
    return Y


def step12_box(g: Grid, lam, pos, nrm, area, j0, j1, k0, k1, threads=None):
    """Steps 1-2 on rows [j0,j1) of planes [k0,k1) only -> Y[(k1-k0),(j1-j0),nx,3] (bounded CPU-baseline sample)."""
    if threads is None:
        threads = max_threads()
    Y = np.zeros((k1 - k0, j1 - j0, g.nx, 3), dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    nrm = np.ascontiguousarray(nrm, dtype=np.float64)
    area = np.ascontiguousarray(area, dtype=np.float64)
    bmin = np.ascontiguousarray(g.bmin, dtype=np.float64)
    _lib().oracle_step12_box(g.nx, g.ny, j0, j1, k0, k1, _dp(bmin), g.cell, lam, len(area), _dp(pos), _dp(nrm),
                             _dp(area), _dp(Y), int(threads))
    return Y


def step12_aswritten(g: Grid, lam, V, F, nrm, area, threads=1, k0=0, k1=None):
    if k1 is None:
        k1 = g.nz
    Y = np.zeros(3 * g.N, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.int32)
    nrm = np.ascontiguousarray(nrm, dtype=np.float64)
    area = np.ascontiguousarray(area, dtype=np.float64)
    bmin = np.ascontiguousarray(g.bmin, dtype=np.float64)
    _lib().oracle_step12_aswritten(g.nx, g.ny, g.nz, k0, k1, _dp(bmin), g.cell, lam, len(area), _dp(V),
                                   F.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _dp(nrm), _dp(area), _dp(Y),
                                   int(threads))
    return Y


# --------------------------------------------------------------------------------------
# Step 3 operators
# --------------------------------------------------------------------------------------
def gradient_matrix(g: Grid):
    """D (3N x N), src/signed_heat_grid_solver.cpp:336-402: forward differences, backward at
    the far face, all / cell.  Built as a scipy CSR exactly like the triplet list."""
    import scipy.sparse as sp
    nx, ny, nz = g.nx, g.ny, g.nz
    N = g.N
    idx = np.arange(N).reshape(nz, ny, nx)
    rows, cols, vals = [], [], []
    strides = (1, nx, nx * ny)
    dims = (nx, ny, nz)
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cur = (I + J * nx + K * nx * ny).ravel()
    coords = (I.ravel(), J.ravel(), K.ravel())
    for a in range(3):
        far = coords[a] == dims[a] - 1
        nxt = np.where(far, cur, cur + strides[a])
        c = np.where(far, cur - strides[a], cur)
        rows += [3 * cur + a, 3 * cur + a]
        cols += [nxt, c]
        vals += [np.ones(N), -np.ones(N)]
    D = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * N, N)).tocsr()
    del idx
    return D / g.cell


def laplacian_matrix(g: Grid):
    """L (N x N), src/signed_heat_grid_solver.cpp:278-334: 7-point, out-of-range neighbours
    redirected to the node itself (so the diagonal is -(#in-range neighbours)), / cell^2."""
    import scipy.sparse as sp
    nx, ny, nz = g.nx, g.ny, g.nz
    N = g.N
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cur = (I + J * nx + K * nx * ny).ravel()
    coords = (I.ravel(), J.ravel(), K.ravel())
    strides = (1, nx, nx * ny)
    dims = (nx, ny, nz)
    rows, cols, vals = [cur], [cur], [-6.0 * np.ones(N)]
    for a in range(3):
        nxt = np.where(coords[a] == dims[a] - 1, cur, cur + strides[a])
        prv = np.where(coords[a] == 0, cur, cur - strides[a])
        rows += [cur, cur]
        cols += [nxt, prv]
        vals += [np.ones(N), np.ones(N)]
    L = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
    return L / (g.cell * g.cell)


def div_rhs(g: Grid, Y, scrub_nonfinite=True):
    """b = D^T Y (src/signed_heat_grid_solver.cpp:70-74) in stencil form:
    per axis with gl = Y_a/cell along a line of n nodes:
      b[0] = -gl[0]; b[t] = gl[t-1]-gl[t] (1<=t<=n-3); b[n-2] = gl[n-3]-gl[n-2]-gl[n-1];
      b[n-1] = gl[n-2]+gl[n-1].
    Mesh overload zeroes non-finite entries (:72-74); point overload does not (:180)."""
    nx, ny, nz = g.nx, g.ny, g.nz
    Y3 = np.asarray(Y, dtype=np.float64).reshape(nz, ny, nx, 3)
    b = np.zeros((nz, ny, nx))
    for a, ax in ((0, 2), (1, 1), (2, 0)):
        gl = np.moveaxis(Y3[..., a], ax, 0) / g.cell
        n = gl.shape[0]
        ba = np.zeros_like(gl)
        ba[0] = -gl[0]
        ba[1:n - 1] = gl[0:n - 2] - gl[1:n - 1]
        ba[n - 2] -= gl[n - 1]
        ba[n - 1] = gl[n - 2] + gl[n - 1]
        b += np.moveaxis(ba, 0, ax)
    b = b.ravel()
    if scrub_nonfinite:
        b[~np.isfinite(b)] = 0.0
    return b


def apply_K(g: Grid, u, threads=None):
    """K u = -L u (matrix-free), C loop."""
    if threads is None:
        threads = max_threads()
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty_like(u)
    _lib().oracle_apply_K(g.nx, g.ny, g.nz, g.cell, _dp(u), _dp(out), int(threads))
    return out


def trilinear(g: Grid, q):
    """Cell indices, the 8 node indices {000,100,010,001,110,101,011,111} and weights for
    points q[n,3] (src/signed_heat_grid_solver.cpp:433-464)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    d = q - g.bmin
    ijk = np.floor(d / g.cell).astype(np.int64)
    p000 = g.bmin + ijk * g.cell
    t = (q - p000) / g.cell
    tx, ty, tz = t[:, 0], t[:, 1], t[:, 2]
    i, j, k = ijk[:, 0], ijk[:, 1], ijk[:, 2]
    nx, ny = g.nx, g.ny

    def nid(a, b, c):
        return (i + a) + (j + b) * nx + (k + c) * nx * ny

    idx = np.stack([nid(0, 0, 0), nid(1, 0, 0), nid(0, 1, 0), nid(0, 0, 1), nid(1, 1, 0), nid(1, 0, 1), nid(0, 1, 1),
                    nid(1, 1, 1)], axis=1)
    w = np.stack([(1 - tx) * (1 - ty) * (1 - tz), tx * (1 - ty) * (1 - tz), (1 - tx) * ty * (1 - tz),
                  (1 - tx) * (1 - ty) * tz, tx * ty * (1 - tz), tx * (1 - ty) * tz, (1 - tx) * ty * tz, tx * ty * tz],
                 axis=1)
    cell_id = i + j * nx + k * nx * ny
    return cell_id, idx, w


def constraints(g: Grid, pos):
    """Constraint rows (src/signed_heat_grid_solver.cpp:80-100): sources in input order, first
    source per grid cell wins.  Returns (src_index[m], node_idx[m,8], w[m,8])."""
    cell_id, idx, w = trilinear(g, pos)
    _, first = np.unique(cell_id, return_index=True)
    first = np.sort(first)  # row order = order of first occurrence
    return first, idx[first], w[first]


def constraint_matrix(g: Grid, idx, w):
    import scipy.sparse as sp
    m = idx.shape[0]
    rows = np.repeat(np.arange(m), 8)
    return sp.coo_matrix((w.ravel(), (rows, idx.ravel())), shape=(m, g.N)).tocsr()


def evaluate_function(g: Grid, u, q):
    """Trilinear interpolation (src/signed_heat_grid_solver.cpp:405-431)."""
    _, idx, w = trilinear(g, q)
    # the reference evaluates the nested-lerp form; algebraically identical to sum w_i u_i
    return (w * u[idx]).sum(axis=1)


def source_average(g: Grid, u, pos, area):
    """evaluateAverageAlongSourceGeometry (src/signed_heat_grid_solver.cpp:466-496)."""
    return float((area * evaluate_function(g, u, pos)).sum() / area.sum())


# --------------------------------------------------------------------------------------
# Step 3 solvers
# --------------------------------------------------------------------------------------
def solve_kkt_lu(g: Grid, b, idx, w):
    """The reference's Step 3 verbatim (src/signed_heat_grid_solver.cpp:101-108):
    [[L, A^T],[A, 0]] [x; mu] = [b; 0], phi = -x, by direct sparse LU."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    L = laplacian_matrix(g)
    A = constraint_matrix(g, idx, w)
    m = A.shape[0]
    LHS = sp.bmat([[L, A.T], [A, None if m == 0 else sp.csr_matrix((m, m))]], format="csc") if m > 0 else L.tocsc()
    rhs = np.concatenate([b, np.zeros(m)])
    sol = spla.splu(LHS).solve(rhs)
    return -sol[:g.N]


class Projector:
    """P = I - A^T (A A^T)^-1 A with a direct sparse factorisation of A A^T (fp64)."""

    def __init__(self, g: Grid, idx, w):
        import scipy.sparse.linalg as spla
        self.A = constraint_matrix(g, idx, w)
        self.m = self.A.shape[0]
        if self.m > 0:
            self.fac = spla.splu((self.A @ self.A.T).tocsc())

    def __call__(self, v):
        if self.m == 0:
            return v
        return v - self.A.T @ self.fac.solve(self.A @ v)


def solve_projected_cg(g: Grid, b, idx, w, tol=1e-12, maxit=20000, precond=None, callback=None):
    """min 1/2 phi^T K phi - b^T phi  s.t.  A phi = 0 -- the same system as solve_kkt_lu
     -- by CG on the null space of A (projected CG), fp64.
    precond: optional callable z = M^-1 r (symmetric PSD)."""
    P = Projector(g, idx, w)
    x = np.zeros(g.N)
    r = P(np.asarray(b, dtype=np.float64))
    z = P(precond(r)) if precond else r
    p = z.copy()
    rho = float(r @ z)
    rho0 = rho
    its = 0
    while its < maxit and rho > tol * tol * rho0 and rho > 0:
        q = apply_K(g, p)
        alpha = rho / float(p @ q)
        x += alpha * p
        r -= alpha * P(q)
        z = P(precond(r)) if precond else r
        rho_new = float(r @ z)
        p = z + (rho_new / rho) * p
        rho = rho_new
        its += 1
        if callback:
            callback(its, x, rho / rho0)
    return x, its


def integrate_greedily(g: Grid, Y):
    """--fast path (src/signed_heat_grid_solver.cpp:224-275): FIFO BFS from node (0,0,0),
    neighbour order -x,+x,-y,+y,-z,+z; phi[q] = phi[p] + normalize(Y_p+Y_q).(q-p)."""
    from collections import deque
    nx, ny, nz = g.nx, g.ny, g.nz
    Y3 = np.asarray(Y).reshape(-1, 3)
    phi = np.zeros(g.N)
    visited = np.zeros(g.N, dtype=bool)
    dq = deque([(0, 0, 0)])
    visited[0] = True
    dims = (nx, ny, nz)
    strides = (1, nx, nx * ny)
    while dq:
        cur = dq.popleft()
        ci = cur[0] + cur[1] * nx + cur[2] * nx * ny
        Yp = Y3[ci]
        for a in range(3):
            for sgn in (-1, 1):
                if (sgn < 0 and cur[a] > 0) or (sgn > 0 and cur[a] < dims[a] - 1):
                    ni = ci + sgn * strides[a]
                    if not visited[ni]:
                        Ya = Yp + Y3[ni]
                        Ya = Ya / np.linalg.norm(Ya)
                        phi[ni] = phi[ci] + Ya[a] * sgn * g.cell
                        visited[ni] = True
                        nxt = list(cur)
                        nxt[a] += sgn
                        dq.append(tuple(nxt))
    return phi


def integrate_greedily_prefix(g: Grid, Y):
    """The same field as integrate_greedily, computed without the queue: on the full box the FIFO order makes
    (i,j,k-1) the first visitor of (i,j,k) for k > 0, (i,j-1,0) for k = 0 < j and (i-1,0,0) on the x axis, so phi is
    three prefix sums (x on the line j=k=0, y on the plane k=0, z everywhere).  This is the form the CUDA kernels use;
    tests/test_oracle.py checks it against the literal BFS above."""
    nx, ny, nz = g.nx, g.ny, g.nz
    Y3 = np.asarray(Y, dtype=np.float64).reshape(nz, ny, nx, 3)
    phi = np.zeros((nz, ny, nx))

    def step(a, b, axis):
        s = a + b
        return s[..., axis] / np.linalg.norm(s, axis=-1) * g.cell

    phi[0, 0, 1:] = np.cumsum(step(Y3[0, 0, :-1], Y3[0, 0, 1:], 0))
    phi[0, 1:, :] = phi[0, 0, :][None, :] + np.cumsum(step(Y3[0, :-1], Y3[0, 1:], 1), axis=0)
    phi[1:] = phi[0][None] + np.cumsum(step(Y3[:-1], Y3[1:], 2), axis=0)
    return phi.ravel()


# --------------------------------------------------------------------------------------
# End-to-end restatement of computeDistance
# --------------------------------------------------------------------------------------
def compute_distance(pos, nrm, area, h, centroid, radius, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False,
                     scrub_nonfinite=True, step3="lu", tol=1e-12, threads=None, return_all=False):
    """computeDistance (src/signed_heat_grid_solver.cpp:5-114 mesh / :116-222 points) on flat
    source arrays.  step3: 'lu' (the reference's KKT sparse LU) or 'pcg' (fp64 projected CG)."""
    g = make_grid(centroid, radius, hCoef, scale)
    lam = lambda_from_h(h, tCoef)
    Y = step12(g, lam, pos, nrm, area, threads=threads)
    b = div_rhs(g, Y, scrub_nonfinite=scrub_nonfinite)
    its = 0
    if fast:
        phi = integrate_greedily(g, Y)
        src, idx, w = None, None, None
    else:
        src, idx, w = constraints(g, pos)
        if step3 == "lu":
            phi = solve_kkt_lu(g, b, idx, w)
        else:
            phi, its = solve_projected_cg(g, b, idx, w, tol=tol)
    phi = phi - source_average(g, phi, pos, area)
    if return_all:
        return dict(phi=phi, Y=Y, b=b, grid=g, lam=lam, m=0 if idx is None else len(idx), its=its)
    return phi


def compute_distance_mesh(V, faces, **kw):
    s = mesh_sources(V, faces)
    return compute_distance(s["pos"], s["nrm"], s["area"], s["h"], s["centroid"], s["radius"], **kw)


# ------------------------------------------------------------------------------------------------ row N3: isosurface
# The reference's downstream consumer (src/main.cpp:116-128 -> polyscope registerIsosurfaceAsMesh,
# deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228) narrows phi to float32 (volume_grid.ipp:103-106) and runs
# MC::marching_cube (deps/polyscope/deps/MarchingCubeCpp/include/MarchingCube/MC.h:242-315) on it.  Restated below in
# float32 arithmetic, cell by cell in the same traversal order, so that vertex numbering and triangle order coincide.

_MC_CASES = None


def mc_case_table():
    """case -> list of edge triples (oracle/mc_case_table.txt; provenance: tools/make_mc_table.py)."""
    global _MC_CASES
    if _MC_CASES is None:
        import os
        tab = [[] for _ in range(256)]
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_case_table.txt")) as f:
            for line in f:
                if line.startswith("#"):
                    continue
                c, rest = line.split(":")
                tab[int(c)] = [tuple(int(e) for e in t.split()) for t in rest.split(";") if t.strip()]
        _MC_CASES = tab
    return _MC_CASES


def marching_cubes(values, isoval, dims, bound_min=None, bound_max=None):
    """MC::marching_cube as polyscope calls it (MC.h:242-315), plus registerIsosurfaceAsMesh's swizzle/scale/translate
    when bounds are given.  values: index i + j*nx + k*nx*ny.  The library reads its field as (X*ny + Y)*nz + Z, i.e. its
    X is the grid's k, its Z the grid's i (volume_grid_scalar_quantity.cpp:222 swizzles back); traversal is Z outermost,
    X innermost.  Returns (vertices float32[nV,3], triangles uint32[nT,3])."""
    f32 = np.float32
    nx, ny, nz = (int(d) for d in dims)
    SX, SY, SZ = nz, ny, nx          # extents of the library's X, Y, Z (the reference always has nx = ny = nz)
    fld = np.ascontiguousarray(np.asarray(values).ravel(), dtype=f32).reshape(SX, SY, SZ)   # fld[X, Y, Z]
    vs = f32(-f32(isoval)) + fld                                                            # MC.h:258-265
    neg = vs < 0
    cfg = np.zeros((SX - 1, SY - 1, SZ - 1), dtype=np.int32)
    for b in range(8):               # corner b: X + (b & 1), Y + (b >> 1 & 1), Z + (b >> 2 & 1)   (MC.h:267-275)
        dx, dy, dz = b & 1, (b >> 1) & 1, (b >> 2) & 1
        cfg |= neg[dx:SX - 1 + dx, dy:SY - 1 + dy, dz:SZ - 1 + dz].astype(np.int32) << b
    act = np.argwhere((cfg != 0) & (cfg != 255))
    act = act[np.lexsort((act[:, 0], act[:, 1], act[:, 2]))]     # Z outermost, then Y, X innermost (MC.h:252-256)
    cases = mc_case_table()
    verts, tris, edge_vertex = [], [], {}

    def compute_edge(va, vb, axis, x, y, z):                      # MC.h:183-193
        if (va < 0) == (vb < 0):
            return
        v = [f32(x), f32(y), f32(z)]
        with np.errstate(invalid="ignore", over="ignore"):     # inf - inf etc. give NaN like the C arithmetic does
            v[axis] = f32(v[axis] + f32(va / f32(va - vb)))
        edge_vertex[(x, y, z, axis)] = len(verts)
        verts.append(v)

    for x, y, z in act.tolist():
        c = [vs[x + (b & 1), y + ((b >> 1) & 1), z + ((b >> 2) & 1)] for b in range(8)]
        if y == 0 and z == 0:                                     # MC.h:279-304
            compute_edge(c[0], c[1], 0, x, y, z)
        if z == 0:
            compute_edge(c[2], c[3], 0, x, y + 1, z)
        if y == 0:
            compute_edge(c[4], c[5], 0, x, y, z + 1)
        compute_edge(c[6], c[7], 0, x, y + 1, z + 1)
        if x == 0 and z == 0:
            compute_edge(c[0], c[2], 1, x, y, z)
        if z == 0:
            compute_edge(c[1], c[3], 1, x + 1, y, z)
        if x == 0:
            compute_edge(c[4], c[6], 1, x, y, z + 1)
        compute_edge(c[5], c[7], 1, x + 1, y, z + 1)
        if x == 0 and y == 0:
            compute_edge(c[0], c[4], 2, x, y, z)
        if y == 0:
            compute_edge(c[1], c[5], 2, x + 1, y, z)
        if x == 0:
            compute_edge(c[2], c[6], 2, x, y + 1, z)
        compute_edge(c[3], c[7], 2, x + 1, y + 1, z)
        edge_key = [(x, y, z, 0), (x, y + 1, z, 0), (x, y, z + 1, 0), (x, y + 1, z + 1, 0),       # MC.h:306-317
                    (x, y, z, 1), (x + 1, y, z, 1), (x, y, z + 1, 1), (x + 1, y, z + 1, 1),
                    (x, y, z, 2), (x + 1, y, z, 2), (x, y + 1, z, 2), (x + 1, y + 1, z, 2)]
        for t in cases[int(cfg[x, y, z])]:
            tris.append([edge_vertex[edge_key[e]] for e in t])
    V = np.asarray(verts, dtype=f32).reshape(-1, 3)
    T = np.asarray(tris, dtype=np.uint32).reshape(-1, 3)
    if bound_min is not None:        # volume_grid.ipp:72-76 and volume_grid_scalar_quantity.cpp:220-224, float32
        bmin = np.asarray(bound_min, dtype=f32)
        bmax = np.asarray(bound_max, dtype=f32)
        scale = (bmax - bmin) / np.array([nx - 1, ny - 1, nz - 1], dtype=f32)
        V = (V[:, ::-1] * scale + bmin).astype(f32)
    return V, T


def grid_bounds_f32(g: "Grid"):
    """The glm::vec3 bounds the reference registers its volume grid with (src/signed_heat_grid_solver.cpp:20-24,35):
    bboxMin / bboxMax narrowed to float."""
    bmin = np.asarray(g.bmin, dtype=np.float64)
    bmax = bmin + g.cell * (np.array([g.nx, g.ny, g.nz]) - 1)
    return bmin.astype(np.float32), bmax.astype(np.float32)
