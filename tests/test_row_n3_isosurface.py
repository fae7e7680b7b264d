"""Row N3: the consumer of phi -- isosurface extraction identical to the reference's
(polyscope registerIsosurfaceAsMesh = the vendored MarchingCube/MC.h + a vertex transform) and plane slices.

CPU: the numpy restatement against the reference's own library (oracle/_ref/libshm_mc_ref.so, compiled from the
vendored header) and the golden fixture; the product's case table against the header and against its intrinsic
properties; the DEVICE LOGIC (isosurface_core.h, the file the kernels are built from) run thread-by-thread on the host
(tests/csrc/mc_emulate.cpp) over the kernels' own launch geometry and scan chunking, against the oracle.
GPU: the kernels through the C ABI against the oracle -- bit-identical vertices, numbering and triangle order."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import shm3d
from conftest import GOLDEN, ROOT, icosphere, load_golden
from oracle import reference_build as rb
from oracle import shm_oracle as o

CSRC = os.path.join(ROOT, "signed-heat-3d_b200", "csrc")


def sphere_field(n, seed, noise=0.05):
    """phi[k,j,i] (flattened: i + j*n + k*n*n) of a squashed off-centre sphere plus noise"""
    rng = np.random.default_rng(seed)
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt((X * 1.1) ** 2 + (Y + 0.1) ** 2 + (Z * 0.9) ** 2).transpose(2, 1, 0) - 0.6
    return (r + noise * rng.standard_normal((n, n, n))).ravel()


CASES = [("noise12", 12, 0.1), ("sphere16", 16, 0.0), ("sphere33", 33, 0.07), ("sphere24neg", 24, -0.2)]
BMIN, BMAX = (-1.0, -1.5, -1.0), (1.0, 1.2, 1.7)


def case_field(name, n):
    if name.startswith("noise"):
        return np.random.default_rng(5).standard_normal(n ** 3)
    return sphere_field(n, seed=n)


def parse_table_inc():
    words = []
    with open(os.path.join(CSRC, "mc_table.inc")) as f:
        for line in f:
            if line.lstrip().startswith("//"):
                continue
            words += [int(w.strip().rstrip("ul"), 16) for w in line.split(",") if w.strip()]
    return np.asarray(words, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("name,n,iso", CASES)
def test_oracle_equals_the_reference_library(name, n, iso):
    if not (rb.build() and rb.mc_available()):
        pytest.skip("no oracle/_ref/libshm_mc_ref.so")
    phi = case_field(name, n)
    for bounds in ((BMIN, BMAX), None):
        Vr, Tr = rb.isosurface(phi, iso, (n, n, n), BMIN, BMAX, world=bounds is not None)
        Vo, To = o.marching_cubes(phi, iso, (n, n, n), *(bounds or (None, None)))
        assert len(Vr) > 100 and np.array_equal(Vr, Vo) and np.array_equal(Tr, To)     # bit for bit, same order


def test_oracle_equals_the_golden_isosurface_of_the_reference():
    z, _ = load_golden("bunny_small")
    gi = np.load(os.path.join(GOLDEN, "iso_bunny_small_h1.npz"))
    nx = int(gi["nx"])
    for tag in ("iso0", "iso1"):
        V, T = o.marching_cubes(z["h1_phi"], float(gi[tag + "_isoval"]), (nx, nx, nx), gi["bound_min"], gi["bound_max"])
        assert np.array_equal(V, gi[tag + "_V"]) and np.array_equal(T, gi[tag + "_T"])
    g = o.Grid(nx, nx, nx, z["h1_bmin"], float(z["h1_cell"]))
    bmin, bmax = o.grid_bounds_f32(g)
    assert np.array_equal(bmin, gi["bound_min"]) and np.array_equal(bmax, gi["bound_max"])


def test_case_table_copies_agree_with_the_reference_header():
    tab = parse_table_inc()
    assert tab.shape == (256,)
    cases = o.mc_case_table()
    for c in range(256):
        w = int(tab[c])
        t = w & 0xF
        assert [tuple((w >> (4 + 4 * (3 * i + a))) & 0xF for a in range(3)) for i in range(t)] == cases[c]
        assert w >> (4 + 12 * t) == 0
    if rb.build() and rb.mc_available():
        assert np.array_equal(tab, rb.mc_table())


def test_case_table_intrinsic_properties():
    """No reference needed: every triangle corner lies on an edge whose ends differ in sign, complementary cases use
    the same edges, and the resulting surface of a closed level set is a closed, consistently oriented 2-manifold."""
    ends = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
    cases = o.mc_case_table()
    assert cases[0] == [] and cases[255] == []
    for c in range(256):
        crossing = {e for e, (a, b) in enumerate(ends) if ((c >> a) ^ (c >> b)) & 1}
        used = {e for t in cases[c] for e in t}
        assert used == crossing and len(cases[c]) <= 5
        assert all(len(set(t)) == 3 for t in cases[c])
        assert {e for t in cases[255 - c] for e in t} == crossing
    n = 20
    V, T = o.marching_cubes(sphere_field(n, 0, noise=0.0), 0.0, (n, n, n))
    half = {}
    for a, b, c in T.tolist():
        for u, v in ((a, b), (b, c), (c, a)):
            assert (u, v) not in half
            half[(u, v)] = 1
    assert all((v, u) in half for (u, v) in half)                 # every edge twice, in opposite directions
    assert len(V) - len(half) // 2 + len(T) == 2                  # Euler characteristic of a sphere


# ------------------------------------------------------------------------------- CPU: the device logic, on the host
@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("mc") / "libmc_emulate.so")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-I" + CSRC, "-o", out,
                           os.path.join(ROOT, "tests", "csrc", "mc_emulate.cpp")])
    L = C.CDLL(out)
    fp, up, ip = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
    L.mc_emulate.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp, fp, C.c_int64, up, C.c_int64, ip, ip]

    def run(phi, iso, dims, bmin=None, bmax=None):
        v = np.ascontiguousarray(np.asarray(phi).ravel(), dtype=np.float32)
        bm = None if bmin is None else np.asarray(bmin, dtype=np.float32)
        bM = None if bmax is None else np.asarray(bmax, dtype=np.float32)
        head = (v.ctypes.data_as(fp), dims[0], dims[1], dims[2], np.float32(iso), None if bm is None else bm.ctypes.data_as(fp),
                None if bM is None else bM.ctypes.data_as(fp))
        nv, nt = C.c_int64(), C.c_int64()
        assert L.mc_emulate(*head, None, 0, None, 0, C.byref(nv), C.byref(nt)) == 0
        V = np.empty((nv.value, 3), dtype=np.float32)
        T = np.empty((nt.value, 3), dtype=np.uint32)
        assert L.mc_emulate(*head, V.ctypes.data_as(fp), nv.value, T.ctypes.data_as(up), nt.value, C.byref(nv), C.byref(nt)) == 0
        return V, T
    return run


@pytest.mark.parametrize("name,n,iso", CASES)
def test_device_logic_on_the_host_equals_oracle(emulator, name, n, iso):
    phi = case_field(name, n)
    for bounds in ((BMIN, BMAX), (None, None)):
        Vo, To = o.marching_cubes(phi, iso, (n, n, n), *bounds)
        Ve, Te = emulator(phi, iso, (n, n, n), *bounds)
        assert np.array_equal(Vo, Ve) and np.array_equal(To, Te)


def test_device_logic_non_cubic_and_degenerate(emulator):
    rng = np.random.default_rng(11)
    for dims in ((9, 14, 6), (2, 2, 2), (2, 17, 3), (31, 2, 5), (70, 33, 41), (34, 35, 3)):          # (nx, ny, nz)
        phi = rng.standard_normal(dims[0] * dims[1] * dims[2])
        Vo, To = o.marching_cubes(phi, 0.0, dims, BMIN, BMAX)
        Ve, Te = emulator(phi, 0.0, dims, BMIN, BMAX)
        assert np.array_equal(Vo, Ve) and np.array_equal(To, Te)
    phi = np.ones(6 ** 3)                                                   # no crossing at all
    Ve, Te = emulator(phi, 0.0, (6, 6, 6))
    assert Ve.shape == (0, 3) and Te.shape == (0, 3)
    phi = np.zeros(6 ** 3)                                                  # exactly on the level: 0 < 0 is false -> empty
    assert emulator(phi, 0.0, (6, 6, 6))[0].shape == (0, 3)
    phi = rng.standard_normal(7 ** 3)
    phi[::5] = np.nan                                                       # NaN compares like a non-negative value
    Vo, To = o.marching_cubes(phi, 0.0, (7, 7, 7))
    Ve, Te = emulator(phi, 0.0, (7, 7, 7))
    assert np.array_equal(To, Te) and np.array_equal(Vo, Ve, equal_nan=True)


def test_device_logic_fuzz_against_the_reference_library(emulator):
    """Seeded fuzz: small random grids with many exact ties (integer-valued fields, value == isoval, +-0), huge and tiny
    magnitudes, infinities -- the device logic must agree with the reference's library bit for bit on all of them."""
    if not (rb.build() and rb.mc_available()):
        pytest.skip("no oracle/_ref/libshm_mc_ref.so")
    rng = np.random.default_rng(2024)
    for trial in range(60):
        dims = tuple(int(d) for d in rng.integers(2, 9, size=3))
        n = dims[0] * dims[1] * dims[2]
        kind = trial % 6
        if kind == 0:
            phi = rng.integers(-2, 3, size=n).astype(np.float64)            # ties with isoval 0 / 1
        elif kind == 1:
            phi = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30)
        elif kind == 2:
            phi = np.where(rng.random(n) < 0.3, 0.0, rng.standard_normal(n)) * np.where(rng.random(n) < 0.5, 1.0, -1.0)  # +-0
        elif kind == 3:
            phi = rng.standard_normal(n)
            phi[rng.integers(0, n, size=max(1, n // 10))] = np.inf
        elif kind == 4:
            phi = np.round(rng.standard_normal(n) * 4) / 4                   # quarter steps: ties with isoval 0.25
        else:
            phi = rng.standard_normal(n).astype(np.float32).astype(np.float64)
        iso = float(rng.choice([0.0, 1.0, 0.25, -0.5, 1e-3]))
        if dims[0] == dims[1] == dims[2]:                                    # the reference library assumes a cube (nx = ny = nz)
            Vr, Tr = rb.isosurface(phi, iso, dims, BMIN, BMAX)
        else:
            Vr, Tr = o.marching_cubes(phi, iso, dims, BMIN, BMAX)
        Ve, Te = emulator(phi, iso, dims, BMIN, BMAX)
        assert np.array_equal(Tr, Te), (trial, dims, iso)
        assert np.array_equal(Vr, Ve, equal_nan=True), (trial, dims, iso)


def test_device_logic_equals_golden_isosurface_of_the_reference(emulator):
    z, _ = load_golden("bunny_small")
    gi = np.load(os.path.join(GOLDEN, "iso_bunny_small_h1.npz"))
    nx = int(gi["nx"])
    for tag in ("iso0", "iso1"):
        V, T = emulator(z["h1_phi"], float(gi[tag + "_isoval"]), (nx, nx, nx), gi["bound_min"], gi["bound_max"])
        assert np.array_equal(V, gi[tag + "_V"]) and np.array_equal(T, gi[tag + "_T"])


# ---------------------------------------------------------------------------------------------------- GPU: kernels
def _params(dims, bmin=(0.0, 0.0, 0.0), cell=1.0):
    p = shm3d.Params()
    p.nx, p.ny, p.nz = dims
    for a in range(3):
        p.bbox_min[a] = bmin[a]
    p.cell = cell
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,iso", CASES)
def test_gpu_isosurface_equals_oracle(gpu_ctx, name, n, iso):
    phi = case_field(name, n)
    p = _params((n, n, n))
    Vo, To = o.marching_cubes(phi, iso, (n, n, n), BMIN, BMAX)
    V, T, st = gpu_ctx.isosurface(p, phi, iso, BMIN, BMAX)                                   # host doubles, as PHI is
    assert st.n_vertices == len(Vo) and st.n_triangles == len(To) and st.gpu_launches >= 4
    assert np.array_equal(V, Vo) and np.array_equal(T, To)                                   # bit for bit, same order
    V2, T2, _ = gpu_ctx.isosurface(p, phi.astype(np.float32), iso, BMIN, BMAX)               # host float32
    assert np.array_equal(V2, Vo) and np.array_equal(T2, To)
    Vl, Tl, _ = gpu_ctx.isosurface(p, phi, iso, lattice=True)
    Vlo, Tlo = o.marching_cubes(phi, iso, (n, n, n))
    assert np.array_equal(Vl, Vlo) and np.array_equal(Tl, Tlo)


@pytest.mark.gpu
def test_gpu_isosurface_from_a_device_field_and_non_cubic(gpu_ctx):
    import torch
    rng = np.random.default_rng(11)
    for dims in ((9, 14, 6), (2, 2, 2), (70, 33, 41)):
        phi = rng.standard_normal(dims[0] * dims[1] * dims[2])
        d = torch.from_numpy(phi.astype(np.float32)).cuda()
        torch.cuda.synchronize()
        p = _params(dims, bmin=(-1.0, 0.5, 2.0), cell=0.25)
        V, T, _ = gpu_ctx.isosurface(p, d.data_ptr(), 0.3)                                   # bounds derived from p
        g = o.Grid(dims[0], dims[1], dims[2], np.array([-1.0, 0.5, 2.0]), 0.25)
        Vo, To = o.marching_cubes(phi, 0.3, dims, *o.grid_bounds_f32(g))
        assert np.array_equal(V, Vo) and np.array_equal(T, To)
    p = _params((6, 6, 6))
    V, T, st = gpu_ctx.isosurface(p, np.ones(216), 0.0)                                      # empty level set
    assert V.shape == (0, 3) and T.shape == (0, 3) and st.n_vertices == 0


@pytest.mark.gpu
def test_gpu_contour_of_a_solved_field(gpu_ctx):
    """The reference's flow (src/main.cpp:90-99 then :116-128): solve, then contour PHI -- here with phi staying in HBM."""
    import torch
    z, F = load_golden("bunny_small")
    solver = shm3d.SignedHeatGridSolver(context=gpu_ctx)
    phi = np.array(solver.computeDistance(z["V"], F, shm3d.SignedHeat3DOptions(hCoef=1)))
    V, T = solver.isosurface(phi, 0.0)
    p = solver.params
    g = o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)
    Vo, To = o.marching_cubes(phi, 0.0, (p.nx, p.ny, p.nz), *o.grid_bounds_f32(g))
    assert np.array_equal(V, Vo) and np.array_equal(T, To)
    # the level set of the GPU's phi is the reference's level set up to the solve tolerance
    gi = np.load(os.path.join(GOLDEN, "iso_bunny_small_h1.npz"))
    assert abs(len(V) - len(gi["iso0_V"])) <= 0.05 * len(gi["iso0_V"])
    from scipy.spatial import cKDTree
    d, _ = cKDTree(gi["iso0_V"]).query(V)
    assert d.max() < 0.5 * p.cell
    # device-resident hand-off: shm3d_solve_device leaves float32 phi in HBM; contouring it there gives the same mesh
    pr, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=1)
    dpos, dnrm, darea = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (pos, nrm, area))
    dphi = torch.empty(pr.N, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.solve_device(pr, dpos.data_ptr(), dnrm.data_ptr(), darea.data_ptr(), dphi.data_ptr(), len(area))
    Vd, Td, st = gpu_ctx.isosurface(pr, dphi.data_ptr(), 0.0)
    Vh, Th = o.marching_cubes(dphi.cpu().numpy(), 0.0, (pr.nx, pr.ny, pr.nz), *o.grid_bounds_f32(g))
    assert np.array_equal(Vd, Vh) and np.array_equal(Td, Th) and st.gpu_launches == 5  # count, scan (2), vertices, triangles


@pytest.mark.gpu
def test_gpu_isosurface_256_sphere_properties(gpu_ctx):
    """Size-independent properties at a size the Python oracle does not loop over: closed, consistently oriented
    2-manifold with Euler characteristic 2, vertices on grid edges within the interpolation error of the sphere, and
    -- where the prebuilt reference library travelled along -- equality with it."""
    n = 256
    g = np.linspace(-1, 1, n, dtype=np.float32)
    r2 = (g[None, None, :] ** 2 + g[None, :, None] ** 2 + g[:, None, None] ** 2)
    phi = (np.sqrt(r2) - np.float32(0.7)).astype(np.float32).ravel()
    p = _params((n, n, n), bmin=(-1.0, -1.0, -1.0), cell=2.0 / (n - 1))
    V, T, st = gpu_ctx.isosurface(p, phi, 0.0, (-1, -1, -1), (1, 1, 1))
    assert len(T) == 2 * len(V) - 4 and len(V) > 100000
    e = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * len(V) + e[:, 1]
    assert len(np.unique(key)) == len(key)                                   # no directed edge twice
    assert np.array_equal(np.sort(key), np.sort(e[:, 1] * len(V) + e[:, 0]))  # each has its opposite
    rad = np.linalg.norm(V.astype(np.float64), axis=1)
    assert np.abs(rad - 0.7).max() < (2.0 / (n - 1)) ** 2                     # linear interpolation error ~ cell^2 / (8 r)
    if rb.mc_available():
        Vr, Tr = rb.isosurface(phi, 0.0, (n, n, n), (-1, -1, -1), (1, 1, 1))
        assert np.array_equal(V, Vr) and np.array_equal(T, Tr)


@pytest.mark.gpu
def test_gpu_slice_equals_evaluate_function(gpu_ctx):
    n = 24
    phi = sphere_field(n, 3)
    bmin, cell = np.array([-1.0, -1.0, -1.0]), 2.0 / (n - 1)
    p = _params((n, n, n), bmin=bmin, cell=cell)
    origin, du, dv = np.array([-1.1, -0.9, 0.13]), np.array([0.031, 0.002, 0.001]), np.array([-0.001, 0.029, 0.004])
    S = gpu_ctx.slice(p, phi, origin, du, dv, 70, 60)
    g = o.Grid(n, n, n, bmin, cell)
    phi32 = phi.astype(np.float32).astype(np.float64)
    a, b = np.meshgrid(np.arange(70), np.arange(60))
    q = origin + a[..., None] * du + b[..., None] * dv
    idx = np.floor((q - bmin) / cell)
    inside = ((idx >= 0) & (idx < n - 1)).all(axis=-1)
    assert inside.any() and (~inside).any()
    assert np.isnan(S[~inside]).all() and np.isfinite(S[inside]).all()
    ref = o.evaluate_function(g, phi32, q[inside])
    assert np.abs(S[inside] - ref).max() < 1e-6
