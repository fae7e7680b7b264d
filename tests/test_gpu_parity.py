"""GPU parity tests: the CUDA path (through the C ABI) against the fp64 oracle and the committed golden fixtures.

Tolerances: north_star asks for <= 1e-4 relative L2 on the distance field; the tests assert that bar on phi and
tighter, fp32-level bars on the intermediate fields (Y: 3e-5 absolute on unit vectors; rhs: 1e-6 relative)."""
import numpy as np
import pytest

import shm3d
from conftest import icosphere, load_golden
from oracle import shm_oracle as o

pytestmark = pytest.mark.gpu
PHI_TOL = 1e-4  # north_star parity bar (relative L2 on the distance field)


def rel(a, b):
    return np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b)


def oracle_grid(p):
    return o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)


CASES = [("bunny_small", 0), ("bunny_small", 1), ("knot", 1), ("polygon-bear", 0)]


@pytest.mark.parametrize("name,hc", CASES)
@pytest.mark.parametrize("tau", [float("inf"), 0.0])  # brute force, and 0 = the library's default (tau = 10)
def test_step12_matches_golden(gpu_ctx, name, hc, tau):
    z, F = load_golden(name)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    p.cull_tau = tau
    Y, st = gpu_ctx.step12(p, pos, nrm, area)
    Yg = z[f"h{hc}_Y"].reshape(p.N, 3).T
    assert np.isfinite(Y).all()
    assert np.abs(Y - Yg).max() < 3e-5
    assert np.abs(np.linalg.norm(Y, axis=0) - 1).max() < 1e-5
    assert st.pairs_bruteforce == p.N * len(area)
    if np.isinf(tau):
        assert st.pairs_evaluated >= st.pairs_bruteforce  # partial tiles may add padding nodes
    else:
        assert 0 < st.pairs_evaluated <= st.pairs_bruteforce


@pytest.mark.parametrize("name,hc", CASES)
def test_rhs_matches_oracle(gpu_ctx, name, hc):
    z, F = load_golden(name)
    p, *_ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    Yg = z[f"h{hc}_Y"].reshape(p.N, 3)
    b = gpu_ctx.rhs(p, Yg.T.astype(np.float32))
    bo = o.div_rhs(oracle_grid(p), Yg.astype(np.float64).ravel()) * p.cell ** 2
    assert rel(b, bo) < 1e-6


@pytest.mark.parametrize("name,hc", CASES)
@pytest.mark.parametrize("mg", [True, False])
def test_step3_matches_kkt_lu(gpu_ctx, name, hc, mg):
    """Step 3 alone, from the oracle's right-hand side, against the reference's own formulation (sparse LU of the
    KKT matrix) -- with the constrained multigrid preconditioner and as plain projected CG."""
    z, F = load_golden(name)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    if not mg:
        p.flags |= shm3d.FLAG_NO_MG
    b = (z[f"h{hc}_b"] * p.cell ** 2).astype(np.float32)
    phi, st = gpu_ctx.step3(p, pos, area, b)
    assert rel(phi, z[f"h{hc}_phi"]) < PHI_TOL / 10
    assert st.m_constraints == int(z[f"h{hc}_m"])
    assert 0 < st.cg_iters < (60 if mg else 600)


@pytest.mark.parametrize("name,hc", CASES)
def test_solve_end_to_end_matches_golden(gpu_ctx, name, hc):
    z, F = load_golden(name)
    solver = shm3d.SignedHeatGridSolver(context=gpu_ctx)
    phi = solver.computeDistance(z["V"], F, shm3d.SignedHeat3DOptions(hCoef=hc))
    assert phi.dtype == np.float64 and phi.shape == (solver.params.N,)
    assert rel(phi, z[f"h{hc}_phi"]) < PHI_TOL
    st = z[f"h{hc}_phi_stats"]
    assert abs(phi.min() - st[0]) < 1e-3 * abs(st[1]) and abs(phi.max() - st[1]) < 1e-3 * abs(st[1])
    assert solver.stats.kernel_launches > 0


def test_culling_error_is_far_below_parity_bar(gpu_ctx):
    z, F = load_golden("bunny_small")
    p, pos, nrm, area, _ = shm3d.prepare_mesh(z["V"], F, hCoef=2)  # 64^3
    p.cull_tau = float("inf")
    phi_bf, st_bf = gpu_ctx.solve(p, pos, nrm, area)
    p.cull_tau = 0.0  # the shipped default (tau = 10)
    phi_c, st_c = gpu_ctx.solve(p, pos, nrm, area)
    assert st_c.pairs_evaluated < st_bf.pairs_evaluated  # coarse mesh (lambda*r_obj ~ 16): little to cull
    assert rel(phi_c, phi_bf) < 1e-5
    # survey known answer for bunny_small hCoef=2: min / max / L2
    assert abs(phi_bf.min() + 0.5525524688) < 2e-3 and abs(phi_bf.max() - 4.5160730887) < 2e-3
    assert abs(np.linalg.norm(phi_bf) / 1122.5031640345 - 1) < PHI_TOL


def test_step3_vs_oracle_projected_cg_64(gpu_ctx):
    """A size the direct LU no longer reaches comfortably: compare with the oracle's fp64 projected CG."""
    V, F = icosphere(3)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=2)
    g = oracle_grid(p)
    s = o.mesh_sources(V, F)
    lam = o.lambda_from_h(s["h"])
    Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"])
    b = o.div_rhs(g, Y)
    _, idx, w = o.constraints(g, s["pos"])
    ref, _ = o.solve_projected_cg(g, b, idx, w, tol=1e-11)
    ref = ref - o.source_average(g, ref, s["pos"], s["area"])
    phi, st = gpu_ctx.solve(p, pos, nrm, area)
    assert rel(phi, ref) < PHI_TOL
    # the constraints hold on the GPU result: interpolated phi at the pinned sources, before the shift, is ~0
    A = o.constraint_matrix(g, idx, w)
    assert np.abs(A @ (phi + st.shift)).max() < 1e-4 * np.abs(phi).max()


def test_point_cloud_overload(gpu_ctx):
    """Point overload: caller-supplied areas and h, no non-finite scrub (src/signed_heat_grid_solver.cpp:116-222)."""
    V, F = icosphere(2)
    s = o.mesh_sources(V, F)
    P, Nn = s["pos"], s["nrm"]
    areas = s["area"] * 1.3
    h = 0.2
    solver = shm3d.SignedHeatGridSolver(context=gpu_ctx)
    phi = solver.computeDistancePoints(P, Nn, areas, h, shm3d.SignedHeat3DOptions(hCoef=1))
    c = P.sum(axis=0) / len(P)
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    ref = o.compute_distance(P, Nn, areas, h, c, r, hCoef=1, scrub_nonfinite=False)
    assert rel(phi, ref) < PHI_TOL
    assert not (solver.params.flags & shm3d.FLAG_SCRUB_NONFINITE)


def test_large_grid_properties_sphere_128(gpu_ctx):
    """BASELINE-size style check through size-independent properties (no oracle run): finite, constraint and
    shift identities, sign, |grad phi| ~ 1 near the surface, similarity invariance, determinism."""
    V, F = icosphere(5)  # 20480 triangles
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=3)  # 128^3
    phi, st = gpu_ctx.solve(p, pos, nrm, area)
    assert np.isfinite(phi).all() and st.cg_iters < 120
    g = oracle_grid(p)
    # zero weighted mean on the source geometry (the shift)
    assert abs(o.source_average(g, phi, pos, area)) < 1e-5
    # distance-like: |x| - 1 within a fraction of a cell near the surface, right sign everywhere
    I, J, K = np.meshgrid(np.arange(p.nx), np.arange(p.ny), np.arange(p.nz), indexing="ij")
    X = g.bmin + g.cell * np.stack([I, J, K], -1).reshape(-1, 3)
    d = np.empty(p.N)
    d[(I + J * p.nx + K * p.nx * p.ny).ravel()] = np.linalg.norm(X, axis=1) - 1.0
    band = np.abs(d) < 0.25
    assert np.abs(phi - d)[band].max() < 0.02
    assert (phi[d > 0.05] > 0).all() and (phi[d < -0.05] < 0).all()
    # determinism: same inputs -> bit-identical output
    phi2, _ = gpu_ctx.solve(p, pos, nrm, area)
    assert np.array_equal(phi, phi2)
    # similarity: scale by 3 and translate -> phi scales by 3
    p3, pos3, nrm3, area3, _ = shm3d.prepare_mesh(V * 3.0 + np.array([5.0, -2.0, 1.0]), F, hCoef=3)
    phi3, _ = gpu_ctx.solve(p3, pos3, nrm3, area3)
    assert rel(phi3 / 3.0, phi) < PHI_TOL


def test_non_power_of_two_grid_uses_plain_projected_cg(gpu_ctx):
    V, F = icosphere(2)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=0)
    # shift the lattice off the sphere centre: with cell = 0.2 a node would sit exactly at the centre, where X cancels
    # by symmetry (|X| ~ 1e-15 of its terms) and the direction Y = X/|X| is rounding noise in ANY arithmetic
    for a in range(3):
        p.bbox_min[a] -= 0.031 * (a + 1)
    p.nx, p.ny, p.nz = 18, 21, 17           # odd sizes: no multigrid hierarchy
    p.cell = 4.0 / 20
    g = oracle_grid(p)
    s = o.mesh_sources(V, F)
    Y = o.step12(g, p.lambda_, s["pos"], s["nrm"], s["area"])
    b = o.div_rhs(g, Y)
    _, idx, w = o.constraints(g, s["pos"])
    ref = o.solve_kkt_lu(g, b, idx, w)
    ref = ref - o.source_average(g, ref, s["pos"], s["area"])
    phi, st = gpu_ctx.solve(p, pos, nrm, area)
    assert rel(phi, ref) < PHI_TOL, (rel(phi, ref), st.cg_iters, st.cg_rel_residual)


@pytest.mark.parametrize("name,hc", [("bunny_small", 0), ("bunny_small", 1), ("polygon-bear", 0)])
def test_fast_integration_matches_oracle_bfs(gpu_ctx, name, hc):
    """SignedHeat3DOptions.fastIntegration: greedy BFS integration (src/signed_heat_grid_solver.cpp:77-78, :224-275)."""
    z, F = load_golden(name)
    solver = shm3d.SignedHeatGridSolver(context=gpu_ctx)
    phi = solver.computeDistance(z["V"], F, shm3d.SignedHeat3DOptions(hCoef=hc, fastIntegration=True))
    ref = o.compute_distance_mesh(z["V"], F, hCoef=hc, fast=True)
    assert np.isfinite(phi).all() and solver.stats.cg_iters == 0
    assert rel(phi, ref) < PHI_TOL


def test_step12_at_arbitrary_query_points(gpu_ctx):
    """Row N4: the same Steps 1-2 sum at arbitrary points (the tet solver's barycentre queries,
    src/signed_heat_tet_solver.cpp:54-72) against a direct fp64 evaluation; lambda*r up to ~600 must not underflow."""
    z, F = load_golden("bunny_small")
    s = o.mesh_sources(z["V"], F)
    rng = np.random.default_rng(5)
    for lam in (o.lambda_from_h(s["h"]), 120.0):
        Q = rng.uniform(-1, 1, size=(777, 3)) * 2.0 * s["radius"] + s["centroid"]
        Y = gpu_ctx.step12_points(lam, s["pos"], s["nrm"], s["area"], Q)
        d = Q[:, None, :] - s["pos"][None, :, :]
        r = np.linalg.norm(d, axis=2)
        w = s["area"][None, :] * np.exp(-lam * (r - r.min(axis=1, keepdims=True))) / r   # shifted: same direction
        X = (w[:, :, None] * s["nrm"][None, :, :]).sum(axis=1)
        Yref = X / np.linalg.norm(X, axis=1, keepdims=True)
        assert np.isfinite(Y).all()
        assert np.abs(Y - Yref).max() < 3e-5


# ---------------------------------------------------------------- error behaviour
def test_nonfinite_source_is_rejected(gpu_ctx):
    V, F = icosphere(1)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=0)
    bad = nrm.copy()
    bad[5, 1] = np.nan
    with pytest.raises(shm3d.Shm3dError) as e:
        gpu_ctx.solve(p, pos, bad, area)
    assert e.value.code == shm3d.ERR_NONFINITE


def test_node_on_source_scrubbed_for_mesh_and_error_for_points(gpu_ctx):
    """A source exactly on a grid node gives r = 0 -> Inf/NaN in Y like the reference.  The mesh overload
    zeroes the affected rhs entries (:72-74); the point overload does not and geometry-central's solveSquare throws."""
    V, F = icosphere(2)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=0)
    node = np.array(p.bbox_min) + p.cell * np.array([5, 6, 7])
    pos2 = pos.copy()
    pos2[0] = node
    Y, _ = gpu_ctx.step12(p, pos2, nrm, area)
    idx = 5 + 6 * p.nx + 7 * p.nx * p.ny
    assert not np.isfinite(Y[:, idx]).all()
    phi, _ = gpu_ctx.solve(p, pos2, nrm, area)  # scrub flag set by prepare_mesh
    assert np.isfinite(phi).all()
    p.flags &= ~shm3d.FLAG_SCRUB_NONFINITE
    with pytest.raises(shm3d.Shm3dError) as e:
        gpu_ctx.solve(p, pos2, nrm, area)
    assert e.value.code == shm3d.ERR_NONFINITE


def test_invalid_arguments(gpu_ctx):
    V, F = icosphere(1)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=0)
    with pytest.raises(shm3d.Shm3dError) as e:
        gpu_ctx.solve(p, pos[:0], nrm[:0], area[:0])
    assert e.value.code == shm3d.ERR_INVALID_ARG
    far = pos.copy()
    far[0] += 50.0
    with pytest.raises(shm3d.Shm3dError) as e:
        gpu_ctx.solve(p, far, nrm, area)
    assert e.value.code == shm3d.ERR_INVALID_ARG
    q = shm3d.Params.from_buffer_copy(p)
    q.cell = -1.0
    with pytest.raises(shm3d.Shm3dError):
        gpu_ctx.solve(q, pos, nrm, area)
    # the context stays usable after errors
    phi, _ = gpu_ctx.solve(p, pos, nrm, area)
    assert np.isfinite(phi).all()
