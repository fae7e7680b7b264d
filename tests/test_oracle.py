"""CPU tests of the oracle itself: known answers (re-derived in tests/golden), KKT-LU == projected
CG, operator identities from the reference's matrix definitions."""
import numpy as np
import pytest

from conftest import icosphere, load_golden
from oracle import shm_oracle as o

# (input, hCoef) -> nx, cell, m, phi min, phi max, ||phi||_2 : SURVEY.md Appendix B known answers
KNOWN = {
    ("bunny_small", 0): (16, 0.3963979558, 79, -0.1879134733, 4.6074565907, 142.7527526488),
    ("bunny_small", 1): (32, 0.1918054625, 316, -0.4558871669, 4.5379467811, 395.4522338177),
    ("knot", 1): (32, 7.0169801820, 969, -4.4043545122, 146.9276958461, 11951.1634037721),
}


@pytest.mark.parametrize("name,hc", list(KNOWN))
def test_golden_matches_survey_known_answers(name, hc):
    z, _ = load_golden(name)
    nx, cell, m, lo, hi, l2 = KNOWN[(name, hc)]
    t = f"h{hc}"
    assert int(z[t + "_nx"]) == nx and int(z[t + "_m"]) == m
    assert abs(float(z[t + "_cell"]) - cell) < 1e-9
    st = z[t + "_phi_stats"]
    assert abs(st[0] - lo) < 1e-8 and abs(st[1] - hi) < 1e-8 and abs(st[2] - l2) < 1e-6


def test_oracle_reproduces_golden_bunny16():
    z, F = load_golden("bunny_small")
    phi = o.compute_distance_mesh(z["V"], F, hCoef=0)
    assert np.abs(phi - z["h0_phi"]).max() < 1e-10


def test_mesh_scalars_bunny():
    z, F = load_golden("bunny_small")
    s = o.mesh_sources(z["V"], F)
    assert len(z["V"]) == 1430 and len(F) == 2856  # stripUnusedVertices (meshio.cpp:22-29)
    assert abs(s["h"] - 0.0950010) < 1e-6 and abs(s["radius"] - 1.48649) < 1e-5
    assert abs(o.lambda_from_h(s["h"]) - 10.5262) < 1e-4
    # outward orientation: generalized signed distance is positive outside
    g = o.make_grid(s["centroid"], s["radius"], 0)
    assert g.nx == 16


def test_lu_equals_projected_cg():
    z, F = load_golden("bunny_small")
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], 0)
    b = z["h0_b"]
    _, idx, w = o.constraints(g, s["pos"])
    lu = o.solve_kkt_lu(g, b, idx, w)
    cg, its = o.solve_projected_cg(g, b, idx, w, tol=1e-13)
    assert its > 10
    assert np.linalg.norm(lu - cg) / np.linalg.norm(lu) < 1e-9
    # the constraints hold: A phi = 0
    A = o.constraint_matrix(g, idx, w)
    assert np.abs(A @ lu).max() < 1e-9


def test_operators_match_matrix_definitions():
    V, F = icosphere(1)
    s = o.mesh_sources(V, F)
    g = o.make_grid(s["centroid"], s["radius"], 0)
    rng = np.random.default_rng(1)
    Y = rng.standard_normal(3 * g.N)
    D = o.gradient_matrix(g)
    assert np.abs(D.T @ Y - o.div_rhs(g, Y)).max() < 1e-12  # stencil form of D^T Y
    L = o.laplacian_matrix(g)
    u = rng.standard_normal(g.N)
    assert np.abs(-(L @ u) - o.apply_K(g, u)).max() < 1e-10
    # L is symmetric, negative semi-definite with the constants in its null space
    assert abs(L - L.T).max() < 1e-12 and np.abs(L @ np.ones(g.N)).max() < 1e-9
    diag = L.diagonal() * g.cell ** 2
    assert diag.min() == -6 and diag.max() == -3
    # trilinear rows sum to one and reproduce linear functions
    q = rng.uniform(-0.5, 0.5, size=(50, 3))
    _, idx, w = o.trilinear(g, q)
    assert np.abs(w.sum(axis=1) - 1).max() < 1e-12
    I, J, K = np.meshgrid(np.arange(g.nx), np.arange(g.ny), np.arange(g.nz), indexing="ij")
    lin = np.zeros(g.N)
    lin[(I + J * g.nx + K * g.nx * g.ny).ravel()] = (2 * I - 3 * J + 0.5 * K).ravel()
    exact = ((q - g.bmin) / g.cell) @ np.array([2, -3, 0.5])
    assert np.abs(o.evaluate_function(g, lin, q) - exact).max() < 1e-9


def test_sphere_distance_is_signed_distance():
    """physics check: for a sphere the generalized signed distance is ~ |x| - R near the surface."""
    V, F = icosphere(2)
    r = o.compute_distance_mesh(V, F, hCoef=1, return_all=False)
    s = o.mesh_sources(V, F)
    g = o.make_grid(s["centroid"], s["radius"], 1)
    I, J, K = np.meshgrid(np.arange(g.nx), np.arange(g.ny), np.arange(g.nz), indexing="ij")
    X = g.bmin + g.cell * np.stack([I, J, K], -1).reshape(-1, 3)
    d = np.zeros(g.N)
    d[(I + J * g.nx + K * g.nx * g.ny).ravel()] = np.linalg.norm(X - s["centroid"], axis=1) - 1.0
    band = np.abs(d) < 0.5
    assert np.abs(r - d)[band].max() < 0.08
    assert r[d < -0.3].max() < 0 and r[d > 0.3].min() > 0


def test_fast_integration_runs():
    V, F = icosphere(1)
    r = o.compute_distance_mesh(V, F, hCoef=0, fast=True)
    assert np.isfinite(r).all()


def test_greedy_bfs_equals_prefix_sums():
    """The reference's FIFO breadth-first integration (src/signed_heat_grid_solver.cpp:224-275) visits every node first
    from (i,j,k-1) / (i,j-1,0) / (i-1,0,0): the closed form the GPU path uses must reproduce the literal BFS."""
    rng = np.random.default_rng(3)
    for dims in [(8, 9, 7), (5, 4, 11), (16, 16, 16)]:
        g = o.Grid(dims[0], dims[1], dims[2], np.zeros(3), 0.37)
        Y = rng.standard_normal((g.N, 3))
        Y /= np.linalg.norm(Y, axis=1, keepdims=True)
        a = o.integrate_greedily(g, Y.ravel())
        b = o.integrate_greedily_prefix(g, Y.ravel())
        assert np.abs(a - b).max() < 1e-12
