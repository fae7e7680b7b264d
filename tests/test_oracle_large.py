"""The large-grid fp64 helpers (oracle/shm_oracle_large.py, used to generate the BASELINE-size fixtures) are the same
algorithm as the plain oracle: pinned here against it on small grids.  CPU only; needs AVX-512."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import shm_oracle as o
from oracle import shm_oracle_large as ol

pytestmark = pytest.mark.skipif(not ol.available(), reason="needs AVX-512 (oracle/csrc/shm_oracle_large.c)")


def _bunny(hCoef):
    z, F = load_golden("bunny_small")
    s = o.mesh_sources(z["V"], F)
    return s, o.make_grid(s["centroid"], s["radius"], hCoef), o.lambda_from_h(s["h"])


def test_brick_culled_step12_equals_plain_loop():
    s, g, lam = _bunny(1)
    Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"]).reshape(-1, 3)
    Yb, st = ol.step12_bricks(g, lam, s["pos"], s["nrm"], s["area"])
    assert st["worst_skipped_ratio"] <= 1e-13 and np.abs(Yb - Y).max() < 1e-12


@pytest.mark.parametrize("tau,lam_scale", [(44.0, 8.0), (12.0, 4.0)])
def test_brick_culling_skips_only_what_is_provably_negligible(tau, lam_scale):
    """A 5120-triangle sphere at 64^3 with lambda scaled up (a shorter diffusion time) so that lambda * box diagonal
    exceeds tau: clusters are skipped; with tau = 12 the a-posteriori bound rejects bricks and they are redone in full
    -- either way the plain loop's result to 1e-12."""
    from conftest import icosphere
    V, F = icosphere(4)
    s = o.mesh_sources(V, F.tolist())
    g, lam = o.make_grid(s["centroid"], s["radius"], 2), lam_scale * o.lambda_from_h(s["h"])
    for k0, k1 in ((0, 8), (24, 32)):
        Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"], k0=k0, k1=k1).reshape(-1, 3)[k0 * g.nx * g.ny:k1 * g.nx * g.ny]
        Yb, st = ol.step12_bricks(g, lam, s["pos"], s["nrm"], s["area"], tau=tau, eps=1e-13, k0=k0, k1=k1)
        assert st["worst_skipped_ratio"] <= 1e-13
        assert np.abs(Yb - Y).max() < 1e-12
        assert st["pairs"] < 0.95 * len(Yb) * len(s["area"])
        if tau < 20:
            assert st["bricks_redone"] > 0


def test_brick_step12_reproduces_the_underflow_artefact():
    """SprayBottle at 16^3: far nodes where the reference's X.norm() squares to zero are non-finite in both."""
    d = np.load(os.path.join(GOLDEN, "spraybottle_mesh.npz"))
    s = o.mesh_sources(d["V"], d["F"].tolist())
    g, lam = o.make_grid(s["centroid"], s["radius"], 0), o.lambda_from_h(s["h"])
    Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"]).reshape(-1, 3)
    Yb, st = ol.step12_bricks(g, lam, s["pos"], s["nrm"], s["area"])
    bad, badb = ~np.isfinite(Y).all(axis=1), ~np.isfinite(Yb).all(axis=1)
    assert bad.any() and np.array_equal(bad, badb)
    # gradual-underflow shell: |Y| != 1 where the squares are subnormal; both evaluate the same expression
    assert np.abs(Yb[~bad] - Y[~bad]).max() < 1e-9
    b0 = o.div_rhs(g, Y.reshape(-1))
    b1, nbad = ol.div_rhs(g, Yb)
    assert nbad > 0 and np.abs(b1 - b0).max() <= 1e-9 * np.abs(b0).max()


def test_div_rhs_and_projected_cg_equal_plain_oracle():
    s, g, lam = _bunny(0)
    Y = o.step12(g, lam, s["pos"], s["nrm"], s["area"])
    b0 = o.div_rhs(g, Y)
    b1, nbad = ol.div_rhs(g, Y)
    assert nbad == 0 and np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max()
    src, idx, w = o.constraints(g, s["pos"])
    x0, it0 = o.solve_projected_cg(g, b0, idx, w, tol=1e-11)
    x1, it1 = ol.solve_projected_cg(g, b0, idx, w, tol=1e-11)
    assert abs(it0 - it1) <= 2
    assert np.linalg.norm(x1 - x0) <= 1e-8 * np.linalg.norm(x0)
    xl = o.solve_kkt_lu(g, b0, idx, w)
    assert np.linalg.norm(x1 - xl) <= 1e-7 * np.linalg.norm(xl)
