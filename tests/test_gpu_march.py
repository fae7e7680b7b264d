"""The TMA-staged z-marching stencil kernels (csrc/grid_march.cuh) against the row-streaming kernels they replace
(csrc/grid_rows.cuh, themselves covered by the oracle parity tests): same arithmetic in the same order, so the fields
must be IDENTICAL; the fp64 reductions are summed in another order (1e-9 relative).  Shapes: the bench grid's fine levels, a
non-cubic grid, the 640 / 768 grids of the multi-GPU runs (box width 128 / 256, 5 / 3 boxes per row), z-slabs with ghost
planes (garbage -- including NaN -- in ghost planes beyond the physical boundary must be ignored)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OPS = {0: "update_p_stencil", 1: "smooth", 2: "smooth_dot", 3: "residual", 4: "smooth01"}
SHAPES = [((128, 128, 128), 0, 128), ((256, 256, 256), 0, 256), ((512, 512, 64), 0, 64), ((256, 128, 96), 0, 96),
          ((640, 640, 40), 0, 40), ((768, 768, 24), 0, 24), ((1024, 1024, 16), 0, 16),
          ((256, 256, 256), 64, 128), ((256, 256, 256), 0, 100), ((256, 256, 256), 200, 256)]


@pytest.mark.parametrize("op", sorted(OPS))
@pytest.mark.parametrize("dims,k0,k1", SHAPES)
def test_marching_kernels_equal_row_kernels(gpu_ctx, op, dims, k0, k1):
    nx, ny, nz = dims
    rng = np.random.default_rng(op * 100 + nx + k0)
    shape = (k1 - k0 + 2, ny, nx)
    a = rng.standard_normal(shape, dtype=np.float32)
    b = rng.standard_normal(shape, dtype=np.float32)
    w = rng.standard_normal(shape, dtype=np.float32)
    if k0 == 0:          # ghost plane beyond the physical boundary: never read as data
        a[0] = np.nan
        b[0] = np.nan
    if k1 == nz:
        a[-1] = np.nan
        b[-1] = np.nan
    scal = {0: [0.013, 0.37], 1: [0.013, 0.81], 2: [0.013, 0.81], 3: [0.013], 4: [0.013, 0.55, 1.7]}[op]
    r0, r1, rr = gpu_ctx.debug_stencil_op(op, dims, k0, k1, a, b, w, scal, use_tma=False)
    t0, t1, tr = gpu_ctx.debug_stencil_op(op, dims, k0, k1, a, b, w, scal, use_tma=True)
    assert np.isfinite(r0[1:-1]).all()
    assert np.array_equal(r0[1:-1], t0[1:-1]), OPS[op]
    if op == 0:
        assert np.array_equal(r1[1:-1], t1[1:-1])
    if op in (0, 2):
        assert np.all(np.abs(rr - tr) <= 1e-9 * np.abs(rr).max()) and np.abs(rr).max() > 0


def test_solver_with_and_without_tma_agree(gpu_ctx):
    """End to end: the same solve through both kernel families (identical fields -> the PCG takes the same path up to the
    summation order of its dot products)."""
    import shm3d
    from conftest import icosphere
    V, F = icosphere(3)
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=3)
    phi, st = gpu_ctx.solve(p, pos, nrm, area)
    q, _, _, _, _ = shm3d.prepare_mesh(V, F, hCoef=3)
    q.flags |= shm3d.FLAG_NO_TMA
    phi2, st2 = gpu_ctx.solve(q, pos, nrm, area)
    assert abs(st.cg_iters - st2.cg_iters) <= 4  # (convergence is checked every 4th iteration)
    assert np.linalg.norm(phi - phi2) <= 2e-5 * np.linalg.norm(phi2)
