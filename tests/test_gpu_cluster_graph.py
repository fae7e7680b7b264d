"""The CUDA-graph replay of the PCG iteration (csrc/solver.cu run_pcg), the programmatic dependent launch of the projector
sweeps, and the opt-in single-launch tail program (csrc/mg_tail.cuh: the V-cycle from 16^3 down as ONE launch) against the
fully serialised kernel-by-kernel path they replace: same operations in the same order (only the dense coarsest solve sums in another order), so the fields
agree to fp32 rounding and the iteration counts match.  The parity tests against the oracle run the default (cluster
programs + graph) path; this file pins the two paths to each other and checks that the fast path is really taken."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(ctx, V, F, hCoef, flags=0, **kw):
    import shm3d
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=hCoef)
    p.flags |= flags
    for k, v in kw.items():
        setattr(p, k, v)
    return ctx.solve(p, pos, nrm, area)


@pytest.mark.parametrize("hCoef", [0, 1, 2, 3])
def test_cluster_tail_and_graph_equal_launch_by_launch(gpu_ctx, hCoef):
    import shm3d
    from conftest import icosphere
    V, F = icosphere(3)
    phi, st = _solve(gpu_ctx, V, F, hCoef, shm3d.FLAG_TAIL_PROGRAM)
    ref, st0 = _solve(gpu_ctx, V, F, hCoef, shm3d.FLAG_NO_GRAPH | shm3d.FLAG_NO_PDL)
    assert st0.tail_ops == 0 and st0.graph_replays == 0
    assert st.tail_ops > 0, "the V-cycle tail did not run as a cluster program"
    assert st.graph_replays >= st.cg_iters - 2 > 0, "the PCG iterations were not replayed from the captured graph"
    assert abs(st.cg_iters - st0.cg_iters) <= 4  # (convergence is checked every 4th iteration)
    err = np.linalg.norm(phi - ref) / np.linalg.norm(ref)
    print(f"hCoef {hCoef}: its {st.cg_iters}/{st0.cg_iters}, tail ops {st.tail_ops}, launches {st.kernel_launches} vs "
          f"{st0.kernel_launches}, rel-L2 {err:.2e}")
    assert err <= 5e-5
    assert st.kernel_launches < st0.kernel_launches


def test_each_switch_alone(gpu_ctx):
    import shm3d
    from conftest import icosphere
    V, F = icosphere(3)
    ref, st0 = _solve(gpu_ctx, V, F, 2, shm3d.FLAG_NO_GRAPH | shm3d.FLAG_NO_PDL)
    a, sa = _solve(gpu_ctx, V, F, 2, shm3d.FLAG_NO_GRAPH | shm3d.FLAG_TAIL_PROGRAM)
    b, sb = _solve(gpu_ctx, V, F, 2, shm3d.FLAG_NO_PDL)
    assert sa.tail_ops > 0 and sa.graph_replays == 0
    assert sb.tail_ops == 0 and sb.graph_replays > 0
    # graph replay runs the very same kernels on the same buffers (only the order of the forward sweep's fp64 atomics
    # is not fixed)
    assert np.linalg.norm(b - ref) <= 1e-6 * np.linalg.norm(ref)
    assert np.linalg.norm(a - ref) <= 5e-5 * np.linalg.norm(ref)


def test_graph_is_reused_and_updated_across_solves(gpu_ctx):
    """Consecutive solves on one context with different sources / sizes: the executable graph of the previous solve is
    updated in place (same topology) or rebuilt (other hierarchy); results equal a fresh kernel-by-kernel solve."""
    import shm3d
    from conftest import icosphere
    for sub, h in ((3, 2), (2, 2), (3, 1), (3, 2)):
        V, F = icosphere(sub, radius=1.0 + 0.1 * sub)
        phi, st = _solve(gpu_ctx, V, F, h)
        ref, _ = _solve(gpu_ctx, V, F, h, shm3d.FLAG_NO_GRAPH | shm3d.FLAG_NO_PDL)
        assert st.graph_replays > 0
        assert np.linalg.norm(phi - ref) <= 5e-5 * np.linalg.norm(ref)


def test_profiled_solve_equals_unprofiled(gpu_ctx):
    """SHM3D_FLAG_PROFILE runs the first iterations kernel by kernel under CUDA events, the rest from the graph."""
    import shm3d
    from conftest import icosphere
    V, F = icosphere(3)
    phi, st = _solve(gpu_ctx, V, F, 3)
    phi2, st2 = _solve(gpu_ctx, V, F, 3, shm3d.FLAG_PROFILE)
    assert st2.pcg_vcycles > 0 and st2.ms_pcg_vcycle > 0 and st2.pcg_stencil_launches > 0
    assert st2.graph_replays > 0
    assert st.cg_iters == st2.cg_iters
    assert np.linalg.norm(phi - phi2) <= 1e-6 * np.linalg.norm(phi)
