"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the host half
of the reference interface (rows a4-a6) matches the oracle, constraint assembly and the nested-dissection
factor match scipy, and the library fails loudly without a GPU.  No compute kernels run here."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import shm3d
from conftest import ROOT, icosphere, load_golden
from oracle import shm_oracle as o


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "shm3d_grid.h")).read()
    declared = set(re.findall(r"\b(shm3d_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"shm3d_ctx"}
    L = ctypes.CDLL(shm3d.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/shm3d_grid.h but not exported"
    assert set(shm3d.EXPORTS) <= declared
    assert b"sm_100a" in shm3d.lib().shm3d_version()


def test_struct_layouts_match_header():
    # sizes the C compiler produces for the two ABI structs (natural alignment)
    assert ctypes.sizeof(shm3d.Params) == 96
    assert ctypes.sizeof(shm3d.Stats) == 184


def test_header_is_plain_c_and_links(tmp_path):
    """include/shm3d_grid.h must be consumable from C (the FFI boundary): compile a C99 program against it with gcc,
    link it to the shared library, and compare the struct layouts with the ctypes mirror."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "shm3d_grid.h"
int main(void) {
    int k0 = -1, k1 = -1;
    if (shm3d_slab_range(1, 4, 512, &k0, &k1) != SHM3D_OK) return 2;
    printf("%zu %zu %zu %zu %zu %d %d %s\\n", sizeof(shm3d_params), sizeof(shm3d_stats), offsetof(shm3d_params, cell),
           offsetof(shm3d_params, cull_tau), offsetof(shm3d_stats, kernel_launches), k0, k1, shm3d_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(shm3d.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-lshm3d_grid", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == ctypes.sizeof(shm3d.Params) and int(out[1]) == ctypes.sizeof(shm3d.Stats)
    assert int(out[2]) == shm3d.Params.cell.offset and int(out[3]) == shm3d.Params.cull_tau.offset
    assert int(out[4]) == shm3d.Stats.kernel_launches.offset
    assert (int(out[5]), int(out[6])) == (128, 256)
    assert "sm_100a" in " ".join(out[7:])


@pytest.mark.parametrize("name,hc", [("bunny_small", 1), ("polygon-bear", 0), ("knot", 2)])
def test_prepare_mesh_matches_oracle(name, hc):
    z, F = load_golden(name)
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], hc)
    p, pos, nrm, area, h = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    assert (p.nx, p.ny, p.nz) == (g.nx, g.ny, g.nz)
    assert abs(p.cell - g.cell) < 1e-13 * g.cell and np.abs(np.array(p.bbox_min) - g.bmin).max() < 1e-12
    assert abs(h - s["h"]) < 1e-12 and abs(p.lambda_ - o.lambda_from_h(s["h"])) < 1e-9
    assert np.abs(pos - s["pos"]).max() < 1e-12 and np.abs(area - s["area"]).max() < 1e-12
    assert np.abs(nrm - s["nrm"]).max() < 1e-9
    assert p.flags & shm3d.FLAG_SCRUB_NONFINITE


def test_fractional_hcoef_truncates_like_reference():
    V, F = icosphere(1)
    p, *_ = shm3d.prepare_mesh(V, F, hCoef=0.5)
    assert p.nx == int(2 * 2 ** 3.5)  # (size_t)(2*2^(hCoef+3)), src/signed_heat_grid_solver.cpp:24


@pytest.mark.parametrize("name,hc", [("bunny_small", 0), ("bunny_small", 1), ("knot", 2)])
def test_constraint_rows_match_oracle(name, hc):
    z, F = load_golden(name)
    p, pos, *_ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    g = o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)
    src, node, w = shm3d.debug_constraints(p, pos)
    src2, idx2, w2 = o.constraints(g, pos)
    assert np.array_equal(src, src2) and np.array_equal(node, idx2)
    assert np.abs(w - w2).max() < 1e-12
    if f"h{hc}_m" in z:
        assert len(src) == int(z[f"h{hc}_m"])


def test_first_source_per_cell_wins_depends_on_order():
    V, F = icosphere(2)
    p, pos, *_ = shm3d.prepare_mesh(V, F, hCoef=0)
    src, _, _ = shm3d.debug_constraints(p, pos)
    src_r, _, _ = shm3d.debug_constraints(p, pos[::-1].copy())
    assert len(src) == len(src_r) < len(pos)           # several barycentres share a cell at 16^3
    assert not np.array_equal(np.sort(src), np.sort(len(pos) - 1 - src_r))


def test_source_outside_grid_is_an_error():
    V, F = icosphere(1)
    p, pos, *_ = shm3d.prepare_mesh(V, F, hCoef=0)
    bad = pos.copy()
    bad[3] += 100.0
    with pytest.raises(shm3d.Shm3dError) as e:
        shm3d.debug_constraints(p, bad)
    assert e.value.code == shm3d.ERR_INVALID_ARG


@pytest.mark.parametrize("name,hc,uniform", [("bunny_small", 1, True), ("bunny_small", 1, False), ("knot", 3, True),
                                             ("polygon-bear", 2, False)])
def test_nested_dissection_factor_matches_scipy(name, hc, uniform):
    z, F = load_golden(name)
    p, pos, *_ = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
    g = o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)
    _, idx, w = o.constraints(g, pos)
    A = o.constraint_matrix(g, idx, w)
    if uniform:
        S = (A @ A.T).tocsc()
    else:
        d = np.full(g.N, 6.0)  # constraints never touch the domain boundary here
        S = (A @ sp.diags(1 / d) @ A.T).tocsc()
    rng = np.random.default_rng(0)
    v = rng.standard_normal(A.shape[0])
    x, mb, height = shm3d.debug_factor_solve(p, pos, v, uniform)
    xr = spla.splu(S).solve(v)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
    assert height >= 1 and mb > 0


def test_slab_ranges_tile_the_grid():
    for nz in (16, 32, 100, 1024):
        for world in (1, 2, 3, 4, 8):
            edges = [shm3d.slab_range(r, world, nz) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nz
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (this box has none; on the GPU box the test is skipped)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(shm3d.Shm3dError) as e:
        shm3d.Context(0)
    assert e.value.code == shm3d.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_drop_in_defaults_follow_the_reference_underflow_and_scrub_rules():
    """the drop-in must return what the reference returns.  The mesh overload scrubs non-finite rhs
    entries (:72-74) and both overloads evaluate X /= X.norm() in double (:61 / :171) -- so the host half sets
    SCRUB + FP64_UNDERFLOW for meshes and FP64_UNDERFLOW alone for point clouds, and the adapter, the C++ mirror and the
    Python mirror do the same by default.  No GPU needed: flags and sources only."""
    V, F = icosphere(1)
    p, *_ = shm3d.prepare_mesh(V, F)
    assert p.flags & shm3d.FLAG_SCRUB_NONFINITE and p.flags & shm3d.FLAG_FP64_UNDERFLOW
    q = shm3d.prepare_points(V, 0.3)
    assert q.flags & shm3d.FLAG_FP64_UNDERFLOW and not (q.flags & shm3d.FLAG_SCRUB_NONFINITE)
    adapter = open(os.path.join(ROOT, "adapter", "signed_heat_grid_solver_b200.cpp")).read()
    flag_lines = [ln for ln in adapter.splitlines() if "p.flags =" in ln]
    assert len(flag_lines) == 2 and all("SHM3D_FLAG_FP64_UNDERFLOW" in ln for ln in flag_lines)
    assert sum("SHM3D_FLAG_SCRUB_NONFINITE" in ln for ln in flag_lines) == 1       # mesh overload only
    mirror = open(os.path.join(ROOT, "include", "shm3d", "signed_heat_grid_solver.hpp")).read()
    assert "bool referenceUnderflow = true;" in mirror
    src = open(os.path.join(ROOT, "signed-heat-3d_b200", "shm3d", "__init__.py")).read()
    assert "self.reference_underflow = True" in src
    # every diagnostics flag of the header has its mirror constant with the same value
    hdr = open(os.path.join(ROOT, "include", "shm3d_grid.h")).read()
    for name, val in re.findall(r"#define SHM3D_FLAG_([A-Z0-9_]+) (\d+)u", hdr):
        assert getattr(shm3d, "FLAG_" + name) == int(val), name
