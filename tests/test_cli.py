"""The headless driver (tools/shm3d_cli.cpp, built on the C++ mirror include/shm3d/signed_heat_grid_solver.hpp):
CPU: its OBJ / .pc readers and host set-up agree with the oracle (--dry-run, no device work);
GPU: a full solve through the C++ class reproduces the golden field."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import shm_oracle as o

CLI = os.path.join(ROOT, "signed-heat-3d_b200", "bin", "shm3d_cli")


def write_obj(path, V, F, junk_vertices=0):
    with open(path, "w") as fh:
        fh.write("# written by tests/test_cli.py\n")
        for _ in range(junk_vertices):  # unreferenced vertices must be stripped (meshio.cpp:22-29)
            fh.write("v 1000.0 1000.0 1000.0\n")
        for v in V:
            fh.write("v %.17g %.17g %.17g\n" % tuple(v))
        for f in F:
            fh.write("f " + " ".join(f"{i + 1 + junk_vertices}/1/1" for i in f) + "\n")


def test_cli_exists_and_prints_usage():
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage: shm3d_cli" in r.stderr
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 1 and "Please specify a mesh file" in r.stderr  # src/main.cpp:253-256


@pytest.mark.parametrize("name,hc", [("bunny_small", 1), ("polygon-bear", 0)])
def test_cli_dry_run_matches_oracle(tmp_path, name, hc):
    z, F = load_golden(name)
    path = str(tmp_path / (name + ".obj"))
    write_obj(path, z["V"], F, junk_vertices=3)
    r = subprocess.run([CLI, path, "--grid", "--h", str(hc), "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], hc)
    assert d["nx"] == g.nx and d["sources"] == len(F) and d["vertices"] == len(z["V"])
    assert abs(d["h"] - s["h"]) < 1e-12 and abs(d["cell"] - g.cell) < 1e-12
    assert abs(d["lambda"] - o.lambda_from_h(s["h"])) < 1e-9
    assert np.abs(np.array(d["bbox_min"]) - g.bmin).max() < 1e-12


TRICKY_OBJ = """# comments, an unreferenced vertex, all four index syntaxes, a quad (no negative indices: the reference's loader has none)
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 5 5 5
v 0.5 0.5 1
vn 0 0 1
vt 0.5 0.5
f 1/1/1 2/1/1 3/1/1 4/1/1
f 1//1 2//1 6//1
f 2 3 6
f 3/1 4/1 6/1
f 4 1 6
"""


def test_cli_obj_reader_equals_geometry_central_readSurfaceMesh(tmp_path):
    """The CLI restates geometry-central's OBJ loader; here the REAL loader (readSurfaceMesh as src/main.cpp:269 calls it:
    SimplePolygonMesh parser + stripUnusedVertices, compiled from the reference tree into oracle/_ref/libshm_gc_ref.so)
    reads the same files and the reference's own host code runs on the result."""
    from oracle import reference_build as rb
    if not (rb.build() and rb.gc_available()):
        pytest.skip("no oracle/_ref/libshm_gc_ref.so")
    paths = []
    p = str(tmp_path / "tricky.obj")
    with open(p, "w") as fh:
        fh.write(TRICKY_OBJ)
    paths.append(p)
    for name in ("bunny_small", "polygon-bear"):
        z, F = load_golden(name)
        p = str(tmp_path / (name + ".obj"))
        write_obj(p, z["V"], F, junk_vertices=2)
        paths.append(p)
    for name in ("SprayBottle", "knot", "chair", "rocker"):      # the reference's own sample files, where present
        p = os.path.join("/root/reference/data", name + ".obj")
        if os.path.exists(p):
            paths.append(p)
    for p in paths:
        g = rb.gc_read_mesh(p)
        r = subprocess.run([CLI, p, "--dry-run"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        d = json.loads(r.stdout)
        assert d["vertices"] == g["n_vertices"] and d["sources"] == g["n_faces"], p
        assert abs(d["h"] - g["h"]) < 1e-13 * g["h"], p
        assert np.abs(np.array(d["bbox_min"]) - (g["centroid"] - 2.0 * g["radius"])).max() < 1e-12 * (1 + g["radius"]), p


def test_cli_reads_pc_files(tmp_path):
    z, F = load_golden("bunny_small")
    s = o.mesh_sources(z["V"], F)
    path = str(tmp_path / "cloud.pc")
    with open(path, "w") as fh:
        for q, n in zip(s["pos"], s["nrm"]):
            fh.write("v %.17g %.17g %.17g\nvn %.17g %.17g %.17g\n" % (*q, *n))
    r = subprocess.run([CLI, path, "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout)
    assert d["sources"] == len(F) and d["nx"] == 16
    # h = mean edge length of the local-Delaunay triangle soup (row N1 without the tufted-cover flips)
    import shm3d
    _, h, _ = shm3d.point_weights(s["pos"], s["nrm"])
    assert abs(d["h"] - h) < 1e-12 * h


def test_cli_point_clouds_of_the_reference_get_geometry_centrals_h():
    """The reference's sample .pc files through the CLI's reader and the library's tufted-cover weights, against
    geometry-central's own pipeline (oracle/_ref/libshm_gc_ref.so) on the same points: same mean edge length, hence the
    same lambda and grid.  Only where the reference tree is present."""
    from oracle import reference_build as rb
    if not (rb.build() and rb.gc_available()):
        pytest.skip("no oracle/_ref/libshm_gc_ref.so")
    seen = 0
    for name in ("bunny", "chair", "rocker", "knot", "SprayBottle"):
        path = os.path.join("/root/reference/data", name + ".pc")
        if not os.path.exists(path):
            continue
        seen += 1
        P, N = o.read_pc(path)
        _, h_ref, _, _ = rb.gc_point_weights(P, N)
        r = subprocess.run([CLI, path, "--dry-run"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        d = json.loads(r.stdout)
        assert d["sources"] == len(P) and abs(d["h"] - h_ref) < 1e-12 * h_ref, name
        assert abs(d["lambda"] - 1.0 / h_ref) < 1e-9 / h_ref, name
    if seen == 0:
        pytest.skip("no reference data files here")


def test_cli_bad_input_fails_cleanly(tmp_path):
    r = subprocess.run([CLI, str(tmp_path / "missing.obj"), "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 3 and "cannot read mesh" in r.stderr


@pytest.mark.gpu
def test_cli_solve_matches_golden(tmp_path):
    z, F = load_golden("bunny_small")
    path = str(tmp_path / "bunny_small.obj")
    out = str(tmp_path / "phi.npy")
    write_obj(path, z["V"], F)
    r = subprocess.run([CLI, path, "--grid", "--h", "1", "-V", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "min:" in r.stderr and "Solve time (s)" in r.stderr
    phi = np.load(out)
    assert phi.shape == (32, 32, 32)
    ref = z["h1_phi"]
    assert np.linalg.norm(phi.ravel() - ref) / np.linalg.norm(ref) < 1e-4


@pytest.mark.gpu
def test_cli_fast_flag(tmp_path):
    z, F = load_golden("bunny_small")
    path = str(tmp_path / "b.obj")
    out = str(tmp_path / "phi.npy")
    write_obj(path, z["V"], F)
    r = subprocess.run([CLI, path, "--fast", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = o.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)
    phi = np.load(out).ravel()
    assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-4


@pytest.mark.gpu
def test_cli_contour_and_export(tmp_path):
    """The GUI's Contour + "Export isosurface" buttons (src/main.cpp:116-128, :160-190) as flags: the OBJ holds the mesh
    the reference's consumer would extract from the phi the same run wrote."""
    z, F = load_golden("bunny_small")
    path = str(tmp_path / "bunny_small.obj")
    out, iso = str(tmp_path / "phi.npy"), str(tmp_path / "iso.obj")
    write_obj(path, z["V"], F)
    r = subprocess.run([CLI, path, "--grid", "--h", "1", "-o", out, "--isoval", "0.25", "--iso-out", iso],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "isosurface phi = 0.25" in r.stderr
    phi = np.load(out)
    g = o.Grid(32, 32, 32, z["h1_bmin"], float(z["h1_cell"]))
    Vo, To = o.marching_cubes(phi.ravel(), 0.25, (32, 32, 32), *o.grid_bounds_f32(g))
    V = np.array([[float(t) for t in ln.split()[1:]] for ln in open(iso) if ln.startswith("v ")], dtype=np.float32)
    T = np.array([[int(t) - 1 for t in ln.split()[1:]] for ln in open(iso) if ln.startswith("f ")], dtype=np.uint32)
    # %.9g round-trips float32; the grid origin comes from the CLI's own host set-up (equal to the fixture's to ~1 ulp of double)
    assert np.array_equal(T, To) and V.shape == Vo.shape and np.abs(V - Vo).max() < 1e-5
