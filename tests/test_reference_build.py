"""The oracle against the REFERENCE'S OWN SOURCE: oracle/_ref/libshm_ref.so is src/signed_heat_grid_solver.cpp +
src/signed_heat_3d.cpp of the reference, compiled unmodified against oracle/ref_shim (a stand-in for the slices of
geometry-central / Eigen / polyscope they use; the sparse LU behind solveSquare is a scipy SuperLU callback).
Built on demand where /root/reference exists (this container); skipped where neither the tree nor the prebuilt library is."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import icosphere, load_golden
from oracle import reference_build as rb
from oracle import shm_oracle as o

pytestmark = pytest.mark.skipif(not rb.build(), reason="no reference tree and no prebuilt oracle/_ref/libshm_ref.so")


@pytest.mark.parametrize("name,hc", [("bunny_small", 0), ("polygon-bear", 0), ("bunny_small", 1)])
def test_oracle_equals_reference_source_mesh(name, hc):
    z, F = load_golden(name)
    phi, info = rb.compute_distance_mesh(z["V"], F, hCoef=hc, return_info=True)
    ora = o.compute_distance_mesh(z["V"], F, hCoef=hc)
    assert np.abs(phi - ora).max() < 1e-10 * np.abs(ora).max()
    # ... and therefore the committed golden vectors are the reference's own numbers
    assert np.abs(phi - z[f"h{hc}_phi"]).max() < 1e-10 * np.abs(ora).max()
    # the side effect main.cpp relies on: registerVolumeGrid("domain", {nx,ny,nz}, bboxMin, bboxMax) (:35)
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], hc)
    assert list(info["dims"]) == [g.nx, g.ny, g.nz]
    assert np.abs(info["bbox"][:3] - g.bmin.astype(np.float32)).max() < 1e-6
    assert np.abs(info["bbox"][3:] - (g.bmin + g.cell * (g.nx - 1)).astype(np.float32)).max() < 1e-5


def test_kkt_system_assembled_by_the_reference_matches_the_oracle():
    """[[L, A^T],[A, 0]] and [D^T Y; 0] exactly as src/signed_heat_grid_solver.cpp:80-106 hands them to solveSquare."""
    z, F = load_golden("bunny_small")
    _, info = rb.compute_distance_mesh(z["V"], F, hCoef=0, return_info=True)
    assert len(info["solves"]) == 1
    K = info["solves"][0]
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], 0)
    L = o.laplacian_matrix(g)
    src, idx, w = o.constraints(g, s["pos"])
    A = o.constraint_matrix(g, idx, w)
    m = A.shape[0]
    assert K["n"] == g.N + m == 4096 + 79
    LHS = sp.bmat([[L, A.T], [A, sp.csr_matrix((m, m))]], format="csc")
    diff = (K["A"] - LHS).tocoo()
    assert np.abs(diff.data).max() if diff.nnz else 0.0 < 1e-12 * np.abs(LHS.data).max()
    assert (K["A"] != 0).nnz == (LHS != 0).nnz
    lam = o.lambda_from_h(s["h"])
    b = o.div_rhs(g, o.step12(g, lam, s["pos"], s["nrm"], s["area"]))
    assert np.abs(K["rhs"][:g.N] - b).max() < 1e-9 * np.abs(b).max()
    assert np.all(K["rhs"][g.N:] == 0)


def test_helper_functions_match():
    """centroid / radius / meanEdgeLength / setFaceVectorAreas / yukawaPotential (src/signed_heat_3d.cpp)."""
    for name in ("bunny_small", "polygon-bear"):
        z, F = load_golden(name)
        r = rb.mesh_scalars(z["V"], F)
        s = o.mesh_sources(z["V"], F)
        assert abs(r["h"] - s["h"]) < 1e-13 and abs(r["radius"] - s["radius"]) < 1e-13
        assert np.abs(r["centroid"] - s["centroid"]).max() < 1e-13
        assert np.abs(r["area"] - s["area"]).max() < 1e-13 and np.abs(r["nrm"] - s["nrm"]).max() < 1e-12
    x, y = np.array([0.3, -1.2, 2.0]), np.array([1.0, 0.5, -0.25])
    rr = np.linalg.norm(x - y)
    assert abs(rb.yukawa(x, y, 3.7) - np.exp(-3.7 * rr) / rr) < 1e-16


def test_fast_integration_matches_reference_bfs():
    z, F = load_golden("bunny_small")
    ref = rb.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)
    assert np.abs(ref - o.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)).max() < 1e-12
    # the closed form the GPU path uses
    s = o.mesh_sources(z["V"], F)
    g = o.make_grid(s["centroid"], s["radius"], 0)
    Y = o.step12(g, o.lambda_from_h(s["h"]), s["pos"], s["nrm"], s["area"])
    pre = o.integrate_greedily_prefix(g, Y)
    pre = pre - o.source_average(g, pre, s["pos"], s["area"])
    assert np.abs(ref - pre).max() < 1e-12


def test_point_cloud_overload_matches():
    V, F = icosphere(2)
    s = o.mesh_sources(V, F)
    P, Nn = s["pos"], s["nrm"]
    areas = s["area"] * 1.3
    h = 0.2
    ref = rb.compute_distance_points(P, Nn, areas, h, hCoef=0)
    c = P.sum(axis=0) / len(P)
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    ora = o.compute_distance(P, Nn, areas, h, c, r, hCoef=0, scrub_nonfinite=False)
    assert np.abs(ref - ora).max() < 1e-10 * np.abs(ora).max()


def test_product_host_half_equals_reference_source():
    """The product's own host code (shm3d_prepare_mesh in libshm3d_grid.so: rows a4-a6) against the reference's
    centroid / radius / meanEdgeLength / setFaceVectorAreas, and its grid against what the reference registered."""
    import shm3d
    for name, hc in (("bunny_small", 1), ("polygon-bear", 0)):
        z, F = load_golden(name)
        r = rb.mesh_scalars(z["V"], F)
        p, pos, nrm, area, h = shm3d.prepare_mesh(z["V"], F, hCoef=hc)
        assert abs(h - r["h"]) < 1e-13
        assert np.abs(area - r["area"]).max() < 1e-13 and np.abs(nrm - r["nrm"]).max() < 1e-12
        s = r["radius"] * 2.0
        assert np.abs(np.array(p.bbox_min) - (r["centroid"] - s)).max() < 1e-13
        assert abs(p.cell - 2.0 * s / (p.nx - 1)) < 1e-15
        _, info = rb.compute_distance_mesh(z["V"], F, hCoef=hc, fast=True, return_info=True)
        assert list(info["dims"]) == [p.nx, p.ny, p.nz]


def test_shim_containers_equal_the_real_geometry_central():
    """The shim that stands in for geometry-central when the reference's two translation units are compiled restates its
    mesh containers (face order, vertex order inside a face, unique edges).  Here the REAL geometry-central sources
    (oracle/_ref/libshm_gc_ref.so: SurfaceMesh built with makeSurfaceMeshAndGeometry as src/main.cpp does, the reference's
    src/signed_heat_3d.cpp on top) give the same host quantities -- to the last bit or two for the shim, to rounding for the oracle
    and the product -- on triangle, polygon and non-trivial meshes; in particular the source order, which decides the
    constraint rows, is the input order."""
    import shm3d
    if not rb.gc_available():
        pytest.skip("no oracle/_ref/libshm_gc_ref.so")
    for name in ("bunny_small", "polygon-bear", "knot"):
        z, F = load_golden(name)
        g = rb.gc_mesh_sources(z["V"], F)
        sh = rb.mesh_scalars(z["V"], F)                      # the shim + the same reference source
        assert sh["h"] == g["h"] and sh["radius"] == g["radius"] and np.array_equal(sh["centroid"], g["centroid"])
        assert np.array_equal(sh["area"], g["area"]) and np.abs(sh["nrm"] - g["nrm"]).max() < 1e-15   # last-bit: N / |N|
        s = o.mesh_sources(z["V"], F)                        # the oracle
        assert np.array_equal(s["pos"], g["pos"])
        assert abs(s["h"] - g["h"]) < 1e-13 * g["h"] and np.abs(s["area"] - g["area"]).max() < 1e-13
        assert np.abs(s["nrm"] - g["nrm"]).max() < 1e-13
        p, pos, nrm, area, h = shm3d.prepare_mesh(z["V"], F, hCoef=0)   # the product
        assert np.array_equal(pos, g["pos"]) and abs(h - g["h"]) < 1e-13 * g["h"]
        assert np.abs(area - g["area"]).max() < 1e-12 and np.abs(nrm - g["nrm"]).max() < 1e-11


def test_reference_source_on_the_real_geometry_central():
    """oracle/_ref/libshm_ref_gc.so: the reference's own src/signed_heat_grid_solver.cpp + src/signed_heat_3d.cpp linked
    with geometry-central's REAL sources from the reference tree (mesh containers, point-cloud pipeline, solveSquare /
    PositiveDefiniteSolver / horizontalStack / verticalStack) -- only Eigen (a stub with a working sparse container; its
    SparseLU hands the system to scipy SuperLU) and polyscope's registerVolumeGrid are stand-ins -- driven like
    src/main.cpp drives the class.  Same fields as the committed golden vectors, the shim build and the oracle: mesh,
    polygon mesh, fastIntegration, and the point-cloud overload END TO END (geometry-central's own tufted-cover weights)."""
    if not rb.ref_gc_available():
        pytest.skip("no oracle/_ref/libshm_ref_gc.so")
    z, F = load_golden("bunny_small")
    phi, info = rb.ref_gc_compute_distance_mesh(z["V"], F, hCoef=0, return_info=True)
    assert np.linalg.norm(phi - z["h0_phi"]) / np.linalg.norm(z["h0_phi"]) < 1e-12
    assert np.abs(phi - rb.compute_distance_mesh(z["V"], F, hCoef=0)).max() < 1e-11
    assert list(info["dims"]) == [16, 16, 16] and len(info["solves"]) == 1
    K = info["solves"][0]["A"]                                   # the KKT matrix geometry-central's stacking produced
    g = o.make_grid(z["centroid"], float(z["radius"]), hCoef=0)
    s = o.mesh_sources(z["V"], F)
    _, idx, w = o.constraints(g, s["pos"])
    A = o.constraint_matrix(g, idx, w)
    Lm = o.laplacian_matrix(g)
    ref = sp.bmat([[Lm, A.T], [A, None]], format="csc")
    assert K.shape == ref.shape and abs(K - ref).max() < 1e-12
    phif = rb.ref_gc_compute_distance_mesh(z["V"], F, hCoef=0, fast=True)
    assert np.abs(phif - o.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)).max() < 1e-12
    zb, Fb = load_golden("polygon-bear")
    phib = rb.ref_gc_compute_distance_mesh(zb["V"], Fb, hCoef=0)
    refb = o.compute_distance_mesh(zb["V"], Fb, hCoef=0)
    assert np.linalg.norm(phib - refb) / np.linalg.norm(refb) < 1e-12
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    P, N = d["P"], d["N"]
    phip = rb.ref_gc_compute_distance_points(P, N, hCoef=0)
    wts = np.load(os.path.join(GOLDEN, "point_weights_gc.npz"))
    c = P.mean(axis=0)
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    refp = o.compute_distance(P, N, wts["bunny_pc_areas"], float(wts["bunny_pc_h"]), c, r, hCoef=0, scrub_nonfinite=False)
    assert np.linalg.norm(phip - refp) / np.linalg.norm(refp) < 1e-12


def test_constraint_rows_and_rhs_at_64_equal_the_reference():
    """A grid far beyond the reference's LU (64^3): the system it ASSEMBLES (on the real geometry-central, factorisation
    skipped) against the product's constraint rows -- same m, same nodes, same trilinear weights -- and the oracle's
    right-hand side D^T Y."""
    import shm3d
    if not rb.ref_gc_available():
        pytest.skip("no oracle/_ref/libshm_ref_gc.so")
    z, F = load_golden("polygon-bear")
    K, rhs = rb.ref_gc_assemble_mesh(z["V"], F, hCoef=2)
    N = 64 ** 3
    A_ref = K.tocsr()[N:, :N]
    p, pos, nrm, area, h = shm3d.prepare_mesh(z["V"], F, hCoef=2)
    src, node, w = shm3d.debug_constraints(p, pos)
    assert A_ref.shape[0] == len(src)
    A = sp.coo_matrix((w.ravel(), (np.repeat(np.arange(len(src)), 8), node.ravel())), shape=(len(src), N)).tocsr()
    assert abs(A_ref - A).max() < 1e-13
    g = o.make_grid(z["centroid"], float(z["radius"]), hCoef=2)
    s = o.mesh_sources(z["V"], F)
    b = o.div_rhs(g, o.step12(g, o.lambda_from_h(s["h"]), s["pos"], s["nrm"], s["area"]))
    assert np.linalg.norm(rhs[:N] - b) / np.linalg.norm(b) < 1e-12 and not rhs[N:].any()


def test_knot_golden_is_the_reference_sources_output():
    """data/knot.obj at hCoef 1 (30 504 faces x 32^3 nodes, ~1 min single-threaded: the reference recomputes every
    barycentre per pair): the committed golden field the GPU tests compare against is the reference's own result."""
    z, F = load_golden("knot")
    phi = rb.compute_distance_mesh(z["V"], F, hCoef=1)
    assert np.abs(phi - z["h1_phi"]).max() < 1e-9 * np.abs(phi).max()


def test_adapter_compiles_against_the_reference_headers(tmp_path):
    """adapter/signed_heat_grid_solver_b200.cpp -- the drop-in replacement of src/signed_heat_grid_solver.cpp -- must
    compile against the reference's own unchanged headers (the shim stands in for geometry-central / Eigen / polyscope)
    and link against libshm3d_grid.so together with the reference's src/signed_heat_3d.cpp."""
    import os
    import subprocess
    import shm3d
    from conftest import ROOT
    if not os.path.isdir(os.path.join(rb.REF_ROOT, "include")):
        pytest.skip("needs the reference tree")
    libdir = os.path.dirname(shm3d.LIB_PATH)
    out = str(tmp_path / "libadapter.so")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-fPIC", "-shared", "-w",
                           "-I" + os.path.join(ROOT, "oracle", "ref_shim", "include"),
                           "-I" + os.path.join(rb.REF_ROOT, "include"), "-I" + os.path.join(ROOT, "include"),
                           "-o", out, os.path.join(ROOT, "adapter", "signed_heat_grid_solver_b200.cpp"),
                           os.path.join(rb.REF_ROOT, "src", "signed_heat_3d.cpp"),
                           os.path.join(ROOT, "oracle", "ref_shim", "shim_defs.cpp"),
                           "-L" + libdir, "-lshm3d_grid", "-Wl,-rpath," + libdir, "-Wl,--no-undefined"])
    import ctypes
    L = ctypes.CDLL(out)   # resolves every symbol: the class is fully defined by the adapter + signed_heat_3d.cpp
    assert L is not None
