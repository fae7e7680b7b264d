"""CPU tests of the point-cloud source weights (row N1: geometry-central's tufted-cover pipeline restated in
csrc/point_weights.cpp).  Pinned to geometry-central's OWN sources: oracle/_ref/libshm_gc_ref.so is its point-cloud /
tufted-cover code compiled from the reference tree against an Eigen interface stub (oracle/Makefile), called exactly as
the reference calls it; the product must reproduce its vertex dual areas and mean edge length to rounding (live where the
library exists, and through the committed fixture tests/golden/point_weights_gc.npz everywhere).  Independent checks:
tangent-plane local Delaunay rings against scipy's Delaunay triangulation, invariants of the cover (closed manifold, area
preserved by the intrinsic flips, intrinsically Delaunay at the end), closed forms on sampled spheres, and the bunny point
cloud (= the vertices of data/bunny_small.obj) against that mesh's own mean edge length."""
import os
import sys

import numpy as np
import pytest
from scipy.spatial import ConvexHull, Delaunay, cKDTree

import shm3d
from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def golden_clouds():
    """the clouds of tests/golden/make_golden_point_weights.py"""
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    yield "bunny_pc", d["P"], d["N"]
    rng = np.random.default_rng(3)
    P = rng.standard_normal((3000, 3))
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    yield "random_sphere_3000", P, P.copy()
    u, v = rng.uniform(0, 2 * np.pi, 4000), rng.uniform(0, 2 * np.pi, 4000)
    P = np.stack([(2 + 0.7 * np.cos(v)) * np.cos(u), (2 + 0.7 * np.cos(v)) * np.sin(u), 0.7 * np.sin(v)], axis=1)
    N = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], axis=1)
    yield "random_torus_4000", P, N


def test_weights_equal_geometry_central_golden():
    """vertexDualAreas and meanEdgeLength(tuftedGeom) as geometry-central's own code computes them (fixture)."""
    g = np.load(os.path.join(GOLDEN, "point_weights_gc.npz"))
    for name, P, N in golden_clouds():
        areas, h, n_tris, diag = shm3d.point_weights(P, N, diagnostics=True)
        assert abs(h - float(g[name + "_h"])) < 1e-12 * h
        assert np.abs(areas - g[name + "_areas"]).max() < 1e-11 * areas.max()
        nf, ne = g[name + "_faces_edges"]
        assert 2 * n_tris == nf and 3 * n_tris == ne          # the cover doubles the soup; closed: E = 3F/2


def test_weights_equal_geometry_central_live():
    """Same comparison against the library itself on clouds the fixture does not hold: lattice ties (cocircular points),
    a jittered plane with boundary, the knot's vertices with averaged face normals."""
    from oracle import reference_build as rb
    if not (rb.build() and rb.gc_available()):
        pytest.skip("no oracle/_ref/libshm_gc_ref.so")
    from conftest import load_golden
    rng = np.random.default_rng(7)
    gx, gy = np.meshgrid(np.arange(30.0), np.arange(30.0))
    lattice = np.stack([gx.ravel(), gy.ravel(), np.zeros(900)], axis=1)
    up = np.tile([0.0, 0.0, 1.0], (900, 1))
    z, F = load_golden("knot")
    V = z["V"]
    Nk = np.zeros_like(V)
    for f in F:
        n = np.cross(V[f[1]] - V[f[0]], V[f[2]] - V[f[0]])
        for v in f:
            Nk[v] += n
    Nk /= np.linalg.norm(Nk, axis=1, keepdims=True)
    for name, P, N in (("lattice", lattice, up), ("jittered", lattice + 1e-3 * rng.standard_normal(lattice.shape), up),
                       ("knot", V, Nk)):
        a_ref, h_ref, nf, ne = rb.gc_point_weights(P, N)
        a, h, n_tris = shm3d.point_weights(P, N)
        assert abs(h - h_ref) < 1e-12 * h_ref, name
        assert np.abs(a - a_ref).max() < 1e-10 * a_ref.max(), name
        assert 2 * n_tris == nf, name


def test_weights_equal_geometry_central_on_degenerate_clouds():
    """Structured and damaged inputs, where every discrete decision is a tie or nearly one: surface lattices of a cube
    (coincident points along its edges), cylinder, latitude-longitude sphere, hexagonal plane; exact duplicates;
    coordinates quantised to 1e-3 (activates the intrinsic mollification).  Same soup, same areas, same h."""
    from oracle import reference_build as rb
    if not (rb.build() and rb.gc_available()):
        pytest.skip("no oracle/_ref/libshm_gc_ref.so")
    rng = np.random.default_rng(42)
    clouds = []
    g = np.linspace(-1, 1, 21)
    pts, nrm = [], []
    for ax in range(3):
        for sgn in (-1.0, 1.0):
            u, v = np.meshgrid(g, g)
            p = np.zeros((u.size, 3))
            p[:, ax], p[:, (ax + 1) % 3], p[:, (ax + 2) % 3] = sgn, u.ravel(), v.ravel()
            n = np.zeros_like(p)
            n[:, ax] = sgn
            pts.append(p)
            nrm.append(n)
    clouds.append(("cube lattice", np.concatenate(pts), np.concatenate(nrm)))
    th, zz = np.meshgrid(np.linspace(0, 2 * np.pi, 64, endpoint=False), np.linspace(-1, 1, 33))
    cyl = np.stack([np.cos(th).ravel(), np.sin(th).ravel(), zz.ravel()], axis=1)
    clouds.append(("cylinder lattice", cyl, cyl * [1.0, 1.0, 0.0]))
    th, ph = np.meshgrid(np.linspace(0, 2 * np.pi, 48, endpoint=False), np.linspace(0.05, np.pi - 0.05, 40))
    sph = np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], axis=-1).reshape(-1, 3)
    clouds.append(("lat-long sphere", sph, sph.copy()))
    a, b = np.meshgrid(np.arange(40), np.arange(40))
    hexp = np.stack([a.ravel() + 0.5 * (b.ravel() % 2), b.ravel() * np.sqrt(3) / 2, 0.0 * a.ravel()], axis=1)
    clouds.append(("hex lattice", hexp, np.tile([0.0, 0.0, 1.0], (len(hexp), 1))))
    P = rng.standard_normal((1500, 3))
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    P = np.concatenate([P, P[:200]])
    clouds.append(("200 duplicated points", P, P.copy()))
    P = rng.standard_normal((3000, 3))
    P = np.round(P / np.linalg.norm(P, axis=1, keepdims=True), 3)
    clouds.append(("quantised coordinates", P, P / np.linalg.norm(P, axis=1, keepdims=True)))
    for name, P, N in clouds:
        a_ref, h_ref, nf, ne = rb.gc_point_weights(P, N)
        a, h, n_tris = shm3d.point_weights(P, N)
        assert 2 * n_tris == nf, name
        assert abs(h - h_ref) < 1e-12 * h_ref, name
        assert np.abs(a - a_ref).max() < 1e-10 * a_ref.max(), name


def fib_points(n):
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def test_local_ring_equals_the_delaunay_star_of_the_centre():
    """local_triangulation.cpp:10-210 restated: for an interior centre the surviving ring is the 1-ring of the centre in
    the 2-D Delaunay triangulation of {centre} + neighbours."""
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(8, 31))
        pts = rng.standard_normal((n, 2))
        ring, tri = shm3d.debug_local_ring(pts)
        D = Delaunay(np.vstack([[0.0, 0.0], pts]))
        star = set()
        for s in D.simplices:
            if 0 in s:
                star |= set(int(v) - 1 for v in s if v != 0)
        on_hull = 0 in set(D.convex_hull.ravel())
        assert set(ring.tolist()) == star, (trial, sorted(ring.tolist()), sorted(star))
        if not on_hull:
            assert tri.all() and len(ring) >= 3
        # counter-clockwise order
        ang = np.arctan2(pts[ring, 1], pts[ring, 0])
        assert (np.diff(ang) > 0).all()


def test_ring_drops_collinear_far_points_and_handles_half_planes():
    pts = np.array([[1.0, 0.0], [2.0, 0.0], [0.0, 1.0], [-1.0, 0.2], [0.0, -1.0]])
    ring, tri = shm3d.debug_local_ring(pts)
    assert 1 not in ring.tolist()            # (2,0) hides behind (1,0)
    half = np.array([[1.0, 0.1], [0.5, 1.0], [-0.5, 1.0], [-1.0, 0.1]])   # all neighbours in the upper half-plane
    ring, tri = shm3d.debug_local_ring(half)
    assert len(ring) == 4 and tri.sum() == 3  # no triangle across the empty half-plane


def test_sphere_weights():
    n = 4000
    P = fib_points(n)
    areas, h, ntri = shm3d.point_weights(P, P)
    assert ntri > 5 * n
    # every surface triangle shows up in the local triangulation of each of its 3 corners, on 2 sheets of the cover
    assert abs(areas.sum() / (6 * 4 * np.pi) - 1) < 0.02
    assert areas.std() / areas.mean() < 0.1
    hull = ConvexHull(P)
    e = np.vstack([hull.simplices[:, [0, 1]], hull.simplices[:, [1, 2]], hull.simplices[:, [2, 0]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    h_hull = np.linalg.norm(P[e[:, 0]] - P[e[:, 1]], axis=1).mean()
    assert abs(h / h_hull - 1) < 0.03
    # any positive rescaling of the normals leaves the weights unchanged (only the tangent plane matters)
    a2, h2, _ = shm3d.point_weights(P, 3.0 * P)
    assert np.allclose(a2, areas, rtol=1e-9) and abs(h2 - h) < 1e-12


def test_bunny_pc_weights_and_errors():
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    P, N = d["P"], d["N"]
    areas, h, ntri = shm3d.point_weights(P, N)
    assert np.isfinite(areas).all() and (areas > 0).mean() > 0.99 and ntri > 4 * len(P)
    nn = cKDTree(P).query(P, k=2)[0][:, 1].mean()
    assert 0.8 * nn < h < 2.5 * nn      # a mean Delaunay edge is somewhat longer than the mean nearest-neighbour distance
    with pytest.raises(shm3d.Shm3dError):
        shm3d.point_weights(P[:20], N[:20])   # k + 1 = 31 > 20 points (knn.cpp:53)
    bad = N.copy()
    bad[3, 0] = np.nan
    with pytest.raises(shm3d.Shm3dError):
        shm3d.point_weights(P, bad)


def test_knn_is_exact():
    rng = np.random.default_rng(2)
    P = rng.standard_normal((1500, 3)) * np.array([1.0, 0.3, 2.0])
    Nn = np.tile([0.0, 0.0, 1.0], (len(P), 1))
    a30, h30, _ = shm3d.point_weights(P, Nn, k=30)
    # planar projection with a common normal: the soup is made of 2-D Delaunay stars; compare one point's ring with the
    # ring built from its TRUE 30 nearest neighbours
    tree = cKDTree(P)
    for i in (0, 17, 733):
        idx = tree.query(P[i], k=31)[1][1:]
        v = P[idx] - P[i]
        v = v - np.outer(v @ Nn[i], Nn[i])
        # tangent basis of (0,0,1): basisX = cross((1,0,0), n) normalised = (0,-1,0); basisY = cross(n, basisX) = (1,0,0)
        coords = np.stack([-v[:, 1], v[:, 0]], axis=1)
        ring, tri = shm3d.debug_local_ring(coords)
        assert len(ring) >= 3


def test_tufted_cover_invariants():
    """The intrinsic flips preserve the total area of the cover and end with every edge Delaunay (cotan weight >= -1e-6,
    simple_idt.cpp); the weights are exactly the thirds of the final face areas."""
    rng = np.random.default_rng(4)
    P = fib_points(3000) + 0.004 * rng.standard_normal((3000, 3))      # noisy sphere: the local stars disagree
    Nn = P / np.linalg.norm(P, axis=1, keepdims=True)
    areas, h, ntri, d = shm3d.point_weights(P, Nn, diagnostics=True)
    assert d["flips"] > 0
    assert abs(areas.sum() / d["area_before"] - 1) < 1e-10
    assert d["min_cotan"] >= -1e-6
    assert (areas > 0).all()


def test_bunny_cloud_mean_edge_length_matches_its_mesh():
    """bunny.pc holds the vertices of bunny_small.obj: the tufted triangulation's mean intrinsic edge length must be
    close to the mesh's mean edge length (0.0950) -- it sets lambda for the point overload."""
    from conftest import load_golden
    from oracle import shm_oracle as o
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    z, F = load_golden("bunny_small")
    assert np.abs(d["P"] - z["V"]).max() < 1e-5
    areas, h, ntri, diag = shm3d.point_weights(d["P"], d["N"], diagnostics=True)
    assert abs(h / o.mesh_sources(z["V"], F)["h"] - 1) < 0.02
    assert diag["min_cotan"] >= -1e-6 and abs(areas.sum() / diag["area_before"] - 1) < 1e-10
    # six copies of the surface (3 local stars x 2 sheets), give or take the disagreement between neighbouring stars
    assert 0.9 < areas.sum() / (6 * o.mesh_sources(z["V"], F)["area"].sum()) < 1.25


def _dual_areas(P, tris):
    a = np.zeros(len(P))
    for t in tris:
        p0, p1, p2 = P[t[0]], P[t[1]], P[t[2]]
        A = 0.5 * np.linalg.norm(np.cross(p1 - p0, p2 - p0))
        a[list(t)] += A / 3
    return a


def test_cover_of_a_delaunay_manifold_mesh_needs_no_flips():
    """A closed manifold Delaunay mesh (convex hull of sphere points): the cover is two glued copies, nothing to flip,
    vertex areas = 2 x the mesh's barycentric dual areas, mean edge length unchanged."""
    P = fib_points(800)
    hull = ConvexHull(P)
    tris = hull.simplices.astype(np.int64)
    a, b, c = P[tris[:, 0]], P[tris[:, 1]], P[tris[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    areas, h, d = shm3d.debug_tufted_weights(P, tris)
    assert d["flips"] == 0
    assert np.abs(areas - 2 * _dual_areas(P, tris)).max() < 1e-9       # (mollification adds ~1e-7 relative)
    e = np.unique(np.sort(np.vstack([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), axis=1), axis=0)
    assert abs(h - np.linalg.norm(P[e[:, 0]] - P[e[:, 1]], axis=1).mean()) < 1e-6


def test_intrinsic_flips_recover_the_planar_delaunay_triangulation():
    """A planar triangulation spoiled by random edge flips has the same intrinsic (flat) metric as the Delaunay
    triangulation of its vertices, which is unique in general position: after the intrinsic flips on the cover the vertex
    areas must equal 2 x the Delaunay dual areas, whatever the starting triangulation, boundary included."""
    rng = np.random.default_rng(9)
    pts = rng.uniform(0, 1, size=(300, 2))
    D = Delaunay(pts)
    tris = D.simplices.astype(np.int64).copy()
    # spoil it: flip interior edges whose quadrilateral is strictly convex, a few hundred times
    def orient(t):
        p = pts[t]
        return (p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[1, 1] - p[0, 1]) * (p[2, 0] - p[0, 0])
    for _ in range(400):
        edge_faces = {}
        for fi, t in enumerate(tris):
            for k in range(3):
                edge_faces.setdefault(tuple(sorted((t[k], t[(k + 1) % 3]))), []).append(fi)
        inner = [(e, f) for e, f in edge_faces.items() if len(f) == 2]
        e, (f0, f1) = inner[rng.integers(len(inner))]
        o0 = [v for v in tris[f0] if v not in e][0]
        o1 = [v for v in tris[f1] if v not in e][0]
        t0, t1 = np.array([o0, e[0], o1]), np.array([o0, o1, e[1]])
        if orient(t0) * orient(t1) <= 1e-12:      # not strictly convex: skip
            continue
        if orient(t0) < 0:
            t0, t1 = t0[[0, 2, 1]], t1[[0, 2, 1]]
        tris[f0], tris[f1] = t0, t1
    P3 = np.column_stack([pts, np.zeros(len(pts))])
    assert abs(_dual_areas(P3, tris).sum() - _dual_areas(P3, D.simplices).sum()) < 1e-12   # still a triangulation of the hull
    areas, h, d = shm3d.debug_tufted_weights(P3, tris)
    assert d["flips"] > 50 and d["min_cotan"] >= -1e-6
    assert abs(areas.sum() / d["area_before"] - 1) < 1e-10
    ref = 2 * _dual_areas(P3, D.simplices.astype(np.int64))
    rel = np.abs(areas - ref) / ref.max()
    db = np.minimum(pts, 1 - pts).min(axis=1)          # distance to the boundary of the square
    # Away from the boundary the planar Delaunay areas come back (to the ~1e-4 the mollification of the skinny spoiled
    # triangles perturbs the metric by).  Along the boundary the cover folds front onto back -- the double of the convex
    # hull, with cone points at the hull vertices -- and the intrinsic Delaunay triangulation of THAT surface legitimately
    # flips obtuse boundary triangles across the fold: only vertices close to the boundary may differ.
    assert (db > 0.15).sum() > 100
    assert rel[db > 0.15].max() < 5e-4
    assert (db[rel > 5e-4] < 0.12).all()
    # uniqueness: starting from the unspoiled Delaunay triangulation gives the same final cover, vertex by vertex
    areas0, _, d0 = shm3d.debug_tufted_weights(P3, D.simplices.astype(np.int64))
    assert d0["flips"] > 0                               # the flips across the boundary fold
    assert (np.abs(areas - areas0) / ref.max()).max() < 2e-3


# ---------------------------------------------------------------------------------------------------- GPU: config[2]
@pytest.mark.gpu
def test_gpu_point_overload_with_tufted_weights_matches_oracle(gpu_ctx):
    """BASELINE config[2] (data/bunny.pc, point overload) with the row-N1 weights instead of the surrogate: at 32^3 against
    the fp64 oracle fed the same weights (tolerance of the north star: 1e-4 relative L2), then the 256^3 configuration
    through the class mirror with its automatic weights."""
    from oracle import shm_oracle as o
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    P, N = d["P"], d["N"]
    areas, h, _ = shm3d.point_weights(P, N)
    c = P.mean(axis=0)
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    ref = o.compute_distance(P, N, areas, h, c, r, hCoef=1, scrub_nonfinite=False)
    solver = shm3d.SignedHeatGridSolver(context=gpu_ctx)
    phi = np.array(solver.computeDistancePoints(P, N, options=shm3d.SignedHeat3DOptions(hCoef=1)))   # weights: N1
    assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-4
    # ... and against the reference's own end-to-end output (its grid solver linked with the real geometry-central,
    # geometry-central's own weights; fixture made by tests/golden/make_golden_point_weights.py)
    gold = np.load(os.path.join(GOLDEN, "point_weights_gc.npz"))["bunny_pc_h1_phi"]
    assert np.linalg.norm(phi - gold) / np.linalg.norm(gold) < 1e-4
    phi = solver.computeDistancePoints(P, N, options=shm3d.SignedHeat3DOptions(hCoef=4))
    p, st = solver.params, solver.stats
    assert p.nx == 256 and np.isfinite(phi).all() and abs(p.lambda_ - 1.0 / h) < 1e-9 / h
    g = o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)
    v = o.evaluate_function(g, np.asarray(phi), P)
    assert abs((areas * v).sum() / areas.sum()) < 1e-5                  # the shift (:216-217), area-weighted with the N1 areas
    src, _, _ = shm3d.debug_constraints(p, P)
    assert len(src) == st.m_constraints
    assert np.abs(v[src] + st.shift).max() < 2e-4 * np.abs(phi).max()   # pinned cells interpolate to the common level
    assert phi[0] > 0 and phi[-1] > 0 and o.evaluate_function(g, np.asarray(phi), c[None, :])[0] < 0


def test_kdtree_port_equals_independent_knn_where_no_distances_tie():
    """Cross-check of the nanoflann restatement: an independent cell-list kNN (ties by point index) must give the very same
    weights on clouds without exactly equidistant neighbours -- random samples and the bunny cloud.  (On data/SprayBottle.pc,
    a structured mesh's vertices, the two differ: phi moves by 1.8e-4, which is why the port is the product path.)"""
    L = shm3d.lib()
    rng = np.random.default_rng(5)
    clouds = []
    X = rng.standard_normal((6000, 3))
    X /= np.linalg.norm(X, axis=1)[:, None]
    clouds.append((X * [1.0, 0.8, 1.3], X / [1.0, 0.8, 1.3]))
    d = np.load(os.path.join(GOLDEN, "bunny_pc.npz"))
    clouds.append((d["P"], d["N"]))
    try:
        for P, N in clouds:
            L.shm3d_debug_knn_mode(0)
            a0, h0, nt0 = shm3d.point_weights(P, N)
            L.shm3d_debug_knn_mode(1)
            a1, h1, nt1 = shm3d.point_weights(P, N)
            assert nt0 == nt1 and h0 == h1 and np.array_equal(a0, a1)
    finally:
        L.shm3d_debug_knn_mode(0)
