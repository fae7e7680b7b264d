"""Deterministic synthetic inputs for benchmarks and large-size tests (no RNG, no files).

fibonacci_sphere(n_tri): unit sphere with exactly n_tri outward-oriented, near-uniform triangles -- the convex hull of
(n_tri + 4) / 2 Fibonacci-lattice points (a hull of N points in general position on a sphere has 2N - 4 facets).
SURVEY.md section 8(d) input 5: 1e5 triangles, radius 1, centre 0 -> h ~ 0.017, lambda ~ 58.6."""
import numpy as np


def fibonacci_sphere(n_tri=100000, radius=1.0, center=(0.0, 0.0, 0.0)):
    from scipy.spatial import ConvexHull
    assert n_tri % 2 == 0 and n_tri >= 8
    n = (n_tri + 4) // 2
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    P = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)
    hull = ConvexHull(P)
    F = hull.simplices.astype(np.int64)
    assert len(F) == n_tri, (len(F), n_tri)
    # outward orientation
    a, b, c = P[F[:, 0]], P[F[:, 1]], P[F[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    F[flip] = F[flip][:, [0, 2, 1]]
    # deterministic face order (ConvexHull's order is an implementation detail): sort by barycentre Morton-free key
    bary = (P[F[:, 0]] + P[F[:, 1]] + P[F[:, 2]]) / 3
    order = np.lexsort((bary[:, 0], bary[:, 1], bary[:, 2]))
    return P * radius + np.asarray(center, dtype=np.float64), F[order]
