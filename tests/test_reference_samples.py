"""The reference's own sample data (data/*.obj, data/*.pc) through the reference's own code -- its grid solver linked with
the real geometry-central (oracle/_ref/libshm_ref_gc.so), driven like src/main.cpp at its default resolution (hCoef 0,
16^3) -- against the oracle fed by the PRODUCT's host half (shm3d_prepare_mesh; shm3d_point_weights).  Runs only where the
reference tree is present (this container); elsewhere the committed fixtures carry the same information."""
import os

import numpy as np
import pytest

import shm3d
from oracle import reference_build as rb
from oracle import shm_oracle as o

DATA = "/root/reference/data"
pytestmark = pytest.mark.skipif(not (os.path.isdir(DATA) and rb.build() and rb.ref_gc_available()),
                                reason="needs the reference tree and oracle/_ref/libshm_ref_gc.so")


@pytest.mark.parametrize("name", ["bunny_small", "polygon-bear", "chair", "rocker", "knot"])
def test_mesh_samples_end_to_end(name):
    V, F = o.read_obj(os.path.join(DATA, name + ".obj"))
    phi = rb.ref_gc_compute_distance_mesh(V, F, hCoef=0)
    p, pos, nrm, area, h = shm3d.prepare_mesh(V, F, hCoef=0)                 # the product's host half
    s = o.mesh_sources(V, F)
    assert np.array_equal(pos, s["pos"]) and abs(h - s["h"]) < 1e-13 * h
    ref = o.compute_distance(pos, nrm, area, h, s["centroid"], s["radius"], hCoef=0)
    assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-11
    phif = rb.ref_gc_compute_distance_mesh(V, F, hCoef=0, fast=True)
    reff = o.compute_distance(pos, nrm, area, h, s["centroid"], s["radius"], hCoef=0, fast=True)
    assert np.linalg.norm(phif - reff) / np.linalg.norm(reff) < 1e-11


@pytest.mark.parametrize("name", ["bunny", "chair", "rocker", "knot", "SprayBottle"])
def test_point_cloud_samples_end_to_end(name):
    """geometry-central's own tufted-cover weights + the reference's solver vs the product's weights + the oracle."""
    P, N = o.read_pc(os.path.join(DATA, name + ".pc"))
    phi = rb.ref_gc_compute_distance_points(P, N, hCoef=0)
    areas, h, _ = shm3d.point_weights(P, N)
    c = P.mean(axis=0)
    r = np.sqrt(((P - c) ** 2).sum(axis=1)).max()
    ref = o.compute_distance(P, N, areas, h, c, r, hCoef=0, scrub_nonfinite=False)
    assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-11
