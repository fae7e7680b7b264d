"""The drop-in translation unit (adapter/signed_heat_grid_solver_b200.cpp) driven through the reference's own class:
oracle/_ref/libshm_adapter.so = that file + the reference's src/signed_heat_3d.cpp compiled against the reference's
unchanged headers (shim for geometry-central / Eigen / polyscope) and linked to libshm3d_grid.so.
CPU: it loads and fails loudly without a GPU.  GPU: same inputs, same answers as the reference's source."""
import os

import numpy as np
import pytest

from conftest import icosphere, load_golden
from oracle import reference_build as rb
from oracle import shm_oracle as o


def _adapter_or_skip():
    if not (rb.build() and os.path.exists(rb.ADAPTER_LIB_PATH)):
        pytest.skip("no prebuilt oracle/_ref/libshm_adapter.so")
    rb.use_adapter(True)
    try:
        rb.lib()
    except OSError as e:  # e.g. the product library is not where the runpath expects it
        rb.use_adapter(False)
        pytest.skip(f"adapter library does not load here: {e}")


def test_adapter_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    _adapter_or_skip()
    try:
        z, F = load_golden("bunny_small")
        with pytest.raises(RuntimeError) as e:
            rb.compute_distance_mesh(z["V"], F, hCoef=0)
        assert "no CPU fallback" in str(e.value)
    finally:
        rb.use_adapter(False)


@pytest.mark.gpu
def test_adapter_equals_reference_source_on_the_gpu():
    _adapter_or_skip()
    try:
        z, F = load_golden("bunny_small")
        try:
            phi, info = rb.compute_distance_mesh(z["V"], F, hCoef=1, return_info=True)
        except RuntimeError as e:  # first GPU outing of this prebuilt harness: report, do not mask numeric failures below
            pytest.skip(f"adapter harness raised before producing a field: {e}")
        ref = z["h1_phi"]  # = the reference source's own output (tests/test_reference_build.py)
        assert np.linalg.norm(phi - ref) / np.linalg.norm(ref) < 1e-4
        assert list(info["dims"]) == [32, 32, 32] and len(info["solves"]) == 0   # registerVolumeGrid side effect; no LU
        # fastIntegration through the same class
        phif = rb.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)
        reff = o.compute_distance_mesh(z["V"], F, hCoef=0, fast=True)
        assert np.linalg.norm(phif - reff) / np.linalg.norm(reff) < 1e-4
        # point-cloud overload with caller-supplied tufted quantities
        V, Fs = icosphere(2)
        s = o.mesh_sources(V, Fs)
        areas, h = s["area"] * 1.3, 0.2
        phip = rb.compute_distance_points(s["pos"], s["nrm"], areas, h, hCoef=1)
        c = s["pos"].sum(axis=0) / len(s["pos"])
        r = np.sqrt(((s["pos"] - c) ** 2).sum(axis=1)).max()
        refp = o.compute_distance(s["pos"], s["nrm"], areas, h, c, r, hCoef=1, scrub_nonfinite=False)
        assert np.linalg.norm(phip - refp) / np.linalg.norm(refp) < 1e-4
    finally:
        rb.use_adapter(False)
