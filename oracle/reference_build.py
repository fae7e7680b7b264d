"""
oracle/reference_build.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/_ref/libshm_ref.so: the REFERENCE's own src/signed_heat_grid_solver.cpp and
src/signed_heat_3d.cpp, compiled unmodified (recipe: oracle/Makefile) against oracle/ref_shim -- a stand-in for the
slices of geometry-central / Eigen / polyscope those files use.  Everything first-party on the hot path (Step 1-2 loops,
laplacian(), gradient(), constraint selection, trilinear weights, shift, integrateGreedily) therefore runs as written;
the one thing the shim cannot provide is Eigen's SparseLU behind solveSquare, which is replaced by a callback that
solves the assembled KKT system with scipy's SuperLU.

Only tests/ (and bench.py's CPU-reference legs) may import this module.  The library exists only where it was built
from /root/reference (this container); it travels to the GPU box as a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libshm_ref.so")
# the same C entry points around the PRODUCT's drop-in translation unit (adapter/signed_heat_grid_solver_b200.cpp compiled
# against the reference's unchanged headers and linked to libshm3d_grid.so): needs a GPU at run time
ADAPTER_LIB_PATH = os.path.join(_HERE, "_ref", "libshm_adapter.so")
# the marching-cubes routine of the reference's downstream consumer (row N3): polyscope's vendored MarchingCube/MC.h + glm
MC_LIB_PATH = os.path.join(_HERE, "_ref", "libshm_mc_ref.so")
# geometry-central's own point-cloud pipeline (row N1), compiled from the reference's vendored sources against an Eigen
# interface stub (oracle/ref_shim/eigen_stub)
GC_LIB_PATH = os.path.join(_HERE, "_ref", "libshm_gc_ref.so")
# the product's drop-in TU compiled against the REAL geometry-central sources (Eigen / polyscope stubbed); needs a GPU
ADAPTER_GC_LIB_PATH = os.path.join(_HERE, "_ref", "libshm_adapter_gc.so")
# the REFERENCE's own src/signed_heat_grid_solver.cpp + src/signed_heat_3d.cpp against the real geometry-central sources
# (Eigen: stub with a working sparse container, SparseLU -> callback; polyscope: stub) -- CPU
REF_GC_LIB_PATH = os.path.join(_HERE, "_ref", "libshm_ref_gc.so")
REF_ROOT = "/root/reference"
_LIB = None

SOLVE_FN = C.CFUNCTYPE(None, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                       C.POINTER(C.c_double), C.POINTER(C.c_double))


def build(force: bool = False) -> bool:
    """(Re)build oracle/_ref/libshm_ref.so when the reference tree is present; returns whether the library exists."""
    if os.path.isdir(os.path.join(REF_ROOT, "src")) and (force or not os.path.exists(LIB_PATH)
                                                         or not os.path.exists(ADAPTER_LIB_PATH)
                                                         or not os.path.exists(MC_LIB_PATH)
                                                         or not os.path.exists(GC_LIB_PATH)
                                                         or not os.path.exists(ADAPTER_GC_LIB_PATH)
                                                         or not os.path.exists(REF_GC_LIB_PATH)):
        subprocess.check_call(["make", "-j4", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


_ADAPTER = None


def use_adapter(flag: bool):
    """Route the compute_distance_* calls below to the product's drop-in TU (GPU) instead of the reference's source."""
    global _LIB, _ADAPTER
    _ADAPTER = bool(flag)
    _LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(ADAPTER_LIB_PATH if _ADAPTER else LIB_PATH)
        dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_float)
        L.ref_compute_distance_mesh.argtypes = [dp, C.c_int64, ip, ip, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int,
                                                SOLVE_FN, dp, C.c_int64, ip, fp]
        L.ref_compute_distance_points.argtypes = [dp, dp, dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                                  C.c_int, SOLVE_FN, dp, C.c_int64, ip, fp]
        L.ref_mesh_scalars.argtypes = [dp, C.c_int64, ip, ip, C.c_int64, dp, dp, dp, dp, dp]
        L.ref_yukawa.argtypes = [dp, dp, C.c_double]
        L.ref_yukawa.restype = C.c_double
        L.ref_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _flatten(faces):
    off = np.zeros(len(faces) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f) for f in faces])
    return np.asarray([v for f in faces for v in f], dtype=np.int64), off


class _Solver:
    """The stand-in for Eigen::SparseLU: scipy SuperLU on the column-compressed matrix the reference assembled."""

    def __init__(self):
        self.calls = []
        self.fn = SOLVE_FN(self._solve)

    def _solve(self, n, nnz, colptr, rowidx, val, rhs, x):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        cp = np.ctypeslib.as_array(colptr, shape=(n + 1,)).copy()
        ri = np.ctypeslib.as_array(rowidx, shape=(nnz,)).copy()
        v = np.ctypeslib.as_array(val, shape=(nnz,)).copy()
        b = np.ctypeslib.as_array(rhs, shape=(n,)).copy()
        A = sp.csc_matrix((v, ri, cp), shape=(n, n))
        self.calls.append(dict(n=int(n), nnz=int(nnz), A=A, rhs=b))
        np.ctypeslib.as_array(x, shape=(n,))[:] = spla.splu(A).solve(b)


class _AssembleOnly(_Solver):
    """Records the system the reference assembled and returns x = 0 without factorising it: lets a test look at the KKT
    matrix and right-hand side of grids far too large for the LU."""

    def _solve(self, n, nnz, colptr, rowidx, val, rhs, x):
        import scipy.sparse as sp
        cp = np.ctypeslib.as_array(colptr, shape=(n + 1,)).copy()
        ri = np.ctypeslib.as_array(rowidx, shape=(nnz,)).copy()
        v = np.ctypeslib.as_array(val, shape=(nnz,)).copy()
        self.calls.append(dict(n=int(n), nnz=int(nnz), A=sp.csc_matrix((v, ri, cp), shape=(n, n)),
                               rhs=np.ctypeslib.as_array(rhs, shape=(n,)).copy()))
        np.ctypeslib.as_array(x, shape=(n,))[:] = 0.0


def ref_gc_assemble_mesh(V, faces, tCoef=1.0, hCoef=0.0, scale=2.0):
    """The KKT matrix and right-hand side the REFERENCE (on the real geometry-central) hands to solveSquare."""
    s = _AssembleOnly()
    _gc_harness_mesh(REF_GC_LIB_PATH, "reference (geometry-central)", V, faces, tCoef, hCoef, scale, False, s)
    return s.calls[0]["A"], s.calls[0]["rhs"]


def compute_distance_mesh(V, faces, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False, return_info=False):
    """SignedHeatGridSolver::computeDistance(VertexPositionGeometry&, options) of the reference itself."""
    V = np.ascontiguousarray(V, dtype=np.float64)
    fv, fo = _flatten(faces)
    nx = int(2 * 2.0 ** (hCoef + 3))
    phi = np.empty(nx ** 3)
    dims = np.zeros(3, dtype=np.int64)
    bbox = np.zeros(6, dtype=np.float32)
    s = _Solver()
    rc = lib().ref_compute_distance_mesh(_dp(V), len(V), _ip(fv), _ip(fo), len(fo) - 1, tCoef, hCoef, scale, int(fast), s.fn,
                                         _dp(phi), phi.size, _ip(dims), bbox.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    if return_info:
        return phi, dict(dims=dims, bbox=bbox, solves=s.calls)
    return phi


def compute_distance_points(P, normals, areas, h, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False, return_info=False):
    """computeDistance(PointPositionNormalGeometry&, options) of the reference; the tufted-triangulation quantities it
    reads (vertex dual areas, mean edge length) are the caller's."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    Nn = np.ascontiguousarray(normals, dtype=np.float64)
    A = np.ascontiguousarray(areas, dtype=np.float64)
    nx = int(2 * 2.0 ** (hCoef + 3))
    phi = np.empty(nx ** 3)
    dims = np.zeros(3, dtype=np.int64)
    bbox = np.zeros(6, dtype=np.float32)
    s = _Solver()
    rc = lib().ref_compute_distance_points(_dp(P), _dp(Nn), _dp(A), len(P), float(h), tCoef, hCoef, scale, int(fast), s.fn,
                                           _dp(phi), phi.size, _ip(dims), bbox.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    if return_info:
        return phi, dict(dims=dims, bbox=bbox, solves=s.calls)
    return phi


def mesh_scalars(V, faces):
    """centroid / radius / meanEdgeLength / setFaceVectorAreas of src/signed_heat_3d.cpp."""
    V = np.ascontiguousarray(V, dtype=np.float64)
    fv, fo = _flatten(faces)
    nF = len(fo) - 1
    c = np.zeros(3)
    r = C.c_double()
    h = C.c_double()
    area = np.zeros(nF)
    nrm = np.zeros((nF, 3))
    rc = lib().ref_mesh_scalars(_dp(V), len(V), _ip(fv), _ip(fo), nF, _dp(c), C.byref(r), C.byref(h), _dp(area), _dp(nrm))
    if rc != 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return dict(centroid=c, radius=r.value, h=h.value, area=area, nrm=nrm)


def yukawa(x, y, lam):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return float(lib().ref_yukawa(_dp(x), _dp(y), float(lam)))


# ------------------------------------------------------------------------------------------------ row N3: isosurface
_MC = None


def mc_available() -> bool:
    return os.path.exists(MC_LIB_PATH)


def mc_lib():
    global _MC
    if _MC is None:
        L = C.CDLL(MC_LIB_PATH)
        fp, up, ip = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
        L.ref_isosurface.argtypes = [fp, C.c_float, up, fp, fp, C.c_int, ip, ip]
        L.ref_isosurface_copy.argtypes = [fp, C.c_int64, up, C.c_int64]
        L.ref_mc_table.restype = C.POINTER(C.c_uint64)
        _MC = L
    return _MC


def mc_table():
    """The 256-entry packed case table of MarchingCube/MC.h (low nibble = triangle count, then one nibble per corner)."""
    return np.ctypeslib.as_array(mc_lib().ref_mc_table(), shape=(256,)).copy()


def isosurface(values, isoval, dims, bound_min, bound_max, world=True):
    """registerIsosurfaceAsMesh of the reference's consumer (deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228):
    values = the node scalars (any float type; narrowed to float32 as polyscope stores them), index i + j*nx + k*nx*ny.
    Returns (vertices float32[nV,3], triangles uint32[nT,3]) in the order MC::marching_cube produced them."""
    v = np.ascontiguousarray(np.asarray(values).ravel(), dtype=np.float32)
    d = np.asarray(dims, dtype=np.uint32)
    assert v.size == int(d[0]) * int(d[1]) * int(d[2])
    bmin = np.ascontiguousarray(bound_min, dtype=np.float32)
    bmax = np.ascontiguousarray(bound_max, dtype=np.float32)
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    nv, ni = C.c_int64(), C.c_int64()
    L = mc_lib()
    L.ref_isosurface(v.ctypes.data_as(fp), np.float32(isoval), d.ctypes.data_as(up), bmin.ctypes.data_as(fp),
                     bmax.ctypes.data_as(fp), int(bool(world)), C.byref(nv), C.byref(ni))
    verts = np.empty((nv.value, 3), dtype=np.float32)
    idx = np.empty(ni.value, dtype=np.uint32)
    if L.ref_isosurface_copy(verts.ctypes.data_as(fp), nv.value, idx.ctypes.data_as(up), ni.value) != 0:
        raise RuntimeError("reference marching cubes: capacity")
    return verts, idx.reshape(-1, 3)


# ------------------------------------------------------------------------------------------------ row N1: point weights
_GC = None


def gc_available() -> bool:
    return os.path.exists(GC_LIB_PATH)


def _gc():
    global _GC
    if _GC is None:
        L = C.CDLL(GC_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.gcref_point_weights.argtypes = [dp, dp, C.c_int64, dp, dp, ip, ip]
        L.gcref_mesh_sources.argtypes = [dp, C.c_int64, ip, ip, C.c_int64, dp, dp, dp, dp, dp, dp, ip]
        L.gcref_read_mesh.argtypes = [C.c_char_p, ip, ip, dp, dp, dp, dp, dp, dp]
        L.gcref_last_error.restype = C.c_char_p
        _GC = L
    return _GC


def gc_read_mesh(path):
    """geometry-central's own readSurfaceMesh (src/main.cpp:269) + the host quantities of the grid solver on the result."""
    L = _gc()
    nV, nF = C.c_int64(), C.c_int64()
    c = np.zeros(3)
    r, h = C.c_double(), C.c_double()
    if L.gcref_read_mesh(path.encode(), C.byref(nV), C.byref(nF), _dp(c), C.byref(r), C.byref(h), None, None, None) != 0:
        raise RuntimeError("geometry-central: " + L.gcref_last_error().decode())
    area, nrm, bary = np.zeros(nF.value), np.zeros((nF.value, 3)), np.zeros((nF.value, 3))
    if L.gcref_read_mesh(path.encode(), C.byref(nV), C.byref(nF), _dp(c), C.byref(r), C.byref(h), _dp(area), _dp(nrm),
                         _dp(bary)) != 0:
        raise RuntimeError("geometry-central: " + L.gcref_last_error().decode())
    return dict(n_vertices=nV.value, n_faces=nF.value, centroid=c, radius=r.value, h=h.value, area=area, nrm=nrm, pos=bary)


def gc_point_weights(P, normals):
    """What the reference reads from geometry-central for the point-cloud overload (src/main.cpp:277-285,
    src/signed_heat_grid_solver.cpp:149-151,165), run through geometry-central's own sources: returns
    (vertexDualAreas[nP], meanEdgeLength(tuftedGeom), n_faces, n_edges of the tufted mesh)."""
    L = _gc()
    P = np.ascontiguousarray(P, dtype=np.float64)
    Nn = np.ascontiguousarray(normals, dtype=np.float64)
    areas = np.empty(len(P))
    h = C.c_double()
    nf, ne = C.c_int64(), C.c_int64()
    rc = L.gcref_point_weights(_dp(P), _dp(Nn), len(P), _dp(areas), C.byref(h), C.byref(nf), C.byref(ne))
    if rc != 0:
        raise RuntimeError("geometry-central: " + L.gcref_last_error().decode())
    return areas, h.value, nf.value, ne.value


def gc_mesh_sources(V, faces):
    """Rows a4-a6 through the real geometry-central containers + the reference's src/signed_heat_3d.cpp: centroid, radius,
    mean edge length, and per face (in mesh.faces() order) area, unit normal, barycentre."""
    L = _gc()
    V = np.ascontiguousarray(V, dtype=np.float64)
    fv, fo = _flatten(faces)
    nF = len(fo) - 1
    c = np.zeros(3)
    r, h = C.c_double(), C.c_double()
    area, nrm, bary = np.zeros(nF), np.zeros((nF, 3)), np.zeros((nF, 3))
    ne = C.c_int64()
    rc = L.gcref_mesh_sources(_dp(V), len(V), _ip(fv), _ip(fo), nF, _dp(c), C.byref(r), C.byref(h), _dp(area), _dp(nrm),
                              _dp(bary), C.byref(ne))
    if rc != 0:
        raise RuntimeError("geometry-central: " + L.gcref_last_error().decode())
    return dict(centroid=c, radius=r.value, h=h.value, area=area, nrm=nrm, pos=bary, n_edges=ne.value)


# ------------------------------------------------------------- the drop-in TU behind the real geometry-central (GPU)
_GCAD = {}


def _gcad(path=None):
    path = path or ADAPTER_GC_LIB_PATH
    if path not in _GCAD:
        L = C.CDLL(path)
        dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_float)
        L.gcad_compute_distance_mesh.argtypes = [dp, C.c_int64, ip, ip, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int,
                                                 dp, C.c_int64, ip, fp]
        L.gcad_compute_distance_points.argtypes = [dp, dp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int, dp,
                                                   C.c_int64, ip, fp]
        L.gcad_set_solver.argtypes = [SOLVE_FN]
        L.gcad_last_error.restype = C.c_char_p
        _GCAD[path] = L
    return _GCAD[path]


def _gc_harness_mesh(path, what, V, faces, tCoef, hCoef, scale, fast, solver=None):
    V = np.ascontiguousarray(V, dtype=np.float64)
    fv, fo = _flatten(faces)
    nx = int(2 * 2.0 ** (hCoef + 3))
    phi = np.empty(nx ** 3)
    dims = np.zeros(3, dtype=np.int64)
    bbox = np.zeros(6, dtype=np.float32)
    L = _gcad(path)
    if solver is not None:
        L.gcad_set_solver(solver.fn)
    rc = L.gcad_compute_distance_mesh(_dp(V), len(V), _ip(fv), _ip(fo), len(fo) - 1, tCoef, hCoef, scale, int(fast), _dp(phi),
                                      phi.size, _ip(dims), bbox.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError(what + ": " + L.gcad_last_error().decode())
    return phi, dims, bbox


def _gc_harness_points(path, what, P, normals, tCoef, hCoef, scale, fast, solver=None):
    P = np.ascontiguousarray(P, dtype=np.float64)
    Nn = np.ascontiguousarray(normals, dtype=np.float64)
    nx = int(2 * 2.0 ** (hCoef + 3))
    phi = np.empty(nx ** 3)
    dims = np.zeros(3, dtype=np.int64)
    bbox = np.zeros(6, dtype=np.float32)
    L = _gcad(path)
    if solver is not None:
        L.gcad_set_solver(solver.fn)
    rc = L.gcad_compute_distance_points(_dp(P), _dp(Nn), len(P), tCoef, hCoef, scale, int(fast), _dp(phi), phi.size,
                                        _ip(dims), bbox.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise RuntimeError(what + ": " + L.gcad_last_error().decode())
    return phi, dims, bbox


def gc_adapter_compute_distance_mesh(V, faces, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False):
    """adapter/signed_heat_grid_solver_b200.cpp driven like src/main.cpp drives the class, on a real geometry-central
    SurfaceMesh / VertexPositionGeometry (GPU).  Returns (phi, dims, bbox)."""
    return _gc_harness_mesh(ADAPTER_GC_LIB_PATH, "adapter (geometry-central)", V, faces, tCoef, hCoef, scale, fast)


def gc_adapter_compute_distance_points(P, normals, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False):
    """The point-cloud overload of the drop-in TU on a real PointPositionNormalGeometry (geometry-central's own
    tufted-cover weights; GPU)."""
    return _gc_harness_points(ADAPTER_GC_LIB_PATH, "adapter (geometry-central)", P, normals, tCoef, hCoef, scale, fast)


def ref_gc_available() -> bool:
    return os.path.exists(REF_GC_LIB_PATH)


def ref_gc_compute_distance_mesh(V, faces, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False, return_info=False):
    """The REFERENCE's computeDistance(VertexPositionGeometry&) on the real geometry-central (CPU; KKT solve: scipy
    SuperLU through the Eigen stub's SparseLU)."""
    s = _Solver()
    phi, dims, bbox = _gc_harness_mesh(REF_GC_LIB_PATH, "reference (geometry-central)", V, faces, tCoef, hCoef, scale, fast, s)
    return (phi, dict(dims=dims, bbox=bbox, solves=s.calls)) if return_info else phi


def ref_gc_compute_distance_points(P, normals, tCoef=1.0, hCoef=0.0, scale=2.0, fast=False, return_info=False):
    """The REFERENCE's computeDistance(PointPositionNormalGeometry&) end to end on the real geometry-central: its own
    tufted-cover weights, its own grid solver."""
    s = _Solver()
    phi, dims, bbox = _gc_harness_points(REF_GC_LIB_PATH, "reference (geometry-central)", P, normals, tCoef, hCoef, scale, fast, s)
    return (phi, dict(dims=dims, bbox=bbox, solves=s.calls)) if return_info else phi
